"""TEST INFRASTRUCTURE -- the DISTRIBUTION of the reference's stochastic estimator (run in the build container):

    python -m oracle.make_golden_meanimage

The CUDA path draws keyed random numbers, the reference draws from torch's global generator: single renders can only be
compared through the oracle (KeyedRNG vs TorchRNG modes).  This fixture pins the estimator itself: the UNMODIFIED reference
renders the same 160 rays of tests/golden/microfacet_g40.pt under 48 different seeds; mean and standard deviation of every
radiance map per ray go to tests/golden/microfacet_g40_meanimage.pt.  tests/test_gpu_parity.py compares the mean of 48
keyed-seed renders of the CUDA path against it within Monte-Carlo error."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import make_golden  # noqa: E402

GOLDEN = make_golden.GOLDEN_DIR
KEYS = ("rgb_map", "spec", "diffuse", "tint")

if __name__ == "__main__":
    fix = torch.load(os.path.join(GOLDEN, "microfacet_g40.pt"), weights_only=False)
    meta = dict(aabb=fix["aabb"], near_far=fix["near_far"], grid_size=[fix["grid_size"]] * 3, bg_resolution=fix["bg_resolution"])
    ref = make_golden.load_scene_into_reference(fix["state"], meta, "microfacet_tensorf2")
    assert torch.equal(ref.sampler.alphaMask.alpha_volume.reshape(-1).to(torch.uint8), fix["alpha_volume"].reshape(-1))
    rays = fix["rays"][:160].contiguous()
    n_seeds = 48
    acc = {k: [] for k in KEYS}
    for s in range(n_seeds):
        ims, _ = make_golden.reference_render(ref, rays, fix["focal"], 1000 + s)
        for k in KEYS:
            acc[k].append(ims[k].float())
    out = dict(n_rays=160, n_seeds=n_seeds, seeds=list(range(1000, 1000 + n_seeds)))
    for k in KEYS:
        st = torch.stack(acc[k])
        out[k + "_mean"], out[k + "_std"] = st.mean(0), st.std(0)
    torch.save(out, os.path.join(GOLDEN, "microfacet_g40_meanimage.pt"))
    print({k: (float(out[k + "_mean"].mean()), float(out[k + "_std"].mean())) for k in KEYS})
