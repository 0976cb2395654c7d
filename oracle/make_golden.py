"""TEST INFRASTRUCTURE -- pins oracle/nmf_oracle.py against the real reference and writes fixtures.

Run inside the build container (needs /root/reference):

    python -m oracle.make_golden            # pin + write tests/golden/*.pt
    python -m oracle.make_golden --full     # additionally pin one 4096-ray chunk at G=300 (no file)

For each case it (1) builds a synthetic scene (nmf_b200/synthetic.py), (2) loads it into the
*unmodified* reference ``TensorNeRF`` (oracle/ref_harness.py), lets the reference build its own
occupancy volume (``updateAlphaMask``) and renders the rays with ``torch.manual_seed(seed)``,
(3) renders the same rays with the restatement using ``TorchRNG`` under the same seed and asserts
agreement, (4) stores inputs + reference outputs as a fixture.  The fixture is what
``tests/test_oracle_golden.py`` replays on machines without the reference.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nmf_b200 import synthetic  # noqa: E402
from oracle import keyed_rng, nmf_oracle, ref_harness  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
IMAGE_KEYS = ["rgb_map", "acc_map", "depth", "world_normal", "normal", "termination_xyz", "surf_width",
              "cross_section", "diffuse", "tint", "roughness", "spec", "albedo"]


def reference_render(model, rays, focal, seed):
    torch.manual_seed(seed)
    with torch.no_grad():
        ims, stats = model(rays, focal, is_train=False, ndc_ray=False, N_samples=-1)
    return ims, stats


def load_scene_into_reference(state, meta, model_name):
    gs = [int(g) for g in meta["grid_size"]]
    t = ref_harness.build_reference_model(meta["aabb"], list(meta["near_far"]), grid_size=[gs[0]] * 3,
                                          bg_resolution=meta["bg_resolution"], model_name=model_name)
    if len(set(gs)) > 1:
        # the reference only builds cubic factors; non-cubic grids arise from its own resolution change
        # (fields/tensoRF.py:408-413 upsample_volume_grid: planes (grid[mat1], grid[mat0]), lines grid[vec])
        t.rf.upsample_volume_grid(torch.tensor(gs))
    missing = t.load_state_dict({k: v for k, v in state.items()}, strict=False)
    bad = [k for k in missing.unexpected_keys]
    assert not bad, bad
    t.sampler.update(t.rf, init=True)
    t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
    t.eval()
    return t


def compare(ref_ims, ref_stats, ims, stats, tol):
    worst = {}
    for k, v in ref_ims.items():
        a, b = v.float(), ims[k].float()
        assert a.shape == b.shape, (k, a.shape, b.shape)
        err = (a - b).abs().max().item() if a.numel() else 0.0
        worst[k] = err
    assert list(ref_stats["n_samples"]) == list(stats["n_samples"]), (ref_stats["n_samples"], stats["n_samples"])
    bad = {k: e for k, e in worst.items() if e > tol.get(k, tol["default"])}
    assert not bad, f"oracle disagrees with the reference: {bad}"
    return worst


def plain_state(meta, seed):
    """weights of MLPRender_Fea(viewpe=2, feape=2, featureC=128) for the model=tensorf plumbing case"""
    g = torch.Generator().manual_seed(77 + seed)
    dims = [(128, 24 + 3 + 96 + 12), (128, 128), (3, 128)]
    st = {}
    for li, (o, i) in zip((0, 2, 4), dims):
        st[f"model.diffuse_module.mlp.{li}.weight"] = (torch.rand(o, i, generator=g) * 2 - 1) / (i ** 0.5)
        st[f"model.diffuse_module.mlp.{li}.bias"] = torch.zeros(o)
    return st


def run_case(name, scene_name, G, bg_res, n_rays, crop, model_name, seed, write, full_density=False):
    state, meta = synthetic.make_scene(scene_name, grid_size=G, bg_resolution=bg_res, full_density=full_density)
    if model_name == "tensorf":
        state = {k: v for k, v in state.items() if not k.startswith("model.")}
        state.update(plain_state(meta, seed))
    ref = load_scene_into_reference(state, meta, model_name)
    poses = synthetic.hemisphere_poses(4, seed=1)
    focal = synthetic.focal_for(800)
    rays = synthetic.camera_rays(poses[1], 800, 800, focal, crop=crop)
    g = torch.Generator().manual_seed(5)
    rays = rays[torch.randperm(rays.shape[0], generator=g)[:n_rays]].contiguous()
    # edge cases the domain has: an axis-parallel ray (d == 0 components) and a ray that misses the box
    rays[0] = torch.tensor([0.3, -4.0, 0.2, 0.0, 1.0, 0.0])
    rays[1] = torch.tensor([5.0, 5.0, 5.0, 0.0, 0.0, 1.0])
    t0 = time.time()
    ref_ims, ref_stats = reference_render(ref, rays, focal, seed)
    t_ref = time.time() - t0
    alpha = ref.sampler.alphaMask.alpha_volume.detach().clone()
    hp = dict(model="microfacet" if model_name == "microfacet_tensorf2" else "plain")
    sc = nmf_oracle.Scene(state, meta["aabb"], meta["near_far"], meta["grid_size"], alpha_volume=alpha, **hp)
    assert sc.n_samples == ref.sampler.nSamples and float(sc.stepsize) == float(ref.sampler.stepsize)
    torch.manual_seed(seed)
    t0 = time.time()
    ims, stats = nmf_oracle.render_chunk(sc, rays, focal, keyed_rng.TorchRNG())
    t_or = time.time() - t0
    tol = dict(default=2e-5, termination_xyz=1e-6, surf_width=0, depth=1e-4, spec=5e-5)   # spec: O(1-3) radiance means
    worst = compare(ref_ims, ref_stats, ims, stats, tol)
    # the oracle's own occupancy rebuild must equal the reference's, voxel for voxel
    mine = nmf_oracle.build_alpha_volume(sc)
    assert torch.equal(mine.reshape(-1), alpha.reshape(-1)), "occupancy volume mismatch"
    print(f"[{name}] rays={n_rays} G={G} n_samples={ref_stats['n_samples']} ref {t_ref:.2f}s oracle {t_or:.2f}s")
    print("   max |oracle - reference| per map:", {k: f"{e:.2e}" for k, e in worst.items()})
    if write:
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        fix = dict(name=name, scene=scene_name, grid_size=G, bg_resolution=bg_res, model=model_name, seed=seed,
                   focal=focal, rays=rays, state={k: v.clone() for k, v in state.items()},
                   aabb=meta["aabb"], near_far=meta["near_far"], alpha_volume=alpha.to(torch.uint8),
                   ref_images={k: v.clone() for k, v in ref_ims.items()}, ref_n_samples=list(ref_stats["n_samples"]),
                   torch_version=torch.__version__, max_abs_err_vs_reference=worst)
        path = os.path.join(GOLDEN_DIR, f"{name}.pt")
        torch.save(fix, path)
        print(f"   wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
    return worst


STAT_KEYS = ["ori_loss", "prediction_loss", "envmap_reg", "brdf_reg", "diffuse_reg", "distortion_loss"]


def run_stats_case(name, write):
    """A19: the reference forward without debug maps (is_train=False, draw_debug=False) returns the regulariser
    inputs (modules/tensor_nerf.py:567-649).  Replays an existing fixture's scene and rays through the reference,
    checks the oracle against it and stores the reference numbers in tests/golden/<name>_stats.pt."""
    fix = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)
    gsz = fix["grid_size"]
    meta = dict(aabb=fix["aabb"], near_far=fix["near_far"], grid_size=[gsz] * 3 if isinstance(gsz, int) else list(gsz),
                bg_resolution=fix["bg_resolution"])
    ref = load_scene_into_reference(fix["state"], meta, fix["model"])
    assert torch.equal(ref.sampler.alphaMask.alpha_volume.reshape(-1).to(torch.uint8), fix["alpha_volume"].reshape(-1))
    torch.manual_seed(fix["seed"])
    with torch.no_grad():
        ims, st = ref(fix["rays"], fix["focal"], is_train=False, ndc_ray=False, N_samples=-1, draw_debug=False)
    hp = dict(model="microfacet" if fix["model"] == "microfacet_tensorf2" else "plain")
    sc = nmf_oracle.Scene(fix["state"], fix["aabb"], fix["near_far"], meta["grid_size"], alpha_volume=fix["alpha_volume"].float(), **hp)
    torch.manual_seed(fix["seed"])
    oi, os_ = nmf_oracle.render_chunk(sc, fix["rays"], fix["focal"], keyed_rng.TorchRNG(), draw_debug=False)
    out = {}
    for k in STAT_KEYS:
        a, b = float(st[k]), float(os_[k])
        assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), (k, a, b)
        out[k] = a
    assert (ims["rgb_map"] - oi["rgb_map"]).abs().max() <= 2e-5
    print(f"[{name}] statistics (reference == oracle):", {k: f"{v:.6g}" for k, v in out.items()})
    if write:
        torch.save(dict(name=name, ref_statistics=out, n_rays=int(fix["rays"].shape[0])), os.path.join(GOLDEN_DIR, f"{name}_stats.pt"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--no-write", action="store_true")
    ap.add_argument("--stats-only", action="store_true", help="only (re)generate the <name>_stats.pt fixtures")
    ap.add_argument("--only", default="", help="'noncubic': only (re)generate the non-cubic fixture")
    a = ap.parse_args()
    assert ref_harness.available(), "needs /root/reference"
    w = not a.no_write
    ap_only = a.only
    if ap_only == "noncubic":
        run_case("microfacet_noncubic", "lego", [36, 48, 42], 32, 256, (330, 470, 330, 470), "microfacet_tensorf2", 11, w,
                 full_density=True)
        run_stats_case("microfacet_noncubic", w)
        return
    if a.stats_only:
        for name in ("microfacet_g40", "microfacet_g56_ship", "plain_g64"):
            run_stats_case(name, w)
        return
    run_case("microfacet_g40", "lego", 40, 32, 384, (330, 470, 330, 470), "microfacet_tensorf2", 20211200, w)
    run_case("microfacet_g56_ship", "ship", 56, 48, 256, (300, 500, 300, 500), "microfacet_tensorf2", 7, w)
    run_case("plain_g64", "lego", 64, 32, 4096, (368, 432, 368, 432), "tensorf", 20211200, w)
    run_case("microfacet_noncubic", "lego", [36, 48, 42], 32, 256, (330, 470, 330, 470), "microfacet_tensorf2", 11, w,
             full_density=True)
    for name in ("microfacet_g40", "microfacet_g56_ship", "plain_g64", "microfacet_noncubic"):
        run_stats_case(name, w)
    if a.full:
        run_case("microfacet_g300_full", "lego", 300, 512, 4096, None, "microfacet_tensorf2", 20211200, False)


if __name__ == "__main__":
    main()
