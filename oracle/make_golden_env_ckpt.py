"""TEST INFRASTRUCTURE -- fixtures that need the real reference (run in the build container):

    python -m oracle.make_golden_env_ckpt

  tests/golden/forest_env.pt      lookups of the UNMODIFIED reference IntegralEquirect (modules/integral_equirect.py:409-504)
                                  on backgrounds/forest.th -- the 1024 x 2048 map of BASELINE configs #3 / #5 -- at seeded
                                  directions / mip levels incl. the poles and the +-pi seam, its SH irradiance
                                  (get_spherical_harmonics, :324-360), and the same for a 128 x 256 area-averaged copy of the
                                  map whose state_dict is stored in the fixture (so the GPU box can run a forest-derived
                                  case without the 25 MB file; the full-size case runs when baseline/_ref/backgrounds/forest.th
                                  travelled with the snapshot)
  tests/golden/ref_ckpt_g24.th    a checkpoint WRITTEN BY the reference's own TensorNeRF.save (modules/tensor_nerf.py:120-134)
                                  of a small synthetic scene with calibrated (non-default) biases and an occupancy volume
  tests/golden/ref_ckpt_g24_render.pt   rays + the reference's render of them from that model (deterministic maps)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nmf_b200 import config, synthetic  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def plain(node):
    if isinstance(node, dict):
        return {k: plain(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return [plain(v) for v in node]
    return node


def probe_dirs(n, seed):
    g = torch.Generator().manual_seed(seed)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    special = torch.tensor([[0, 0, 1.0], [0, 0, -1.0], [1e-4, 0, 1.0], [-1.0, 1e-7, 0.0], [-1.0, -1e-7, 0.0], [-1.0, 0, 0.2],
                            [1.0, 0, 0], [0, 1.0, 0], [0.02, 0.01, -0.9997], [-0.7, 1e-5, 0.7]])
    d[:special.shape[0]] = torch.nn.functional.normalize(special, dim=-1)
    mip = torch.rand(n, generator=g) * 11 - 9          # solid angles from sub-texel to a large part of the sphere
    return d.contiguous(), mip.contiguous()


def make_forest():
    ref_harness.install_stubs()
    from modules.integral_equirect import IntegralEquirect
    path = os.path.join(ref_harness.REFERENCE_ROOT, "backgrounds", "forest.th")
    sd = torch.load(path, map_location="cpu", weights_only=False)
    mk = lambda res: IntegralEquirect(bg_resolution=res, mipbias=0, activation="exp", lr=0.001, init_val=-1.897, mul_lr=0.001,
                                      brightness_lr=0, betas=[0.0, 0.0], mul_betas=[0.9, 0.9], mipbias_lr=1e-4, mipnoise=0.0)
    full = mk(1024)
    full.load_state_dict(sd, strict=False)                # forest.th predates the sh_A buffer (train.py loads it the same way)
    dirs, mip = probe_dirs(4096, 3)
    with torch.no_grad():
        out_full = full(dirs, mip.reshape(-1, 1))
        sh_full = full.get_spherical_harmonics(100)[1]
        act = full.activation_fn(full.bg_mat)                                      # (1,3,1024,2048) radiance
        small_act = torch.nn.functional.avg_pool2d(act, kernel_size=8)             # 128 x 256, area average
        small_sd = {"bg_mat": torch.log(small_act.clip(min=1e-8)), "mipbias": sd["mipbias"].clone(),
                    "brightness": torch.tensor(0.0, dtype=torch.float64), "mul": torch.tensor(1.0, dtype=torch.float64)}
        small = mk(128)
        small.load_state_dict(small_sd, strict=False)
        out_small = small(dirs, mip.reshape(-1, 1))
        sh_small = small.get_spherical_harmonics(100)[1]
    fix = dict(dirs=dirs, mip=mip, out_full=out_full, sh_conv_full=sh_full, small_state=small_sd, out_small=out_small,
               sh_conv_small=sh_small, full_scalars={k: float(sd[k]) for k in ("mipbias", "brightness", "mul")},
               torch_version=torch.__version__)
    torch.save(fix, os.path.join(GOLDEN, "forest_env.pt"))
    print("forest_env.pt:", {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in fix.items() if k != "small_state"})


def make_ckpt():
    G, bg = 24, 16
    state, meta = synthetic.make_scene("materials", grid_size=G, bg_resolution=bg)
    t = ref_harness.build_reference_model(meta["aabb"], list(meta["near_far"]), grid_size=[G] * 3, bg_resolution=bg)
    t.load_state_dict(state, strict=False)
    t.sampler.update(t.rf, init=True)
    t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
    t.eval()
    # "calibrated" biases: plain attributes, carried by the checkpoint's config only (tensor_nerf.py:138-146)
    t.model.brdf.bias, t.model.diffuse_module.diffuse_bias, t.model.diffuse_module.roughness_bias = 0.31, -1.07, 0.42
    over = [f"field.grid_size=[{G},{G},{G}]", f"model.arch.bg_module.bg_resolution={bg}"]
    cfg = plain(config.to_plain(config.compose(over).model.arch))
    cfg["model"]["brdf"]["bias"] = 0.31
    cfg["model"]["diffuse_module"]["diffuse_bias"] = -1.07
    cfg["model"]["diffuse_module"]["roughness_bias"] = 0.42
    path = os.path.join(GOLDEN, "ref_ckpt_g24.th")
    t.save(path, cfg)                                                              # the reference's own writer
    focal = synthetic.focal_for(800)
    rays = synthetic.camera_rays(synthetic.hemisphere_poses(4, seed=1)[2], 800, 800, focal)
    rays = rays[torch.randperm(rays.shape[0], generator=torch.Generator().manual_seed(9))[:192]].contiguous()
    torch.manual_seed(4)
    with torch.no_grad():
        ims, stats = t(rays, focal, is_train=False, ndc_ray=False, N_samples=-1)
    keep = ("acc_map", "depth", "world_normal", "albedo", "roughness", "surf_width", "rgb_map", "diffuse")
    torch.save(dict(rays=rays, focal=focal, near_far=list(meta["near_far"]), n_samples=[int(x) for x in stats["n_samples"]],
                    ref_images={k: ims[k].clone() for k in keep}, overrides=over, torch_version=torch.__version__),
               os.path.join(GOLDEN, "ref_ckpt_g24_render.pt"))
    ck = torch.load(path, weights_only=False)
    print("ref_ckpt_g24.th:", sorted(ck.keys()), len(ck["state_dict"]), "state_dict entries;", os.path.getsize(path) // 1024, "KB;",
          "n_samples", stats["n_samples"])


if __name__ == "__main__":
    assert ref_harness.available(), "needs the reference tree (/root/reference)"
    os.makedirs(GOLDEN, exist_ok=True)
    make_forest()
    make_ckpt()
