"""Times the fused training step of model=tensorf (nmf_train_plain) on the synthetic G=300 scene and runs a short
optimiser loop.  Run under gpurun:  python tools/train_bench.py [--rays 4096] [--steps 20] [--iters 60]
Prints one JSON line (rays/s and valid samples/s of the forward+backward step, CUDA-event timed; MSE of the loop)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nmf_b200 import ops, synthetic, train  # noqa: E402
from nmf_b200.scene import DeviceScene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--grid", type=int, default=300)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--iters", type=int, default=60)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    state, meta = synthetic.make_scene("lego", grid_size=a.grid, bg_resolution=32)
    state.update(synthetic.plain_mlp_state(0))
    mk = lambda st, vol=None: DeviceScene(st, meta["aabb"], meta["near_far"], meta["grid_size"], alpha_volume=vol, device=dev,
                                          model="plain")
    sc = mk(state)
    vol = sc.update_alpha_mask()
    H = W = 800
    focal = synthetic.focal_for(W)
    pose = synthetic.hemisphere_poses(4)[1]
    g = torch.Generator().manual_seed(0)
    pix = torch.randperm(H * W, generator=g)[:a.rays]
    rays = synthetic.camera_rays(pose, H, W, focal)[pix].contiguous().to(dev)
    gt = ops.render_rays(sc, rays, focal, chunk=a.rays, skip_eps=0.0, t_cut=0.0)[0]["rgb_map"].clone()
    out = train.train_plain(sc, rays, gt, focal=focal, seed=1)
    bufs = out["buffers"]
    for _ in range(3):
        train.train_plain(sc, rays, gt, focal=focal, seed=1, buffers=bufs, check_errors=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        train.train_plain(sc, rays, gt, focal=focal, seed=1, buffers=bufs, check_errors=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    # short optimiser loop from perturbed appearance weights
    gg = torch.Generator().manual_seed(2)
    st2 = {k: v.clone() for k, v in state.items()}
    for k in train.PLAIN_PARAM_KEYS:
        if ("app_rf" in k) or ("mlp" in k):
            st2[k] = st2[k] + 0.05 * st2[k].abs().mean() * torch.randn(st2[k].shape, generator=gg)
    tr = train.PlainTrainer(st2, meta["aabb"], meta["near_far"], meta["grid_size"], alpha_volume=vol, device=dev)
    mse = [tr.step(rays, gt)["mse"] for _ in range(a.iters)]
    print(json.dumps(dict(what="nmf_train_plain forward+backward, model=tensorf", grid=a.grid, rays=a.rays,
                          n_samples=out["n_samples"], ms_per_step=ms, rays_per_s=a.rays / ms * 1e3,
                          samples_per_s=out["n_samples"] / ms * 1e3, loop_mse_first=mse[0], loop_mse_last=mse[-1],
                          loop_iters=a.iters)))


if __name__ == "__main__":
    main()
