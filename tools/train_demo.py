"""End-to-end check of the training loop for model=tensorf (BASELINE config #1's model): a freshly initialised field and
view MLP (the reference's initialisers, via the plugin mirrors) are fitted with PlainTrainer.fit -- the device-side
restatement of train.py:497-813 -- to images rendered from the synthetic lego scene, with the reference's coarse-to-fine
schedule shape (occupancy updates + upsampling + optimiser re-creation); reports held-out PSNR.
Run under gpurun:  python tools/train_demo.py [--iters 1500] [--res 200] [--views 24]
Prints one JSON line.  Inputs are synthetic (there is no dataset in the container); the teacher is a TensoRF of the same
family, so the fit is realisable."""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nmf_b200 import config, ops, synthetic, train  # noqa: E402
from nmf_b200.scene import DeviceScene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=1500)
    ap.add_argument("--res", type=int, default=200)
    ap.add_argument("--views", type=int, default=24)
    ap.add_argument("--test-views", type=int, default=4)
    ap.add_argument("--teacher-grid", type=int, default=128)
    ap.add_argument("--alpha-thres", type=float, default=4e-4,
                    help="AlphaGridSampler.alphaMask_thres.  The reference culls by 1 - exp(-sigma * step) >= thres (alphagrid.py:222, no "
                         "distance_scale) but renders with 1 - exp(-sigma * step * 25): its default 1e-3 keeps only voxels whose RENDER alpha "
                         "per step is >= 2.5 %%, which deletes the soft Gaussian blobs of the synthetic teacher (PSNR 42 -> 20 dB at the first "
                         "rebuild after an upsampling, and with rebuilds that respect the current mask -- as the reference's do -- it never comes back: "
                         "profiles/r02_c_train_demo_thres1e-3.json); 4e-5 = 1e-3 / 25 keeps the student's fog everywhere (no culling, test PSNR "
                         "21 dB); 4e-4 (render alpha 1 %%) culls the fog and keeps the blobs: held-out 33.6 dB")
    ap.add_argument("--ups", type=str, default="", help="upsampling iterations, comma separated (default: 20/40/60 %% of --iters)")
    ap.add_argument("--upd", type=str, default="", help="occupancy rebuild iterations (default: 13/30/50/70 %% of --iters)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    # ---- teacher: the synthetic lego field under a model=tensorf view MLP, rendered by the eval path ----
    state, meta = synthetic.make_scene("lego", grid_size=a.teacher_grid, bg_resolution=32)
    state.update(synthetic.plain_mlp_state(0))
    teacher = DeviceScene(state, meta["aabb"], meta["near_far"], meta["grid_size"], device=dev, model="plain")
    teacher.update_alpha_mask()
    H = W = a.res
    focal = synthetic.focal_for(W)
    poses = synthetic.hemisphere_poses(a.views + a.test_views)
    rays, rgbs = [], []
    for pose in poses:
        r = synthetic.camera_rays(pose, H, W, focal).to(dev)
        im, _ = ops.render_rays(teacher, r, focal, chunk=4096, skip_eps=0.0, t_cut=0.0)
        rays.append(r)
        rgbs.append(im["rgb_map"].clone())
    train_rays, train_rgbs = torch.cat(rays[:a.views]), torch.cat(rgbs[:a.views])
    # ---- student: fresh initialisation through the plugin mirrors (reference initialisers), 64^3 start ----
    torch.manual_seed(20211200)                                     # configs/default.yaml:35
    t, _ = config.build_model(["model=tensorf", "field.grid_size=[64,64,64]"], aabb=meta["aabb"], near_far=list(meta["near_far"]))
    init = {k: v.detach().clone() for k, v in t.state_dict().items()}
    n_it = a.iters
    ups = [int(v) for v in a.ups.split(",")] if a.ups else [int(n_it * f) for f in (0.2, 0.4, 0.6)]
    n_vox = [96 ** 3, 128 ** 3, 160 ** 3][:len(ups)]
    upd = [int(v) for v in a.upd.split(",")] if a.upd else [int(n_it * f) for f in (0.13, 0.3, 0.5, 0.7)]
    # every occupancy rebuild is recorded: the fraction of lattice points whose mask alpha 1 - exp(-sigma * step) reaches the
    # threshold (samplers/alphagrid.py:222: NO distance_scale) next to the fraction whose RENDER alpha
    # 1 - exp(-sigma * step * distance_scale) does -- the reference culls by the first and renders with the second
    rebuilds = []
    dense_alpha0 = ops.dense_alpha

    def dense_alpha_logged(scene, gs):
        al = dense_alpha0(scene, gs)
        x = -torch.log1p(-al.clamp(max=1 - 1e-7))
        thr = scene.hp["alpha_mask_thres"]
        rebuilds.append(dict(grid=list(gs), stepsize=float(scene.c.stepsize), distance_scale=float(scene.c.distance_scale),
                             frac_mask_alpha=float((al >= thr).float().mean()),
                             frac_render_alpha=float(((1 - torch.exp(-x * scene.c.distance_scale)) >= thr).float().mean()),
                             frac_render_alpha_1pct=float(((1 - torch.exp(-x * scene.c.distance_scale)) >= 0.01).float().mean())))
        return al
    ops.dense_alpha = dense_alpha_logged
    tr = train.PlainTrainer(init, meta["aabb"], meta["near_far"], [64, 64, 64], alpha_volume=None, device=dev,
                            max_samples=400000, seed=1, params=dict(n_iters=n_it), alpha_mask_thres=a.alpha_thres)
    log = []

    def cb(rec):
        if rec["iteration"] % max(n_it // 15, 1) == 0 or rec.get("reinit"):
            log.append(dict(it=rec["iteration"], psnr=-10 * math.log10(max(rec["mse"], 1e-12)), lbatch=rec["lbatch_size"],
                            subs=rec["sub_batches"], samples=rec["n_samples"], grid=rec["grid"][0], reinit=bool(rec.get("reinit")),
                            occupied=None if tr.alpha_volume is None else round(float(torch.as_tensor(tr.alpha_volume).float().mean()), 5),
                            occ_shape=None if tr.alpha_volume is None else list(torch.as_tensor(tr.alpha_volume).shape[-3:])))

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hist = tr.fit(train_rays, train_rgbs, n_iters=n_it, upsamp_list=ups, n_voxel_list=n_vox, update_list=upd, callback=cb)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    # ---- held-out views ----
    psnr = []
    for r, gt in zip(rays[a.views:], rgbs[a.views:]):
        im, _ = ops.render_rays(tr.scene, r, focal, chunk=4096, skip_eps=0.0, t_cut=0.0)
        psnr.append(-10 * math.log10(float(((im["rgb_map"] - gt) ** 2).mean())))
    first = sum(h["mse"] for h in hist[:10]) / 10
    last = sum(h["mse"] for h in hist[-10:]) / 10
    print(json.dumps(dict(what="PlainTrainer.fit from a fresh initialisation (model=tensorf), synthetic lego teacher",
                          iters=n_it, train_views=a.views, res=a.res, alpha_mask_thres=a.alpha_thres, final_grid=tr.meta["grid_size"],
                          train_psnr_first10=-10 * math.log10(first), train_psnr_last10=-10 * math.log10(last),
                          test_psnr=psnr, test_psnr_mean=sum(psnr) / len(psnr), wall_s=wall, ms_per_iter=wall / n_it * 1e3,
                          rays_seen=sum(h["kept_rays"] for h in hist), upsamp_list=ups, update_list=upd, rebuilds=rebuilds, log=log)))


if __name__ == "__main__":
    main()
