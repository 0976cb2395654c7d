"""BASELINE config #5 on N GPUs: the relight sweep -- `render_only=True` of {ficus, helmet, toaster} checkpoints under 6
environment maps (train.py:54-188), 18 independent jobs dealt round-robin to the ranks (scene-parallel replicas, no
data-path collective).  No datasets or trained checkpoints exist offline: the three scenes are the synthetic
stand-ins of nmf_b200/synthetic.py at G=300, saved and re-loaded through the reference's checkpoint wire format
(TensorNeRF.save / relight.load_for_render), the environments are five procedural HDR maps at 512x1024 plus
backgrounds/forest.th (1024x2048) when the staged reference is present.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 \
      tools/relight_sweep_bench.py --views 4

Timed on the device per rank (CUDA events around the rank's jobs, checkpoint load and environment swap included), max over
ranks; rank 0 prints ONE JSON line."""
import argparse
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=4)
    ap.add_argument("--size", type=int, default=800)
    ap.add_argument("--grid", type=int, default=300)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=240))
    from nmf_b200 import config, relight, synthetic
    from nmf_b200.plugins import IntegralEquirect
    scenes = ["ficus", "helmet", "toaster"]
    envs = {}
    for i in range(5):
        e = IntegralEquirect(bg_resolution=512, init_val=-0.6, activation="exp", mipbias=0.0)
        with torch.no_grad():
            e.bg_mat.copy_(synthetic.procedural_env_log_radiance(512, seed=10 + i).reshape(e.bg_mat.shape))
        envs[f"procedural{i}"] = e.state_dict()
    forest = os.path.join(ROOT, "baseline", "_ref", "backgrounds", "forest.th")
    if os.path.exists(forest):
        envs["forest"] = forest
    else:
        e = IntegralEquirect(bg_resolution=1024, init_val=-0.6, activation="exp", mipbias=0.0)
        with torch.no_grad():
            e.bg_mat.copy_(synthetic.procedural_env_log_radiance(1024, seed=99).reshape(e.bg_mat.shape))
        envs["procedural_1024"] = e.state_dict()
    jobs = [(s, e) for s in scenes for e in envs]
    mine = relight.job_slice(len(jobs), rank, world)
    tmp = tempfile.mkdtemp(prefix=f"relight_r{rank}_")
    # every rank writes only the checkpoints of its own jobs (a trained checkpoint would simply be a path)
    ckpts = {}
    G = a.grid
    for s in sorted({jobs[j][0] for j in mine}):
        state, meta = synthetic.make_scene(s, grid_size=G)
        t, cfg = config.build_model([f"field.grid_size=[{G},{G},{G}]"], aabb=meta["aabb"], near_far=list(meta["near_far"]))
        t.load_state_dict(state, strict=False)
        t = t.cuda().eval()
        t.sampler.update(t.rf, init=True)
        t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
        p = os.path.join(tmp, f"{s}.th")
        t.save(p, cfg.model.arch)
        ckpts[s] = p
        near_far = list(meta["near_far"])
        del t
    torch.cuda.empty_cache()
    H = W = a.size
    focal = synthetic.focal_for(W)
    flip = torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))
    poses = [torch.as_tensor(q, dtype=torch.float32) @ flip for q in synthetic.hemisphere_poses(a.views, seed=3)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res, cache = {}, {}
    for j in mine:
        s, e = jobs[j]
        if s not in cache:
            cache = {s: relight.load_for_render(ckpts[s], near_far=near_far)}
        t = relight.swap_env(cache[s], envs[e])
        from nmf_b200 import renderer
        res[(s, e)] = renderer.evaluate_views(t, poses, H, W, focal)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    stats = torch.tensor([float(len(mine)), float(sum(float(r["images"][v]["rgb_map"].mean()) for r in res.values() for v in range(a.views)))],
                         device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats)
    if rank == 0:
        n_rays = len(jobs) * a.views * H * W
        finite = all(bool(torch.isfinite(r["images"][v]["rgb_map"]).all()) for r in res.values() for v in range(a.views))
        print(json.dumps({
            "what": "relight sweep (BASELINE config #5): render_only over 3 scenes x 6 environment maps, scene-parallel replicas, "
                    "no data-path collective; checkpoint load + env swap + ray generation + render inside the timed region",
            "n_gpus": world, "jobs": len(jobs), "jobs_rank0": len(mine), "views_per_job": a.views, "image": [H, W], "grid": G,
            "envs": list(envs), "scenes": scenes, "seconds": float(ms[0]) / 1e3, "rays_per_s": n_rays / (float(ms[0]) / 1e3),
            "images_per_s": len(jobs) * a.views / (float(ms[0]) / 1e3), "jobs_done": int(stats[0]), "mean_rgb_sum": float(stats[1]),
            "all_finite_rank0": finite, "timing": "CUDA events per rank, max over ranks"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
