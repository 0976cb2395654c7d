"""Where one MicrofacetTrainer iteration spends its time (1 GPU): wall clock with a synchronise around every part
(= CPU launch work + GPU work of that part) next to the CUDA-event time of the un-instrumented iteration.
usage: python tools/mf_iter_bench.py [--steps 20] [--grid 300]"""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmf_b200 import ops, synthetic, train

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--grid", type=int, default=300)
ap.add_argument("--rays", type=int, default=4096)
a = ap.parse_args()
dev = torch.device("cuda:0")
state, meta = synthetic.make_scene("lego", grid_size=a.grid)
tr = train.MicrofacetTrainer(state, meta["aabb"], meta["near_far"], meta["grid_size"], device=dev, max_samples=200000, seed=7,
                             params=dict(train.MICROFACET_REFERENCE_PARAMS))
tr.alpha_volume = tr.scene.update_alpha_mask()
tr.update_n_samples = lambda n: None          # fixed re-trace budget (the adaptive controller would move it to ~38 000 after 20 iterations)
H = W = 800
focal = synthetic.focal_for(W)
pose = synthetic.hemisphere_poses(8)[1]
pix = torch.randperm(H * W, generator=torch.Generator().manual_seed(0))[:a.rays]
rays = synthetic.camera_rays(pose, H, W, focal)[pix].contiguous().to(dev)
gt = ops.render_rays(tr.scene, rays, focal, chunk=a.rays, skip_eps=0.0, t_cut=0.0)[0]["rgb_map"].clone()

def timed(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t) * 1e3, r

parts = {}
def add(k, ms): parts[k] = parts.get(k, 0.0) + ms
p = tr.params
for it in range(a.steps + 3):
    rec = it >= 3
    ms, out = timed(lambda: tr.accumulate(rays, gt, first=True));
    if rec: add("accumulate (zero + nmf_train_microfacet)", ms)
    ms, _ = timed(lambda: tr.grads.finish(p["bg_module.bg_mat"].data, None, None, scalars_dev=tr._env_dev()))
    if rec: add("grads.finish (env scans, stencil adjoint)", ms)
    ms, _ = timed(lambda: tr.grads.copy_into({k: q.grad for k, q in p.items()}))
    if rec: add("gradient hand-over -> bucket (nmf_transpose_batch)", ms)
    def l1():
        if tr.l1_weight > 0:
            tr.l1_sum.zero_()
            for k, q in p.items():
                if ".density_rf." in k:
                    train.l1_reg(q.data, tr.l1_weight, q.grad, tr.l1_sum)
    ms, _ = timed(l1)
    if rec: add("density L1", ms)
    ms, _ = timed(lambda: tr.optimizer.step(grad_scale=1.0 / a.rays))
    if rec: add("FusedAdam", ms)
    st = dict(tr.state); st.update({k: q.detach() for k, q in p.items()})
    ms, _ = timed(lambda: tr.scene._pack_factors(st, derivatives=True))
    if rec: add("repack: factors", ms)
    ms, _ = timed(lambda: tr.scene._pack_shading(st))
    if rec: add("repack: shading", ms)
    ms, _ = timed(lambda: tr.scene._set_env(st, None, dev_scalars=tr._env_dev()))
    if rec: add("repack: env", ms)
parts = {k: round(v / a.steps, 4) for k, v in parts.items()}
# un-instrumented iterations, CUDA events and wall clock
def loop(body):
    for _ in range(3): body()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); t0 = time.perf_counter(); ev[0].record()
    for it in range(a.steps): body()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / a.steps, (time.perf_counter() - t0) * 1e3 / a.steps

def serial():
    out = tr.accumulate(rays, gt, first=True); tr.apply(out["n_rays"], out["loss_photo"])
ev_serial, wall_serial = loop(serial)
ev_step, wall_step = loop(lambda: tr.step(rays, gt))
print(json.dumps({"max_retrace": list(tr.scene.hp["max_retrace_rays"]), "parts_wall_ms_with_sync": parts, "sum_parts": round(sum(parts.values()), 3),
                  "iteration_serial_event_ms": ev_serial, "iteration_serial_wall_ms": wall_serial,
                  "iteration_step_event_ms": ev_step, "iteration_step_wall_ms": wall_step}))
