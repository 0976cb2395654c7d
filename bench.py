#!/usr/bin/env python
"""Headline benchmark: rays/s of the NMF forward render (model=microfacet_tensorf2, field=tensorf at G=300,
800x800 image = 640 000 rays in 4096-ray chunks) on N B200s.   python bench.py --gpus N --steps K --warmup W

A "step" is one full pass of the hot path over one synthetic 800x800 image (157 chunks).  `value` is timed with
the rays resident in HBM; `e2e` goes through the reference-facing API (nmf_b200.renderer.chunk_renderer ->
nmf_render_rays_host) with pinned HOST rays in and every output map copied back to the host inside the timed
region.  `--impl reference` times the reference algorithm on the host cores (the oracle port of the reference's
PyTorch path; the reference itself is not installable on the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rays/sec at 800x800 lego (forward render, microfacet_tensorf2, G=300)"
UNIT = "rays/s"
# algorithmic bytes (fp32, reference layouts, no cache reuse)
B_CAND, B_DENSITY, B_APP, B_NORMAL, B_ENV, B_RAY = 32, 1152, 1728, 1920, 192, 156


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=300)
    ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--scene", default="lego")
    ap.add_argument("--cpu-chunks", type=int, default=16, help="chunks of the image the CPU baseline renders")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-train", action="store_true", help="skip the extra train_step entry (training slice, SURVEY 8f row 1)")
    ap.add_argument("--skip-eps", type=float, default=None)
    ap.add_argument("--t-cut", type=float, default=None)
    return ap.parse_args()


def workload(a, rank):
    """Synthetic Blender-format scene + the rays of one 800x800 view, shuffled as BundleRender does
    (renderer.py:130-132) so that every chunk is a random subset of the image."""
    from nmf_b200 import synthetic
    state, meta = synthetic.make_scene(a.scene, grid_size=a.grid, bg_resolution=512)
    poses = synthetic.hemisphere_poses(200, seed=1)
    focal = synthetic.focal_for(a.res)
    rays = synthetic.camera_rays(poses[rank % 200], a.res, a.res, focal)
    perm = torch.randperm(rays.shape[0], generator=torch.Generator().manual_seed(20211200 + rank))
    return state, meta, rays[perm].contiguous(), focal


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *exc):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_traffic.py from the capture named there); None when absent."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        return d["kernels"].get(kernel), d.get("capture")
    except Exception:
        return None, None


def oracle_chunks(state, meta, alpha_volume, rays, focal, chunk, n_chunks, seed, warm=1):
    """The reference algorithm on the host cores: the oracle port, `n_chunks` chunks after `warm` warm-up chunks."""
    from oracle import keyed_rng, nmf_oracle
    torch.set_num_threads(os.cpu_count())
    sc = nmf_oracle.Scene(state, meta["aabb"], meta["near_far"], meta["grid_size"], alpha_volume=alpha_volume)
    rng = keyed_rng.KeyedRNG()
    outs, t_tot, n_tot = [], 0.0, 0
    with torch.no_grad():
        for c in range(warm + n_chunks):
            idx = c - warm if c >= warm else 0
            r = rays[idx * chunk:(idx + 1) * chunk]
            t0 = time.perf_counter()
            ims, _ = nmf_oracle.render_rays(sc, r, focal, rng, chunk=chunk, seed=seed, ray_id0=idx * chunk)
            dt = time.perf_counter() - t0
            if c >= warm:
                t_tot += dt
                n_tot += r.shape[0]
                outs.append(ims)
    return n_tot / t_tot, t_tot, outs


def run_reference(a):
    """--impl reference: the reference path on the host CPU (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    state, meta, rays, focal = workload(a, 0)
    alpha = reference_alpha_volume(state, meta)
    from oracle import keyed_rng, nmf_oracle
    torch.set_num_threads(os.cpu_count())
    sc = nmf_oracle.Scene(state, meta["aabb"], meta["near_far"], meta["grid_size"], alpha_volume=alpha)
    rng = keyed_rng.KeyedRNG()
    times = []
    with torch.no_grad():
        for s in range(a.warmup + a.steps):
            r = rays[(s % 8) * a.chunk:(s % 8 + 1) * a.chunk]
            t0 = time.perf_counter()
            nmf_oracle.render_rays(sc, r, focal, rng, chunk=a.chunk, seed=20211200, ray_id0=(s % 8) * a.chunk)
            if s >= a.warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    v = a.steps * a.chunk / total
    sample = f"{a.steps} steps of one {a.chunk}-ray chunk of the {a.res}x{a.res} image (after {a.warmup} warm-up chunks)"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(a, rays.shape[0]),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def reference_alpha_volume(state, meta):
    """Occupancy volume for the CPU arm.  With a GPU present it is built by the CUDA path (parity-tested against the
    oracle's rebuild); without one the oracle rebuilds it itself (slow at G=300)."""
    if torch.cuda.is_available():
        from nmf_b200.scene import DeviceScene
        sc = DeviceScene(state, meta["aabb"], meta["near_far"], meta["grid_size"], device="cuda:0")
        vol = sc.update_alpha_mask().cpu()
        del sc
        torch.cuda.empty_cache()
        return vol
    from oracle import nmf_oracle
    sc = nmf_oracle.Scene(state, meta["aabb"], meta["near_far"], meta["grid_size"])
    return nmf_oracle.build_alpha_volume(sc)


def config_dict(a, n_rays):
    return {"workload": f"model=microfacet_tensorf2 field=tensorf dataset={a.scene}(synthetic) {a.res}x{a.res} forward render, "
                        f"{a.chunk}-ray chunks", "grid": a.grid, "rays_per_step": n_rays, "chunk": a.chunk,
            "env": "IntegralEquirect 512x1024", "rays_per_ray": 128, "max_retrace_rays": 1000,
            "l2": "working set >> L2: the per-step scratch (several GB of sample / bounce-ray records) is streamed "
                  "through HBM every step, factor planes (78 MB) compete with it for the 126 MB L2"}


_REAL_STDOUT = None


def _claim_stdout():
    """Rank 0 must print exactly ONE JSON line: everything else that writes to fd 1 during the run (the NCCL version
    banner, library chatter) is sent to stderr; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    a = parse()
    _claim_stdout()
    if a.impl == "reference":
        return run_reference(a)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from nmf_b200 import _lib, ops, renderer
    from nmf_b200.scene import DeviceScene
    _lib.lib()
    state, meta, rays_host, focal = workload(a, rank)
    n = rays_host.shape[0]
    scene = DeviceScene(state, meta["aabb"], meta["near_far"], meta["grid_size"], device=dev)
    alpha = scene.update_alpha_mask()
    kw = {}
    if a.skip_eps is not None:
        kw["skip_eps"] = a.skip_eps
    if a.t_cut is not None:
        kw["t_cut"] = a.t_cut
    seed = 20211200
    rays = rays_host.to(dev)
    bufs = ops.RenderBuffers(scene, n, a.chunk, ops.image_keys(scene))
    ops.profile_enable(True)

    def step():
        ops.render_rays(scene, rays, focal, chunk=a.chunk, seed=seed, buffers=bufs, check_errors=False, **kw)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    stats = ops.read_counters(bufs, n, a.chunk)           # raises on a device-side overflow
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    phases = ops.profile_read()
    ops.read_counters(bufs, n, a.chunk)

    # ---- e2e: the reference-facing API with host buffers (H2D of the rays, D2H of every map, sync per step) ----
    host = renderer.HostRenderer(scene, n, a.chunk)
    rays_pinned = rays_host.pin_memory()
    for _ in range(2):
        host.render(rays_pinned, focal, seed=seed, **kw)
    barrier()
    phase_acc = {k: 0.0 for k in phases}
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ims_host, _ = host.render(rays_pinned, focal, seed=seed, **kw)
        for k, v in ops.profile_read().items():
            phase_acc[k] += v / a.steps
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)

    times = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(times[0]), float(times[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = world * n * a.steps / (ms / 1e3)
    e2e_v = world * n * a.steps / (e2e_ms / 1e3)
    # ---- roofline of the dominant kernels (algorithmic bytes of SURVEY 8d / measured launch duration) ----
    peak, peak_src = measured_peak()
    M0, M1 = sum(stats["n_samples0"]), sum(stats["n_samples1"])
    cand = sum(stats["n_cand"])
    n1 = sum(stats["n_retrace"])
    sh0, sh1 = stats["n_shaded"]
    kernels = {
        "march0": dict(bytes=B_CAND * cand * M0 / max(M0 + M1, 1) + B_DENSITY * M0 + 24 * n, touched=None),
        "shade0": dict(bytes=(B_APP + B_NORMAL) * M0, touched=(B_APP + B_NORMAL) * sh0),
        "march1": dict(bytes=B_DENSITY * M1 + 24 * n1, touched=None),
        "shade1": dict(bytes=(B_APP + B_NORMAL) * M1, touched=(B_APP + B_NORMAL) * sh1),
    }
    for k, d in kernels.items():
        t = phase_acc[k]
        d["ms"] = t
        d["gbs"] = d["bytes"] / (t * 1e-3) / 1e9 if t > 0 else 0.0
    fused_bytes = B_CAND * cand + (B_DENSITY + B_APP + B_NORMAL) * (M0 + M1) + B_ENV * (sum(stats["n_bounce_rays0"]) - n1 + sum(stats["n_bounce_rays1"])) + B_RAY * n
    fused_ms = sum(kernels[k]["ms"] for k in kernels)
    dom = max(("march0", "shade0"), key=lambda k: kernels[k]["ms"])
    traffic, capture = ncu_traffic(f"k_{dom[:-1]}<0>")
    n_bray = sum(stats["n_bounce_rays0"]) + sum(stats["n_bounce_rays1"])
    mlp_ms = phase_acc["bounce0"] + phase_acc["bounce1"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1600.0)))
    roof = {"bound": "hbm", "kernel": f"k_{dom[:-1]}<0>", "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
            "frac": kernels[dom]["gbs"] / peak, "traffic": traffic, "traffic_capture": capture, "peak_source": peak_src,
            "note": "achieved = algorithmic bytes of SURVEY 8d (reference layouts, no reuse, every valid sample) / measured "
                    "launch time; the factor planes are L2-resident (traffic << algorithmic bytes) and samples below the "
                    "weight cut are not shaded, so frac > 1 is cache reuse + pruning, not an HBM rate",
            "mlp": {"bound": "tensor", "kernel": "k_bounce<0>+k_bounce<1>", "flop_per_ray": 17152, "rays": n_bray,
                    "achieved": 17152.0 * n_bray / max(mlp_ms, 1e-9) / 1e9, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": 17152.0 * n_bray / max(mlp_ms, 1e-9) / 1e9 / tf_peak, "ms": mlp_ms,
                    "note": "fp16 tcgen05 GEMMs of the 66-64-64-4 BRDF MLP; the kernels also sample GGX, encode and look up the environment"},
            "launch_ms": kernels[dom]["ms"], "algorithmic_bytes_per_launch": kernels[dom]["bytes"],
            "march_plus_query": {"ms": fused_ms, "algorithmic_bytes": (B_CAND * cand + (B_DENSITY + B_APP + B_NORMAL) * (M0 + M1)),
                                 "achieved": (B_CAND * cand + (B_DENSITY + B_APP + B_NORMAL) * (M0 + M1)) / max(fused_ms, 1e-9) / 1e6,
                                 "frac": (B_CAND * cand + (B_DENSITY + B_APP + B_NORMAL) * (M0 + M1)) / max(fused_ms, 1e-9) / 1e6 / peak},
            "whole_step": {"algorithmic_bytes": fused_bytes, "achieved": fused_bytes / (ms / a.steps) / 1e6,
                           "frac": fused_bytes / (ms / a.steps) / 1e6 / peak},
            "per_kernel": {k: {"ms": round(d["ms"], 4), "GBps_algorithmic": round(d["gbs"], 1),
                               "GBps_touched": (round(d["touched"] / (d["ms"] * 1e-3) / 1e9, 1) if d["touched"] and d["ms"] > 0 else None)}
                           for k, d in kernels.items()},
            "phase_ms": {k: round(v, 4) for k, v in phase_acc.items()},
            "phase_note": "CUDA-event phase times averaged over the e2e (host-buffer) steps: same kernels as the device-resident "
                          "steps; 'finish' there also holds the last staged device-to-host copies (k_finish0 itself: ~40 us)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(a, n), "clocks": clk.summary(),
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": host.h2d_bytes, "d2h_bytes_per_step": host.d2h_bytes,
                    "ms_per_step": e2e_ms / a.steps},
            "gpu_launches": 15 * a.steps, "roofline": roof,
            "samples": {"valid_primary": M0, "valid_secondary": M1, "candidates": cand, "shaded_primary": sh0,
                        "shaded_secondary": sh1, "bounce_rays0": sum(stats["n_bounce_rays0"]),
                        "bounce_rays1": sum(stats["n_bounce_rays1"]), "retraced": n1}}
    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample; also gives PSNR vs oracle ----
    if not a.no_cpu and world == 1:
        v, t_tot, outs = oracle_chunks(state, meta, alpha.cpu(), rays_host, focal, a.chunk, a.cpu_chunks, seed)
        ref = torch.cat([o["rgb_map"] for o in outs])
        mine = ims_host["rgb_map"][:ref.shape[0]]
        mse = float(((mine - ref).clip(-1, 1) ** 2).mean())
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{a.cpu_chunks} chunks of {a.chunk} rays of the same image ({t_tot:.1f} s) after 1 warm-up chunk; "
                                          "oracle = CPU restatement of the reference's PyTorch path (torch CPU ops, all cores)"}
        line["psnr_vs_oracle_db"] = (10 * math.log10(1.0 / mse)) if mse > 0 else 99.0
    # ---- extra, outside the timed region and not part of `value`: the fused training step of the model=tensorf slice ----
    if not a.no_train and not a.no_cpu and world == 1:
        try:
            from nmf_b200 import train
            dev_s = f"cuda:{torch.cuda.current_device()}"
            line["train_step"] = train.benchmark_plain(a.grid, 4096, steps=10, iters=8, device=dev_s)
            line["train_forward_microfacet"] = train.benchmark_microfacet_forward(a.grid, 4096, steps=10, device=dev_s)
        except Exception as e:                      # never lets the extra entry take the bench line down
            line["train_step"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
