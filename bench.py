#!/usr/bin/env python
"""Headline benchmark: rays/s of the NMF forward render (model=microfacet_tensorf2, field=tensorf at G=300,
800x800 image = 640 000 rays in 4096-ray chunks) on N B200s.   python bench.py --gpus N --steps K --warmup W

A "step" is one full pass of the hot path over one synthetic 800x800 image (157 chunks).  `value` is timed with
the rays resident in HBM; `e2e` goes through the C-ABI host-buffer call (renderer.HostRenderer -> nmf_render_rays_host)
with pinned HOST rays in and every output map copied back to the host inside the timed region; `e2e_plugin` is the same
through the plugin stack a reference user calls (config.build_model -> TensorNeRF -> renderer.chunk_renderer(
render2completion=True)).  `--impl reference` times the reference algorithm on the host cores (the oracle port of the
reference's PyTorch path, GPU-free: no CUDA library is mapped) on a bounded sample of the same workload;
`--impl reference-cuda` times the UNMODIFIED reference (baseline/_ref, staged by __graft_entry__.build()) through its own
PyTorch-CUDA path on the same B200 -- the renderer the >= 10x target of BASELINE.json is defined against.
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
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--ref-chunks", type=int, default=12, help="chunks of the image the reference-cuda leg renders")
    ap.add_argument("--no-refcuda", action="store_true", help="skip the reference_cuda leg of the b200 arm")
    ap.add_argument("--sustain-s", type=float, default=5.0, help="seconds of back-to-back steps for the `sustained` key (0 = skip)")
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
    alpha = reference_alpha_volume(state, meta, a)
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


def reference_alpha_volume(state, meta, a=None):
    """Occupancy volume for the CPU arm, GPU-free: the copy the ORACLE computed at build time
    (oracle/_cache, __graft_entry__.stage_reference) or, failing that, the oracle's own rebuild here (~20 s at G=300)."""
    import numpy as np
    cache = os.path.join(ROOT, "oracle", "_cache", f"alpha_{a.scene if a else 'lego'}_g{a.grid if a else 300}.pt")
    if os.path.exists(cache):
        d = torch.load(cache, weights_only=False)
        n = int(np.prod(d["shape"]))
        return torch.from_numpy(np.unpackbits(d["bits"].numpy())[:n].reshape(d["shape"])).float()
    from oracle import nmf_oracle
    sc = nmf_oracle.Scene(state, meta["aabb"], meta["near_far"], meta["grid_size"])
    return nmf_oracle.build_alpha_volume(sc)


def reference_cuda(a, state, meta, rays, focal, n_chunks):
    """The UNMODIFIED reference (TensorNeRF + its plugins, imported from baseline/_ref or /root/reference through
    oracle/ref_harness.py's stub modules) on the GPU through its own PyTorch-CUDA path: the loop of renderer.chunk_renderer
    (renderer.py:72-104: one forward per 4096-ray chunk, every output copied to the host), CUDA-event timed per chunk.
    Returns a dict or {"unavailable": why}."""
    root = None
    for c in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(c, "modules")):
            root = c
            break
    if root is None:
        return {"unavailable": "reference tree not staged (baseline/_ref is written by __graft_entry__.build() where /root/reference exists)"}
    if not torch.cuda.is_available():
        return {"unavailable": "no CUDA device"}
    os.environ["NMF_REFERENCE_ROOT"] = root
    try:
        from oracle import ref_harness
        ref_harness.REFERENCE_ROOT = root
        dev = torch.device("cuda", torch.cuda.current_device())
        gs = [int(g) for g in meta["grid_size"]]
        t = ref_harness.build_reference_model(meta["aabb"], list(meta["near_far"]), grid_size=gs, bg_resolution=meta["bg_resolution"])
        t.load_state_dict({k: v for k, v in state.items()}, strict=False)
        t = t.to(dev)
        t.sampler.update(t.rf, init=True)
        t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
        t.eval()
        torch.manual_seed(20211200)
        times = []
        with torch.no_grad():
            for c in range(n_chunks + 2):
                r = rays[(c % 64) * a.chunk:(c % 64 + 1) * a.chunk].to(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ims, stats = t(r, focal, is_train=False, ndc_ray=False, N_samples=-1)
                host = {k: v.cpu() for k, v in ims.items() if torch.is_tensor(v)}        # renderer.py:88-97
                e1.record()
                torch.cuda.synchronize(dev)
                if c >= 2:
                    times.append(e0.elapsed_time(e1))
        times.sort()
        med = times[len(times) // 2]
        return {"value": a.chunk / (med * 1e-3), "unit": UNIT, "ms_per_chunk_median": med, "ms_per_chunk_mean": sum(times) / len(times),
                "chunks": n_chunks, "n_samples_last": [int(x) for x in stats["n_samples"]], "root": os.path.relpath(root, ROOT) if root.startswith(ROOT) else root,
                "what": "unmodified reference TensorNeRF.forward per 4096-ray chunk on cuda + D2H of every output (its chunk_renderer loop), "
                        "CUDA events, median over chunks after 2 warm-up chunks"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def run_reference_cuda(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    state, meta, rays, focal = workload(a, 0)
    r = reference_cuda(a, state, meta, rays, focal, max(a.steps, 1) * 4)
    if "unavailable" in r:
        emit({"impl": "reference-cuda", "unavailable": r["unavailable"]})
        return
    emit({"impl": "reference-cuda", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
          "ms_per_step": r["ms_per_chunk_median"] * math.ceil(rays.shape[0] / a.chunk), "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(a, rays.shape[0]), "reference_cuda": r,
          "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": a.chunk * 24, "d2h_bytes_per_step": None}, "gpu_launches": 0})


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
    if a.impl == "reference-cuda":
        return run_reference_cuda(a)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=150))
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

    # ---- e2e_plugin: the same image through the plugin stack a reference user calls (config.build_model ->
    # TensorNeRF -> renderer.chunk_renderer(render2completion=True): H2D of the rays, one launch sequence, D2H of every map)
    plug_ms, plug_err = None, None
    try:
        plug_ms = plugin_e2e(a, state, meta, alpha, rays_pinned, focal, dev, barrier)
    except Exception as e:                          # the extra leg never takes the bench line down
        plug_err = f"{type(e).__name__}: {e}"[:200]

    # ---- sustained: back-to-back device-resident steps for >= --sustain-s seconds (clocks sag under seconds-long load) ----
    sus = None
    if a.sustain_s > 0:
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_sus = max(a.steps, int(math.ceil(a.sustain_s * 1e3 / max(ms / a.steps, 1e-3))))
        with ClockSampler(local) as clk2:
            s0.record()
            for _ in range(n_sus):
                step()
            s1.record()
            barrier()
        sus = (s0.elapsed_time(s1), n_sus, clk2.summary())

    times = torch.tensor([ms, e2e_ms, plug_ms or 0.0, sus[0] if sus else 0.0], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(times[0]), float(times[1])
    plug_ms = float(times[2]) if plug_ms else None
    extra_n = multi_gpu_extras(a, dist, rank, world, dev, state, meta, alpha, focal, barrier) if dist is not None else None
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
    mlp_ms = phase_acc["bounce0"] + phase_acc["bounce1"]          # k_tile_prefix + k_bounce<L> only (k_incoming<1> has its own phase)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1600.0)))
    # the gather kernels' factor set is L2-resident: their ceiling is the rate of independent 16-byte taps the L2 serves,
    # measured here (nmf_bench_gather over a 78 MB set = the factor set of G=300), not the HBM copy rate
    l2_by_seg = {16 * g: ops.gather_peak(78 << 20, group=g) for g in (1, 4, 8)}
    l2_peak = l2_by_seg[128]
    dram_gather = ops.gather_peak(8 << 30, taps=32, group=4)
    touched = kernels[dom]["touched"] or kernels[dom]["bytes"]
    t_dom = kernels[dom]["ms"] * 1e-3
    roof = {"bound": "l2", "bound_note": "the factor set is L2-resident (DRAM traffic = 6 % of the copy rate), so the ceiling of the dominant "
                                       "kernel is the measured L2 gather rate; the HBM view of the same launch is under 'dram' (real DRAM "
                                       "bytes) and 'hbm_reference_equivalent' (SURVEY 8d bytes), the tensor-core view under 'mlp'",
            "kernel": f"k_{dom[:-1]}<0>", "achieved": touched / t_dom / 1e9, "peak": l2_peak, "unit": "GB/s",
            "frac": touched / t_dom / 1e9 / l2_peak,
            "peak_source": "measured in this run: nmf_bench_gather, independent random 128-byte segments (8 lanes x 16 B) over a 78 MB "
                           "(L2-resident) set; l2_gather_GBps_by_segment_bytes has the 16- and 64-byte figures",
            "l2_gather_GBps_by_segment_bytes": l2_by_seg,
            "achieved_note": "bytes of the taps the kernel actually issues (16-byte factor taps of the samples it shades) / "
                             "CUDA-event launch time",
            "traffic": traffic, "traffic_capture": capture,
            "dram": {"achieved": (traffic / t_dom / 1e9) if traffic else None, "peak": peak, "frac": (traffic / t_dom / 1e9 / peak) if traffic else None,
                     "peak_source": peak_src, "random_64B_gather_GBps": dram_gather,
                     "note": "real DRAM bytes per launch (ncu) / launch time: the factor planes stay in L2"},
            "hbm_reference_equivalent": {"achieved": kernels[dom]["gbs"], "peak": peak, "frac": kernels[dom]["gbs"] / peak,
                                         "note": "SURVEY 8d algorithmic bytes (reference layouts, no reuse, EVERY valid sample) / launch "
                                                 "time; > 1 = L2 reuse + samples below the weight cut are not shaded -- not an HBM rate"},
            "mlp": {"bound": "tensor", "kernel": "k_bounce<0>+k_bounce<1>", "flop_per_ray": 17152, "rays": n_bray,
                    "achieved": 17152.0 * n_bray / max(mlp_ms, 1e-9) / 1e9, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": 17152.0 * n_bray / max(mlp_ms, 1e-9) / 1e9 / tf_peak, "ms": mlp_ms, "operands": "fp16, fp32 accumulate (tcgen05)",
                    "note": "tcgen05 GEMMs of the 66-64-64-4 BRDF MLP; the kernels also draw the GGX samples and encode them"},
            "launch_ms": kernels[dom]["ms"], "algorithmic_bytes_per_launch": kernels[dom]["bytes"], "touched_bytes_per_launch": touched,
            "march_plus_query": {"ms": fused_ms, "algorithmic_bytes": (B_CAND * cand + (B_DENSITY + B_APP + B_NORMAL) * (M0 + M1)),
                                 "achieved": (B_CAND * cand + (B_DENSITY + B_APP + B_NORMAL) * (M0 + M1)) / max(fused_ms, 1e-9) / 1e6,
                                 "frac_of_hbm_copy": (B_CAND * cand + (B_DENSITY + B_APP + B_NORMAL) * (M0 + M1)) / max(fused_ms, 1e-9) / 1e6 / peak},
            "whole_step": {"algorithmic_bytes": fused_bytes, "achieved": fused_bytes / (ms / a.steps) / 1e6,
                           "frac_of_hbm_copy": fused_bytes / (ms / a.steps) / 1e6 / peak},
            "per_kernel": {k: {"ms": round(d["ms"], 4), "GBps_algorithmic": round(d["gbs"], 1),
                               "GBps_touched": (round(d["touched"] / (d["ms"] * 1e-3) / 1e9, 1) if d["touched"] and d["ms"] > 0 else None),
                               "frac_of_l2_gather": (round(d["touched"] / (d["ms"] * 1e-3) / 1e9 / l2_peak, 3) if d["touched"] and d["ms"] > 0 else None)}
                           for k, d in kernels.items()},
            "phase_ms": {k: round(v, 4) for k, v in phase_acc.items()},
            "phase_note": "CUDA-event phase times averaged over the e2e (host-buffer) steps: same kernels as the device-resident "
                          "steps; 'finish' there also holds the last staged device-to-host copies (k_finish0 itself: ~40 us)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "dtype_note": "fp32 everywhere except the operands of the BRDF-MLP GEMMs (fp16, fp32 accumulate in TMEM)",
            "data": "synthetic", "config": config_dict(a, n), "clocks": clk.summary(),
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": host.h2d_bytes, "d2h_bytes_per_step": host.d2h_bytes,
                    "ms_per_step": e2e_ms / a.steps, "api": "renderer.HostRenderer.render -> nmf_render_rays_host (C ABI, host buffers)"},
            "e2e_plugin": ({"value": world * n * a.steps / (plug_ms / 1e3), "unit": UNIT, "ms_per_step": plug_ms / a.steps,
                            "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": host.d2h_bytes,
                            "api": "config.build_model -> TensorNeRF -> renderer.chunk_renderer(render2completion=True)"}
                           if plug_ms else {"error": plug_err}),
            "sustained": ({"value": world * n * sus[1] / (float(times[3]) / 1e3), "unit": UNIT, "seconds": float(times[3]) / 1e3, "steps": sus[1],
                           "clocks": sus[2]} if sus else None),
            "gpu_launches": 16 * a.steps, "roofline": roof,
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
            line["train_step"] = train.benchmark_microfacet_train(a.grid, 4096, steps=10, device=dev_s)
            line["train_step_plain"] = train.benchmark_plain(a.grid, 4096, steps=10, iters=8, device=dev_s)
        except Exception as e:                      # never lets the extra entry take the bench line down
            line["train_step"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    if extra_n is not None:
        line.update(extra_n)
    if not a.no_refcuda and world == 1:
        line["reference_cuda"] = reference_cuda(a, state, meta, rays_host, focal, a.ref_chunks)
        if "value" in line["reference_cuda"]:
            line["reference_cuda"]["ratio_e2e"] = e2e_v / line["reference_cuda"]["value"]
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def plugin_e2e(a, state, meta, alpha, rays_pinned, focal, dev, barrier):
    """ms for a.steps images through the reference-facing plugin stack (host rays in, host maps out)."""
    from nmf_b200 import config, renderer
    gs = meta["grid_size"]
    t, _ = config.build_model([f"field.grid_size=[{gs[0]},{gs[1]},{gs[2]}]", f"model.arch.bg_module.bg_resolution={meta['bg_resolution']}"],
                              aabb=meta["aabb"], near_far=list(meta["near_far"]))
    t.load_state_dict(state, strict=False)
    t = t.to(dev).eval()
    t.sampler.update(t.rf, init=True)
    from nmf_b200.plugins import AlphaGridMask
    t.sampler.alphaMask = AlphaGridMask(t.rf.aabb, alpha.to(dev))
    run = lambda: renderer.chunk_renderer(rays_pinned.to(dev, non_blocking=True), t, focal, keys=None, chunk=a.chunk, render2completion=True)
    for _ in range(2):
        run()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0)


def multi_gpu_extras(a, dist, rank, world, dev, state, meta, alpha, focal, barrier):
    """N > 1 only -- measurements that can fail:
      strong_scaling  ONE 800x800 image sharded over the ranks by distributed.shard_chunks (whole chunks, global ray ids),
                      gathered on rank 0 with NCCL, timed end to end (max over ranks)
      train_sharded   a ray-sharded nmf_train_microfacet step per rank + ONE flat fp32 gradient all-reduce (NCCL) + FusedAdam,
                      with the all-reduce time and its bus bandwidth broken out"""
    from nmf_b200 import distributed, ops, synthetic, train
    from nmf_b200.scene import DeviceScene
    out = {}
    try:
        scene = DeviceScene(state, meta["aabb"], meta["near_far"], meta["grid_size"], alpha_volume=alpha, device=dev)
        poses = synthetic.hemisphere_poses(200, seed=1)
        rays_all = synthetic.camera_rays(poses[0], a.res, a.res, focal)
        perm = torch.randperm(rays_all.shape[0], generator=torch.Generator().manual_seed(20211200))
        rays_all = rays_all[perm].contiguous().to(dev)
        n = rays_all.shape[0]
        lo, hi = distributed.shard_chunks(n, a.chunk, rank, world)
        mine = rays_all[lo:hi].contiguous()
        bufs = ops.RenderBuffers(scene, max(hi - lo, 1), a.chunk, ["rgb_map", "acc_map", "depth"])
        counts = [c[1] - c[0] for c in (distributed.shard_chunks(n, a.chunk, r, world) for r in range(world))]
        stage = [None]

        def one():
            ims, _ = ops.render_rays(scene, mine, focal, chunk=a.chunk, seed=20211200, ray_id0=lo, buffers=bufs, check_errors=False)
            _, stage[0] = distributed.gather_ragged(ims["rgb_map"], counts, rank, world, buffers=stage[0])
        for _ in range(2):
            one()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            one()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["strong_scaling"] = {"value": n * a.steps / (float(t[0]) / 1e3), "unit": UNIT, "ms_per_image": float(t[0]) / a.steps,
                                 "what": "one 800x800 image sharded by whole chunks over the ranks, rgb gathered on rank 0 (NCCL), max over ranks"}
        del bufs
        torch.cuda.empty_cache()
    except Exception as e:
        out["strong_scaling"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    try:
        out["train_sharded"] = train.benchmark_sharded_train(a.grid, 4096, steps=10, device=dev)
    except Exception as e:
        out["train_sharded"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    return out


if __name__ == "__main__":
    main()
