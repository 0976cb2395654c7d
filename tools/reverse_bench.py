"""Times the reverse-pass kernels of the microfacet model (csrc/nmf_env_bwd.cu, nmf_normals_bwd.cu, nmf_shade_bwd.cu) on the
bench scene (G=300, 512x1024 environment map) with CUDA events and prints one JSON line: ms per call and GB/s against the
algorithmic bytes stated in DESIGN.md section 9.  python tools/reverse_bench.py [--iters N]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--grid", type=int, default=300)
    ap.add_argument("--lookups", type=int, default=300000)      # bounce rays of one 4096-ray training batch (both levels)
    ap.add_argument("--samples", type=int, default=226000)      # shaded samples of one training batch
    args = ap.parse_args()
    from nmf_b200 import ops, synthetic
    from nmf_b200.scene import DeviceScene
    dev = torch.device("cuda:0")
    state, meta = synthetic.make_scene("lego", grid_size=args.grid, bg_resolution=512)
    aabb = torch.as_tensor(meta["aabb"], dtype=torch.float32)
    sc = DeviceScene(state, meta["aabb"], meta["near_far"], meta["grid_size"], device=dev)
    g = torch.Generator().manual_seed(0)
    out = {"grid": args.grid, "iters": args.iters}
    # environment map
    n = args.lookups
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
    mip = (torch.rand(n, generator=g) * 8 - 9).to(dev)
    up = torch.randn(n, 3, generator=g).to(dev)
    acc = ops.EnvMapGrad(sc)
    bg, br, mul = state["bg_module.bg_mat"].to(dev), float(state.get("bg_module.brightness", 0.0)), float(state.get("bg_module.mul", 1.0))
    h, w = acc.h, acc.w
    ms = timed(lambda: acc.scatter(d, mip, up), args.iters)
    out["env_scatter_plus_mipbias"] = {"ms": ms, "lookups": n, "Mlookups_per_s": n / ms / 1e3}
    ms = timed(lambda: acc.finish(bg, br, mul), args.iters)
    byt = 2 * 2 * h * w * 16 + 2 * 3 * h * w * 4
    out["env_finish"] = {"ms": ms, "algorithmic_bytes": byt, "GBps": byt / ms / 1e6, "note": "includes the wrapper's zero-fills"}
    # normals
    m = args.samples
    lo, hi = aabb[0], aabb[1]
    xyz = torch.cat([lo + (hi - lo) * (0.3 + 0.4 * torch.rand(m, 3, generator=g)), torch.zeros(m, 1)], dim=1).to(dev)
    dn = torch.randn(m, 3, generator=g).to(dev)
    ng = ops.NormalsGrad(sc)
    ms = timed(lambda: ng.scatter(xyz, dn), args.iters)
    out["normals_scatter"] = {"ms": ms, "samples": m, "Msamples_per_s": m / ms / 1e3}
    ms = timed(lambda: ng.finish(), args.iters)
    byt = sum((192 + 128) * t.shape[0] * t.shape[1] for t in ng.gpack)
    out["normals_finish"] = {"ms": ms, "algorithmic_bytes": byt, "GBps": byt / ms / 1e6, "note": "includes the wrapper's zero-fills and permutes"}
    # heads
    feat = (torch.randn(m, 24, generator=g) * 0.5).to(dev)
    ga, gf, gr = torch.randn(m, 3, generator=g).to(dev), torch.randn(m, 3, generator=g).to(dev), torch.randn(m, generator=g).to(dev)
    dw, db = torch.zeros(11, 24, device=dev), torch.zeros(11, device=dev)
    ms = timed(lambda: ops.material_heads_bwd(sc, feat, ga, gf, gr, dw, db), args.iters)
    out["heads_bwd"] = {"ms": ms, "samples": m, "algorithmic_bytes": 220 * m, "GBps": 220 * m / ms / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
