"""GPU check of the tcgen05 BRDF MLP against the oracle (prints error statistics)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_fixture, oracle_scene, device_scene
from nmf_b200 import ops
from oracle import nmf_oracle as O
fix = load_fixture("microfacet_g40")
osc = oracle_scene(fix)
g = torch.Generator().manual_seed(1)
for n in (128, 1000, 30000):
    half = O.unit(torch.randn(n, 3, generator=g)); diff = O.unit(torch.randn(n, 3, generator=g))
    r = torch.rand(n, 1, generator=g) * 0.49 + 0.01
    feat = torch.randn(n, 24, generator=g) * 0.3
    ref = O.brdf_mlp(osc, feat, half, diff, r)
    for mode in ("fp32", "f16"):
        dsc = device_scene(fix, "cuda:0", mlp=mode)
        bw = ops.brdf_mlp(dsc, feat.cuda(), half.cuda(), diff.cuda(), r.cuda()).cpu()
        e = (bw - ref).abs()
        print(n, mode, "max", float(e.max()), "mean", float(e.mean()), flush=True)
