"""Times nmf_train_microfacet (forward + loss + reverse pass of microfacet_tensorf2) on the synthetic G=300 scene: BASELINE
config #3's shape (4096-ray batches, max_samples 200000, one re-traced level).  Run under gpurun:
  python tools/mf_train_bench.py [--rays 4096] [--steps 20] [--retrace 1000,38000] [--mlp f16]
Prints one JSON line per re-trace budget (CUDA-event timed)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nmf_b200 import train  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--grid", type=int, default=300)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--retrace", default="1000")
    ap.add_argument("--mlp", default="f16")
    ap.add_argument("--detach-N", type=int, default=0)
    a = ap.parse_args()
    for r in a.retrace.split(","):
        print(json.dumps(train.benchmark_microfacet_train(a.grid, a.rays, a.steps, max_retrace=int(r), detach_N=bool(a.detach_N),
                                                          mlp=a.mlp)), flush=True)
