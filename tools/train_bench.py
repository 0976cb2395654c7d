"""Times the fused training step of model=tensorf (nmf_train_plain) on the synthetic G=300 scene and runs a short
optimiser loop.  Run under gpurun:  python tools/train_bench.py [--rays 4096] [--steps 20] [--iters 60]
Prints one JSON line (rays/s and valid samples/s of the forward+backward step, CUDA-event timed; per-iteration time and
MSE of the Adam loop)."""
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
    ap.add_argument("--iters", type=int, default=60)
    a = ap.parse_args()
    print(json.dumps(train.benchmark_plain(a.grid, a.rays, a.steps, a.iters)))
    print(json.dumps(train.benchmark_microfacet_forward(a.grid, a.rays, a.steps)))
