"""Diagnostic (run under gpurun): per-parameter gradient error of nmf_train_plain against the oracle's autograd."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import device_scene, load_fixture  # noqa: E402
from test_hostmath import oracle_train_plain  # noqa: E402
from nmf_b200 import train  # noqa: E402

fix = load_fixture("plain_g64")
dsc = device_scene(fix, "cuda:0")
for n, seed in ((256, 21), (256, 22), (1000, 21)):
    rays = fix["rays"][:n].contiguous()
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    ref = oracle_train_plain(fix, rays, gt, seed, np.arange(n).astype(np.uint64), -1, 0.001)
    outs = [train.train_plain(dsc, rays.cuda(), gt.cuda(), focal=fix["focal"], seed=seed, lambda_pred=0.001) for _ in range(2)]
    gs = [{k: v.cpu() for k, v in o["grads"].reference_layout().items()} for o in outs]
    err = {k.split(".", 1)[1]: f"{float((gs[0][k] - g).abs().max()) / (float(g.abs().max()) + 1e-12):.1e}"
           for k, g in ref["grads"].items() if k in gs[0]}
    rep = {k.split(".", 1)[1]: f"{float((gs[0][k] - gs[1][k]).abs().max()) / (float(gs[0][k].abs().max()) + 1e-12):.1e}" for k in gs[0]}
    print(n, seed, "rgb", float((outs[0]["rgb_map"].cpu() - ref["rgb_map"]).abs().max()), "loss", outs[0]["loss_photo"], ref["photo"])
    print("  vs oracle:", err)
    print("  run-to-run:", rep)
    # where is the worst entry of app_line.0?
    k = "rf.app_rf.app_line.0"
    d = (gs[0][k] - ref["grads"][k]).abs()
    i = int(d.argmax())
    print("  worst app_line.0 entry", np.unravel_index(i, d.shape), float(gs[0][k].reshape(-1)[i]), float(ref["grads"][k].reshape(-1)[i]),
          "max", float(ref["grads"][k].abs().max()))
