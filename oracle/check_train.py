"""TEST INFRASTRUCTURE -- pins the TRAINING forward and its gradients of oracle/nmf_oracle.py against the reference.

    python -m oracle.check_train [--write]         (build container only: needs /root/reference)

For a golden fixture's scene and rays it runs the unmodified reference ``TensorNeRF.forward(is_train=True)`` under
``torch.manual_seed``, forms the training loss of train.py:586-650 (squared error of the clipped colour + the
regularisers with non-zero weights) and back-propagates; then does the same through the oracle (``TorchRNG``,
``Scene(requires_grad=True)``) and compares images, statistics and the gradient of EVERY parameter.  With --write it
stores the reference's loss, statistics and per-parameter gradient norms / probes in tests/golden/<name>_train.pt for
tests/test_oracle_golden.py.  This is the oracle for SURVEY.md section 8f row 1 (training), built ahead of the kernels.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import keyed_rng, make_golden, nmf_oracle  # noqa: E402

STAT_W = dict(ori_loss=0.1, diffuse_reg=0.01, brdf_reg=0.01, prediction_loss=0.001)


def training_loss(ims, stats, target):
    """train.py:586-650 with charbonier_loss = False, hdr = False; the lambdas are test values (all paths exercised)."""
    rgb = ims["rgb_map"].clip(max=1)
    loss = ((rgb.clip(0, 1) - target[stats["whole_valid"]].clip(0, 1)) ** 2).sum()
    for k, w in STAT_W.items():
        loss = loss + w * torch.as_tensor(stats[k]).sum()
    return loss


def probes(g, n=16):
    """a fixed, spread-out subset of a gradient's entries"""
    flat = g.reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, min(n, flat.numel())).long()
    return idx, flat[idx].clone()


def run(name, detach_N, write, retrace=True):
    """retrace=False: Microfacet.max_retrace_rays = [] -- one shading level, every bounce ray goes to the environment
    (microfacet.py:475-476 `else` branch); the configuration the host-composed reverse pass of DESIGN.md section 9 covers."""
    fix = torch.load(os.path.join(make_golden.GOLDEN_DIR, f"{name}.pt"), weights_only=False)
    gsz = fix["grid_size"]
    grid = [gsz] * 3 if isinstance(gsz, int) else list(gsz)
    meta = dict(aabb=fix["aabb"], near_far=fix["near_far"], grid_size=grid, bg_resolution=fix["bg_resolution"])
    ref = make_golden.load_scene_into_reference(fix["state"], meta, fix["model"])
    ref.train()
    if hasattr(ref.model, "detach_N"):
        ref.model.detach_N = detach_N
    if not retrace:
        ref.model.max_retrace_rays = []
    rays, focal, seed = fix["rays"], fix["focal"], fix["seed"]
    target = torch.rand(rays.shape[0], 3, generator=torch.Generator().manual_seed(99))
    torch.manual_seed(seed)
    ims, st = ref(rays, focal, is_train=True, ndc_ray=False, N_samples=-1)
    loss = training_loss(ims, st, target)
    loss.backward()
    ref_grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters() if p.grad is not None}

    model = "microfacet" if fix["model"] == "microfacet_tensorf2" else "plain"
    sc = nmf_oracle.Scene(fix["state"], fix["aabb"], fix["near_far"], grid, alpha_volume=fix["alpha_volume"].float(),
                          requires_grad=True, model=model, **({} if retrace else dict(max_retrace_rays=())))
    torch.manual_seed(seed)
    oi, os_ = nmf_oracle.render_chunk(sc, rays, focal, keyed_rng.TorchRNG(), draw_debug=False, is_train=True,
                                      detach_N=detach_N, max_samples=ref.sampler.max_samples)
    oloss = training_loss(oi, os_, target)
    oloss.backward()
    print(f"[{name} detach_N={detach_N} retrace={retrace}] loss ref {float(loss):.6f} oracle {float(oloss):.6f}  n_samples {st['n_samples']} {os_['n_samples']}")
    assert list(st["n_samples"]) == list(os_["n_samples"])
    assert (ims["rgb_map"] - oi["rgb_map"]).abs().max() < 2e-5
    worst = {}
    for k, g in ref_grads.items():
        og = sc.params.get(k)
        if og is None:
            assert float(g.abs().max()) == 0.0, f"{k}: the reference has a gradient the oracle does not model"
            continue
        og = og.grad if og.grad is not None else torch.zeros_like(g)
        denom = float(g.abs().max()) + 1e-12
        worst[k] = float((g - og.to(g.dtype)).abs().max()) / denom
    # 0-dim parameters (mipbias, brightness, mul) are sums of ~1e5 signed terms in fp32: 1e-2 there, 2e-3 elsewhere
    bad = {k: v for k, v in worst.items() if v > (1e-2 if ref_grads[k].numel() == 1 else 2e-3)}
    print("   max |grad_ref - grad_oracle| / max |grad_ref| :", {k.split(".", 1)[1] if "." in k else k: f"{v:.1e}" for k, v in worst.items()})
    assert not bad, bad
    if write:
        out = dict(name=name, detach_N=detach_N, retrace=retrace, loss=float(loss), target_seed=99, stat_weights=STAT_W,
                   max_samples=int(ref.sampler.max_samples),
                   statistics={k: float(torch.as_tensor(st[k]).sum()) for k in STAT_W},
                   grads={k: dict(norm=float(g.norm()), max=float(g.abs().max()), probes=probes(g)) for k, g in ref_grads.items() if k in sc.params})
        path = os.path.join(make_golden.GOLDEN_DIR, f"{name}_train{'' if retrace else '_noretrace'}{'_dN' if detach_N else ''}.pt")
        torch.save(out, path)
        print("   wrote", path)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true")
    ap.add_argument("--only-noretrace", action="store_true")
    a = ap.parse_args()
    for dn in (True, False):
        run("microfacet_g40", dn, a.write, retrace=False)
    if a.only_noretrace:
        sys.exit(0)
    for nm in ("microfacet_g40", "microfacet_noncubic"):
        for dn in (True, False):
            run(nm, dn, a.write)
    run("plain_g64", True, a.write)
