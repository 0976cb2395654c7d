"""GPU: the reverse pass of the MICROFACET model (BASELINE configs #3 / #4) through the C ABI
(nmf_train_microfacet) against torch autograd through the oracle's training forward.

The oracle's loss and the gradient of every parameter are pinned to the UNMODIFIED reference by oracle/check_train.py
(tests/golden/microfacet_*_train*.pt, replayed by tests/test_oracle_golden.py); here the oracle consumes the keyed random
numbers the kernels draw (KeyedRNG), so sample counts are compared exactly and gradients to a stated tolerance
(relative L2 per parameter: 3e-3, 5e-3 for the density factors whose gradient divides by 1 - alpha + 1e-10 and, with
detach_N off, runs through the 5x5 stencil adjoint; 1.5e-2 for the scalar mipbias, a sum of cancelling box-size terms;
measured on a B200: 1e-5 .. 1e-3, mipbias 7e-5 .. 3.5e-3; x4 with the fp16 tensor-core forward MLP)."""
import numpy as np
import pytest
import torch

from conftest import device_scene, load_fixture, oracle_scene

pytestmark = pytest.mark.gpu

LAM_PRED, LAM_ORI = 3e-4, 0.1          # configs/model/microfacet_tensorf2.yaml:209,212


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from nmf_b200 import _lib
    _lib.lib()          # raises if the extension is missing: no fallback
    return torch.device("cuda:0")


def oracle_grads(fix, rays, gt, seed, id0, detach_N, max_samples, min_rough=0.0, **hp):
    from oracle import keyed_rng as KR
    from oracle import nmf_oracle as O
    osc = oracle_scene(fix, requires_grad=True, min_rough=min_rough, **hp)
    keys = KR.primary_ray_keys(seed, np.arange(id0, id0 + rays.shape[0]).astype(np.uint64))
    ims, st = O.render_chunk(osc, rays, fix["focal"], KR.KeyedRNG(), keys, draw_debug=False, is_train=True, detach_N=detach_N,
                             max_samples=max_samples)
    whole = st["whole_valid"]
    photo = ((ims["rgb_map"].clip(0, 1) - gt[whole].clip(0, 1)) ** 2).sum()
    (photo + LAM_PRED * st["prediction_loss"] + LAM_ORI * st["ori_loss"]).backward()
    return osc, ims, st, float(photo.detach())


def compare_grads(got, P, tol=5e-3, tol_density=2e-2, tol_scalar=1.5e-2, min_checked=20, tol_brdf=None):
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
    report = {}
    for k, p in P.items():
        if k not in got:
            continue
        g = got[k].detach().cpu().double().reshape(p.shape) if p.dim() else got[k].detach().cpu().double().reshape(())
        if p.grad is None or float(p.grad.abs().max()) == 0.0:
            assert float(g.abs().max()) < 1e-6, (k, float(g.abs().max()))
            continue
        report[k] = rel(g, p.grad.double())
    assert len(report) >= min_checked, (len(report), sorted(report))
    lim = lambda k: (tol_density if "density_rf" in k else
                     (tol_scalar if k in ("bg_module.mipbias", "bg_module.brightness", "bg_module.mul") else
                      (tol_brdf if (tol_brdf is not None and k.startswith("model.brdf.")) else tol)))
    bad = {k: v for k, v in report.items() if v > lim(k)}
    return report, bad


CASES = [
    # name, n rays, detach_N, max_samples fraction, min_rough, retrace, mlp
    ("microfacet_g40", 96, True, None, 0.0, False, "fp32"),
    ("microfacet_noncubic", 96, False, 0.6, 0.0, False, "fp32"),
    ("microfacet_g40", 24, True, None, 0.0, True, "fp32"),
    ("microfacet_g40", 24, False, None, 0.1, True, "fp32"),
    ("microfacet_noncubic", 24, False, None, 0.0, True, "fp32"),
    ("microfacet_g40", 96, False, None, 0.0, False, "f16"),
    ("microfacet_g40", 24, False, None, 0.0, True, "f16"),
    # rays cast from INSIDE the object (near = 0.05): they leave through back-facing surfaces, so ori_loss > 0 and its
    # gradient (to the weights and, through the normal, to the density factors) is a material part of the step
    ("inside:microfacet_g40", 64, True, None, 0.0, False, "fp32"),
    ("inside:microfacet_noncubic", 64, False, None, 0.0, False, "fp32"),
]


@pytest.mark.parametrize("name,n,detach_N,frac,min_rough,retrace,mlp", CASES)
def test_microfacet_train_step_matches_oracle_gradients(env, name, n, detach_N, frac, min_rough, retrace, mlp):
    """Loss, sample counts and the gradient of EVERY parameter of one nmf_train_microfacet call: one shading level
    (max_retrace_rays = ()) and with every bounce ray re-traced (max_retrace_rays above their number: the no_grad top-k
    selection is the identity), detach_N on / off, with the dynamic batch truncation and min_rough, fp32 and tcgen05 fp16
    forward MLP (the latter with a looser tolerance: the forward BRDF weights carry ~1e-3 of fp16 rounding)."""
    from nmf_b200 import train
    from oracle import keyed_rng as KR
    from oracle import nmf_oracle as O
    inside = name.startswith("inside:")
    fix = load_fixture(name.split(":")[-1])
    hp = dict(max_retrace_rays=(8192,), max_brdf_rays=(650000, 20000)) if retrace else dict(max_retrace_rays=())
    rays = fix["rays"][40:40 + n].contiguous()
    if inside:
        fix = dict(fix, near_far=(0.05, 6.0))
        gen = torch.Generator().manual_seed(11)
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
        rays = torch.cat([0.05 * torch.randn(n, 3, generator=gen), d], dim=1).contiguous()
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    seed, id0 = 21, 300
    ms = -1
    if frac is not None:
        keys = KR.primary_ray_keys(seed, np.arange(id0, id0 + n).astype(np.uint64))
        _, valid, _, _, _ = O.sample_rays(oracle_scene(fix), rays, fix["focal"], None, True, KR.KeyedRNG(), keys, -1)
        ms = int(int(valid.sum()) * frac)
    osc, ims, st, photo = oracle_grads(fix, rays, gt, seed, id0, detach_N, ms, min_rough=min_rough, **hp)
    dsc = device_scene(fix, env, mlp=mlp, **hp)
    out = train.train_microfacet(dsc, rays.cuda(), gt.cuda(), focal=fix["focal"], seed=seed, ray_id0=id0, max_samples=ms,
                                 min_rough=min_rough, lambda_pred=LAM_PRED, lambda_ori=LAM_ORI, detach_N=detach_N)
    kept = int(st["whole_valid"].sum())
    assert out["n_rays"] == kept and (frac is None or 0 < kept < n)
    assert torch.equal(out["whole_valid"].cpu(), st["whole_valid"])
    assert out["n_samples"][0] == st["n_samples"][0]
    if retrace:
        assert len(st["n_samples"]) == 2 and st["n_samples"][1] > 500
        assert abs(out["n_samples"][1] - st["n_samples"][1]) <= max(8, 0.002 * st["n_samples"][1]), (out["n_samples"], st["n_samples"])
        assert out["counters"]["n_retrace"][0] == out["counters"]["n_bounce_rays0"][0]      # every bounce ray was re-traced
    loose = mlp == "f16"
    assert float((out["rgb_map"].cpu() - ims["rgb_map"].detach()).abs().max()) < (2e-3 if loose else 3e-4)
    assert abs(out["loss_photo"] - photo) <= (2e-3 if loose else 2e-4) * max(1.0, photo)
    assert abs(2 * out["sum_acc"] - float(st["prediction_loss"])) <= 1e-4 * float(st["prediction_loss"])
    assert abs(out["ori_loss"] - float(st["ori_loss"])) <= 2e-3 * float(st["ori_loss"]) + 1e-9
    assert (float(st["ori_loss"]) > 1e-3) == inside, float(st["ori_loss"])
    g = out["grads"]
    g.finish(dsc_bg(fix), *env_scalars(fix))
    got = g.reference_views()
    scale = 4.0 if loose else 1.0
    # tcgen05 reverse MLP (mlp="f16"): BF16 operands (8 mantissa bits) through a three-GEMM chain -- measured 2e-3 (last
    # layer) .. 1.3e-2 (first layer) on the BRDF weights, 2.8e-3 on what flows on into the appearance factors
    report, bad = compare_grads(got, osc.params, tol=3e-3 * scale, tol_density=5e-3 * scale, tol_scalar=1.5e-2 * scale,
                                tol_brdf=2.5e-2 if loose else None)
    print(name, "detach_N" if detach_N else "live_N", "retrace" if retrace else "env", mlp, {k: f"{v:.1e}" for k, v in report.items()})
    assert not bad, bad


def dsc_bg(fix):
    return torch.as_tensor(fix["state"]["bg_module.bg_mat"]).float().cuda()


def env_scalars(fix):
    st = fix["state"]
    return float(st.get("bg_module.brightness", 0.0)), float(st.get("bg_module.mul", 1.0))


def test_microfacet_train_step_accumulates_and_is_linear(env):
    """Size-independent properties at a larger batch: two sub-batches accumulated into one set of buffers equal the sum of
    the two separate calls (gradients are sums over rays; the selection of re-traced rays is per call, as in the reference)."""
    from nmf_b200 import train
    fix = load_fixture("microfacet_g56_ship")
    dsc = device_scene(fix, env)
    rays = fix["rays"].contiguous().cuda()
    n = rays.shape[0]
    h = n // 2
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(1)).cuda()
    a = train.train_microfacet(dsc, rays[:h], gt[:h], focal=fix["focal"], seed=3, ray_id0=0, detach_N=False)
    ga = {k: v.clone() for k, v in a["grads"].t.items()}
    b = train.train_microfacet(dsc, rays[h:], gt[h:], focal=fix["focal"], seed=3, ray_id0=h, detach_N=False)
    gb = {k: v.clone() for k, v in b["grads"].t.items()}
    acc = train.MicrofacetGradBuffers(dsc)
    train.train_microfacet(dsc, rays[:h], gt[:h], focal=fix["focal"], seed=3, ray_id0=0, detach_N=False, grads=acc, zero_grads=True)
    train.train_microfacet(dsc, rays[h:], gt[h:], focal=fix["focal"], seed=3, ray_id0=h, detach_N=False, grads=acc, zero_grads=False)
    for k in ("a_plane0", "a_line2", "d_plane0", "d_line0", "basis_t", "head_w", "w0t", "w1t", "b2", "gsat", "gpack0"):
        ref = ga[k] + gb[k]
        err = float((acc.t[k] - ref).abs().max())
        assert err <= 2e-4 * float(ref.abs().max()) + 1e-9, (k, err, float(ref.abs().max()))
        assert float(ref.abs().max()) > 0, k


def test_gradient_hand_over_in_one_launch_matches_the_views(env):
    """nmf_transpose_batch (channel-last kernel layouts -> the reference's parameter layouts, every tensor in one launch)
    writes exactly what copying MicrofacetGradBuffers.reference_views() tensor by tensor writes: bit-equal."""
    from nmf_b200 import train
    fix = load_fixture("microfacet_noncubic")
    dsc = device_scene(fix, env)
    rays = fix["rays"][:128].contiguous().cuda()
    gt = torch.rand(128, 3, generator=torch.Generator().manual_seed(2)).cuda()
    out = train.train_microfacet(dsc, rays, gt, focal=fix["focal"], seed=1, detach_N=False)
    g = out["grads"]
    g.finish(dsc_bg(fix), *env_scalars(fix))
    views = g.reference_views()
    dst = {k: torch.full(tuple(v.shape), float("nan"), device="cuda") for k, v in views.items()}
    g.copy_into(dst)
    assert len(dst) == 31
    for k, v in views.items():
        assert torch.equal(dst[k], v.contiguous()), k
    assert sum(float(v.abs().max()) > 0 for v in views.values()) >= 25


def test_microfacet_train_step_edge_batches(env):
    """Edge cases of one nmf_train_microfacet call: a batch whose rays all miss the box (no samples: the loss is the photometric
    term against the white background, every gradient is exactly zero, nothing overflows or divides by zero), a single ray, and
    a batch size that is not a multiple of anything (ragged tiles everywhere)."""
    from nmf_b200 import train
    fix = load_fixture("microfacet_g40")
    dsc = device_scene(fix, env, max_retrace_rays=(64,), max_brdf_rays=(650000, 20000))
    n = 37
    o = torch.tensor([0.0, 0.0, 4.0]).repeat(n, 1)
    d = torch.nn.functional.normalize(torch.tensor([0.0, 0.0, 1.0]) + 0.01 * torch.randn(n, 3, generator=torch.Generator().manual_seed(0)), dim=-1)
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(1))
    out = train.train_microfacet(dsc, torch.cat([o, d], 1).cuda(), gt.cuda(), focal=fix["focal"], seed=1)
    assert out["n_samples"][0] == 0 and out["n_rays"] == n
    assert abs(out["loss_photo"] - float(((1.0 - gt) ** 2).sum())) < 1e-4
    assert all(float(v.abs().max()) == 0.0 for k, v in out["grads"].t.items() if k != "gsat"), \
        [k for k, v in out["grads"].t.items() if float(v.abs().max()) != 0.0]
    for m in (1, 131):
        rays = fix["rays"][:m].contiguous().cuda()
        g1 = torch.rand(m, 3, generator=torch.Generator().manual_seed(2)).cuda()
        o1 = train.train_microfacet(dsc, rays, g1, focal=fix["focal"], seed=1, detach_N=False)
        assert o1["n_rays"] == m and o1["rgb_map"].shape == (m, 3) and bool(torch.isfinite(o1["rgb_map"]).all())
        assert all(bool(torch.isfinite(v).all()) for v in o1["grads"].t.values())
