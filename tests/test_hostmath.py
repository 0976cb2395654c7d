"""The kernels' per-element math (nmf_b200/csrc/nmf_math.cuh, nmf_field.cuh), compiled for the host by
tests/hostcheck, against the oracle.  No GPU needed: this catches arithmetic / layout mistakes before a GPU run.
The same header functions are what the CUDA kernels call, the warp-level orchestration is covered by `-m gpu`."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from conftest import device_scene, load_fixture, oracle_scene
from oracle import keyed_rng as KR
from oracle import nmf_oracle as O


def ptr(t):
    return C.c_void_p(t.data_ptr())


@pytest.fixture(scope="module", params=["microfacet_g40", "microfacet_noncubic"])
def scenes(request):
    fix = load_fixture(request.param)
    osc = oracle_scene(fix)
    dsc = device_scene(fix, "cpu", sh_conv=O.sh_irradiance_coeffs(osc))
    return fix, osc, dsc


def test_keyed_rng(hostcheck):
    g = np.random.RandomState(0)
    a = g.randint(0, 2 ** 63, size=1000, dtype=np.int64).astype(np.uint64)
    b = g.randint(0, 5000, size=1000, dtype=np.int64).astype(np.uint64)
    out = np.zeros(1000, dtype=np.uint64)
    hostcheck.hc_mix64(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), 1000, out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, KR.mix64(a, b))
    u = torch.zeros(1000)
    hostcheck.hc_uniform(a.ctypes.data_as(C.c_void_p), C.c_uint32(KR.STREAM_BOUNCE), 1000, ptr(u))
    assert torch.equal(u, KR.uniform(a, KR.STREAM_BOUNCE))
    n = torch.zeros(1000)
    hostcheck.hc_normal(a.ctypes.data_as(C.c_void_p), C.c_uint32(3), C.c_uint32(67), 1000, ptr(n))
    assert torch.allclose(n, KR.normal(a, 3, 67), atol=2e-6)
    nz = torch.zeros(1000, 24)
    hostcheck.hc_noise24(a.ctypes.data_as(C.c_void_p), 1000, ptr(nz))
    ref = KR.noise24(a)
    assert torch.allclose(nz, ref, atol=3e-6)
    assert abs(float(ref.mean())) < 0.02 and abs(float(ref.std()) - 1.0) < 0.02          # 24 000 standard normals
    c = torch.corrcoef(ref.T)
    assert float((c - torch.eye(24)).abs().max()) < 0.15                                   # no cross-feature structure


@pytest.mark.parametrize("name", ["microfacet_g40", "microfacet_g56_ship", "plain_g64", "microfacet_noncubic"])
def test_sampler_mask_bit_exact(hostcheck, name):
    fix = load_fixture(name)
    osc = oracle_scene(fix)
    dsc = device_scene(fix, "cpu", sh_conv=torch.zeros(9, 3))
    rays = fix["rays"][:512].contiguous()
    for override in (None, 3 * float(osc.stepsize)):
        _, valid, z, _ = O.sample_rays(osc, rays, fix["focal"], override)
        S = osc.n_samples
        assert dsc.n_steps == S and dsc.stepsize == float(osc.stepsize)
        v = torch.zeros(rays.shape[0], S, dtype=torch.uint8)
        zz = torch.zeros(rays.shape[0], S)
        hostcheck.hc_sample_rays(dsc.ref(), ptr(rays), rays.shape[0], C.c_float(-1.0 if override is None else override),
                                 ptr(v), ptr(zz))
        assert torch.equal(zz, z)
        assert torch.equal(v.bool(), valid), (v.bool() != valid).sum()
        assert valid.sum() > 1000


def test_sampler_mask_on_lattice_points(hostcheck, scenes):
    """rays that run exactly along lattice planes (zero trilinear weights) and axis-parallel rays"""
    fix, osc, dsc = scenes
    from conftest import grid_of
    gx, gy, gz = grid_of(fix)
    lx, ly, lz = (torch.linspace(-1.5, 1.5, g) for g in (gx, gy, gz))
    rays = []
    for i in range(0, min(gx, gy, gz), 3):
        rays.append([float(lx[i]), -4.0, float(lz[(i * 7) % gz]), 0.0, 1.0, 0.0])
        rays.append([-4.0, float(ly[i]), float(lz[(i * 5) % gz]), 1.0, 0.0, 0.0])
        rays.append([float(lx[i]), float(ly[(i * 3) % gy]), 4.0, 0.0, 0.0, -1.0])
    rays = torch.tensor(rays)
    _, valid, z, _ = O.sample_rays(osc, rays, fix["focal"])
    v = torch.zeros(rays.shape[0], osc.n_samples, dtype=torch.uint8)
    zz = torch.zeros(rays.shape[0], osc.n_samples)
    hostcheck.hc_sample_rays(dsc.ref(), ptr(rays), rays.shape[0], C.c_float(-1.0), ptr(v), ptr(zz))
    assert torch.equal(v.bool(), valid)
    assert valid.sum() > 100


def test_vm_queries(hostcheck, scenes):
    fix, osc, dsc = scenes
    xyz, valid, z, _ = O.sample_rays(osc, fix["rays"][:256], fix["focal"])
    xyz = xyz.contiguous()
    n = xyz.shape[0]
    sig = torch.zeros(n)
    hostcheck.hc_vm_density(dsc.ref(), ptr(xyz), n, 4, 1, ptr(sig))
    ref = O.feature2density(osc, O.density_feature(osc, xyz))
    assert torch.allclose(sig, ref, rtol=2e-5, atol=1e-6), (sig - ref).abs().max()
    feat = torch.zeros(n, 24)
    hostcheck.hc_vm_appfeature(dsc.ref(), ptr(xyz), n, 4, ptr(feat))
    assert torch.allclose(feat, O.app_feature(osc, xyz), atol=2e-6)
    nrm = torch.zeros(n, 3)
    hostcheck.hc_vm_normals(dsc.ref(), ptr(xyz), n, 4, ptr(nrm))
    refn = O.vm_normals(osc, xyz)
    # normals of near-empty space are normalised noise; compare where the gradient is well defined
    sel = ref > 1e-2
    assert sel.sum() > 100
    assert torch.allclose(nrm[sel], refn[sel], atol=2e-4), (nrm[sel] - refn[sel]).abs().max()


def test_env_lookup(hostcheck, scenes):
    fix, osc, dsc = scenes
    g = torch.Generator().manual_seed(0)
    n = 20000
    d = O.unit(torch.randn(n, 3, generator=g))
    # include poles, the seam and axis directions
    d[:6] = torch.tensor([[0, 0, 1.0], [0, 0, -1.0], [-1.0, 1e-4, 0.0], [-1.0, -1e-4, 0.0], [1.0, 0, 0], [0, 1.0, 0]])
    d[6:200, 2] = d[6:200, 2].sign() * 0.999
    d[6:200] = O.unit(d[6:200])
    mip = torch.rand(n, generator=g) * 14 - 10
    out = torch.zeros(n, 3)
    hostcheck.hc_env_lookup(dsc.ref(), ptr(d.contiguous()), ptr(mip), n, ptr(out))
    ref = O.env_lookup(osc, d, mip)
    err = (out - ref).abs() / (ref.abs() + 1e-2)
    # fp32 SAT differences cancel catastrophically for sub-pixel boxes (SURVEY section 7 hard part 4)
    assert err.max() < 5e-2 and err.mean() < 2e-4, (err.max(), err.mean())


def test_ggx_and_bases(hostcheck):
    g = torch.Generator().manual_seed(1)
    n = 20000
    N = O.unit(torch.randn(n, 3, generator=g))
    N[:50] = torch.tensor([0.0, 0.0, 1.0])
    N[50:100] = torch.tensor([0.0, 0.0, -1.0])
    V = O.unit(torch.randn(n, 3, generator=g))
    N = N * (V * N).sum(-1, keepdim=True).sign()
    r = (torch.rand(n, 1, generator=g) * 0.49 + 0.01)
    r[:10] = 0.01
    u = torch.rand(n, 2, generator=g)
    mask = torch.ones(n, 1, dtype=torch.bool)
    L, cols, lpdf = O.ggx_sample(u[:, :1], u[:, 1:], V, N, r, mask)
    H = O.unit((V + L) / 2)
    to_local = cols.permute(0, 2, 1)
    diff_l = torch.matmul(to_local, L.unsqueeze(-1)).squeeze(-1)
    half_l = torch.matmul(to_local, H.unsqueeze(-1)).squeeze(-1)
    oL, olp, oh, od = torch.zeros(n, 3), torch.zeros(n), torch.zeros(n, 3), torch.zeros(n, 3)
    hostcheck.hc_ggx(ptr(u.contiguous()), ptr(V.contiguous()), ptr(N.contiguous()), ptr(r.contiguous()), n,
                     ptr(oL), ptr(olp), ptr(oh), ptr(od))
    ok = (oL - L).abs().max(dim=1).values < 1e-3
    assert ok.float().mean() > 0.999        # grazing reflections amplify rounding
    assert (oL - L).abs().median() < 1e-6
    assert (olp - lpdf).abs().median() < 1e-5 and ((olp - lpdf).abs() < 1e-2).float().mean() > 0.999
    assert (oh - half_l).abs().median() < 1e-6 and (od - diff_l).abs().median() < 1e-6
    ish = torch.zeros(n, 18)
    hostcheck.hc_ish18(ptr(half_l.contiguous()), ptr(r.contiguous()), n, ptr(ish))
    assert torch.allclose(ish, O.ish18(half_l, r), atol=2e-6)
    sh = torch.zeros(n, 9)
    hostcheck.hc_sh9(ptr(N.contiguous()), n, ptr(sh))
    assert torch.allclose(sh, O.sh9(N), atol=1e-6)
    x = torch.rand(n, generator=g) * 2
    y = torch.zeros(n)
    hostcheck.hc_srgb(ptr(x), n, ptr(y))
    assert torch.allclose(y, O.srgb(x, noclip=True), atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
# training slice: nmf_train.cuh on the host against the oracle (pinned to the reference's forward
# AND gradients by tests/test_oracle_golden.py::test_oracle_training_gradients_reproduce_reference)
# ---------------------------------------------------------------------------------------------------------------
def u64ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", ["plain_g64", "microfacet_g40", "microfacet_noncubic"])
def test_train_sampler_bit_exact(hostcheck, name):
    """AlphaGridSampler.sample(is_train=True): jittered cumulative steps, validity mask, bit for bit"""
    fix = load_fixture(name)
    osc = oracle_scene(fix)
    dsc = device_scene(fix, "cpu", sh_conv=torch.zeros(9, 3))
    rays = fix["rays"][:384].contiguous()
    ids = np.arange(1000, 1000 + rays.shape[0]).astype(np.uint64)
    keys = KR.primary_ray_keys(11, ids)
    _, valid, z, dists, whole = O.sample_rays(osc, rays, fix["focal"], None, True, KR.KeyedRNG(), keys, -1)
    S = osc.n_samples
    v = torch.zeros(rays.shape[0], S, dtype=torch.uint8)
    zz = torch.zeros(rays.shape[0], S)
    hostcheck.hc_sample_rays_train(dsc.ref(), ptr(rays), rays.shape[0], C.c_float(-1.0), C.c_uint64(11), C.c_uint64(1000),
                                   None, ptr(v), ptr(zz))
    assert torch.equal(zz, z)
    assert torch.equal(v.bool(), valid)
    assert valid.sum() > 1000 and bool(whole.all())
    # the same ids passed explicitly
    v2, z2 = torch.zeros_like(v), torch.zeros_like(zz)
    hostcheck.hc_sample_rays_train(dsc.ref(), ptr(rays), rays.shape[0], C.c_float(-1.0), C.c_uint64(11), C.c_uint64(0),
                                   u64ptr(ids), ptr(v2), ptr(z2))
    assert torch.equal(z2, zz) and torch.equal(v2, v)


def plain_variant(name):
    """a golden fixture's field (factors, occupancy, rays) under the model=tensorf view MLP: turns the microfacet
    fixtures -- non-cubic grids, density in all three plane/line pairs -- into training cases"""
    from nmf_b200 import synthetic
    fix = dict(load_fixture(name))
    if fix["model"] != "tensorf":
        fix["state"] = dict(fix["state"])
        fix["state"].update(synthetic.plain_mlp_state(3))
        fix["model"] = "tensorf"
    return fix


def oracle_train_plain(fix, rays, gt, seed, ids, max_samples, lambda_pred):
    """loss and gradients of the oracle's training forward (KeyedRNG jitter), reference state_dict keys"""
    osc = oracle_scene(fix, requires_grad=True)
    keys = KR.primary_ray_keys(seed, ids)
    ims, st = O.render_chunk(osc, rays, fix["focal"], KR.KeyedRNG(), keys, draw_debug=False, is_train=True,
                             max_samples=max_samples)
    whole = st["whole_valid"]
    rgb = ims["rgb_map"].clip(max=1)
    photo = ((rgb.clip(0, 1) - gt[whole].clip(0, 1)) ** 2).sum()
    loss = photo + lambda_pred * st["prediction_loss"]
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in osc.params.items()}
    return dict(photo=float(photo.detach()), acc=float(ims["acc_map"].detach().sum()), whole=whole, n_samples=st["n_samples"][0],
                rgb_map=ims["rgb_map"].detach(), grads=grads)


def check_plain_grads(mine, ref, tol=5e-3, tol_density=1e-2):
    """max |g - g_ref| <= tol * max |g_ref| per parameter.  Typical agreement is 1e-5 .. 1e-6 (tools/train_diag.py); the
    tolerance is set by two discontinuities of the gradient itself: a ReLU whose pre-activation is within fp32 rounding
    of zero takes the other branch on the GPU (the forward value does not change, that sample's gradient does by a finite
    amount -- measured 2e-3 of the largest entry for one flipped unit in 17 000 samples), and the density factors'
    gradient divides by (1 - alpha + 1e-10) (cumprod backward, tensor_nerf.py:19-35), which amplifies fp32 rounding on
    opaque surfaces (run-to-run 5e-4 with unordered atomics)."""
    bad = {}
    for k, g in ref.items():
        if k not in mine:
            assert float(g.abs().max()) == 0.0 or k.startswith("bg_module"), k
            continue
        scale = float(g.abs().max()) + 1e-12
        err = float((mine[k].cpu() - g).abs().max()) / scale
        if not err <= (tol_density if "density_rf" in k else tol):
            bad[k] = err
    assert not bad, bad


@pytest.mark.parametrize("name,max_samples", [("plain_g64", -1), ("plain_g64", 2500), ("microfacet_noncubic", -1)])
def test_train_plain_host_gradients(hostcheck, name, max_samples):
    """the per-element forward + backward math of nmf_train_plain (host build) against the oracle's autograd"""
    from nmf_b200 import _lib
    from nmf_b200.train import PlainGradBuffers
    fix = plain_variant(name)
    dsc = device_scene(fix, "cpu")
    n = 96
    rays = fix["rays"][:n].contiguous()
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    ids = np.arange(n).astype(np.uint64)
    ref = oracle_train_plain(fix, rays, gt, 21, ids, max_samples, 0.001)
    gb = PlainGradBuffers(dsc)
    tp = _lib.NmfTrain(n_rays=n, focal=float(fix["focal"]), seed=21, ray_id0=0, ray_ids=None, max_samples=max_samples,
                       cap_samples=1 << 20, lambda_pred=0.001, white_bg=1)
    rgb_map, acc_map = torch.zeros(n, 3), torch.zeros(n)
    whole = torch.zeros(n, dtype=torch.uint8)
    loss = torch.zeros(2, dtype=torch.float64)
    kept = torch.zeros(2, dtype=torch.int32)
    hostcheck.hc_train_plain(dsc.ref(), C.byref(tp), ptr(rays), ptr(gt), C.byref(gb.c), ptr(rgb_map), ptr(acc_map), ptr(whole),
                             ptr(loss), ptr(kept))
    assert torch.equal(whole.bool(), ref["whole"])
    assert int(kept[0]) == int(ref["whole"].sum()) and int(kept[1]) == ref["n_samples"]
    if max_samples > 0:
        assert 0 < int(kept[0]) < n and int(kept[1]) < max_samples
    nk = int(kept[0])
    assert float((rgb_map[:nk] - ref["rgb_map"]).abs().max()) < 2e-5
    assert abs(float(loss[0]) - ref["photo"]) <= 1e-4 * max(1.0, ref["photo"])
    assert abs(float(loss[1]) - ref["acc"]) <= 1e-4 * max(1.0, ref["acc"])
    check_plain_grads(gb.reference_layout(), ref["grads"])


@pytest.mark.parametrize("shape,size", [((16, 40, 40), (56, 56)), ((24, 36, 48), (42, 77)), ((16, 40, 1), (93, 1)),
                                        ((3, 17, 9), (17, 9)), ((2, 300, 1), (128, 1))])
def test_upsample_matches_interpolate(hostcheck, shape, size):
    """TensoRF.upsample (fields/tensoRF.py:208-227): F.interpolate(bilinear, align_corners=True), tap indices exact"""
    src = torch.randn(shape, generator=torch.Generator().manual_seed(0))
    ref = torch.nn.functional.interpolate(src[None], size=size, mode="bilinear", align_corners=True)[0]
    out = torch.zeros(shape[0], *size)
    hostcheck.hc_upsample(ptr(src), shape[0], shape[1], shape[2], ptr(out), size[0], size[1])
    assert float((out - ref).abs().max()) <= 4e-7 * float(src.abs().max())


# ---------------------------------------------------------------------------------------------------------------
# optimiser step (nmf_adam_step / nmf_l1_reg per-element math on the host) against torch's own implementations
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wd,eps,clip,scale", [(0.0, 1e-8, 0.0, 1.0), (1e-6, 1e-15, 10.0, 1.0 / 4096), (1e-2, 1e-15, 0.5, 1.0)])
def test_adam_host_matches_torch(hostcheck, wd, eps, clip, scale):
    """train.py:443-467, 752-755: clip_grad_norm_ + Adam(weight_decay) over several updates with a changing learning rate"""
    g = torch.Generator().manual_seed(3)
    n = 5003
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=0.02, betas=(0.9, 0.99), eps=eps, weight_decay=wd)
    p, m, v = p0.clone(), torch.zeros(n), torch.zeros(n)
    for step in range(1, 8):
        grad = torch.randn(n, generator=g) * (50.0 if step % 2 else 1e-3) / scale
        lr = 0.02 * (0.5 + 0.1 * step)
        ref.grad = grad * scale
        if clip > 0:
            torch.nn.utils.clip_grad_norm_([ref], clip)
        for gr in opt.param_groups:
            gr["lr"] = lr
        opt.step()
        hostcheck.hc_adam(ptr(p), ptr(grad), ptr(m), ptr(v), n, C.c_float(lr), C.c_float(0.9), C.c_float(0.99), C.c_float(eps),
                          C.c_float(wd), step, C.c_float(scale), C.c_float(clip))
        assert torch.allclose(p, ref.detach(), rtol=2e-6, atol=2e-7), (step, (p - ref.detach()).abs().max())
    st = opt.state[ref]
    assert torch.allclose(m, st["exp_avg"], rtol=1e-5, atol=1e-6 * float(m.abs().max()))
    assert torch.allclose(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-6 * float(v.abs().max()))


def test_l1_host_matches_autograd(hostcheck):
    """fields/tensoRF.py:332-340 density_L1 term: weight * mean|p| and its gradient"""
    g = torch.Generator().manual_seed(4)
    p = torch.randn(1, 16, 9, 7, generator=g)
    p.view(-1)[::5] = 0.0
    q = p.clone().requires_grad_(True)
    (8e-5 * q.abs().mean()).backward()
    grad = torch.zeros_like(p)
    hostcheck.hc_l1.restype = C.c_double
    s = hostcheck.hc_l1(ptr(p), p.numel(), C.c_float(8e-5 / p.numel()), ptr(grad))
    assert abs(s - float(p.abs().sum())) < 1e-6 * float(p.abs().sum())
    assert torch.allclose(grad, q.grad, rtol=1e-6, atol=1e-14)
    hostcheck.hc_l1(ptr(p), p.numel(), C.c_float(8e-5 / p.numel()), ptr(grad))        # accumulates
    assert torch.allclose(grad, 2 * q.grad, rtol=1e-6, atol=1e-14)


def test_learning_rate_decay_is_the_reference_formula():
    """utils.py:318-359 (log_lerp with the reverse-cosine delay), configs/model/tensorf.yaml:108-111"""
    from nmf_b200 import train
    for step in (0, 1, 50, 100, 101, 15000, 30000, 40000):
        delay = 0.1 + 0.9 * np.sin(0.5 * np.pi * np.clip(step / 100, 0, 1))
        want = delay * np.exp(np.clip(step / 30000, 0, 1) * (np.log(1e-3) - np.log(1.0)) + np.log(1.0))
        got = train.learning_rate_decay(step, max_steps=30000, **train.REFERENCE_PARAMS)
        assert abs(got - want) < 1e-12
    assert train.learning_rate_decay(7, lr_init=2.0, lr_final=2.0, max_steps=10) == pytest.approx(2.0)


# ---------------------------------------------------------------------------------------------------------------
# microfacet backward, first stage (csrc/nmf_microfacet_bwd.cuh) against the oracle's autograd
# ---------------------------------------------------------------------------------------------------------------
def _ggx_inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    N = O.unit(torch.randn(n, 3, generator=g))
    N[:4] = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0], [0.02, 0.0, 0.9998], [1.0, 0.0, 0.0]])   # frame switch at |n_z| >= 0.999
    V = O.unit(torch.randn(n, 3, generator=g))
    V = torch.where((V * N).sum(-1, keepdim=True) < 0, -V, V)            # the shaded normal faces the viewer (microfacet.py:354-356)
    r = torch.rand(n, 1, generator=g) * 0.49 + 0.01
    r[:8] = 0.01                                                           # roughness at its clip
    u = torch.rand(n, 2, generator=g)
    return u, V, N, r


def test_ggx_roughness_derivative_matches_autograd(hostcheck):
    """d L / d roughness and d H / d roughness of the GGX VNDF sample (brdf_samplers/ggx.py:61-226 with its detach points)
    as forward-mode arithmetic, against torch autograd through the oracle's ggx_sample (one bounce ray per row)."""
    n = 4000
    u, V, N, r = _ggx_inputs(n, 7)
    rr = r.clone().requires_grad_(True)
    L, cols, _ = O.ggx_sample(u[:, :1], u[:, 1:], V, N, rr, torch.ones(n, 1, dtype=torch.bool))
    H = O.unit((V + L) / 2)
    want_dL = torch.stack([torch.autograd.grad(L[:, c].sum(), rr, retain_graph=True)[0].reshape(-1) for c in range(3)], dim=1)
    want_dH = torch.stack([torch.autograd.grad(H[:, c].sum(), rr, retain_graph=True)[0].reshape(-1) for c in range(3)], dim=1)
    oL, odL, oH, odH = (torch.zeros(n, 3) for _ in range(4))
    hostcheck.hc_ggx_dr(ptr(u.contiguous()), ptr(V.contiguous()), ptr(N.contiguous()), ptr(r.reshape(-1).contiguous()), n,
                        ptr(oL), ptr(odL), ptr(oH), ptr(odH))
    assert (oL - L.detach()).abs().max() < 2e-5 and (oH - H.detach()).abs().max() < 2e-5
    # derivatives are O(1 / roughness): relative tolerance on the vector, a handful of samples sit on a branch
    # (u2 at the disk split, L on the horizon) where fp32 rounding picks the other side
    for got, want in ((odL, want_dL), (odH, want_dH)):
        err = (got - want).norm(dim=1) / (want.norm(dim=1) + 1e-2)
        assert (err < 2e-3).float().mean() > 0.995, float((err < 2e-3).float().mean())
        assert float(err.median()) < 1e-5
    assert float(want_dL.norm(dim=1).median()) > 0.1               # the derivative is not trivially zero


def test_fresnel_mix_backward_matches_autograd(hostcheck):
    """models/microfacet.py:584-600: comb = F L_in brdf + (1 - F) diffuse with F = R0 + (1 - R0) clip(1 - |v.h|, 0, 1)^5"""
    g = torch.Generator().manual_seed(2)
    n = 3000
    leaf = lambda *s: torch.rand(*s, generator=g).requires_grad_(True)
    R0, inc, bw, diff, cost = leaf(n, 3), leaf(n, 3), leaf(n, 3), leaf(n, 3), leaf(n, 1)
    with torch.no_grad():
        cost[:5] = torch.tensor([[0.0], [1.0], [1e-8], [0.5], [0.999]])
    up = torch.randn(n, 3, generator=g)
    fres = R0 + (1 - R0) * (1 - cost).clip(min=0, max=1) ** 5
    comb = fres * inc * bw + (1 - fres) * diff
    (comb * up).sum().backward()
    outs = [torch.zeros(n, 3) for _ in range(4)] + [torch.zeros(n)]
    c = lambda t: t.detach().contiguous()
    hostcheck.hc_fresnel_mix_bwd(ptr(c(R0)), ptr(c(cost).reshape(-1)), ptr(c(inc)), ptr(c(bw)), ptr(c(diff)), ptr(up.contiguous()), n,
                                 *[ptr(o) for o in outs])
    for got, leaf_t in zip(outs, (R0, inc, bw, diff, cost)):
        assert torch.allclose(got.reshape(-1), leaf_t.grad.reshape(-1), rtol=1e-5, atol=1e-6)


def test_material_heads_backward_matches_autograd(hostcheck, scenes):
    """modules/render_modules.py:553-560 through the oracle's material_heads on a scene with autograd leaves: gradients of the
    three heads the path uses w.r.t. their weights, biases and the feature"""
    fix = load_fixture("microfacet_g40")
    osc = oracle_scene(fix, requires_grad=True)
    g = torch.Generator().manual_seed(3)
    n = 500
    feat = (torch.randn(n, 24, generator=g) * 0.5).requires_grad_(True)
    feat.data[:3] *= 30                                             # drives the roughness head into its clip
    albedo, tint, f0, r1 = O.material_heads(osc, feat)
    ga, gf, gr = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g), torch.randn(n, 1, generator=g)
    ((albedo * ga).sum() + (f0 * gf).sum() + (r1 * gr).sum()).backward()
    names = ("diffuse", "tint", "f0", "roughness")
    W = torch.cat([fix["state"][f"model.diffuse_module.{h}_mlp.0.weight"] for h in names]).float().contiguous()
    b = torch.cat([fix["state"][f"model.diffuse_module.{h}_mlp.0.bias"] for h in names]).float().contiguous()
    dW, db, dfeat = torch.zeros(11, 24), torch.zeros(11), torch.zeros(n, 24)
    hp = osc.hp
    hostcheck.hc_heads_bwd(ptr(feat.detach().contiguous()), ptr(W), ptr(b), C.c_float(hp["diffuse_mul"]), C.c_float(hp["diffuse_bias"]),
                           C.c_float(hp["f0_bias"]), C.c_float(hp["roughness_bias"]), ptr(ga.contiguous()), ptr(gf.contiguous()),
                           ptr(gr.reshape(-1).contiguous()), n, ptr(dW), ptr(db), ptr(dfeat))
    assert torch.allclose(dfeat, feat.grad, rtol=1e-4, atol=1e-6)
    rows = {"diffuse": slice(0, 3), "tint": slice(3, 6), "f0": slice(6, 9), "roughness": slice(9, 11)}
    for h in names:
        gw = osc.params[f"model.diffuse_module.{h}_mlp.0.weight"].grad
        gb = osc.params[f"model.diffuse_module.{h}_mlp.0.bias"].grad
        gw = torch.zeros_like(W[rows[h]]) if gw is None else gw
        gb = torch.zeros_like(b[rows[h]]) if gb is None else gb
        assert torch.allclose(dW[rows[h]], gw, rtol=1e-4, atol=1e-5 * max(1.0, float(gw.abs().max()))), h
        assert torch.allclose(db[rows[h]], gb, rtol=1e-4, atol=1e-5 * max(1.0, float(gb.abs().max()))), h


@pytest.mark.parametrize("full_size", [False, True])
def test_env_map_gradient_matches_autograd(hostcheck, full_size):
    """d loss / d bg_mat of IntegralEquirect lookups (modules/integral_equirect.py:263-273, 409-504): per-lookup scatter with
    the forward's own box walk (nmf_env_lookup1_bwd_map), then the adjoint of the double cumsum and the exp activation as
    whole-map passes -- against torch autograd through the oracle's env_lookup (boxes of every mip level, wrap-around at
    the seam, pole overhangs, pole rows)."""
    fix = load_fixture("microfacet_g40")
    if full_size:                                                  # the 512 x 1024 map of configs/model/microfacet_tensorf2.yaml
        fix["state"]["bg_module.bg_mat"] = torch.randn(1, 3, 512, 1024, generator=torch.Generator().manual_seed(11)) * 0.7 - 0.5
    osc = oracle_scene(fix, requires_grad=True)
    g = torch.Generator().manual_seed(4)
    n = 20000
    d = O.unit(torch.randn(n, 3, generator=g))
    d[:6] = torch.tensor([[0, 0, 1.0], [0, 0, -1.0], [-1.0, 1e-4, 0.0], [-1.0, -1e-4, 0.0], [1.0, 0, 0], [0, 1.0, 0]])
    d[6:200, 2] = d[6:200, 2].sign() * 0.97                        # near the poles: overhang boxes
    d = O.unit(d)
    sa = torch.rand(n, generator=g) * 14 - 10
    up = torch.randn(n, 3, generator=g)
    (O.env_lookup(osc, d, sa) * up).sum().backward()
    want = osc.params["bg_module.bg_mat"].grad[0]                  # (3,h,w)
    h, w = want.shape[-2:]
    gsat, g_top, g_bot = torch.zeros(h, w, 4), torch.zeros(3), torch.zeros(3)
    hostcheck.hc_env_bwd_map(h, w, C.c_float(float(osc.mipbias.detach())), ptr(d.contiguous()), ptr(sa.contiguous()), ptr(up.contiguous()), n,
                             ptr(gsat), ptr(g_top), ptr(g_bot))
    assert float(g_top.abs().sum()) > 0 and float(g_bot.abs().sum()) > 0
    gs = gsat[..., :3].permute(2, 0, 1).double()                   # (3,h,w)
    dact = gs.flip(1).cumsum(1).flip(1).flip(2).cumsum(2).flip(2)  # adjoint of cumsum over y then x
    dact[:, 0, :] += g_top.double()[:, None] / w                   # pole rows: mean over the row
    dact[:, -1, :] += g_bot.double()[:, None] / w
    with torch.no_grad():
        x = (osc.brightness + osc.mul * osc.bg_mat)[0].double()
        act = torch.exp(x.clip(max=20))
        got = dact * act * float(osc.mul) * (x <= 20)
    scale = float(want.abs().max())
    assert scale > 0
    err = (got.float() - want).abs()
    assert float(err.max()) < 2e-3 * scale and float(err.mean()) < 2e-5 * scale, (float(err.max()) / scale, float(err.mean()) / scale)
    # the same finishing pass as the reverse-pass kernels will run it (fp32, in place)
    fin = torch.zeros(3, h, w)
    hostcheck.hc_env_map_grad_finish(ptr(gsat), h, w, ptr(g_top), ptr(g_bot), ptr(osc.bg_mat.detach()[0].contiguous()),
                                     C.c_float(float(osc.brightness.detach())), C.c_float(float(osc.mul.detach())), ptr(fin))
    err = (fin - want).abs()
    assert float(err.max()) < 2e-3 * scale and float(err.mean()) < 2e-5 * scale, (float(err.max()) / scale, float(err.mean()) / scale)


@pytest.mark.parametrize("name", ["microfacet_g40", "microfacet_g56_ship"])
def test_env_mipbias_gradient_matches_autograd(hostcheck, name):
    """d loss / d IntegralEquirect.mipbias (the box size moves with the bias, integral_equirect.py:373-397): the forward-mode
    pass with the unit tangent on the bias (nmf_env_lookup1_dmipbias, the per-lookup body of k_env_bwd_mipbias) against
    autograd through the oracle's env_lookup; tolerance 2e-3 relative (measured: <= 5e-4)."""
    fix = load_fixture(name)
    dsc = device_scene(fix, "cpu", sh_conv=O.sh_irradiance_coeffs(oracle_scene(fix)))
    g = torch.Generator().manual_seed(4)
    n = 20000
    d = O.unit(torch.randn(n, 3, generator=g))
    d[6:200, 2] = d[6:200, 2].sign() * 0.97
    d = O.unit(d).contiguous()
    sa = (torch.rand(n, generator=g) * 14 - 10).contiguous()
    hostcheck.hc_env_mipbias_grad.restype = C.c_double
    for up in (torch.randn(n, 3, generator=g).abs().contiguous(), torch.randn(n, 3, generator=g).contiguous()):
        osc = oracle_scene(fix, requires_grad=True)                # the scene caches the SAT's graph: one backward per scene
        (O.env_lookup(osc, d, sa) * up).sum().backward()
        want = float(osc.params["bg_module.mipbias"].grad.detach())
        got = hostcheck.hc_env_mipbias_grad(dsc.ref(), ptr(d), ptr(sa), ptr(up), n)
        assert abs(got - want) < 2e-3 * abs(want), (got, want)


def test_env_direction_derivative_matches_autograd(hostcheck, scenes):
    """Directional derivative of an environment lookup along a tangent of the direction (how the bounce radiance moves with
    the roughness through L): forward-mode restatement (nmf_env_lookup1_d) against J^T products of torch autograd through
    the oracle's env_lookup, with the reference's damped atan2 gradient, clip gates and grid_sample coordinate gradient."""
    fix, osc, dsc = scenes
    g = torch.Generator().manual_seed(6)
    n = 6000
    d = O.unit(torch.randn(n, 3, generator=g))
    d[:6] = torch.tensor([[0, 0, 1.0], [0, 0, -1.0], [-1.0, 1e-4, 0.0], [-1.0, -1e-4, 0.0], [1.0, 0, 0], [0, 1.0, 0]])
    d[6:200, 2] = d[6:200, 2].sign() * 0.97
    d = O.unit(d)
    mip = torch.rand(n, generator=g) * 14 - 10
    tangent = torch.randn(n, 3, generator=g)
    dd = d.clone().requires_grad_(True)
    ref = O.env_lookup(osc, dd, mip)
    want = torch.stack([(torch.autograd.grad(ref[:, c].sum(), dd, retain_graph=True)[0] * tangent).sum(-1) for c in range(3)], dim=1)
    rgb, drgb = torch.zeros(n, 3), torch.zeros(n, 3)
    hostcheck.hc_env_lookup_d(dsc.ref(), ptr(d.contiguous()), ptr(tangent.contiguous()), ptr(mip), n, ptr(rgb), ptr(drgb))
    out = torch.zeros(n, 3)
    hostcheck.hc_env_lookup(dsc.ref(), ptr(d.contiguous()), ptr(mip), n, ptr(out))
    verr = (rgb - out).abs() / (out.abs() + 1e-2)                  # same values as the forward tap walk, up to the rounding of a
    assert verr.max() < 5e-2 and verr.mean() < 2e-4, (verr.max(), verr.mean())   # differently ordered fp32 evaluation (sub-pixel boxes)
    # The derivative is a difference of SAT slopes divided by the box size: for sub-pixel boxes (mip level 0) the fp32 SAT
    # differences cancel catastrophically on BOTH sides (hard part 4: ~5e-3 absolute noise on O(1) radiance),
    # so the comparison is global (relative L2) plus per lookup for boxes of more than a texel.  The oracle's own gradient
    # is NaN exactly at the poles (sqrt at 0); the restatement gives 0 there (pole rows are constants).
    ok = ~torch.isnan(want).any(1)
    assert int((~ok).sum()) <= 2 and bool(torch.isfinite(drgb).all())
    lw, lh = O.env_mip_levels(osc, d, mip)
    big = ok & ((lw >= 1.5) | (lh >= 1.5))
    rel = lambda sel: float((want[sel] - drgb[sel]).norm() / want[sel].norm())
    assert rel(ok) < 2e-3 and rel(big) < 5e-4, (rel(ok), rel(big))
    scale = want.abs().max(dim=1).values + 1e-2 * ref.detach().abs().max(dim=1).values + 1e-3
    err = ((drgb - want).abs().max(dim=1).values / scale)[big]
    assert (err < 2e-2).float().mean() > 0.97 and float(err.median()) < 1e-4, (float((err < 2e-2).float().mean()), float(err.median()))
    assert float(want[ok].abs().median()) > 1e-4


def test_bounce_sample_backward_matches_autograd(hostcheck):
    """The composed reverse pass of one shading level without re-trace (models/microfacet.py:352-613; nmf_bounce_sample_bwd =
    GGX derivative -> environment direction derivative + Fresnel angle -> mix -> BRDF MLP -> map scatter), n samples with
    m rays each, against torch autograd through the oracle's own functions wired as in shade_microfacet."""
    import math as _m
    fix = load_fixture("microfacet_g40")
    osc = oracle_scene(fix, requires_grad=True)
    dsc = device_scene(fix, "cpu", sh_conv=O.sh_irradiance_coeffs(oracle_scene(fix)))
    n, m = 400, 5
    g = torch.Generator().manual_seed(8)
    N = O.unit(torch.randn(n, 3, generator=g))
    V = O.unit(torch.randn(n, 3, generator=g))
    V = torch.where((V * N).sum(-1, keepdim=True) < 0, -V, V)
    leaf = lambda t: t.clone().requires_grad_(True)
    nfeat = leaf(torch.randn(n, 24, generator=g) * 0.3)
    R0, diffuse = leaf(torch.rand(n, 3, generator=g)), leaf(torch.rand(n, 3, generator=g))
    rr = leaf(torch.rand(n, 1, generator=g) * 0.4 + 0.08)
    u = torch.rand(n, m, 2, generator=g)
    up = torch.randn(n, 3, generator=g)
    L, cols, lpdf = O.ggx_sample(u[..., 0], u[..., 1], V, N, rr, torch.ones(n, m, dtype=torch.bool))
    ri = torch.arange(n).repeat_interleave(m)
    eV = V[ri]
    H = O.unit((eV + L) / 2)
    to_local = cols.permute(0, 2, 1)
    diff_l = torch.matmul(to_local, L.unsqueeze(-1)).squeeze(-1)
    half_l = torch.matmul(to_local, H.unsqueeze(-1)).squeeze(-1)
    mip = -_m.log(m) - lpdf
    bw = O.brdf_mlp(osc, nfeat[ri], half_l.detach(), diff_l.detach(), rr.detach().expand(n, m).reshape(-1))
    inc = O.env_lookup(osc, L, mip)
    cost = (-eV * H).sum(dim=-1, keepdim=True).abs()
    fres = R0[ri] + (1 - R0[ri]) * (1 - cost).clip(min=0, max=1) ** 5
    comb = fres * inc * bw + (1 - fres) * diffuse[ri]
    reflect = comb.reshape(n, m, 3).mean(dim=1)
    (reflect * up).sum().backward()

    z = lambda *s: torch.zeros(*s)
    dR0, ddiff, dr, dfeat = z(n, 3), z(n, 3), z(n), z(n, 24)
    dw0t, db0, dw1t, db1, dw2t, db2 = z(66, 64), z(64), z(64, 64), z(64), z(64, 4), z(4)
    h, w = osc.bg_mat.shape[-2:]
    gsat, g_top, g_bot = z(h, w, 4), z(3), z(3)
    c = lambda t: t.detach().contiguous()
    hostcheck.hc_bounce_samples_bwd(dsc.ref(), ptr(c(nfeat)), ptr(c(V)), ptr(c(N)), ptr(c(R0)), ptr(c(diffuse)), ptr(c(rr).reshape(-1)),
                                    ptr(u.contiguous()), n, m, ptr(up.contiguous()), ptr(dR0), ptr(ddiff), ptr(dr), ptr(dfeat),
                                    ptr(dw0t), ptr(db0), ptr(dw1t), ptr(db1), ptr(dw2t), ptr(db2), ptr(gsat), ptr(g_top), ptr(g_bot))
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-12))
    assert rel(dR0, R0.grad) < 2e-4 and rel(ddiff, diffuse.grad) < 1e-5 and rel(dfeat, nfeat.grad) < 2e-4       # measured 1e-5
    P = osc.params
    for got, key in ((dw0t.t(), "model.brdf.mlp.0.weight"), (db0, "model.brdf.mlp.0.bias"), (dw1t.t(), "model.brdf.mlp.2.weight"),
                     (db1, "model.brdf.mlp.2.bias"), (dw2t.t(), "model.brdf.mlp.4.weight"), (db2, "model.brdf.mlp.4.bias")):
        assert rel(got, P[key].grad) < 2e-4, (key, rel(got, P[key].grad))                                       # measured 2e-5
    # roughness: through the bounce direction only (Fresnel angle + environment direction); the environment part carries the
    # fp32 SAT-cancellation noise of sub-pixel boxes (peaky lobes -> mip level 0)
    assert rel(dr, rr.grad.reshape(-1)) < 2e-3, rel(dr, rr.grad.reshape(-1))                                          # measured 5e-5
    gs = gsat[..., :3].permute(2, 0, 1).double()
    dact = gs.flip(1).cumsum(1).flip(1).flip(2).cumsum(2).flip(2)
    dact[:, 0, :] += g_top.double()[:, None] / w
    dact[:, -1, :] += g_bot.double()[:, None] / w
    with torch.no_grad():
        x = (osc.brightness + osc.mul * osc.bg_mat)[0].double()
        got_bg = (dact * torch.exp(x.clip(max=20)) * float(osc.mul.detach()) * (x <= 20)).float()
    assert rel(got_bg, P["bg_module.bg_mat"].grad[0]) < 1e-4, rel(got_bg, P["bg_module.bg_mat"].grad[0])             # measured 2e-6


@pytest.mark.parametrize("name,detach_N,max_samples", [("microfacet_g40", True, -1), ("microfacet_noncubic", True, -1),
                                                       ("microfacet_g40", False, -1), ("microfacet_noncubic", False, 2500)])
def test_train_microfacet_host_gradients(hostcheck, name, detach_N, max_samples):
    """The reverse pass of the MICROFACET training forward, composed on the host for one shading level (no re-trace), with
    Microfacet.detach_N on (first iteration) and off (every later one: the bounce direction also moves with the normal, whose
    gradient reaches the density factors through the smoothed-difference planes): loss and the gradient of EVERY parameter
    against autograd through the oracle's render_chunk(is_train=True) -- which oracle/check_train.py pins to the unmodified
    reference -- on the same keyed random numbers."""
    import torch.nn.functional as Fn
    from nmf_b200 import _lib
    from nmf_b200.train import PlainGradBuffers
    fix = load_fixture(name)
    hp = dict(max_retrace_rays=())
    osc = oracle_scene(fix, requires_grad=True, **hp)
    dsc = device_scene(fix, "cpu", sh_conv=O.sh_irradiance_coeffs(oracle_scene(fix)), **hp)
    n, seed = 64, 21
    rays = fix["rays"][:n].contiguous()
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    ids = np.arange(n).astype(np.uint64)
    keys = KR.primary_ray_keys(seed, ids)
    ims, st = O.render_chunk(osc, rays, fix["focal"], KR.KeyedRNG(), keys, draw_debug=False, is_train=True, detach_N=detach_N,
                             max_samples=max_samples)
    assert len(st["n_samples"]) == 1                               # no re-traced level
    whole = st["whole_valid"]
    nk = int(whole.sum())
    assert (max_samples < 0 and nk == n) or (0 < nk < n)           # dynamic batch truncation (alphagrid.py:353-364)
    photo = ((ims["rgb_map"].clip(0, 1) - gt[whole].clip(0, 1)) ** 2).sum()
    # the loss of configs/model/microfacet_tensorf2.yaml:192-218: photometric + pred_lambda * prediction_loss + ori_lambda * ori_loss
    lam_pred, lam_ori = 3e-4, 0.1
    # (ori_loss is 0 on these clean fixtures -- every weighted normal faces the viewer; its gradient path is exercised on its
    # own in test_ori_loss_normal_path_matches_autograd)
    (photo + lam_pred * st["prediction_loss"] + lam_ori * st["ori_loss"]).backward()
    P = osc.params
    gb = PlainGradBuffers(dsc)
    tp = _lib.NmfTrain(n_rays=n, focal=float(fix["focal"]), seed=seed, ray_id0=0, ray_ids=None, max_samples=max_samples,
                       cap_samples=1 << 20, lambda_pred=lam_pred, white_bg=1)
    z = lambda *s: torch.zeros(*s)
    dhw, dhb = z(11, 24), z(11)
    dw0t, db0, dw1t, db1, dw2t, db2 = z(66, 64), z(64), z(64, 64), z(64), z(64, 4), z(4)
    h, w = osc.bg_mat.shape[-2:]
    gsat, g_top, g_bot = z(h, w, 4), z(3), z(3)
    rgb_map, acc_map = z(n, 3), z(n)
    loss = torch.zeros(3, dtype=torch.float64)
    ns = torch.zeros(2, dtype=torch.int32)
    gpack = [torch.zeros_like(dsc.keep[f"dpack{p}"]) for p in range(3)]          # gradient images laid out like dpack / lpack
    glpack = [torch.zeros_like(dsc.keep[f"lpack{p}"]) for p in range(3)]
    parr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
    hostcheck.hc_train_microfacet(dsc.ref(), C.byref(tp), ptr(rays), ptr(gt), C.byref(gb.c), ptr(dhw), ptr(dhb), ptr(dw0t), ptr(db0),
                                  ptr(dw1t), ptr(db1), ptr(dw2t), ptr(db2), ptr(gsat), ptr(g_top), ptr(g_bot), ptr(rgb_map),
                                  ptr(acc_map), ptr(loss), ptr(ns), int(detach_N), parr(gpack), parr(glpack), C.c_float(lam_ori))
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
    assert int(ns[0]) == st["n_samples"][0] and int(ns[1]) == nk
    assert float((rgb_map[:nk] - ims["rgb_map"].detach()).abs().max()) < 2e-4
    assert abs(float(loss[0]) - float(photo.detach())) <= 1e-4 * max(1.0, float(photo.detach()))
    assert abs(float(loss[1]) * 2 - float(st["prediction_loss"].detach())) <= 1e-4 * float(st["prediction_loss"].detach())
    assert abs(float(loss[2]) - float(st["ori_loss"].detach())) <= 1e-3 * float(st["ori_loss"].detach()) + 1e-12
    got = dict(gb.reference_layout())
    if True:
        # normal path (ori_loss always; the bounce direction once detach_N is off): value parts add directly, dx / dy parts
        # go through the adjoint of the 5x5 stencil convolution
        kx, ky = O.derivative_stencils()
        conv = lambda img, k: Fn.conv2d(img.permute(1, 0, 2, 3), k, stride=1, padding=(2, 2)).permute(1, 0, 2, 3)

        def adjoint(shape, k, gimg):
            xz = torch.zeros(shape, requires_grad=True)
            return torch.autograd.grad(conv(xz, k), xz, gimg)[0]
        if not detach_N:
            assert any(float(t.abs().max()) > 0 for t in gpack)
            direct_only = rel(got["rf.density_rf.app_plane.0"], P["rf.density_rf.app_plane.0"].grad)
            assert direct_only > 5e-3, direct_only           # the normal path is a material part of this gradient
        for p in range(3):
            gp = gpack[p].reshape(gpack[p].shape[0], gpack[p].shape[1], 48)
            img = lambda sl: gp[..., sl].permute(2, 0, 1)[None].contiguous()
            key = f"rf.density_rf.app_plane.{p}"
            got[key] = got[key] + img(slice(0, 16)) + adjoint(got[key].shape, kx, img(slice(16, 32))) + adjoint(got[key].shape, ky, img(slice(32, 48)))
            gl = glpack[p].reshape(-1, 4, 8)
            lin_img = lambda sl: gl[:, :, sl].reshape(-1, 16).t()[None, :, :, None].contiguous()
            key = f"rf.density_rf.app_line.{p}"
            got[key] = got[key] + lin_img(slice(0, 4)) + adjoint(got[key].shape, ky, lin_img(slice(4, 8)))
    names = ("diffuse", "tint", "f0", "roughness")
    rows = {"diffuse": slice(0, 3), "tint": slice(3, 6), "f0": slice(6, 9), "roughness": slice(9, 11)}
    for hname in names:
        got[f"model.diffuse_module.{hname}_mlp.0.weight"] = dhw[rows[hname]]
        got[f"model.diffuse_module.{hname}_mlp.0.bias"] = dhb[rows[hname]]
    for i, (wt, b) in zip((0, 2, 4), ((dw0t, db0), (dw1t, db1), (dw2t, db2))):
        got[f"model.brdf.mlp.{i}.weight"] = wt.t()
        got[f"model.brdf.mlp.{i}.bias"] = b
    gs = gsat[..., :3].permute(2, 0, 1).double()
    dact = gs.flip(1).cumsum(1).flip(1).flip(2).cumsum(2).flip(2)
    dact[:, 0, :] += g_top.double()[:, None] / w
    dact[:, -1, :] += g_bot.double()[:, None] / w
    with torch.no_grad():
        x = (osc.brightness + osc.mul * osc.bg_mat)[0].double()
        got["bg_module.bg_mat"] = (dact * torch.exp(x.clip(max=20)) * float(osc.mul.detach()) * (x <= 20)).float()[None]
    report, checked = {}, 0
    for k, p in P.items():
        if p.grad is None or float(p.grad.abs().max()) == 0.0 or k not in got:
            if k in got and p.grad is not None:
                assert float(got[k].abs().max()) < 1e-6 * max(1.0, float(p.grad.abs().max())), k
            continue
        report[k] = rel(got[k].reshape(p.grad.shape), p.grad)
        checked += 1
    print(report)
    assert checked >= 20, checked
    bad = {k: v for k, v in report.items() if v > (2e-2 if "density_rf" in k else 5e-3)}
    assert not bad, bad


def test_ggx_normal_derivative_matches_autograd(hostcheck):
    """d L / d N (the path that opens when Microfacet.detach_N goes off, microfacet.py:352-353): three forward-mode passes,
    one per column, against torch autograd through the oracle's ggx_sample (the tangent frame, the stretched view vector
    and the reflection all move with N)."""
    n = 3000
    u, V, N, r = _ggx_inputs(n, 9)
    N = N.clone()
    N[:4] = O.unit(torch.tensor([[0.3, 0.1, 0.9], [0.0, 0.6, -0.8], [0.5, 0.5, 0.7], [1.0, 0.2, 0.1]]))   # away from the frame switch
    V = torch.where((V * N).sum(-1, keepdim=True) < 0, -V, V)
    NN = N.clone().requires_grad_(True)
    L, _, _ = O.ggx_sample(u[:, :1], u[:, 1:], V, NN, r, torch.ones(n, 1, dtype=torch.bool))
    J = torch.stack([torch.autograd.grad(L[:, a].sum(), NN, retain_graph=True)[0] for a in range(3)], dim=1)   # (n, out a, in c)
    worst = []
    for c in range(3):
        dL, dH = torch.zeros(n, 3), torch.zeros(n, 3)
        hostcheck.hc_ggx_dN(ptr(u.contiguous()), ptr(V.contiguous()), ptr(N.contiguous()), ptr(r.reshape(-1).contiguous()), n, c,
                            ptr(dL), ptr(dH))
        want = J[:, :, c]
        err = (dL - want).norm(dim=1) / (want.norm(dim=1) + 1e-2)
        worst.append(float((err < 2e-3).float().mean()))
        assert float(err.median()) < 1e-5
    assert min(worst) > 0.99, worst


@pytest.mark.parametrize("name", ["microfacet_g40", "microfacet_noncubic"])
def test_ori_loss_normal_path_matches_autograd(hostcheck, name):
    """ori_loss = sum w * min(v.n, 0)^2 (modules/tensor_nerf.py:573-583, ori_lambda = 0.1 in microfacet_tensorf2.yaml) sends a
    gradient through the normal into the density factors: normalisation backward, scatter into dpack / lpack-shaped images,
    adjoint of the 5x5 stencil -- against autograd through the oracle's vm_normals at random points and view directions."""
    import torch.nn.functional as Fn
    fix = load_fixture(name)
    osc = oracle_scene(fix, requires_grad=True)
    dsc = device_scene(fix, "cpu", sh_conv=O.sh_irradiance_coeffs(oracle_scene(fix)))
    g = torch.Generator().manual_seed(12)
    n = 3000
    lo, hi = osc.aabb[0], osc.aabb[1]
    xyz = torch.cat([lo + (hi - lo) * (0.1 + 0.8 * torch.rand(n, 3, generator=g)), torch.zeros(n, 1)], dim=1)
    V = O.unit(torch.randn(n, 3, generator=g))
    wgt = torch.rand(n, generator=g)
    nrm = O.vm_normals(osc, xyz)
    vn = (V * nrm).sum(-1)
    loss = (wgt * vn.clamp(max=0) ** 2).sum()
    assert float(loss.detach()) > 0
    loss.backward()
    gpack = [torch.zeros_like(dsc.keep[f"dpack{p}"]) for p in range(3)]
    glpack = [torch.zeros_like(dsc.keep[f"lpack{p}"]) for p in range(3)]
    parr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
    hostcheck.hc_ori_loss_bwd.restype = C.c_double
    val = hostcheck.hc_ori_loss_bwd(dsc.ref(), ptr(xyz.contiguous()), ptr(V.contiguous()), ptr(wgt.contiguous()), n, parr(gpack), parr(glpack))
    assert abs(val - float(loss.detach())) <= 1e-4 * float(loss.detach())
    kx, ky = O.derivative_stencils()
    conv = lambda img, k: Fn.conv2d(img.permute(1, 0, 2, 3), k, stride=1, padding=(2, 2)).permute(1, 0, 2, 3)

    def adjoint(shape, k, gimg):
        xz = torch.zeros(shape, requires_grad=True)
        return torch.autograd.grad(conv(xz, k), xz, gimg)[0]
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
    for p in range(3):
        want_p, want_l = osc.params[f"rf.density_rf.app_plane.{p}"].grad, osc.params[f"rf.density_rf.app_line.{p}"].grad
        gp = gpack[p].reshape(gpack[p].shape[0], gpack[p].shape[1], 48)
        img = lambda sl: gp[..., sl].permute(2, 0, 1)[None].contiguous()
        got_p = img(slice(0, 16)) + adjoint(want_p.shape, kx, img(slice(16, 32))) + adjoint(want_p.shape, ky, img(slice(32, 48)))
        gl = glpack[p].reshape(-1, 4, 8)
        lin_img = lambda sl: gl[:, :, sl].reshape(-1, 16).t()[None, :, :, None].contiguous()
        got_l = lin_img(slice(0, 4)) + adjoint(want_l.shape, ky, lin_img(slice(4, 8)))
        if float(want_p.abs().max()) > 0:
            assert rel(got_p, want_p) < 2e-3, (p, rel(got_p, want_p))
            assert rel(got_l, want_l) < 2e-3, (p, rel(got_l, want_l))
        # the finishing pass as the reverse-pass kernels will run it (nmf_plane_grad_finish / nmf_line_grad_finish: stencil
        # adjoint per texel, channel-last output added to the d_plane / d_line buffers) equals the autograd adjoint
        H_, W_ = gp.shape[:2]
        fin_p, fin_l = torch.zeros(H_, W_, 16), torch.zeros(gl.shape[0], 16)
        hostcheck.hc_normal_grad_finish(ptr(gpack[p]), H_, W_, ptr(glpack[p]), gl.shape[0], ptr(kx.reshape(-1).contiguous()),
                                        ptr(ky.reshape(-1).contiguous()), ptr(fin_p), ptr(fin_l))
        assert torch.allclose(fin_p.permute(2, 0, 1)[None], got_p, rtol=1e-4, atol=1e-6 * float(got_p.abs().max()) + 1e-12)
        assert torch.allclose(fin_l.t()[None, :, :, None], got_l, rtol=1e-4, atol=1e-6 * float(got_l.abs().max()) + 1e-12)


@pytest.mark.parametrize("name", ["microfacet_g40", "microfacet_g56_ship", "microfacet_noncubic"])
def test_normals_reverse_pass_matches_autograd(hostcheck, name):
    """nmf_normals_bwd_sample + the stencil adjoint -- the per-sample body of k_normals_bwd_scatter and the per-texel body of
    k_normals_bwd_planes / _lines (csrc/nmf_normals_bwd.cu) -- against autograd through the oracle's vm_normals
    (fields/tensor_base.py:107-129) for an arbitrary upstream d loss / d normal."""
    from nmf_b200.scene import derivative_stencils
    fix = load_fixture(name)
    osc = oracle_scene(fix, requires_grad=True)
    dsc = device_scene(fix, "cpu", sh_conv=O.sh_irradiance_coeffs(oracle_scene(fix)))
    g = torch.Generator().manual_seed(21)
    n = 6000
    lo, hi = osc.aabb[0], osc.aabb[1]
    xyz = torch.cat([lo + (hi - lo) * (0.05 + 0.9 * torch.rand(n, 3, generator=g)), torch.zeros(n, 1)], dim=1).contiguous()
    up = torch.randn(n, 3, generator=g)
    up[::7] = 0                                                    # samples without upstream are skipped
    with torch.no_grad():                                          # upstream only where there is density (see the GPU test)
        sig = O.feature2density(osc, O.density_feature(osc, xyz))
    up[sig <= 1e-2] = 0
    assert int((up.abs().sum(1) > 0).sum()) > 500
    (O.vm_normals(osc, xyz) * up).sum().backward()
    gpack = [torch.zeros_like(dsc.keep[f"dpack{p}"]) for p in range(3)]
    glpack = [torch.zeros_like(dsc.keep[f"lpack{p}"]) for p in range(3)]
    parr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
    hostcheck.hc_normals_bwd(dsc.ref(), ptr(xyz), 4, ptr(up), n, parr(gpack), parr(glpack))
    kx, ky = derivative_stencils()
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
    seen = 0
    for p in range(3):
        H_, W_, N_ = gpack[p].shape[0], gpack[p].shape[1], glpack[p].shape[0]
        fin_p, fin_l = torch.zeros(H_, W_, 16), torch.zeros(N_, 16)
        hostcheck.hc_normal_grad_finish(ptr(gpack[p]), H_, W_, ptr(glpack[p]), N_, ptr(kx.reshape(-1).contiguous()),
                                        ptr(ky.reshape(-1).contiguous()), ptr(fin_p), ptr(fin_l))
        want_p, want_l = osc.params[f"rf.density_rf.app_plane.{p}"].grad, osc.params[f"rf.density_rf.app_line.{p}"].grad
        if float(want_p.abs().max()) > 0:
            seen += 1
            assert rel(fin_p.permute(2, 0, 1)[None], want_p) < 1e-4 and rel(fin_l.t()[None, :, :, None], want_l) < 1e-4, p
        else:
            assert float(fin_p.abs().max()) == 0 and float(fin_l.abs().max()) == 0
    assert seen > 0


def test_ggx_view_derivative_matches_autograd(hostcheck):
    """d L / d V and d H / d V: the tangent the shading of a RE-TRACED ray sees (its view vector is minus the parent's bounce
    direction; sample positions are detached in the field, so this is the only way the secondary radiance moves with the
    parent's roughness) -- three forward-mode passes against torch autograd through the oracle's ggx_sample."""
    n = 3000
    u, V, N, r = _ggx_inputs(n, 15)
    N = N.clone()
    N[:4] = O.unit(torch.tensor([[0.3, 0.1, 0.9], [0.0, 0.6, -0.8], [0.5, 0.5, 0.7], [1.0, 0.2, 0.1]]))
    V = torch.where((V * N).sum(-1, keepdim=True) < 0, -V, V)
    VV = V.clone().requires_grad_(True)
    L, _, _ = O.ggx_sample(u[:, :1], u[:, 1:], VV, N, r, torch.ones(n, 1, dtype=torch.bool))
    H = O.unit((VV + L) / 2)
    JL = torch.stack([torch.autograd.grad(L[:, a].sum(), VV, retain_graph=True)[0] for a in range(3)], dim=1)
    JH = torch.stack([torch.autograd.grad(H[:, a].sum(), VV, retain_graph=True)[0] for a in range(3)], dim=1)
    for c in range(3):
        dL, dH = torch.zeros(n, 3), torch.zeros(n, 3)
        hostcheck.hc_ggx_dV(ptr(u.contiguous()), ptr(V.contiguous()), ptr(N.contiguous()), ptr(r.reshape(-1).contiguous()), n, c,
                            ptr(dL), ptr(dH))
        for got, want in ((dL, JL[:, :, c]), (dH, JH[:, :, c])):
            err = (got - want).norm(dim=1) / (want.norm(dim=1) + 1e-2)
            assert (err < 2e-3).float().mean() > 0.99 and float(err.median()) < 1e-5, (c, float((err < 2e-3).float().mean()))


def test_bounce_sample_view_tangent_matches_autograd(hostcheck):
    """d reflect / d V along a tangent (nmf_bounce_sample_tangent): how the radiance a re-traced ray returns moves with the
    parent's bounce direction -- against J^T products of torch autograd through the oracle's functions wired as in
    shade_microfacet (BRDF encodings detached, mip without gradient)."""
    import math as _m
    fix = load_fixture("microfacet_g40")
    osc = oracle_scene(fix)
    dsc = device_scene(fix, "cpu", sh_conv=O.sh_irradiance_coeffs(osc))
    n, m = 300, 5
    g = torch.Generator().manual_seed(16)
    N = O.unit(torch.randn(n, 3, generator=g))
    V = O.unit(torch.randn(n, 3, generator=g))
    V = torch.where((V * N).sum(-1, keepdim=True) < 0, -V, V)
    nfeat = torch.randn(n, 24, generator=g) * 0.3
    R0, diffuse = torch.rand(n, 3, generator=g), torch.rand(n, 3, generator=g)
    rr = torch.rand(n, 1, generator=g) * 0.4 + 0.08
    u = torch.rand(n, m, 2, generator=g)
    tangent = torch.randn(n, 3, generator=g)
    VV = V.clone().requires_grad_(True)
    L, cols, lpdf = O.ggx_sample(u[..., 0], u[..., 1], VV, N, rr, torch.ones(n, m, dtype=torch.bool))
    ri = torch.arange(n).repeat_interleave(m)
    eV = VV[ri]
    H = O.unit((eV + L) / 2)
    to_local = cols.permute(0, 2, 1)
    diff_l = torch.matmul(to_local, L.unsqueeze(-1)).squeeze(-1)
    half_l = torch.matmul(to_local, H.unsqueeze(-1)).squeeze(-1)
    mip = -_m.log(m) - lpdf
    bw = O.brdf_mlp(osc, nfeat[ri], half_l.detach(), diff_l.detach(), rr.expand(n, m).reshape(-1))
    inc = O.env_lookup(osc, L, mip)
    cost = (-eV * H).sum(dim=-1, keepdim=True).abs()
    fres = R0[ri] + (1 - R0[ri]) * (1 - cost).clip(min=0, max=1) ** 5
    reflect = (fres * inc * bw + (1 - fres) * diffuse[ri]).reshape(n, m, 3).mean(dim=1)
    want = torch.stack([(torch.autograd.grad(reflect[:, c].sum(), VV, retain_graph=True)[0] * tangent).sum(-1) for c in range(3)], dim=1)
    refl, drefl = torch.zeros(n, 3), torch.zeros(n, 3)
    c_ = lambda t: t.detach().contiguous()
    hostcheck.hc_bounce_samples_tangent(dsc.ref(), ptr(c_(nfeat)), ptr(c_(V)), ptr(c_(tangent)), ptr(c_(N)), ptr(c_(R0)), ptr(c_(diffuse)),
                                        ptr(c_(rr).reshape(-1)), ptr(u.contiguous()), n, m, ptr(refl), ptr(drefl))
    assert torch.allclose(refl, reflect.detach(), rtol=1e-3, atol=1e-3)
    rel = float((drefl - want).norm() / want.norm())
    assert rel < 2e-3, rel


@pytest.mark.parametrize("name,detach_N", [("microfacet_g40", True), ("microfacet_g40", False), ("microfacet_noncubic", False)])
def test_train_microfacet_retrace_host_gradients(hostcheck, name, detach_N):
    """The reverse pass of the microfacet training forward WITH its re-traced level, composed on the host (tests/hostcheck
    hc_train_microfacet_retrace): every bounce ray of the primary samples is re-traced (max_retrace_rays above their number, so
    the top-k selection is the identity), the secondary rays are marched with their own jitter, shaded with the budgeted
    bounce counts of recur = 1 (pt_selectors.py), composited over the environment; the reverse pass sends each parent ray's
    d L_in into its secondary ray (parameters of both levels) and brings the secondary radiance's dependence on the parent's
    bounce direction back as a tangent (view vector of the level-1 shading + background lookup; positions are detached,
    tensoRF.py:182-183).  Loss, sample counts of both levels and the gradient of EVERY parameter against autograd through
    the oracle's render_chunk(is_train=True), detach_N on and off."""
    import torch.nn.functional as Fn
    from nmf_b200 import _lib
    from nmf_b200.train import PlainGradBuffers
    fix = load_fixture(name)
    hp = dict(max_retrace_rays=(100000,), max_brdf_rays=(650000, 20000))
    osc = oracle_scene(fix, requires_grad=True, **hp)
    dsc = device_scene(fix, "cpu", sh_conv=O.sh_irradiance_coeffs(oracle_scene(fix)), **hp)
    n, seed = 12, 21
    rays = fix["rays"][40:40+n].contiguous()
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    keys = KR.primary_ray_keys(seed, np.arange(n).astype(np.uint64))
    ims, st = O.render_chunk(osc, rays, fix["focal"], KR.KeyedRNG(), keys, draw_debug=False, is_train=True, detach_N=detach_N)
    photo = ((ims["rgb_map"].clip(0, 1) - gt.clip(0, 1)) ** 2).sum()
    photo.backward()
    P = osc.params
    gb = PlainGradBuffers(dsc)
    tp = _lib.NmfTrain(n_rays=n, focal=float(fix["focal"]), seed=seed, ray_id0=0, ray_ids=None, max_samples=-1, cap_samples=1 << 20, lambda_pred=0.0, white_bg=1)
    z = lambda *s: torch.zeros(*s)
    dhw, dhb = z(11, 24), z(11)
    dw0t, db0, dw1t, db1, dw2t, db2 = z(66, 64), z(64), z(64, 64), z(64), z(64, 4), z(4)
    h, w = osc.bg_mat.shape[-2:]
    gsat, g_top, g_bot = z(h, w, 4), z(3), z(3)
    rgb_map = z(n, 3)
    loss = torch.zeros(3, dtype=torch.float64)
    ns = torch.zeros(2, dtype=torch.int32)
    gpack = [torch.zeros_like(dsc.keep[f"dpack{p}"]) for p in range(3)]
    glpack = [torch.zeros_like(dsc.keep[f"lpack{p}"]) for p in range(3)]
    parr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
    hostcheck.hc_train_microfacet_retrace(dsc.ref(), C.byref(tp), ptr(rays), ptr(gt), C.byref(gb.c), ptr(dhw), ptr(dhb), ptr(dw0t), ptr(db0), ptr(dw1t), ptr(db1),
                                   ptr(dw2t), ptr(db2), ptr(gsat), ptr(g_top), ptr(g_bot), ptr(rgb_map), ptr(loss), ptr(ns), int(detach_N), parr(gpack), parr(glpack))
    assert ns.tolist() == list(st["n_samples"]) and len(st["n_samples"]) == 2 and ns[1] > 1000      # both levels, sample counts identical
    assert float((rgb_map - ims["rgb_map"].detach()).abs().max()) < 2e-4
    assert abs(float(loss[0]) - float(photo.detach())) <= 1e-4 * max(1.0, float(photo.detach()))
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
    got = dict(gb.reference_layout())
    kx, ky = O.derivative_stencils()
    conv = lambda img, k: Fn.conv2d(img.permute(1, 0, 2, 3), k, stride=1, padding=(2, 2)).permute(1, 0, 2, 3)
    def adjoint(shape, k, gimg):
        xz = torch.zeros(shape, requires_grad=True)
        return torch.autograd.grad(conv(xz, k), xz, gimg)[0]
    for p in range(3):
        gp = gpack[p].reshape(gpack[p].shape[0], gpack[p].shape[1], 48)
        img = lambda sl: gp[..., sl].permute(2, 0, 1)[None].contiguous()
        key = f"rf.density_rf.app_plane.{p}"
        got[key] = got[key] + img(slice(0, 16)) + adjoint(got[key].shape, kx, img(slice(16, 32))) + adjoint(got[key].shape, ky, img(slice(32, 48)))
        gl = glpack[p].reshape(-1, 4, 8)
        lin_img = lambda sl: gl[:, :, sl].reshape(-1, 16).t()[None, :, :, None].contiguous()
        key = f"rf.density_rf.app_line.{p}"
        got[key] = got[key] + lin_img(slice(0, 4)) + adjoint(got[key].shape, ky, lin_img(slice(4, 8)))
    names = ("diffuse", "tint", "f0", "roughness")
    rows = {"diffuse": slice(0, 3), "tint": slice(3, 6), "f0": slice(6, 9), "roughness": slice(9, 11)}
    for hname in names:
        got[f"model.diffuse_module.{hname}_mlp.0.weight"] = dhw[rows[hname]]
        got[f"model.diffuse_module.{hname}_mlp.0.bias"] = dhb[rows[hname]]
    for i, (wt, b) in zip((0, 2, 4), ((dw0t, db0), (dw1t, db1), (dw2t, db2))):
        got[f"model.brdf.mlp.{i}.weight"] = wt.t(); got[f"model.brdf.mlp.{i}.bias"] = b
    fin = torch.zeros(3, h, w)
    hostcheck.hc_env_map_grad_finish(ptr(gsat), h, w, ptr(g_top), ptr(g_bot), ptr(osc.bg_mat.detach()[0].contiguous()), C.c_float(float(osc.brightness.detach())), C.c_float(float(osc.mul.detach())), ptr(fin))
    got["bg_module.bg_mat"] = fin[None]
    report = {k: rel(got[k].reshape(p.grad.shape), p.grad) for k, p in P.items()
              if p.grad is not None and float(p.grad.abs().max()) > 0.0 and k in got}
    assert len(report) >= 20, len(report)
    bad = {k: v for k, v in report.items() if v > (2e-2 if "density_rf" in k else 5e-3)}
    assert not bad, bad
