"""GPU: the hydra plugin slots (nmf_b200/plugins.py, config.py, renderer.py) against the oracle, used the way the
reference's train.py / renderer.py use them."""
import pytest
import torch

from conftest import load_fixture, oracle_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from nmf_b200 import config
    fix = load_fixture("microfacet_g40")
    G = fix["grid_size"]
    t, cfg = config.build_model([f"field.grid_size=[{G},{G},{G}]", "model.arch.bg_module.bg_resolution=32"],
                                aabb=fix["aabb"], near_far=list(fix["near_far"]))
    t.load_state_dict(fix["state"], strict=False)
    t = t.cuda().eval()
    t.sampler.update(t.rf, init=True)
    return fix, oracle_scene(fix), t


def test_update_alpha_mask_matches_reference_volume(model):
    fix, osc, t = model
    t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
    vol = t.sampler.alphaMask.alpha_volume.reshape(-1).cpu()
    assert torch.equal(vol.to(torch.uint8), fix["alpha_volume"].reshape(-1)), (vol != fix["alpha_volume"].reshape(-1).float()).sum()


def test_update_alpha_mask_respects_the_existing_mask(model):
    """samplers/alphagrid.py:209-224: a rebuild evaluates the density only where the CURRENT mask is set, so the occupied
    set can only shrink (up to the 3^3 dilation).  Second rebuild on a field whose density was raised everywhere (a
    from-scratch rebuild would occupy the whole box) against the oracle's rebuild with its existing mask -- bit-equal."""
    from oracle import nmf_oracle as O
    from conftest import oracle_scene as mk
    fix, osc, t = model
    t.sampler.alphaMask = None
    t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
    first = t.sampler.alphaMask.alpha_volume.reshape(-1).cpu().clone()
    state = {k: v.clone() for k, v in fix["state"].items()}
    for p in range(3):                                   # raise the density feature everywhere
        state[f"rf.density_rf.app_plane.{p}"] = state[f"rf.density_rf.app_plane.{p}"].abs() + 0.5
        state[f"rf.density_rf.app_line.{p}"] = state[f"rf.density_rf.app_line.{p}"].abs() + 0.5
    try:
        with torch.no_grad():
            t.load_state_dict(state, strict=False)
        t.invalidate()
        t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
        second = t.sampler.alphaMask.alpha_volume.reshape(-1).cpu()
    finally:
        with torch.no_grad():
            t.load_state_dict(fix["state"], strict=False)
        t.invalidate()
        t.sampler.alphaMask = None
    osc2 = mk(dict(fix, state=state))
    ref = O.build_alpha_volume(osc2, use_existing_mask=True).reshape(-1)
    full = O.build_alpha_volume(osc2, use_existing_mask=False).reshape(-1)
    assert torch.equal(second, ref), int((second != ref).sum())
    assert int(full.sum()) > int(ref.sum()) and int(second.sum()) < second.numel()       # the mask mattered
    t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
    assert torch.equal(t.sampler.alphaMask.alpha_volume.reshape(-1).cpu(), first)


def test_sampler_and_field_slots(model):
    from oracle import nmf_oracle as O
    fix, osc, t = model
    rays = fix["rays"].cuda()
    xyzs, ray_valid, S, z_vals, dists, whole_valid = t.sampler.sample(rays, fix["focal"], rf=t.rf)
    oxyz, ovalid, oz, odists = O.sample_rays(osc, fix["rays"], fix["focal"])
    assert S == osc.n_samples and torch.equal(ray_valid.cpu(), ovalid) and torch.equal(z_vals.cpu(), oz)
    # positions bit-exact; the 4th coordinate z/focal is a torch division whose CUDA kernel multiplies by 1/focal
    assert torch.equal(xyzs.cpu()[:, :3], oxyz[:, :3]) and torch.allclose(xyzs.cpu()[:, 3], oxyz[:, 3], rtol=1e-6)
    assert torch.equal(dists.cpu(), odists) and bool(whole_valid.all())
    sig = t.rf.compute_densityfeature(xyzs).cpu()
    assert torch.allclose(sig, O.feature2density(osc, O.density_feature(osc, oxyz)), rtol=2e-5, atol=1e-6)
    assert torch.allclose(t.rf.compute_appfeature(xyzs).cpu(), O.app_feature(osc, oxyz), atol=2e-6)
    sel = sig > 1e-2
    assert torch.allclose(t.rf.compute_normals(xyzs).cpu()[sel], O.vm_normals(osc, oxyz)[sel], atol=2e-4)


def test_bg_module_slot(model):
    from oracle import nmf_oracle as O
    fix, osc, t = model
    g = torch.Generator().manual_seed(3)
    d = O.unit(torch.randn(4000, 3, generator=g))
    mip = torch.rand(4000, generator=g) * 10 - 8
    out = t.bg_module(d.cuda(), mip.cuda().reshape(-1, 1)).cpu()
    ref = O.env_lookup(osc, d, mip)
    err = (out - ref).abs() / (ref.abs() + 1e-2)
    assert err.max() < 5e-2 and err.mean() < 2e-4
    coeffs, conv = t.bg_module.get_spherical_harmonics(100)
    assert torch.allclose(conv.cpu(), O.sh_conv_state({k[10:]: v for k, v in fix["state"].items() if k.startswith("bg_module.")}), rtol=1e-4, atol=1e-4)   # = the reference's conv_coeffs / pi
    # integral_equirect.py:286-287 reshapes the (1,3,H,W) map to (-1,3) before the mean (sic): reproduced literally
    assert torch.allclose(t.bg_module.mean_color().detach().cpu(), O.env_tables(osc)[0].reshape(-1, 3).mean(dim=0), rtol=1e-5)


def test_ggx_slot(model):
    from oracle import nmf_oracle as O
    fix, osc, t = model
    g = torch.Generator().manual_seed(4)
    n, m = 300, 7
    N = O.unit(torch.randn(n, 3, generator=g))
    V = O.unit(torch.randn(n, 3, generator=g))
    N = N * (V * N).sum(-1, keepdim=True).sign()
    r = torch.rand(n, 1, generator=g) * 0.4 + 0.02
    u = torch.rand(n, m, 2, generator=g)
    mask = torch.rand(n, m, generator=g) < 0.7
    L, cols, lp = O.ggx_sample(u[..., 0], u[..., 1], V, N, r, mask)
    oL, obasis, olp = t.model.brdf_sampler.sample(u[..., 0].cuda(), u[..., 1].cuda(), V.cuda(), N.cuda(), r.cuda(), r.cuda(), mask.cuda())
    assert (oL.cpu() - L).abs().median() < 1e-6 and ((oL.cpu() - L).abs().max(dim=1).values < 1e-3).float().mean() > 0.99
    assert (olp.cpu() - lp).abs().median() < 1e-5
    assert torch.allclose(obasis.cpu(), cols, atol=1e-5)


def test_forward_and_chunk_renderer(model):
    from nmf_b200 import renderer
    from oracle import keyed_rng as KR
    from oracle import nmf_oracle as O
    from test_gpu_parity import compare_images
    fix, osc, t = model
    t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
    rays = fix["rays"].cuda()
    t.seed, t._calls = 11, 0
    ims, stats = t(rays, fix["focal"], is_train=False, ndc_ray=False, N_samples=-1)      # one chunk, like renderer.py:83
    ref, ns = O.render_rays(osc, fix["rays"], fix["focal"], KR.KeyedRNG(), chunk=rays.shape[0], seed=11)
    assert stats["n_samples"][0] == ns[0][0] and bool(stats["whole_valid"].all()) and stats["recur"] == 0
    report, bad = compare_images(ims, ref)
    assert not bad, bad
    # chunked driver, device outputs and host outputs
    ref2, ns2 = O.render_rays(osc, fix["rays"], fix["focal"], KR.KeyedRNG(), chunk=128, seed=11)
    out, st = renderer.chunk_renderer(rays, t, fix["focal"], keys=["rgb_map", "depth", "n_samples"], chunk=128, is_train=False)
    assert set(out) == {"rgb_map", "depth"} and [c[0] for c in st["n_samples"]] == [c[0] for c in ns2]
    assert (out["rgb_map"].cpu() - ref2["rgb_map"]).abs().max() < 2e-3
    out_h, _ = renderer.chunk_renderer(rays, t, fix["focal"], keys=None, chunk=128, render2completion=True)
    assert not out_h["rgb_map"].is_cuda and torch.allclose(out_h["rgb_map"], out["rgb_map"].cpu(), atol=1e-5)
    # host-buffer entry point (nmf_render_rays_host)
    host = renderer.HostRenderer(t.scene(), rays.shape[0], 128)
    ims_h, st_h = host.render(fix["rays"].contiguous().pin_memory(), fix["focal"], seed=11)
    assert torch.allclose(ims_h["rgb_map"], out["rgb_map"].cpu(), atol=1e-5) and torch.equal(ims_h["surf_width"], ref2["surf_width"])
    assert host.h2d_bytes == rays.shape[0] * 24 and host.d2h_bytes > 0


def test_parameter_update_invalidates_scene(model):
    fix, osc, t = model
    rays = fix["rays"][:64].cuda()
    a, _ = t.render_chunks(rays, fix["focal"], chunk=64)
    a = a["rgb_map"].clone()
    with torch.no_grad():
        t.model.diffuse_module.diffuse_mlp[0].bias.add_(1.0)
    b, _ = t.render_chunks(rays, fix["focal"], chunk=64)
    assert (a - b["rgb_map"]).abs().max() > 1e-3
    with torch.no_grad():
        t.model.diffuse_module.diffuse_mlp[0].bias.sub_(1.0)


def test_checkpoint_round_trip(model, tmp_path):
    from nmf_b200 import config
    from nmf_b200.plugins import TensorNeRF
    fix, osc, t = model
    G = fix["grid_size"]
    cfg = config.compose([f"field.grid_size=[{G},{G},{G}]", "model.arch.bg_module.bg_resolution=32"])
    p = str(tmp_path / "ckpt.th")
    t.save(p, cfg.model.arch)
    ck = torch.load(p, weights_only=False)
    assert set(ck) == {"config", "state_dict"}
    t2 = TensorNeRF.load(ck, near_far=list(fix["near_far"])).cuda().eval()
    rays = fix["rays"][:64].cuda()
    t.seed = t2.seed = 3
    a, _ = t.render_chunks(rays, fix["focal"], chunk=64)
    b, _ = t2.render_chunks(rays, fix["focal"], chunk=64)
    assert torch.allclose(a["rgb_map"], b["rgb_map"], atol=1e-5) and torch.equal(a["surf_width"], b["surf_width"])


def test_shading_sub_plugins_and_calibration(model):
    """model.brdf(...) / model.diffuse_module(...) as free-standing plugin calls (modules/brdf.py:177-261,
    render_modules.py:519-574) against the oracle, and the start-of-training calibration (train.py:403-437,
    models/microfacet.py:79-96, render_modules.py:632-642, brdf.py:141-176): after it the mean albedo / roughness / BRDF
    weight sit at their targets and the fused render picks the new biases up without re-packing the factors."""
    import math
    from nmf_b200 import config, train
    from oracle import nmf_oracle as O
    fix, osc, _ = model
    G = fix["grid_size"]                       # its own module instance: the calibration changes biases and budgets
    t, _ = config.build_model([f"field.grid_size=[{G},{G},{G}]", "model.arch.bg_module.bg_resolution=32"],
                              aabb=fix["aabb"], near_far=list(fix["near_far"]))
    t.load_state_dict(fix["state"], strict=False)
    t = t.cuda().eval()
    t.sampler.update(t.rf, init=True)
    t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
    g = torch.Generator().manual_seed(0)
    n = 4000
    feat = (torch.randn(n, 24, generator=g) * 0.3).cuda()
    a, tint, ex = t.model.diffuse_module(None, None, feat)
    ra, rt, rf0, rr = O.material_heads(osc, feat.cpu())
    assert torch.allclose(a.cpu(), ra, atol=2e-6) and torch.allclose(tint.cpu(), rt, atol=2e-6)
    assert torch.allclose(ex["f0"].cpu(), rf0, atol=2e-6) and torch.allclose(ex["r1"].cpu(), rr[:, :1], atol=2e-6)
    w = t.model.diffuse_module.roughness_mlp[0]
    r2 = (torch.sigmoid(feat @ w.weight[1] + w.bias[1] + t.model.diffuse_module.roughness_bias) / 2).clip(1e-2, 1)
    assert torch.allclose(ex["r2"].reshape(-1), r2, atol=2e-6)
    unit = lambda v: v / v.norm(dim=-1, keepdim=True)
    half_l, diff_l = unit(torch.randn(n, 3, generator=g)), unit(torch.randn(n, 3, generator=g))
    rough = torch.rand(n, generator=g) * 0.49 + 0.01
    bw = t.model.brdf(None, None, None, None, None, half_l.cuda(), diff_l.cuda(), feat, rough.cuda(), rough.cuda())
    ref = O.brdf_mlp(osc, feat.cpu(), half_l, diff_l, rough.reshape(-1, 1))
    assert (bw.cpu() - ref).abs().max() < 1e-3 and (bw.cpu() - ref).abs().mean() < 1e-4
    # ---- calibration ----
    rays = fix["rays"][:128].cuda()
    before, _ = t(rays, fix["focal"])
    before = before["rgb_map"].clone()
    t._calls = 0
    sc0 = t.scene()
    args = config.compose([])
    shift0 = t.rf.density_shift
    t.rf.calibrate = True
    # train.py:403-418: density_shift += log(target_sigma) - log(mean density of 20000 random points) (the reference's
    # formula "assumes exponential activation": with softplus it moves towards the target, it does not land on it)
    g1 = torch.Generator(device="cuda").manual_seed(1)
    xyz = (torch.rand(20000, 3, device="cuda", generator=g1) * 2 - 1) * t.rf.aabb[1].reshape(1, 3)
    sigma0 = float(t.rf.compute_densityfeature(xyz).mean())
    target_sigma = -math.log(1 - 5e-3) / (float(t.sampler.stepsize) * t.rf.distance_scale)
    args = train.calibrate_start(t, args, start_density=5e-3, generator=torch.Generator(device="cuda").manual_seed(1))
    assert args.model.arch.model.brdf.bias == t.model.brdf.bias and args.field.density_shift == t.rf.density_shift
    assert abs(t.rf.density_shift - (shift0 + math.log(target_sigma) - math.log(sigma0))) < 1e-5
    assert float(t.rf.compute_densityfeature(xyz).mean()) < sigma0        # the shifted field is thinner, as intended
    bright = float(t.bg_module.mean_color().detach().mean())
    feat2 = t.rf.compute_appfeature(torch.cat([torch.rand(50000, 3, device="cuda") * 2 - 1, torch.zeros(50000, 1, device="cuda")], 1))
    a2, _, ex2 = t.model.diffuse_module(None, None, feat2)
    inv = lambda x: (x / (1 - x)).log()
    target = min(0.5 / bright, 0.999)
    assert abs(float(inv(a2.clip(1e-6, 1 - 1e-6)).mean()) - math.log(target / (1 - target))) < 0.05
    # (the roughness bias takes the reference's single additive step, render_modules.py:640-642, which is exact only in
    # the logit domain of the un-halved value: its result is pinned by test_calibrate_matches_reference_biases)
    assert t.model.diffuse_module.roughness_bias != -1
    after, _ = t(rays, fix["focal"])
    assert not torch.allclose(after["rgb_map"], before, atol=1e-3)      # new biases and density shift reach the kernels
    # the adaptive retrace controller (models/microfacet.py:241-268) moves the budget; the scene only patches scalars
    sc1 = t.scene()
    t.model.update_n_samples([20000])
    assert t.model.max_retrace_rays != [1000]
    sc2 = t.scene()
    assert sc2 is sc1 and sc2.c.max_retrace == t.model.max_retrace_rays[0]
    ims, st = t(rays, fix["focal"])
    nre = st["n_retrace"][0] if isinstance(st["n_retrace"], (list, tuple)) else st["n_retrace"]
    assert nre <= t.model.max_retrace_rays[0]
    t.model.reset_counter()
    assert t.model.max_retrace_rays == [1000] and t.scene().c.max_retrace == 1000


def test_calibrate_matches_reference_biases(model):
    """RandHydraMLPDiffuse.calibrate / MLPBRDF.calibrate against the biases the unmodified reference arrives at on the same
    weights and features (tests/golden/controller.pt; the BRDF calibration draws random directions on both sides, so it
    is compared as a Monte-Carlo estimate)."""
    from nmf_b200 import plugins
    gold = load_fixture("controller")
    h = gold["heads_calibrate"]
    d = plugins.RandHydraMLPDiffuse(in_channels=24, diffuse_bias=h["diffuse_bias0"], roughness_bias=h["roughness_bias0"]).cuda()
    d.load_state_dict(h["state"])
    d.calibrate(torch.tensor(h["brightness"]), True, h["xyz"].cuda(), None, h["feat"].cuda())
    assert abs(d.diffuse_bias - h["diffuse_bias"]) < 2e-4 and abs(d.roughness_bias - h["roughness_bias"]) < 2e-4
    b = gold["brdf_calibrate"]
    brdf = plugins.MLPBRDF(in_channels=24, h_encoder=plugins.ListISH([0, 1, 2, 4]), d_encoder=plugins.ListISH([0, 1, 2, 4]),
                           bias=b["bias0"]).cuda()
    brdf.load_state_dict(b["state"])
    assert brdf.init_val == b["init_val"]
    torch.manual_seed(3)
    brdf.calibrate(h["feat"].cuda(), torch.tensor(h["brightness"]))
    assert abs(brdf.bias - b["bias"]) < 0.05, (brdf.bias, b["bias"])


@pytest.mark.parametrize("name", ["microfacet_g40", "microfacet_noncubic", "forest"])
def test_repack_kernels_match_the_torch_pack(name):
    """csrc/nmf_repack.cu (what a training run rebuilds from the parameters after EVERY optimiser step) against the torch
    restatement of the reference ops it replaces: factor layouts bit-equal, smoothed-difference planes vs F.conv2d
    (modules/grid_sample_Cinf.py:218-242), the summed-area table vs the fp64-accumulated double cumsum
    (modules/integral_equirect.py:431-433) incl. an HDR map with non-trivial brightness / mul (forest-derived), the pole rows,
    and the occupancy bit-fields (voxels, cell OR, coarse) + 0/1 volume vs max_pool3d + threshold (samplers/alphagrid.py:256-261)."""
    from conftest import device_scene, load_fixture
    from nmf_b200.scene import DeviceScene
    fix = load_fixture("microfacet_g40" if name == "forest" else name)
    if name == "forest":
        fe = load_fixture("forest_env")
        fix = dict(fix, state=dict(fix["state"]))
        for k, v in fe["small_state"].items():
            fix["state"]["bg_module." + k] = v
        fix["state"]["bg_module.brightness"] = torch.tensor(1.3, dtype=torch.float64)
        fix["state"]["bg_module.mul"] = torch.tensor(1.7, dtype=torch.float64)
    a = device_scene(fix, "cuda:0")
    DeviceScene._torch_pack = True
    try:
        b = device_scene(fix, "cuda:0")
        vol_b = b.update_alpha_mask()
    finally:
        del DeviceScene._torch_pack
    vol_a = a.update_alpha_mask()
    for p in range(3):
        for k in ("dval", "lval", "aval", "alval"):
            assert torch.equal(a.keep[f"{k}{p}"], b.keep[f"{k}{p}"]), (k, p)
        for k in ("dpack", "lpack"):
            x, y = a.keep[f"{k}{p}"], b.keep[f"{k}{p}"].reshape(a.keep[f"{k}{p}"].shape)
            assert float((x - y).abs().max()) <= 2e-6 * max(1.0, float(y.abs().max())), (k, p, float((x - y).abs().max()))
    # material heads and BRDF MLP operands (nmf_pack_shading): fp32 copies, fp16 and bf16 tensor-core tiles, bit-equal
    for k in ["head_w", "head_b"] + [f"brdf_{n}{i}{sfx}" for i in range(3) for n, sfx in (("w", "t"), ("b", ""), ("w", "u"), ("w", "b"))]:
        assert a.keep[k].dtype == b.keep[k].dtype and torch.equal(a.keep[k], b.keep[k].reshape(a.keep[k].shape)), k
    sa, sb = a.keep["env_sat"], b.keep["env_sat"]
    rel = float(((sa - sb).abs() / (sb.abs() + 1e-6)).max())
    assert rel < 3e-7, rel                                     # at most the last bit of a prefix (fp64 scan order)
    for i in range(3):
        assert abs(a.c.env_top[i] - b.c.env_top[i]) <= 1e-5 * abs(b.c.env_top[i]) and abs(a.c.env_bot[i] - b.c.env_bot[i]) <= 1e-5 * abs(b.c.env_bot[i])
    assert torch.equal(vol_a, vol_b)
    for k in ("occ_vox", "occ_cell", "occ_coarse"):
        assert torch.equal(a.keep[k], b.keep[k]), k
    assert (a.c.ow, a.c.oh, a.c.od, a.c.opitch, a.c.ocw, a.c.och, a.c.ocd) == (b.c.ow, b.c.oh, b.c.od, b.c.opitch, b.c.ocw, b.c.och, b.c.ocd)
    # 5000 lookups at mip -5 over tables that may differ in the last bit of a prefix (sub-texel boxes on an HDR map amplify it)
    assert float((a.keep["sh_conv"] - b.keep["sh_conv"]).abs().max()) <= 1e-3 * float(b.keep["sh_conv"].abs().max())


def test_device_resident_environment_scalars_match_the_host_path():
    """NmfScene.env_dyn (mipbias and the pole-row means read from device memory, tables built by nmf_env_build_sat_dev from
    device-resident brightness / mul / mipbias) against the host path that passes them by value: the same tables bit for bit,
    the same render, the same environment-map gradient finishing pass."""
    from conftest import device_scene, load_fixture
    from nmf_b200 import ops, train
    fix = load_fixture("microfacet_g40")
    fix = dict(fix, state=dict(fix["state"]))
    fix["state"]["bg_module.brightness"] = torch.tensor(0.4, dtype=torch.float64)
    fix["state"]["bg_module.mul"] = torch.tensor(1.3, dtype=torch.float64)
    fix["state"]["bg_module.mipbias"] = torch.tensor(0.7, dtype=torch.float64)
    a = device_scene(fix, "cuda:0")
    b = device_scene(fix, "cuda:0")
    sc = torch.tensor([0.4, 1.3, 0.7], device="cuda")
    st = {k: torch.as_tensor(v) for k, v in fix["state"].items()}
    b._set_env(st, None, dev_scalars=sc)
    assert b.c.env_dyn and not a.c.env_dyn
    assert torch.equal(a.keep["env_sat"], b.keep["env_sat"])
    dyn = b.keep["env_dyn"].cpu()
    assert abs(float(dyn[0]) - a.c.env_mipbias) < 1e-7
    for i in range(3):
        assert abs(float(dyn[1 + i]) - a.c.env_top[i]) <= 1e-6 * abs(a.c.env_top[i])
        assert abs(float(dyn[4 + i]) - a.c.env_bot[i]) <= 1e-6 * abs(a.c.env_bot[i])
    assert torch.allclose(a.keep["sh_conv"], b.keep["sh_conv"], rtol=1e-6, atol=1e-7)
    rays = fix["rays"][:256].cuda()
    ia, _ = ops.render_rays(a, rays, fix["focal"], chunk=128, seed=3)
    ib, _ = ops.render_rays(b, rays, fix["focal"], chunk=128, seed=3)
    assert float((ia["rgb_map"] - ib["rgb_map"]).abs().max()) <= 2e-6
    # directions at the poles exercise env_top / env_bot read through the pointer
    d = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0], [1e-3, 0.0, 1.0], [0.3, 0.2, -0.9]], device="cuda")
    d = torch.nn.functional.normalize(d, dim=-1)
    mip = torch.full((4,), -3.0, device="cuda")
    assert torch.allclose(ops.env_lookup(a, d, mip), ops.env_lookup(b, d, mip), rtol=1e-6, atol=1e-7)
    # the gradient finishing pass with device scalars
    ga, gb = train.MicrofacetGradBuffers(a), train.MicrofacetGradBuffers(b)
    noise = torch.randn(ga.t["gsat"].shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    ga.t["gsat"].copy_(noise); gb.t["gsat"].copy_(noise)
    bg = torch.as_tensor(fix["state"]["bg_module.bg_mat"]).float().cuda()
    ga.finish(bg, 0.4, 1.3)
    gb.finish(bg, None, None, scalars_dev=sc)
    assert torch.equal(ga.t["d_bg"], gb.t["d_bg"])
    assert torch.allclose(ga.t["d_env_scalars"], gb.t["d_env_scalars"], rtol=1e-5)      # sums of per-warp atomics: order varies
