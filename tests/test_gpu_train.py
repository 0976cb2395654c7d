"""GPU: the training slice (model=tensorf) through the C ABI against the oracle.

The oracle's training forward AND its gradients are pinned to the unmodified reference (oracle/check_train.py,
tests/test_oracle_golden.py::test_oracle_training_gradients_reproduce_reference); here the oracle consumes the keyed
jitter (KeyedRNG) the kernels draw, so sampling is compared bit for bit and losses / gradients to a stated tolerance
(2e-3 of each gradient's largest entry: fp32 atomics accumulate ~1e5 terms per entry in arbitrary order)."""
import numpy as np
import pytest
import torch

from conftest import device_scene, load_fixture, oracle_scene
from test_hostmath import check_plain_grads, oracle_train_plain, plain_variant

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from nmf_b200 import _lib
    _lib.lib()          # raises if the extension is missing: no fallback
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", ["plain_g64", "microfacet_g40", "microfacet_noncubic"])
def test_train_sampler_bit_exact(env, name):
    """A1/A2 in train mode: jittered cumulative steps, validity, per-ray counts, dynamic batch truncation"""
    from nmf_b200 import train
    from oracle import keyed_rng as KR
    from oracle import nmf_oracle as O
    fix = load_fixture(name)
    osc, dsc = oracle_scene(fix), device_scene(fix, env)
    rays = fix["rays"][:1024].contiguous()
    ids = np.arange(500, 500 + rays.shape[0]).astype(np.uint64)
    keys = KR.primary_ray_keys(7, ids)
    _, valid, z, _, _ = O.sample_rays(osc, rays, fix["focal"], None, True, KR.KeyedRNG(), keys, -1)
    total = int(valid.sum())
    for max_samples in (-1, total // 2, total + 10):
        v, zz, nv, whole, kept = train.sample_rays_train(dsc, rays.cuda(), seed=7, ray_id0=500, max_samples=max_samples)
        assert torch.equal(zz.cpu(), z)
        assert torch.equal(v.cpu(), valid)
        assert torch.equal(nv.cpu().long(), valid.sum(1))
        _, ov, _, _, owhole = O.sample_rays(osc, rays, fix["focal"], None, True, KR.KeyedRNG(), keys, max_samples)
        assert torch.equal(whole.cpu(), owhole)
        assert kept.tolist() == [int(owhole.sum()), int(ov.sum())]
        if 0 < max_samples < total:
            assert 0 < int(kept[0]) < rays.shape[0]
    # explicit ray ids and a near override (the secondary-ray form of the call)
    v2, z2, _, _, _ = train.sample_rays_train(dsc, rays.cuda(), seed=7, ray_ids=torch.from_numpy(ids.astype(np.int64)))
    assert torch.equal(z2.cpu(), z) and torch.equal(v2.cpu(), valid)
    near = 3 * float(osc.stepsize)
    _, ovn, ozn, _, _ = O.sample_rays(osc, rays, fix["focal"], near, True, KR.KeyedRNG(), keys, -1)
    v3, z3, _, _, _ = train.sample_rays_train(dsc, rays.cuda(), seed=7, ray_id0=500, override_near=near)
    assert torch.equal(z3.cpu(), ozn) and torch.equal(v3.cpu(), ovn)


def test_sampler_plugin_train_mode(env):
    """AlphaGridSampler.sample(is_train=True) as TensorNeRF.forward calls it (tensor_nerf.py:237-255)"""
    from nmf_b200 import config
    from oracle import keyed_rng as KR
    from oracle import nmf_oracle as O
    fix = load_fixture("microfacet_g40")
    G = fix["grid_size"]
    t, _ = config.build_model([f"field.grid_size=[{G},{G},{G}]", "model.arch.bg_module.bg_resolution=32"],
                              aabb=fix["aabb"], near_far=list(fix["near_far"]))
    t.load_state_dict(fix["state"], strict=False)
    t = t.cuda().eval()
    t.sampler.update(t.rf, init=True)
    t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
    osc = oracle_scene(fix)
    rays = fix["rays"][:512]
    keys = KR.primary_ray_keys(3, np.arange(rays.shape[0]))
    t.sampler.max_samples = 6000
    oxyz, ovalid, oz, odists, owhole = O.sample_rays(osc, rays, fix["focal"], None, True, KR.KeyedRNG(), keys, 6000)
    xyzs, ray_valid, S, z_vals, dists, whole = t.sampler.sample(rays.cuda(), fix["focal"], rf=t.rf, is_train=True, seed=3)
    assert 0 < int(owhole.sum()) < rays.shape[0]
    assert torch.equal(whole.cpu(), owhole) and torch.equal(ray_valid.cpu(), ovalid) and torch.equal(z_vals.cpu(), oz)
    assert torch.equal(dists.cpu(), odists) and torch.equal(xyzs.cpu()[:, :3], oxyz[:, :3])


@pytest.mark.parametrize("name,n,max_samples", [("plain_g64", 256, -1), ("plain_g64", 256, 6000), ("plain_g64", 1000, -1),
                                                ("microfacet_noncubic", 256, -1), ("microfacet_g40", 300, 5000)])
def test_train_plain_matches_oracle_gradients(env, name, n, max_samples):
    """nmf_train_plain: loss, images, whole_valid and the gradient of EVERY parameter against the oracle's autograd
    (plain_g64: the model=tensorf fixture; the microfacet fixtures' fields -- non-cubic grid, density in all three
    plane/line pairs -- under the same view MLP)"""
    from nmf_b200 import train
    fix = plain_variant(name)
    dsc = device_scene(fix, env)
    rays = fix["rays"][:n].contiguous()
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    ids = np.arange(n).astype(np.uint64)
    ref = oracle_train_plain(fix, rays, gt, 21, ids, max_samples, 0.001)
    out = train.train_plain(dsc, rays.cuda(), gt.cuda(), focal=fix["focal"], seed=21, max_samples=max_samples,
                            lambda_pred=0.001, cap_samples=4096)       # small on purpose: exercises the regrow path
    assert torch.equal(out["whole_valid"].cpu(), ref["whole"])
    assert out["n_rays"] == int(ref["whole"].sum()) and out["n_samples"] == ref["n_samples"]
    if max_samples > 0:
        assert 0 < out["n_rays"] <= n
    nk = out["n_rays"]
    assert float((out["rgb_map"][:nk].cpu() - ref["rgb_map"]).abs().max()) < 2e-5
    assert abs(out["loss_photo"] - ref["photo"]) <= 1e-4 * max(1.0, ref["photo"])
    assert abs(out["sum_acc"] - ref["acc"]) <= 1e-4 * max(1.0, ref["acc"])
    check_plain_grads(out["grads"].reference_layout(), ref["grads"])


def test_train_plain_directional_derivative(env):
    """size-independent property: <grad, v> equals the central finite difference of the loss along v (fixed jitter seed)"""
    from nmf_b200 import train
    from nmf_b200.scene import DeviceScene
    from conftest import grid_of
    fix = load_fixture("plain_g64")
    n = 2048
    rays = fix["rays"][:n].cuda()
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(9)).cuda()
    state = {k: v.clone() for k, v in fix["state"].items()}

    def scene_of(st):
        return DeviceScene(st, fix["aabb"], fix["near_far"], grid_of(fix), alpha_volume=fix["alpha_volume"], device=env,
                           model="plain")

    def loss_of(st):
        o = train.train_plain(scene_of(st), rays, gt, focal=fix["focal"], seed=4)
        return o["loss_photo"], o

    _, out = loss_of(state)
    grads = {k: v.cpu().double() for k, v in out["grads"].reference_layout().items()}
    g = torch.Generator().manual_seed(1)
    for keys, eps in ((["model.diffuse_module.mlp.4.bias", "model.diffuse_module.mlp.2.weight"], 2e-3),
                      (["rf.basis_mat.weight"], 2e-3),
                      ([f"rf.app_rf.app_plane.{p}" for p in range(3)] + [f"rf.app_rf.app_line.{p}" for p in range(3)], 2e-3),
                      ([f"rf.density_rf.app_plane.{p}" for p in range(3)], None)):
        if eps is None:
            # a random direction over whole density planes has a derivative below the fp32 noise of the loss difference:
            # use the planes themselves (a multiplicative change of the density factors)
            v, eps = {k: state[k].clone() for k in keys}, 1e-3
        else:
            v = {k: torch.randn(state[k].shape, generator=g) for k in keys}
        plus = dict(state); minus = dict(state)
        for k in keys:
            plus[k] = state[k] + eps * v[k]
            minus[k] = state[k] - eps * v[k]
        fd = (loss_of(plus)[0] - loss_of(minus)[0]) / (2 * eps)
        an = sum(float((grads[k] * v[k].double()).sum()) for k in keys)
        # absolute term: the loss is a sum of ~6000 fp32 terms accumulated with atomics, its run-to-run noise (~5e-6) divided
        # by 2 eps is a few 1e-3 (seen: |fd - an| = 3.5e-3 on the density planes, where the derivative itself is 5e-3)
        assert abs(fd - an) <= 3e-2 * max(abs(an), abs(fd)) + 6e-3, (keys[0], fd, an)


def test_trainer_reduces_loss(env):
    """the optimiser loop (train.PlainTrainer): perturbed weights are pulled back towards the images they came from"""
    from conftest import grid_of
    from nmf_b200 import ops, train
    fix = load_fixture("plain_g64")
    dsc = device_scene(fix, env)
    n = 2048
    rays = fix["rays"][:n].cuda()
    target, _ = ops.render_rays(dsc, rays, fix["focal"], chunk=n, skip_eps=0.0, t_cut=0.0)
    gt = target["rgb_map"].clone()
    g = torch.Generator().manual_seed(2)
    state = {k: v.clone() for k, v in fix["state"].items()}
    for k in train.PLAIN_PARAM_KEYS:
        if "app_" in k and "density" not in k or "mlp" in k:
            state[k] = state[k] + 0.3 * state[k].abs().mean() * torch.randn(state[k].shape, generator=g)
    tr = train.PlainTrainer(state, fix["aabb"], fix["near_far"], grid_of(fix), alpha_volume=fix["alpha_volume"], device=env,
                            lr_grid=2e-2, lr_net=1e-3)
    mse = [tr.step(rays, gt)["mse"] for _ in range(60)]
    assert mse[-1] < 0.6 * mse[0], (mse[0], mse[-1])
    assert all(np.isfinite(mse))


def test_plugin_train_step_fills_grads(env):
    """TensorNeRF.train_step for model=tensorf: gradients land on the plugin's parameters under the reference's names"""
    from nmf_b200 import config
    fix = load_fixture("plain_g64")
    G = fix["grid_size"]
    t, _ = config.build_model(["model=tensorf", f"field.grid_size=[{G},{G},{G}]"], aabb=fix["aabb"],
                              near_far=list(fix["near_far"]))
    missing = t.load_state_dict(fix["state"], strict=False)
    t = t.cuda().train()
    t.sampler.update(t.rf, init=True)
    from nmf_b200.plugins import AlphaGridMask
    t.sampler.alphaMask = AlphaGridMask(t.rf.aabb, fix["alpha_volume"].float().cuda()).cuda()
    n = 512
    rays, gt = fix["rays"][:n].cuda(), torch.rand(n, 3, generator=torch.Generator().manual_seed(5)).cuda()
    loss, images, stats = t.train_step(rays, gt, focal=fix["focal"], lambda_pred=0.0)
    assert np.isfinite(loss) and images["rgb_map"].shape == (n, 3) and stats["n_samples"][0] > 1000
    params = dict(t.named_parameters())
    from nmf_b200.train import PLAIN_PARAM_KEYS
    for k in PLAIN_PARAM_KEYS:
        assert params[k].grad is not None and params[k].grad.shape == params[k].shape, k
        # the fixture's density lives in plane / line 0 only: the other pairs' products -- and their gradients -- are zero
        if not ("density_rf" in k and not k.endswith(".0")):
            assert float(params[k].grad.abs().max()) > 0, k
    with pytest.raises(NotImplementedError):
        t.forward(rays, fix["focal"], is_train=True)


def _ddp_worker(rank, world, port, tmp):
    import torch.distributed as dist
    from conftest import grid_of
    from nmf_b200 import train
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dev = torch.device("cuda", rank)
    fix = load_fixture("plain_g64")
    n = 2048
    rays = fix["rays"][:n]
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    mk = lambda: train.PlainTrainer(fix["state"], fix["aabb"], fix["near_far"], grid_of(fix),
                                    alpha_volume=fix["alpha_volume"], device=dev, seed=3)
    ids = torch.arange(n)
    lo, hi = rank * n // world, (rank + 1) * n // world
    tr = mk()
    for _ in range(5):                                   # each rank: its own rays, global ray ids, one flat all-reduce
        tr.step(rays[lo:hi].to(dev), gt[lo:hi].to(dev), ray_ids=ids[lo:hi])
    sharded = {k: p.detach().cpu() for k, p in tr.params.items()}
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        single = mk()
        for _ in range(5):
            single.step(rays.to(dev), gt.to(dev), ray_ids=ids)
        worst = {k: float((sharded[k] - p.detach().cpu()).abs().max()) for k, p in single.params.items()}
        torch.save(worst, tmp)


def test_ray_sharded_training_two_gpus(env, tmp_path):
    """SURVEY 8e / BASELINE config #4 on the training slice: two ranks train on disjoint halves of the ray batch, ONE flat
    NCCL all-reduce of the gradients per step; the parameters after 5 Adam steps equal the single-GPU run on all rays
    (keyed jitter: every ray draws the same steps wherever it is rendered)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    out = str(tmp_path / "worst.pt")
    mp.spawn(_ddp_worker, args=(2, 29731, out), nprocs=2, join=True)
    worst = torch.load(out)
    # Adam normalises the step: parameters move by ~lr per iteration, a last-bit gradient difference moves them by << lr
    assert max(worst.values()) < 2e-4, worst


def _ddp_worker_mf(rank, world, port, tmp):
    import torch.distributed as dist
    from conftest import grid_of
    from nmf_b200 import train
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dev = torch.device("cuda", rank)
    fix = load_fixture("microfacet_g40")
    n = 256
    rays = fix["rays"][:n]
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    # no re-traced level: the top-k selection of re-traced rays is per forward call (as in the reference), so a sharded
    # batch selects per rank; with it off, a 2-GPU iteration is the same sum over rays as the 1-GPU iteration
    mk = lambda: train.MicrofacetTrainer(fix["state"], fix["aabb"], fix["near_far"], grid_of(fix), alpha_volume=fix["alpha_volume"],
                                         device=dev, seed=3, max_samples=-1, max_retrace_rays=(), mlp="fp32")
    lo, hi = rank * n // world, (rank + 1) * n // world
    tr = mk()
    # (1) the invariant: the all-reduced flat gradient of the sharded batch IS the gradient of the whole batch
    tr._calls = 0
    tr.accumulate(rays[lo:hi].to(dev), gt[lo:hi].to(dev), first=True, ray_id0=lo)
    tr.finish_into_bucket()
    tr.bucket.allreduce(scale=1.0)
    g_sharded = tr.bucket.flat.detach().cpu().clone()
    # (2) a few optimiser iterations run in lock-step on both ranks
    mse = []
    for it in range(4):
        tr._calls = it
        mse.append(tr.step(rays[lo:hi].to(dev), gt[lo:hi].to(dev), ray_id0=1000 * it + lo)["mse"])
        tr.check_schedule(it)
    flat = tr.flat_params.detach().clone()
    other = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(other, flat)
    same = all(torch.equal(o, flat) for o in other)                 # replicas stay bit-identical: same reduced gradient, same update
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        single = mk()
        single._calls = 0
        single.accumulate(rays.to(dev), gt.to(dev), first=True, ray_id0=0)
        single.finish_into_bucket()
        g_single = single.bucket.flat.detach().cpu()
        rel, off = {}, 0
        for k, p in single.params.items():
            a, b = g_sharded[off:off + p.numel()], g_single[off:off + p.numel()]
            off += p.numel()
            if float(b.abs().max()) > 0:
                rel[k] = float((a - b).norm() / b.norm())
        torch.save(dict(rel=rel, same=same, mse=mse), tmp)


def test_microfacet_ray_sharded_training_two_gpus(env, tmp_path):
    """BASELINE config #4 (ray-batch sharded microfacet training, ONE flat NCCL gradient all-reduce per iteration): the
    all-reduced gradient of two ranks on disjoint halves of the batch equals the single-GPU gradient of the whole batch
    (keyed random numbers: every ray draws the same jitter, bounce counts and directions wherever it is rendered; relative
    L2 per parameter < 5e-4: only the order of the fp32 atomic sums differs), and the replicas stay bit-identical over
    optimiser iterations.  (Parameters after Adam are NOT compared across world sizes: Adam moves an entry whose gradient
    is accumulation noise -- most texels of the environment map -- by +-lr whatever its magnitude.)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    out = str(tmp_path / "worst_mf.pt")
    mp.spawn(_ddp_worker_mf, args=(2, 29741, out), nprocs=2, join=True)
    res = torch.load(out)
    print(res)
    assert res["same"]
    assert len(res["rel"]) >= 20 and max(res["rel"].values()) < 5e-4, res["rel"]      # measured on 2 B200s: 6e-8 .. 1.8e-4
    assert all(m == m and m < 1.0 for m in res["mse"])


def test_upsample_schedule(env):
    """resolution schedule: nmf_upsample_bilinear == F.interpolate(align_corners=True); the plugin's check_schedule and the
    trainer's upsample keep rendering the same scene at the new resolution (fields/tensoRF.py:208-227, 408-413)"""
    from conftest import grid_of
    from nmf_b200 import config, ops, train
    g = torch.Generator().manual_seed(0)
    for shape, size in (((1, 16, 40, 40), (56, 56)), ((1, 24, 36, 48), (42, 77)), ((1, 16, 40, 1), (93, 1)), ((1, 24, 300, 300), (300, 300))):
        src = torch.randn(shape, generator=g)
        ref = torch.nn.functional.interpolate(src, size=size, mode="bilinear", align_corners=True)
        out = ops.upsample_bilinear(src.cuda(), size).cpu()
        assert float((out - ref).abs().max()) <= 4e-7 * float(src.abs().max())
    fix = load_fixture("plain_g64")
    n = 1024
    rays = fix["rays"][:n].cuda()
    tr = train.PlainTrainer(fix["state"], fix["aabb"], fix["near_far"], grid_of(fix), alpha_volume=fix["alpha_volume"], device=env)
    before = ops.render_rays(tr.scene, rays, fix["focal"], chunk=n)[0]["rgb_map"].clone()
    tr.upsample([96, 96, 96])
    assert tuple(tr.params["rf.app_rf.app_plane.0"].shape) == (1, 24, 96, 96) and tr.scene.n_steps > 219
    after = ops.render_rays(tr.scene, rays, fix["focal"], chunk=n)[0]["rgb_map"]
    assert float((after - before).abs().mean()) < 5e-2          # same scene, resampled factors and finer steps
    gt = 0.7 * before                                          # a target the scene has to move towards
    mse = [tr.step(rays, gt)["mse"] for _ in range(15)]        # the re-created optimiser keeps training
    assert np.isfinite(mse).all() and mse[-1] < 0.8 * mse[0], (mse[0], mse[-1])
    # plugin slot: TensorVMSplit.check_schedule fires on upsamp_list and returns True (train.py:806-809)
    t, _ = config.build_model(["model=tensorf", "field.grid_size=[64,64,64]", "field.upsamp_list=[5]", "field.N_voxel_final=884736"],
                              aabb=fix["aabb"], near_far=list(fix["near_far"]))
    t = t.cuda()
    assert t.rf.check_schedule(4, 1) is False and t.rf.check_schedule(5, 1) is True
    from nmf_b200.plugins import _n_to_reso
    res = _n_to_reso(t.rf.N_voxel_list[0], t.rf.aabb.cpu())
    assert 95 <= res[0] <= 96 and t.rf.grid_size.tolist() == res
    assert tuple(t.rf.app_rf.app_plane[0].shape) == (1, 24, res[1], res[0])
    assert tuple(t.rf.density_rf.app_line[0].shape) == (1, 16, res[2], 1)


# ---------------------------------------------------------------------------------------------------------------
# microfacet model: the training FORWARD through the fused render kernels (nmf_render_rays_train)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n,frac,min_rough", [("microfacet_g40", 384, None, 0.0), ("microfacet_g40", 384, 0.5, 0.2),
                                                   ("microfacet_noncubic", 256, None, 0.0),
                                                   ("microfacet_g56_ship", 256, 0.6, 0.05)])
def test_microfacet_train_forward_matches_oracle(env, name, n, frac, min_rough):
    """TensorNeRF.forward(is_train=True, draw_debug=False) of the microfacet model: jittered steps at recur 0 and in the
    re-traced rays, dynamic batch truncation at recur 0 only, min_rough, A19 statistics -- against the oracle's
    training forward (pinned to the unmodified reference by oracle/check_train.py) on the same keyed random numbers.
    whole_valid / kept counts / primary sample count: bit-exact; maps and statistics: tolerances of the eval tests."""
    from nmf_b200 import ops
    from oracle import keyed_rng as KR
    from oracle import nmf_oracle as O
    from test_gpu_parity import FLOAT_TOL, compare_images
    fix = load_fixture(name)
    osc, dsc = oracle_scene(fix, min_rough=min_rough), device_scene(fix, env)
    rays = fix["rays"][:n].contiguous()
    seed, id0 = 13, 2000
    keys = KR.primary_ray_keys(seed, np.arange(id0, id0 + n))
    _, valid, _, _, _ = O.sample_rays(osc, rays, fix["focal"], None, True, KR.KeyedRNG(), keys, -1)
    ms = -1 if frac is None else int(int(valid.sum()) * frac)
    ref, rst = O.render_chunk(osc, rays, fix["focal"], KR.KeyedRNG(), keys, draw_debug=False, is_train=True, max_samples=ms)
    ims, st = ops.render_rays_train(dsc, rays.cuda(), fix["focal"], seed=seed, ray_id0=id0, max_samples=ms,
                                    min_rough=min_rough)
    assert torch.equal(st["whole_valid"].cpu(), rst["whole_valid"])
    kept = int(rst["whole_valid"].sum())
    assert st["n_kept"] == kept and (frac is None or 0 < kept < n)
    assert st["n_samples"][0] == rst["n_samples"][0]
    m1 = rst["n_samples"][1] if len(rst["n_samples"]) > 1 else 0
    assert abs(st["n_samples"][1] - m1) <= max(8, 0.002 * m1), (st["n_samples"], rst["n_samples"])
    assert ims["rgb_map"].shape == (kept, 3) and ims["acc_map"].shape == (kept,)
    report, bad = compare_images(ims, ref, {k: FLOAT_TOL[k] for k in ("rgb_map", "acc_map")})
    print(report)
    assert not bad, bad
    for k in ("ori_loss", "diffuse_reg", "brdf_reg", "prediction_loss"):
        a, b = st["statistics"][k], float(rst[k])
        assert abs(a - b) <= 2e-3 * max(abs(b), 1e-3), (k, a, b)
    # the jitter is keyed: the same call again gives the same truncation and the same geometry, another seed does not
    ims2, st2 = ops.render_rays_train(dsc, rays.cuda(), fix["focal"], seed=seed, ray_id0=id0, max_samples=ms,
                                      min_rough=min_rough)
    assert torch.equal(ims2["acc_map"], ims["acc_map"]) and st2["n_samples"][0] == st["n_samples"][0]
    _, st3 = ops.render_rays_train(dsc, rays.cuda(), fix["focal"], seed=seed + 1, ray_id0=id0, max_samples=-1)
    assert st3["n_samples"][0] != st["n_samples"][0] or frac is not None


def test_microfacet_train_forward_plugin(env):
    """TensorNeRF.forward(is_train=True) of the plugin mirror: kept rows, statistics keys of tensor_nerf.py:567-649,
    and the eval forward of the same module is untouched by it."""
    from nmf_b200 import config
    fix = load_fixture("microfacet_g40")
    G = fix["grid_size"]
    t, _ = config.build_model([f"field.grid_size=[{G},{G},{G}]", "model.arch.bg_module.bg_resolution=32"],
                              aabb=fix["aabb"], near_far=list(fix["near_far"]))
    t.load_state_dict(fix["state"], strict=False)
    t = t.cuda().eval()
    t.sampler.update(t.rf, init=True)
    t.sampler.updateAlphaMask(t.rf, t.rf.grid_size)
    rays = fix["rays"][:256].cuda()
    ev, _ = t(rays, fix["focal"])
    ev = ev["rgb_map"].clone()
    t.sampler.max_samples = 4000
    ims, st = t(rays, fix["focal"], is_train=True, draw_debug=False)
    kept = int(st["whole_valid"].sum())
    assert 0 < kept < 256 and ims["rgb_map"].shape == (kept, 3)
    assert st["n_samples"][0] < 4000 and len(st["n_samples"]) == 2
    for k in ("ori_loss", "diffuse_reg", "brdf_reg", "prediction_loss", "envmap_reg"):
        assert np.isfinite(st[k])
    assert abs(st["prediction_loss"] - 2 * float(ims["acc_map"].sum())) < 1e-3 * max(1.0, st["prediction_loss"])
    t._calls = 0
    ev2, _ = t(rays, fix["focal"])
    assert torch.allclose(ev2["rgb_map"], ev, atol=1e-5)
    # train_step = forward + reverse pass on the device (nmf_train_microfacet): every trained parameter gets a gradient
    t.zero_grad()
    loss, ims2, st2 = t.train_step(rays, torch.rand(256, 3, generator=torch.Generator().manual_seed(0)).cuda(), focal=fix["focal"],
                                   lambda_pred=3e-4, lambda_ori=0.1)
    assert np.isfinite(loss) and loss > 0 and ims2["rgb_map"].shape[0] == int(st2["whole_valid"].sum())
    missing = [k for k, p in t.named_parameters() if (p.grad is None or float(p.grad.abs().max()) == 0.0)
               and "tint_mlp" not in k and "dbasis" not in k
               and not (("density_rf" in k) and k[-1] in "12")]      # this fixture's density lives in plane / line 0 only
    assert not missing, missing


# ---------------------------------------------------------------------------------------------------------------
# the optimiser loop of train.py:443-467, 497-813 on the device
# ---------------------------------------------------------------------------------------------------------------
def test_fused_adam_matches_torch_optim(env):
    """FusedAdam (nmf_grad_sq_norm + nmf_adam_step: loss normalisation, clip_grad_norm_, L2 weight decay, Adam, LambdaLR
    in one pass per parameter) against torch.optim.Adam + lr_scheduler.LambdaLR + clip_grad_norm_ on the same gradients;
    odd sizes exercise the unaligned tails."""
    from nmf_b200 import train
    from nmf_b200.distributed import FlatGradBucket
    g = torch.Generator().manual_seed(0)
    shapes = [(1, 16, 37, 41), (1, 16, 37, 1), (3,), (127, 5), (1000003,)]
    _fused_adam_case(shapes, g, flat=False)
    _fused_adam_case(shapes, g, flat=True)
    # without a flat buffer / clipping / schedule: plain Adam with torch's defaults
    a = torch.nn.Parameter(torch.randn(999, generator=g).cuda())
    b = torch.nn.Parameter(a.detach().clone())
    o1, o2 = train.FusedAdam([dict(params=[a], lr=1e-3)]), torch.optim.Adam([b], lr=1e-3, betas=(0.9, 0.99))
    for _ in range(3):
        a.grad = torch.randn(999, generator=g).cuda()
        b.grad = a.grad.clone()
        o1.step()
        o2.step()
    assert torch.allclose(a.detach(), b.detach(), rtol=2e-5, atol=2e-7)
    with pytest.raises(Exception):
        train.FusedAdam([dict(params=[torch.nn.Parameter(torch.zeros(3))], lr=1e-3)])     # CPU tensors: no fallback


def _fused_adam_case(shapes, g, flat):
    """flat=True: the parameters are views of one buffer (PlainTrainer's layout): one launch per optimiser group"""
    from nmf_b200 import train
    from nmf_b200.distributed import FlatGradBucket
    vals = [torch.randn(s, generator=g).cuda() for s in shapes]
    if flat:
        buf = torch.cat([v.reshape(-1) for v in vals])
        offs = np.cumsum([0] + [v.numel() for v in vals])
        mine = [torch.nn.Parameter(buf[offs[i]:offs[i + 1]].view(shapes[i])) for i in range(len(shapes))]
    else:
        mine = [torch.nn.Parameter(v) for v in vals]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    lam = lambda step: train.learning_rate_decay(step, max_steps=50, **train.REFERENCE_PARAMS)
    bucket = FlatGradBucket(mine)
    opt = train.FusedAdam([dict(params=mine[:2], lr=2e-2), dict(params=mine[2:], lr=1e-3)], betas=(0.9, 0.99), eps=1e-15,
                          weight_decay=1e-6, clip_grad=10.0, lr_lambda=lam, flat_grad=bucket.flat)
    topt = torch.optim.Adam([dict(params=ref[:2], lr=2e-2), dict(params=ref[2:], lr=1e-3)], betas=(0.9, 0.99), eps=1e-15,
                            weight_decay=1e-6)
    sched = torch.optim.lr_scheduler.LambdaLR(topt, lam)
    scale = 1.0 / 4096
    for step in range(6):
        mag = 3000.0 if step % 2 == 0 else 1.0            # clipped / not clipped
        for p, r in zip(mine, ref):
            gr = (torch.randn(p.shape, generator=g) * mag).cuda()
            p.grad.copy_(gr)
            r.grad = gr * scale
        torch.nn.utils.clip_grad_norm_(ref, 10.0)
        topt.step()
        sched.step()
        opt.step(grad_scale=scale)
        for p, r in zip(mine, ref):
            assert torch.allclose(p.detach(), r.detach(), rtol=2e-5, atol=2e-7), (step, p.shape, (p - r).abs().max())
    assert opt.n_launches() == (3 if flat else 1 + len(shapes))
    st = topt.state[ref[0]]
    assert torch.allclose(opt.state[id(mine[0])][0], st["exp_avg"], rtol=1e-4, atol=1e-6 * float(st["exp_avg"].abs().max()))


def test_l1_reg_kernel(env):
    from nmf_b200 import train
    g = torch.Generator().manual_seed(1)
    p = torch.randn(1, 16, 33, 29, generator=g).cuda()
    p.view(-1)[::7] = 0.0
    q = p.clone().requires_grad_(True)
    (8e-5 * q.abs().mean()).backward()
    grad = torch.zeros_like(p)
    total = torch.zeros(1, dtype=torch.float64, device="cuda")
    train.l1_reg(p, 8e-5, grad, total)
    train.l1_reg(p, 8e-5, grad, total)
    assert torch.allclose(grad, 2 * q.grad, rtol=1e-6, atol=1e-14)
    assert abs(float(total) - 2 * float(p.double().abs().sum())) < 1e-9 * float(total)


def test_fit_loop_matches_the_reference_loop_on_the_oracle(env):
    """PlainTrainer.fit (train.py:497-813: ray-id sampler, adaptive batch controller with gradient accumulation, density
    L1, 1/lbatch_size, clip_grad_norm_, Adam(weight_decay, eps) + LambdaLR) against the same loop written with the oracle's
    autograd and torch.optim on the CPU.  The controller's decisions (sub-batch sizes, kept rays, sample counts) are
    integer work and must be identical; parameters agree to the tolerance below (Adam's update is ~ lr * sign(g) in the
    first steps, so an element whose gradient is rounding noise around zero may move the other way: a quantile is
    bounded, not the maximum)."""
    from nmf_b200 import train
    from conftest import grid_of
    from test_hostmath import oracle_train_plain
    fix = load_fixture("plain_g64")
    hp = dict(starting_batch_size=48, min_batch_size=160, max_batch_size=256, target_num_samples=2500, batch_size=64,
              n_iters=4)
    n_iters, seed, ms = 4, 17, 3000
    allrays = fix["rays"][:1024].contiguous()
    allrgbs = torch.rand(allrays.shape[0], 3, generator=torch.Generator().manual_seed(2))
    tr = train.PlainTrainer(fix["state"], fix["aabb"], fix["near_far"], grid_of(fix), alpha_volume=fix["alpha_volume"],
                            device=env, max_samples=ms, seed=seed, params=hp)
    hist = tr.fit(allrays, allrgbs, n_iters=n_iters)
    assert any(h["sub_batches"] > 1 for h in hist) and all(h["kept_rays"] <= h["lbatch_size"] for h in hist)

    # ---- the same loop on the CPU: oracle forward + autograd, torch.optim.Adam, LambdaLR, clip_grad_norm_ ----
    H = dict(train.REFERENCE_PARAMS, **hp)
    state = {k: v.clone() for k, v in fix["state"].items()}
    params = {k: torch.nn.Parameter(state[k].float().clone()) for k in train.PLAIN_PARAM_KEYS}
    grid = [params[k] for k in train.PLAIN_PARAM_KEYS if k.startswith("rf.") and "basis" not in k]
    net = [params[k] for k in train.PLAIN_PARAM_KEYS if not (k.startswith("rf.") and "basis" not in k)]
    opt = torch.optim.Adam([dict(params=grid, lr=2e-2), dict(params=net, lr=1e-3)], betas=H["betas"], eps=H["eps"],
                           weight_decay=H["weight_decay"])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: train.learning_rate_decay(s, max_steps=H["n_iters"], **H))
    sampler = train.RayIdSampler(allrays.shape[0], H["batch_size"], env, seed=seed)
    num_rays, prev, calls = H["starting_batch_size"], None, 0
    for it in range(n_iters):
        opt.zero_grad(set_to_none=True)
        lbatch = min(H["min_batch_size"] if num_rays < H["min_batch_size"] else num_rays, H["max_batch_size"])
        remaining, kept, samples, subs = lbatch, 0, 0, 0
        while remaining > 0:
            ln = min(num_rays, remaining)
            remaining -= ln
            ids = sampler.nextids(ln).cpu()
            cur = dict(fix, state=dict(state, **{k: p.detach().clone() for k, p in params.items()}))
            ref = oracle_train_plain(cur, allrays[ids], allrgbs[ids], seed + calls, ids.numpy().astype(np.uint64), ms, 0.0)
            calls += 1
            for k, p in params.items():
                gk = ref["grads"][k].reshape(p.shape).clone()
                if ".density_rf." in k:
                    gk += H["L1_weight_initial"] * torch.sign(p.detach()) / p.numel()
                p.grad = gk / lbatch if p.grad is None else p.grad + gk / lbatch
            nk = int(ref["whole"].sum())
            kept, samples, subs = kept + nk, samples + ref["n_samples"], subs + 1
            ratio = nk / max(ref["n_samples"], 1)
            prev = ratio if prev is None else min(0.1 * ratio + 0.9 * prev, ratio)
            num_rays = int(prev * H["target_num_samples"] + 1)
        torch.nn.utils.clip_grad_norm_(list(params.values()), H["clip_grad"])
        opt.step()
        sched.step()
        h = hist[it]
        assert (h["lbatch_size"], h["sub_batches"], h["kept_rays"], h["n_samples"], h["next_num_rays"]) == \
            (lbatch, subs, kept, samples, num_rays), (it, h)
    n_moved = 0
    for k, p in params.items():
        d = (tr.params[k].detach().cpu() - p.detach()).abs().reshape(-1)
        n_moved += int(float((p.detach() - state[k].float()).abs().max()) > 0)   # all-zero factors with zero gradient stay put
        q = float(d.quantile(0.999)) if d.numel() < 10_000_000 else float(d.max())
        assert q <= 2e-4 and float(d.mean()) <= 2e-5, (k, q, float(d.mean()), float(d.max()))
    assert n_moved >= 12


def test_fit_schedule_events(env):
    """TensorNeRF.check_schedule inside the loop (modules/tensor_nerf.py:177-195, train.py:806-812): the occupancy update
    of `update_list` runs at the field's current resolution, the upsampling of `upsamp_list` re-creates the optimiser
    (Adam moments and the LambdaLR delay restart) and resets the batch controller; training continues afterwards."""
    from nmf_b200 import train
    from nmf_b200.plugins import _n_to_reso
    from conftest import grid_of
    fix = load_fixture("plain_g64")
    hp = dict(starting_batch_size=64, min_batch_size=128, max_batch_size=256, target_num_samples=4000, batch_size=128)
    allrays = fix["rays"][:2048].contiguous()
    tr0 = train.PlainTrainer(fix["state"], fix["aabb"], fix["near_far"], grid_of(fix), alpha_volume=fix["alpha_volume"], device=env)
    from nmf_b200 import ops
    gt = ops.render_rays(tr0.scene, allrays.cuda(), fix["focal"], chunk=2048, skip_eps=0.0, t_cut=0.0)[0]["rgb_map"].clone()
    tr = train.PlainTrainer(fix["state"], fix["aabb"], fix["near_far"], grid_of(fix), alpha_volume=fix["alpha_volume"],
                            device=env, max_samples=50000, seed=3, params=hp)
    n_vox = [80 ** 3]
    hist = tr.fit(allrays, gt, n_iters=8, upsamp_list=[3], n_voxel_list=n_vox, update_list=[2, 5])
    want = _n_to_reso(n_vox[0], torch.as_tensor(fix["aabb"]).float())
    assert [h.get("reinit", False) for h in hist] == [False, False, False, True, False, False, False, False]
    assert hist[3]["grid"] == grid_of(fix) and hist[4]["grid"] == want and tr.meta["grid_size"] == want
    assert hist[3]["next_num_rays"] != hp["starting_batch_size"] and hist[4]["lbatch_size"] == hp["min_batch_size"]
    assert tr.optimizer.t == 4 and abs(hist[4]["lr_factor"] - train.learning_rate_decay(1, max_steps=30000, **train.REFERENCE_PARAMS)) < 1e-9
    assert tr.params["rf.density_rf.app_plane.0"].shape[-2:] == (want[1], want[0])
    assert tuple(tr.alpha_volume.shape) == (want[2], want[1], want[0])          # rebuilt at iteration 5, new resolution
    assert all(np.isfinite(h["mse"]) for h in hist) and hist[-1]["mse"] < 1e-2


def test_microfacet_trainer_fits_a_teacher(env):
    """MicrofacetTrainer (the loop of train.py:497-813 for model=microfacet_tensorf2 on nmf_train_microfacet + FusedAdam):
    from perturbed appearance / material / BRDF weights, a few dozen iterations on rays rendered from the unperturbed
    scene bring the photometric loss down; the schedule state (detach_N, min_rough, re-trace budget) moves as the
    reference's does."""
    from nmf_b200 import ops, train
    fix = load_fixture("microfacet_g40")
    G = fix["grid_size"]
    teacher = device_scene(fix, env)
    rays = fix["rays"].cuda()
    gt = ops.render_rays(teacher, rays, fix["focal"], chunk=rays.shape[0], skip_eps=0.0, t_cut=0.0, seed=5)[0]["rgb_map"].clone()
    g = torch.Generator().manual_seed(2)
    st = {k: v.clone() for k, v in fix["state"].items()}
    for k in train.MICROFACET_PARAM_KEYS:
        if "app_rf" in k or "diffuse_module" in k or "brdf.mlp" in k:
            st[k] = st[k] + 0.3 * st[k].abs().mean() * torch.randn(st[k].shape, generator=g)
    tr = train.MicrofacetTrainer(st, fix["aabb"], fix["near_far"], [G] * 3, alpha_volume=fix["alpha_volume"], device=env,
                                 max_samples=-1, min_rough_start=0.2, seed=3)
    assert tr.detach_N and tr.min_rough == 0.2
    mse = []
    for it in range(40):
        out = tr.step(rays, gt)
        mse.append(out["mse"])
        tr.check_schedule(it)
    assert not tr.detach_N and tr.min_rough < 0.2                      # models/microfacet.py:112-121
    assert np.isfinite(mse).all() and np.mean(mse[-5:]) < 0.6 * np.mean(mse[:3]), (mse[:3], mse[-5:])
    assert tr.scene.hp["max_retrace_rays"][0] != 1000                   # the adaptive re-trace budget moved (microfacet.py:241-268)


def test_microfacet_step_repeats_an_overflowed_iteration(env):
    """MicrofacetTrainer.step has ONE host synchronisation, at its end: the loss normaliser and the overflow flag reach FusedAdam
    through device memory (NmfAdam.control).  With scratch lists that are far too small the first pass overflows, the update is a
    no-op on the device, and the repeated iteration (larger lists, same random numbers) produces the gradients and the step count
    of a trainer that never overflowed; both agree with the serial accumulate + apply path."""
    from nmf_b200 import ops, train
    fix = load_fixture("microfacet_g40")
    G = fix["grid_size"]
    rays = fix["rays"].cuda()
    gt = torch.rand(rays.shape[0], 3, generator=torch.Generator().manual_seed(4)).cuda()
    mk = lambda: train.MicrofacetTrainer({k: v.clone() for k, v in fix["state"].items()}, fix["aabb"], fix["near_far"], [G] * 3,
                                         alpha_volume=fix["alpha_volume"], device=env, max_samples=-1, seed=3)
    a, b, c = mk(), mk(), mk()
    n = rays.shape[0]
    b.buffers = ops.RenderBuffers(b.scene, n, n, ops.TRAIN_KEYS, cap_scale=0.02, train=True)
    oa, ob = a.step(rays, gt), b.step(rays, gt)
    assert b.buffers.cap_scale > 0.02 and a.optimizer.t == b.optimizer.t == 1 and a.iteration == b.iteration == 1
    assert oa["n_rays"] == ob["n_rays"] and oa["n_samples"] == ob["n_samples"] and abs(oa["mse"] - ob["mse"]) <= 1e-5 * oa["mse"]
    oc = c.accumulate(rays, gt, first=True)                      # the serial path: synchronises after the step
    c.apply(oc["n_rays"], oc["loss_photo"])
    rel = lambda x, y: float((x - y).norm() / (y.norm() + 1e-20))
    for k, q in a.params.items():              # gradients per parameter (fp32 atomics: only the summation order differs run to run)
        if float(q.grad.abs().max()) == 0.0:
            continue
        tol = 2e-2 if q.numel() == 1 else 1e-3          # the 0-dim environment scalars are sums of cancelling terms
        assert rel(b.params[k].grad, q.grad) < tol and rel(c.params[k].grad, q.grad) < tol, (k, rel(b.params[k].grad, q.grad), rel(c.params[k].grad, q.grad))
    for k in ("rf.app_rf.app_plane.0", "model.brdf.mlp.0.weight", "bg_module.mipbias"):
        assert bool(torch.isfinite(b.params[k]).all())
        assert float((a.params[k] - c.params[k]).abs().max()) <= 2.1 * a.optimizer.groups[0]["lr"]        # Adam: at most +-lr apart per step
    assert a.scene.c.env_dyn and abs(float(a.scene.keep["env_dyn"][0]) - float(a.params["bg_module.mipbias"])) < 1e-7
