"""Training slice, first model: ``model=tensorf`` (models/tensorf.py + MLPRender_Fea).

``train_plain`` runs ONE fused forward + backward on the device (``nmf_train_plain``: hand-written CUDA for every stage,
no autograd) and returns the loss terms and the gradient of every parameter under its reference state_dict key and in
the reference's layout, so that the reference's optimiser loop (train.py:497-813: Adam over ``get_optparam_groups``)
can consume it unchanged.  ``PlainTrainer`` is that loop for this model: ``fit`` mirrors train.py:497-813 (ray-id
sampler, adaptive batch controller with gradient accumulation, density L1, gradient clipping, Adam + LambdaLR decay,
the resolution / occupancy schedule with optimiser re-creation); the update itself is ``FusedAdam`` -- clip + weight
decay + Adam in one pass over each parameter (``nmf_adam_step``), no torch.optim -- followed by the in-place re-packing
of the updated factors, and -- when ``torch.distributed`` is initialised -- the single flat gradient all-reduce of
ray-sharded training (distributed.FlatGradBucket).

There is no CPU / PyTorch fallback: the step raises when libnmf_b200.so is missing or the tensors are not on a GPU.
"""
import ctypes as C
import math

import torch

from . import _lib
from .ops import _f32, _p, _stream

PLAIN_PARAM_KEYS = ([f"rf.density_rf.app_plane.{p}" for p in range(3)] + [f"rf.density_rf.app_line.{p}" for p in range(3)] +
                    [f"rf.app_rf.app_plane.{p}" for p in range(3)] + [f"rf.app_rf.app_line.{p}" for p in range(3)] +
                    ["rf.basis_mat.weight"] +
                    [f"model.diffuse_module.mlp.{i}.{w}" for i in (0, 2, 4) for w in ("weight", "bias")])


class PlainGradBuffers:
    """Channel-last gradient buffers of nmf_train_plain (struct NmfPlainGrads) for one scene geometry."""

    def __init__(self, scene):
        dev, s = scene.device, scene.c
        self.t = {}
        self.c = _lib.NmfPlainGrads()
        z = lambda *shape: torch.zeros(*shape, device=dev, dtype=torch.float32)
        for p in range(3):
            h, w, n = s.plane_h[p], s.plane_w[p], s.line_n[p]
            for name, t, arr in ((f"d_plane{p}", z(h, w, 16), self.c.d_plane), (f"d_line{p}", z(n, 16), self.c.d_line),
                                 (f"a_plane{p}", z(h, w, 24), self.c.a_plane), (f"a_line{p}", z(n, 24), self.c.a_line)):
                self.t[name] = t
                arr[p] = t.data_ptr()
        for name, shape in (("basis_t", (72, 24)), ("w0t", (135, 128)), ("b0", (128,)), ("w1t", (128, 128)), ("b1", (128,)),
                            ("w2t", (128, 3)), ("b2", (3,))):
            self.t[name] = z(*shape)
            setattr(self.c, name, self.t[name].data_ptr())

    def zero_(self):
        for t in self.t.values():
            t.zero_()

    def reference_views(self):
        """{reference state_dict key: gradient in the parameter's own shape}: planes (1,C,H,W), lines (1,C,N,1),
        basis_mat (24,72), mlp weights (out,in) -- strided VIEWS of the channel-last buffers (no copy: the trainer adds
        them straight into the flat gradient bucket, one pass)."""
        g = {}
        for p in range(3):
            g[f"rf.density_rf.app_plane.{p}"] = self.t[f"d_plane{p}"].permute(2, 0, 1)[None]
            g[f"rf.density_rf.app_line.{p}"] = self.t[f"d_line{p}"].t()[None, :, :, None]
            g[f"rf.app_rf.app_plane.{p}"] = self.t[f"a_plane{p}"].permute(2, 0, 1)[None]
            g[f"rf.app_rf.app_line.{p}"] = self.t[f"a_line{p}"].t()[None, :, :, None]
        g["rf.basis_mat.weight"] = self.t["basis_t"].t()
        for i, li in enumerate((0, 2, 4)):
            g[f"model.diffuse_module.mlp.{li}.weight"] = self.t[f"w{i}t"].t()
            g[f"model.diffuse_module.mlp.{li}.bias"] = self.t[f"b{i}"]
        return g

    def reference_layout(self):
        """reference_views(), materialised (contiguous copies that outlive the next step)."""
        return {k: v.contiguous().clone() if v.is_contiguous() else v.contiguous() for k, v in self.reference_views().items()}


class TrainBuffers:
    """Outputs + scratch of nmf_train_plain for batches of up to n_rays rays and cap_samples valid samples."""

    def __init__(self, scene, n_rays, cap_samples):
        dev = scene.device
        self.n_rays, self.cap_samples = int(n_rays), int(cap_samples)
        self.rgb_map = torch.zeros(n_rays, 3, device=dev)
        self.acc_map = torch.zeros(n_rays, device=dev)
        self.whole_valid = torch.zeros(n_rays, dtype=torch.uint8, device=dev)
        self.loss = torch.zeros(2, dtype=torch.float64, device=dev)
        self.n_kept = torch.zeros(2, dtype=torch.int32, device=dev)
        self.error = torch.zeros(1, dtype=torch.int32, device=dev)
        self.c = _lib.NmfTrainOut(rgb_map=self.rgb_map.data_ptr(), acc_map=self.acc_map.data_ptr(),
                                  whole_valid=self.whole_valid.data_ptr(), loss=self.loss.data_ptr(),
                                  n_kept=self.n_kept.data_ptr(), error=self.error.data_ptr())
        self.grads = PlainGradBuffers(scene)
        nbytes = _lib.lib().nmf_train_workspace_bytes(scene.ref(), self.n_rays, self.cap_samples)
        if nbytes == 0:
            raise _lib.NmfError("nmf_train_workspace_bytes: bad arguments")
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        self.ws_ptr = self.workspace.data_ptr() + (-self.workspace.data_ptr()) % 256
        self.ws_bytes = nbytes


def sample_rays_train(scene, rays, seed=0, ray_id0=0, ray_ids=None, max_samples=-1, override_near=None):
    """AlphaGridSampler.sample(is_train=True) (samplers/alphagrid.py:167-173, 353-364), for ALL rays of the batch:
    (ray_valid (B,S) bool, z_vals (B,S), n_valid (B) int32, whole_valid (B) bool, (rays kept, samples kept))."""
    r = _f32(rays[:, :6], scene.device)
    B, S = r.shape[0], scene.n_steps
    dev = r.device
    valid = torch.empty(B, S, dtype=torch.uint8, device=dev)
    z = torch.empty(B, S, device=dev)
    nv = torch.empty(B, dtype=torch.int32, device=dev)
    whole = torch.empty(B, dtype=torch.uint8, device=dev)
    kept = torch.zeros(2, dtype=torch.int32, device=dev)
    ids = None if ray_ids is None else ray_ids.to(device=dev, dtype=torch.int64).contiguous()
    _lib.check(_lib.lib().nmf_sample_rays_train(scene.ref(), _p(r), B, -1.0 if override_near is None else float(override_near),
                                                int(seed), int(ray_id0), _p(ids), int(max_samples), _p(valid), _p(z), _p(nv),
                                                _p(whole), _p(kept), _stream()), "nmf_sample_rays_train")
    return valid.bool(), z, nv, whole.bool(), kept


def train_plain(scene, rays, gt, focal=1.0, seed=0, ray_id0=0, ray_ids=None, max_samples=-1, lambda_pred=0.0,
                cap_samples=None, buffers=None, check_errors=True):
    """One fused training forward + backward of model=tensorf (nmf_train_plain).  rays (B,>=6), gt (B,3) on the GPU.
    Returns dict(loss_photo, sum_acc, n_rays, n_samples, whole_valid, rgb_map, acc_map, grads (channel-last buffers),
    buffers).  With check_errors the call synchronises and regrows the per-sample scratch when it overflowed."""
    if scene.hp["model"] != "plain":
        raise _lib.NmfError("train_plain: the training slice covers model=tensorf (a 'plain' DeviceScene)")
    r = _f32(rays[:, :6], scene.device)
    g = _f32(gt.reshape(-1, 3), scene.device)
    B = r.shape[0]
    if g.shape[0] != B:
        raise _lib.NmfError("train_plain: gt must hold one colour per ray")
    ids = None if ray_ids is None else ray_ids.to(device=r.device, dtype=torch.int64).contiguous()
    if cap_samples is None:
        cap_samples = int(max_samples) if max_samples > 0 else 64 * B
    while True:
        if buffers is None or buffers.n_rays < B or buffers.cap_samples < cap_samples:
            buffers = TrainBuffers(scene, B, cap_samples)
        tp = _lib.NmfTrain(n_rays=B, focal=float(focal), seed=int(seed), ray_id0=int(ray_id0),
                           ray_ids=None if ids is None else ids.data_ptr(), max_samples=int(max_samples),
                           cap_samples=buffers.cap_samples, lambda_pred=float(lambda_pred), white_bg=1)
        st = _lib.lib().nmf_train_plain(scene.ref(), C.byref(tp), _p(r), _p(g), C.byref(buffers.grads.c), C.byref(buffers.c),
                                        C.c_void_p(buffers.ws_ptr), buffers.ws_bytes, _stream())
        _lib.check(st, "nmf_train_plain")
        out = dict(buffers=buffers, grads=buffers.grads, rgb_map=buffers.rgb_map[:B], acc_map=buffers.acc_map[:B],
                   whole_valid=buffers.whole_valid[:B].bool(), loss=buffers.loss, n_kept=buffers.n_kept)
        if not check_errors:
            return out
        kept = buffers.n_kept.tolist()                        # synchronises
        if int(buffers.error.item()):
            if kept[1] <= buffers.cap_samples:
                raise _lib.NmfOverflow("nmf_train_plain: device error without an overflow")
            cap_samples = int(kept[1] * 1.25) + 1024          # the per-sample scratch was too small: grow, step again
            buffers = None
            continue
        loss = buffers.loss.tolist()
        out.update(loss_photo=loss[0], sum_acc=loss[1], n_rays=kept[0], n_samples=kept[1])
        return out


MICROFACET_PARAM_KEYS = ([f"rf.density_rf.app_plane.{p}" for p in range(3)] + [f"rf.density_rf.app_line.{p}" for p in range(3)] +
                         [f"rf.app_rf.app_plane.{p}" for p in range(3)] + [f"rf.app_rf.app_line.{p}" for p in range(3)] +
                         ["rf.basis_mat.weight"] +
                         [f"model.diffuse_module.{h}_mlp.0.{w}" for h in ("diffuse", "tint", "f0", "roughness") for w in ("weight", "bias")] +
                         [f"model.brdf.mlp.{i}.{w}" for i in (0, 2, 4) for w in ("weight", "bias")] +
                         ["bg_module.bg_mat", "bg_module.mipbias", "bg_module.brightness", "bg_module.mul"])
HEAD_ROWS = {"diffuse": slice(0, 3), "tint": slice(3, 6), "f0": slice(6, 9), "roughness": slice(9, 11)}


class MicrofacetGradBuffers:
    """Gradient buffers of nmf_train_microfacet (struct NmfMicrofacetGrads) for one scene geometry.  Every buffer
    ACCUMULATES over the sub-batches of an optimiser step; `finish` runs the two whole-image passes (environment map,
    normal stencil adjoint) once per step, `reference_views` hands the result out under the reference's state_dict keys."""

    def __init__(self, scene):
        from .scene import derivative_stencils
        dev, s = scene.device, scene.c
        self.scene = scene
        self.t = {}
        self.c = _lib.NmfMicrofacetGrads()
        self.eh, self.ew = int(s.env_h), int(s.env_w)
        # every buffer is a view of ONE flat allocation (16-byte aligned segments): zero_() is a single fill per step
        specs = []
        for p in range(3):
            h, w, n = s.plane_h[p], s.plane_w[p], s.line_n[p]
            specs += [(f"d_plane{p}", (h, w, 16)), (f"d_line{p}", (n, 16)), (f"a_plane{p}", (h, w, 24)), (f"a_line{p}", (n, 24)),
                      (f"gpack{p}", (h, w, 48)), (f"glpack{p}", (n, 4, 8))]
        specs += [("basis_t", (72, 24)), ("head_w", (11, 24)), ("head_b", (11,)), ("w0t", (66, 64)), ("b0", (64,)),
                  ("w1t", (64, 64)), ("b1", (64,)), ("w2t", (64, 4)), ("b2", (4,)), ("gsat", (self.eh * self.ew * 4 + 8,)),
                  ("d_mipbias", (1,)), ("d_bg", (3, self.eh, self.ew)), ("d_env_scalars", (2,))]
        numel = lambda shape: int(math.prod(shape))
        total = sum((numel(sh) + 3) // 4 * 4 for _, sh in specs)
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        for name, sh in specs:
            self.t[name] = self.flat[off:off + numel(sh)].view(*sh)
            off += (numel(sh) + 3) // 4 * 4
        arrays = dict(d_plane=self.c.d_plane, d_line=self.c.d_line, a_plane=self.c.a_plane, a_line=self.c.a_line,
                      gpack=self.c.normals.gpack, glpack=self.c.normals.glpack)
        for name, t in self.t.items():
            if name[:-1] in arrays and name[-1] in "012":
                arrays[name[:-1]][int(name[-1])] = t.data_ptr()
            elif name not in ("d_bg", "d_env_scalars"):
                setattr(self.c, name, t.data_ptr())
        kx, ky = derivative_stencils()
        self.kx = kx.reshape(25).to(device=dev, dtype=torch.float32).contiguous()
        self.ky = ky.reshape(25).to(device=dev, dtype=torch.float32).contiguous()
        self.finished = False

    def zero_(self):
        self.flat.zero_()
        self.finished = False

    def finish(self, bg_mat, brightness, mul, scalars_dev=None):
        """Once per optimiser step, after the last sub-batch: the adjoint of the environment's double cumsum + activation
        (nmf_env_lookup_bwd_finish -> d bg_mat, d brightness, d mul) and of the 5x5 smoothed-difference stencil
        (nmf_vm_normals_bwd_finish, added into the density-factor gradients)."""
        if self.finished:
            raise _lib.NmfError("MicrofacetGradBuffers.finish: already finished (zero_() starts the next step)")
        L, dev = _lib.lib(), self.scene.device
        bg = _f32(bg_mat.reshape(3, self.eh, self.ew), dev)
        with torch.cuda.device(dev):
            sc = self.t["d_env_scalars"]
            if scalars_dev is not None:       # [brightness, mul, ...] as fp32 device memory: no host copy of the parameters
                sd = scalars_dev.detach().to(device=dev, dtype=torch.float32).contiguous()
                _lib.check(L.nmf_env_lookup_bwd_finish_dev(_p(self.t["gsat"]), self.eh, self.ew, _p(bg), _p(sd), _p(self.t["d_bg"]),
                                                           _p(sc[0:1]), _p(sc[1:2]), _stream()), "nmf_env_lookup_bwd_finish_dev")
            else:
                br, mu = (float(v.detach()) if torch.is_tensor(v) else float(v) for v in (brightness, mul))
                _lib.check(L.nmf_env_lookup_bwd_finish(_p(self.t["gsat"]), self.eh, self.ew, _p(bg), br, mu, _p(self.t["d_bg"]),
                                                       _p(sc[0:1]), _p(sc[1:2]), _stream()), "nmf_env_lookup_bwd_finish")
            pp = (C.c_void_p * 3)(*[self.t[f"d_plane{p}"].data_ptr() for p in range(3)])
            lp = (C.c_void_p * 3)(*[self.t[f"d_line{p}"].data_ptr() for p in range(3)])
            _lib.check(L.nmf_vm_normals_bwd_finish(self.scene.ref(), C.byref(self.c.normals), _p(self.kx), _p(self.ky), pp, lp,
                                                   _stream()), "nmf_vm_normals_bwd_finish")
        self.finished = True

    def reference_views(self):
        """{reference state_dict key: gradient in the parameter's own shape} (strided views; call `finish` first)."""
        g = {}
        for p in range(3):
            g[f"rf.density_rf.app_plane.{p}"] = self.t[f"d_plane{p}"].permute(2, 0, 1)[None]
            g[f"rf.density_rf.app_line.{p}"] = self.t[f"d_line{p}"].t()[None, :, :, None]
            g[f"rf.app_rf.app_plane.{p}"] = self.t[f"a_plane{p}"].permute(2, 0, 1)[None]
            g[f"rf.app_rf.app_line.{p}"] = self.t[f"a_line{p}"].t()[None, :, :, None]
        g["rf.basis_mat.weight"] = self.t["basis_t"].t()
        for h, rows in HEAD_ROWS.items():
            g[f"model.diffuse_module.{h}_mlp.0.weight"] = self.t["head_w"][rows]
            g[f"model.diffuse_module.{h}_mlp.0.bias"] = self.t["head_b"][rows]
        for i, li in enumerate((0, 2, 4)):
            g[f"model.brdf.mlp.{li}.weight"] = self.t[f"w{i}t"].t()
            g[f"model.brdf.mlp.{li}.bias"] = self.t[f"b{i}"]
        g["bg_module.bg_mat"] = self.t["d_bg"][None]
        g["bg_module.mipbias"] = self.t["d_mipbias"][0]
        g["bg_module.brightness"] = self.t["d_env_scalars"][0]
        g["bg_module.mul"] = self.t["d_env_scalars"][1]
        return g

    def copy_into(self, grads_by_key):
        """Writes every gradient into `grads_by_key[key]` (contiguous tensors in the reference's parameter shapes, e.g. the
        views of the flat all-reduce bucket) with ONE nmf_transpose_batch launch; the job table is built once per set of
        destination pointers and lives on the device."""
        import numpy as np
        src = {}
        for p in range(3):
            src[f"rf.density_rf.app_plane.{p}"] = (self.t[f"d_plane{p}"], 16)
            src[f"rf.density_rf.app_line.{p}"] = (self.t[f"d_line{p}"], 16)
            src[f"rf.app_rf.app_plane.{p}"] = (self.t[f"a_plane{p}"], 24)
            src[f"rf.app_rf.app_line.{p}"] = (self.t[f"a_line{p}"], 24)
        src["rf.basis_mat.weight"] = (self.t["basis_t"], 24)
        for h, rows in HEAD_ROWS.items():
            src[f"model.diffuse_module.{h}_mlp.0.weight"] = (self.t["head_w"][rows], 1)
            src[f"model.diffuse_module.{h}_mlp.0.bias"] = (self.t["head_b"][rows], 1)
        for i, li in enumerate((0, 2, 4)):
            src[f"model.brdf.mlp.{li}.weight"] = (self.t[f"w{i}t"], self.t[f"w{i}t"].shape[1])
            src[f"model.brdf.mlp.{li}.bias"] = (self.t[f"b{i}"], 1)
        src["bg_module.bg_mat"] = (self.t["d_bg"], 1)
        src["bg_module.mipbias"] = (self.t["d_mipbias"][0:1], 1)
        src["bg_module.brightness"] = (self.t["d_env_scalars"][0:1], 1)
        src["bg_module.mul"] = (self.t["d_env_scalars"][1:2], 1)
        sig = tuple((k, grads_by_key[k].data_ptr()) for k in sorted(grads_by_key))
        if getattr(self, "_jobs_sig", None) != sig:
            jobs = (_lib.NmfTransposeJob * len(grads_by_key))()
            for j, (k, dst) in enumerate(sorted(grads_by_key.items())):
                t, c = src[k]
                assert dst.is_contiguous() and t.is_contiguous() and dst.numel() == t.numel() and dst.dtype == torch.float32, k
                jobs[j].src, jobs[j].dst, jobs[j].n, jobs[j].c = t.data_ptr(), dst.data_ptr(), t.numel() // c, c
            raw = torch.from_numpy(np.frombuffer(bytes(jobs), dtype=np.uint8).copy())
            self._jobs_dev, self._jobs_n, self._jobs_sig = raw.to(self.scene.device), len(grads_by_key), sig
        with torch.cuda.device(self.scene.device):
            _lib.check(_lib.lib().nmf_transpose_batch(_p(self._jobs_dev), self._jobs_n, 128, _stream()), "nmf_transpose_batch")


def train_microfacet(scene, rays, gt, focal=1.0, seed=0, ray_id0=0, max_samples=-1, min_rough=0.0, lambda_pred=3e-4,
                     lambda_ori=0.1, detach_N=True, grads=None, zero_grads=True, buffers=None, check_errors=True):
    """One fused training forward + backward of model=microfacet_tensorf2 on a ray batch (nmf_train_microfacet): the
    forward of TensorNeRF.forward(is_train=True), the loss of train.py:586-657 and the hand-written reverse pass of every
    stage.  rays (B,>=6), gt (B,3) on the GPU (gt row i belongs to ray i; the kept rays are a prefix).  Gradients accumulate
    into `grads` (MicrofacetGradBuffers; zeroed first when zero_grads) and still need `grads.finish(...)` once per
    optimiser step.  Returns dict(loss_photo, sum_acc, ori_loss, n_rays, n_samples, whole_valid, rgb_map, acc_map, grads,
    buffers, statistics)."""
    from . import ops
    if scene.hp["model"] != "microfacet":
        raise _lib.NmfError("train_microfacet: a 'microfacet' DeviceScene is required (model=tensorf trains through train_plain)")
    r = _f32(rays[:, :6], scene.device)
    g = _f32(gt.reshape(-1, 3), scene.device)
    B = r.shape[0]
    if B == 0 or g.shape[0] != B:
        raise _lib.NmfError("train_microfacet: rays (B,6) and gt (B,3) with B > 0 expected")
    if grads is None:
        grads = MicrofacetGradBuffers(scene)
    elif zero_grads:
        grads.zero_()
    # scratch is grow-only: the batch size follows the adaptive controller (train.py:616-626) and the re-trace budget
    # moves (models/microfacet.py:241-268), so the buffers are re-created only when they are too small, with headroom
    cap_scale = 2.0 if buffers is None else buffers.cap_scale
    need = _lib.lib().nmf_render_train_workspace_bytes(scene.ref(), B, cap_scale)
    if buffers is None or not getattr(buffers, "train", False) or buffers.n_rays < B or buffers.ws_bytes < need:
        nr = B if buffers is None else max(B, buffers.n_rays)
        buffers = ops.RenderBuffers(scene, nr, nr, ops.TRAIN_KEYS, cap_scale=cap_scale, train=True,
                                    min_ws_bytes=need if buffers is None else int(need * 1.25))
    while True:
        rp = _lib.NmfRender(n_rays=B, chunk=B, focal=float(focal), seed=int(seed), ray_id0=int(ray_id0), skip_eps=0.0, t_cut=0.0,
                            white_bg=1, cap_scale=buffers.cap_scale)
        tr = _lib.NmfRenderTrain(max_samples=int(max_samples), min_rough=float(min_rough),
                                 whole_valid=buffers.whole_valid.data_ptr(), n_kept=buffers.n_kept.data_ptr())
        tp = _lib.NmfMicrofacetTrain(lambda_pred=float(lambda_pred), lambda_ori=float(lambda_ori), detach_N=int(bool(detach_N)),
                                     loss=buffers.loss3.data_ptr())
        with torch.cuda.device(scene.device):
            st = _lib.lib().nmf_train_microfacet(scene.ref(), C.byref(rp), C.byref(tr), C.byref(tp), _p(r), _p(g), C.byref(grads.c),
                                                 C.byref(buffers.c_images), C.byref(buffers.c_counters),
                                                 C.c_void_p(buffers.ws_ptr), buffers.ws_bytes, _stream())
        _lib.check(st, "nmf_train_microfacet")
        out = dict(buffers=buffers, grads=grads, loss=buffers.loss3, n_kept=buffers.n_kept, n_batch=B)
        if not check_errors:
            return out                                         # the caller reads back later: train_microfacet_readback
        try:
            stats = ops.read_counters(buffers, B, B)           # synchronises
        except _lib.NmfOverflow:
            if buffers.cap_scale >= 32 or not zero_grads:
                raise                                          # (accumulated gradients cannot be rolled back)
            buffers = ops.RenderBuffers(scene, buffers.n_rays, buffers.n_rays, ops.TRAIN_KEYS, cap_scale=buffers.cap_scale * 2,
                                        train=True)
            grads.zero_()
            continue
        return _mf_fill_out(out, buffers, stats, B)


def _mf_fill_out(out, buffers, stats, B):
    from . import ops
    kept, m0 = ops.host_n_kept(buffers)
    loss = ops.host_loss3(buffers)
    out.update(loss_photo=loss[0], sum_acc=loss[1], ori_loss=loss[2], n_rays=kept, n_samples=stats["n_samples"][0],
               whole_valid=buffers.whole_valid[:B].bool(), rgb_map=buffers.images["rgb_map"][:kept],
               acc_map=buffers.images["acc_map"][:kept], statistics=stats["statistics"][0], counters=stats)
    return out


def train_microfacet_readback(out):
    """Second half of train_microfacet(check_errors=False): ONE synchronising device-to-host copy of the counters, loss sums
    and kept counts; raises NmfOverflow when a device-side list overflowed (the accumulated gradients are then incomplete:
    the caller zeroes them and repeats the call with larger buffers)."""
    from . import ops
    B = out["n_batch"]
    stats = ops.read_counters(out["buffers"], B, B)
    return _mf_fill_out(out, out["buffers"], stats, B)


# configs/model/tensorf.yaml:69-111 (`params:`), the values train.py reads for model=tensorf
REFERENCE_PARAMS = dict(L1_weight_initial=8e-5, clip_grad=10.0, weight_decay=1e-6, eps=1e-15, betas=(0.9, 0.99),
                        starting_batch_size=100, min_batch_size=4096, max_batch_size=32000, target_num_samples=400000,
                        n_iters=30000, batch_size=4096, lr_init=1.0, lr_final=1e-3, lr_delay_mult=0.1, lr_delay_steps=100)


def learning_rate_decay(step, lr_init=1.0, lr_final=1e-3, max_steps=30000, lr_delay_steps=0, lr_delay_mult=1.0, **_):
    """utils.learning_rate_decay / log_lerp (utils.py:318-359): the LambdaLR factor of train.py:458-466."""
    if lr_delay_steps > 0:
        delay = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
    else:
        delay = 1.0
    t = min(max(step / max_steps, 0.0), 1.0)
    return delay * math.exp(t * (math.log(lr_final) - math.log(lr_init)) + math.log(lr_init))


@torch.no_grad()
def calibrate_start(tensorf, args=None, start_density=5e-3, n_density=20000, n_model=100000, generator=None):
    """The initial calibration of a training run (train.py:403-437, `args.ckpt is None`, num_pretrain = 0):
      * `rf.calibrate`: density_shift += log(target_sigma) - log(mean density of 20000 random points), with
        target_sigma = -log(1 - start_density) / (stepsize * distance_scale)            (train.py:403-418)
      * `model.calibrate(args, xyz, feat, bg_brightness)` on 100000 random points       (train.py:428-437)
    Field queries and material / BRDF evaluations run on the device through the plugin slots; returns `args`."""
    dev = tensorf.get_device()
    rf = tensorf.rf
    rand = lambda *shape: torch.rand(*shape, device=dev, generator=generator)
    if getattr(rf, "calibrate", False):
        xyz = (rand(n_density, 3) * 2 - 1) * rf.aabb[1].reshape(1, 3)
        sigma_feat = rf.compute_densityfeature(xyz)
        target_sigma = -math.log(1 - start_density) / (float(tensorf.sampler.stepsize) * rf.distance_scale)
        rf.density_shift += math.log(target_sigma) - math.log(float(sigma_feat.mean()))
        if args is not None:
            args.field.density_shift = rf.density_shift
    tensorf.sampler.update(rf, init=True)
    xyz = rand(n_model, 4) * 2 - 1
    xyz[:, 3] *= 0
    feat = rf.compute_appfeature(xyz)
    bg_brightness = tensorf.bg_module.mean_color().detach().mean()
    return tensorf.model.calibrate(args, xyz, feat, bg_brightness)


class FusedAdam:
    """torch.optim.Adam + lr_scheduler.LambdaLR + clip_grad_norm_ as the reference composes them (train.py:443-467,
    752-755), as device updates: one nmf_grad_sq_norm over the flat gradient buffer, then one nmf_adam_step per parameter
    that applies loss normalisation, clip coefficient, L2 weight decay and the Adam update in a single pass (the clip
    coefficient is read from device memory: no host synchronisation).  `groups`: [{"params": [...], "lr": base_lr}]."""

    def __init__(self, groups, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0, clip_grad=None, lr_lambda=None, flat_grad=None):
        self.groups = [dict(params=list(g["params"]), lr=float(g["lr"]), betas=g.get("betas")) for g in groups]
        self.betas, self.eps, self.weight_decay = (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.clip_grad = float(clip_grad) if clip_grad is not None and clip_grad > 0 else 0.0
        self.lr_lambda, self.flat_grad = lr_lambda, flat_grad
        ps = [p for g in self.groups for p in g["params"]]
        if not ps:
            raise ValueError("FusedAdam: no parameters")
        for p in ps:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.NmfError("FusedAdam takes contiguous fp32 CUDA parameters (there is no CPU path)")
        self.device = ps[0].device
        # Parameters (and their gradients) that sit back to back in memory -- views of one flat buffer, as PlainTrainer
        # allocates them -- are updated by ONE launch per run: the step is launch-bound otherwise (19 small tensors).
        self.state = {}                                   # id(param) -> (exp_avg, exp_avg_sq) views
        for g in self.groups:
            runs = []
            for p in g["params"]:
                last = runs[-1][-1] if runs else None
                if (last is not None and p.grad is not None and last.grad is not None
                        and p.data_ptr() == last.data_ptr() + 4 * last.numel()
                        and p.grad.data_ptr() == last.grad.data_ptr() + 4 * last.numel()):
                    runs[-1].append(p)
                else:
                    runs.append([p])
            g["runs"] = []
            for run in runs:
                n = sum(q.numel() for q in run)
                m, v = torch.zeros(n, device=self.device), torch.zeros(n, device=self.device)
                off = 0
                for q in run:
                    self.state[id(q)] = (m[off:off + q.numel()].view_as(q), v[off:off + q.numel()].view_as(q))
                    off += q.numel()
                g["runs"].append((run, n, m, v))
        self.sq_norm = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.t = 0            # updates done by this instance = scheduler epoch (both restart when it is re-created)

    def lr_factor(self):
        return float(self.lr_lambda(self.t)) if self.lr_lambda is not None else 1.0

    def n_launches(self):
        return sum(len(g["runs"]) for g in self.groups) + (1 if self.clip_grad > 0 else 0)

    def step(self, grad_scale=1.0, control=None):
        """control: optional fp32 device tensor [grad_scale, skip] (NmfAdam.control): the loss normaliser is read from device
        memory and skip != 0 makes the whole update a no-op (the caller then also takes back the step count: `self.t -= 1`)."""
        L = _lib.lib()
        with torch.cuda.device(self.device):
            st = _stream()
            clip = self.clip_grad > 0
            if clip:
                self.sq_norm.zero_()
                if self.flat_grad is not None:
                    _lib.check(L.nmf_grad_sq_norm(_p(self.flat_grad), self.flat_grad.numel(), _p(self.sq_norm), st), "nmf_grad_sq_norm")
                else:
                    for g in self.groups:
                        for p in g["params"]:
                            if p.grad is not None:
                                _lib.check(L.nmf_grad_sq_norm(_p(p.grad), p.numel(), _p(self.sq_norm), st), "nmf_grad_sq_norm")
            lam = self.lr_factor()                      # LambdaLR: update k (1-based) runs at base_lr * lambda(k - 1)
            self.t += 1
            for g in self.groups:
                b1, b2 = self.betas if g.get("betas") is None else (float(g["betas"][0]), float(g["betas"][1]))
                a = _lib.NmfAdam(lr=g["lr"] * lam, beta1=b1, beta2=b2, eps=self.eps,
                                 weight_decay=self.weight_decay, step=self.t, grad_scale=float(grad_scale),
                                 max_norm=self.clip_grad, control=None if control is None else control.data_ptr())
                for run, n, m, v in g["runs"]:
                    if run[0].grad is None:
                        continue
                    _lib.check(L.nmf_adam_step(_p(run[0].data), _p(run[0].grad), _p(m), _p(v), n, C.byref(a),
                                               _p(self.sq_norm) if clip else None, st), "nmf_adam_step")


def l1_reg(param, weight, grad=None, sum_abs=None):
    """weight * mean|param| (one term of TensorVMSplit.density_L1, fields/tensoRF.py:332-340): accumulates the value's
    numerator into `sum_abs` (fp64 device scalar) and d/dparam into `grad`."""
    with torch.cuda.device(param.device):
        _lib.check(_lib.lib().nmf_l1_reg(_p(param), param.numel(), float(weight) / param.numel(), _p(grad), _p(sum_abs), _stream()),
                   "nmf_l1_reg")


class RayIdSampler:
    """train.SimpleSampler (train.py:36-51): epochs of a device-side random permutation of the ray ids."""

    def __init__(self, total, batch, device, seed=0):
        self.total, self.batch, self.curr, self.ids = total, batch, total, None
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)
        self.device = device

    def nextids(self, batch=None):
        batch = self.batch if batch is None else batch
        self.curr += batch
        if self.curr + batch > self.total or self.ids is None:
            self.ids = torch.randperm(self.total, dtype=torch.long, device=self.device, generator=self.gen)
            self.curr = 0
        return self.ids[self.curr:self.curr + batch]


class PlainTrainer:
    """The optimiser loop of train.py:497-813 for model=tensorf: parameters are kept in the reference's layout and under
    the reference's state_dict keys (a checkpoint loads / saves unchanged), every step is one nmf_train_plain call, the
    update is torch.optim.Adam (train.py:301-303: betas (0.9, 0.99)), and the device scene is re-packed from the updated
    parameters.  With torch.distributed initialised each rank trains on its own ray ids and the gradients are summed with
    ONE all-reduce over a flat bucket before the update."""

    def __init__(self, state, aabb, near_far, grid_size, alpha_volume=None, device="cuda", lr_grid=2e-2, lr_net=1e-3,
                 max_samples=-1, lambda_pred=0.0, seed=0, params=None, **hp):
        """params: the `params:` block of the model config (REFERENCE_PARAMS = configs/model/tensorf.yaml); None keeps
        a bare Adam (torch defaults, no L1 / clipping / decay), which is what the gradient tests want."""
        from .distributed import FlatGradBucket
        from .scene import DeviceScene
        self.device = torch.device(device)
        self.meta = dict(aabb=aabb, near_far=near_far, grid_size=grid_size, hp=dict(hp, model=self.MODEL))
        self.alpha_volume = alpha_volume
        self.state = {k: torch.as_tensor(v).detach().clone().to(self.device) for k, v in state.items()}
        self.lr_grid, self.lr_net = lr_grid, lr_net
        self.params = self._flat_params({k: self.state[k].float() for k in self.PARAM_KEYS})
        self._FlatGradBucket = FlatGradBucket
        self.max_samples, self.lambda_pred, self.seed = max_samples, lambda_pred, seed
        self.hparams = None if params is None else dict(self.DEFAULT_PARAMS, **params)
        self.l1_weight = 0.0 if params is None else float(self.hparams["L1_weight_initial"])
        self.l1_sum = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.iteration = 0
        self._calls = 0
        self.buffers = None
        self._DeviceScene = DeviceScene
        self._make_optimizer()
        self.repack()

    PARAM_KEYS = PLAIN_PARAM_KEYS
    MODEL = "plain"
    DEFAULT_PARAMS = REFERENCE_PARAMS

    @staticmethod
    def _is_grid(k):
        return k.startswith("rf.") and "basis" not in k

    def _group_defs(self):
        """[(keys, lr, betas or None)] in flat-buffer order: every optimiser group is one contiguous segment."""
        return [([k for k in self.PARAM_KEYS if self._is_grid(k)], self.lr_grid, None),
                ([k for k in self.PARAM_KEYS if not self._is_grid(k)], self.lr_net, None)]

    def _flat_params(self, tensors):
        """Every parameter is a view of ONE flat fp32 buffer, the factor (grid) group first, then the network group --
        the same order as the flat gradient bucket, so that an optimiser group is one contiguous segment."""
        order = [k for keys, _, _ in self._group_defs() for k in keys]
        flat = torch.empty(sum(tensors[k].numel() for k in order), dtype=torch.float32, device=self.device)
        out, off = {}, 0
        for k in order:
            t = tensors[k]
            view = flat[off:off + t.numel()].view(t.shape)
            view.copy_(t)
            out[k] = torch.nn.Parameter(view)
            off += t.numel()
        self.flat_params = flat
        return out

    def _make_optimizer(self):
        defs = self._group_defs()
        self.bucket = self._FlatGradBucket([self.params[k] for keys, _, _ in defs for k in keys])
        groups = [dict(params=[self.params[k] for k in keys], lr=lr, betas=betas) for keys, lr, betas in defs]
        if self.hparams is None:
            self.optimizer = FusedAdam(groups, betas=(0.9, 0.99), flat_grad=self.bucket.flat)
        else:
            h = self.hparams
            lam = lambda step: learning_rate_decay(step, max_steps=h["n_iters"], **h)      # train.py:458-466
            self.optimizer = FusedAdam(groups, betas=h["betas"], eps=h["eps"], weight_decay=h["weight_decay"],
                                       clip_grad=h["clip_grad"], lr_lambda=lam, flat_grad=self.bucket.flat)

    def upsample(self, grid_size, rebuild_occupancy=True):
        """Resolution schedule (fields/tensor_base.py:234-243, fields/tensoRF.py:208-227, 408-413; train.py:806-809): the
        factors are resampled on the device (nmf_upsample_bilinear) and the optimiser is re-created, as the reference does
        when check_schedule fires.  rebuild_occupancy=True also rebuilds the occupancy grid at the new resolution
        (samplers/alphagrid.py:249-276); the reference loop does that on its own schedule (`update_list`), so `fit`
        passes False and keeps the current volume."""
        from . import ops
        res = [int(g) for g in grid_size]
        mat, vec = [[0, 1], [0, 2], [1, 2]], [2, 1, 0]
        new = {}
        for k, p in self.params.items():
            if ".app_plane." in k:
                i = int(k[-1])
                new[k] = ops.upsample_bilinear(p.data, (res[mat[i][1]], res[mat[i][0]]))
            elif ".app_line." in k:
                new[k] = ops.upsample_bilinear(p.data, (res[vec[int(k[-1])]], 1))
            else:
                new[k] = p.data
        self.params = self._flat_params(new)
        self.meta["grid_size"] = res
        if rebuild_occupancy:
            self.alpha_volume = None
        self.buffers = None
        self._make_optimizer()
        self.repack()
        if rebuild_occupancy:
            self.alpha_volume = self.scene.update_alpha_mask(res)

    def repack(self, rebuild=True):
        """rebuild=True: a new DeviceScene (construction, resolution change); False: the updated parameters are packed
        into the existing buffers in place (every optimiser step)."""
        st = dict(self.state)
        st.update({k: p.detach() for k, p in self.params.items()})
        m = self.meta
        if rebuild or getattr(self, "scene", None) is None:
            self.scene = self._DeviceScene(st, m["aabb"], m["near_far"], m["grid_size"], alpha_volume=self.alpha_volume,
                                           device=self.device, **m["hp"])
        else:
            self._refresh_scene(st)

    def _refresh_scene(self, st):
        self.scene.refresh_plain(st)

    def _on_reinit(self):
        """train.py:806-813 after the optimiser was re-created (model.reset_counter for the microfacet model)."""

    def accumulate(self, rays, gt, ray_ids=None, first=True):
        """Forward + backward of one ray sub-batch (the body of the `while num_remaining > 0` loop, train.py:509-712):
        gradients are SUMS over rays, added into the flat bucket (first=True overwrites = optimizer.zero_grad), plus the
        density L1 term the reference adds to every sub-batch's loss (train.py:675-678)."""
        import torch.distributed as dist
        out = train_plain(self.scene, rays, gt, seed=self.seed + self._calls, ray_ids=ray_ids,
                          max_samples=self.max_samples, lambda_pred=self.lambda_pred, buffers=self.buffers)
        self._calls += 1
        self.buffers = out["buffers"]
        grads = out["grads"].reference_views()
        for k, p in self.params.items():                      # p.grad is a view into the flat bucket
            if first:
                p.grad.copy_(grads[k])
            else:
                p.grad.add_(grads[k])
        if self.l1_weight > 0:
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            self.l1_sum.zero_()
            for k, p in self.params.items():
                if ".density_rf." in k:                       # every rank adds its share: the all-reduce sums them
                    l1_reg(p.data, self.l1_weight / world, p.grad, self.l1_sum)
        return out

    def apply(self, n_rays_local, loss_local=0.0, normaliser=None):
        """clip_grad_norm_ + optimizer.step() + scheduler.step() (train.py:752-755) on the accumulated gradients, after
        the ONE flat all-reduce of ray-sharded training; the loss normaliser 1 / lbatch_size (train.py:709) is applied
        inside the fused update.  normaliser: this rank's lbatch_size (summed over the ranks); None: the global number of
        kept rays."""
        import torch.distributed as dist
        tot = torch.tensor([float(n_rays_local), float(loss_local), float(normaliser or 0.0)], device=self.device,
                           dtype=torch.float64)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(tot)                              # global ray count / global lbatch_size = the loss normaliser
        self.bucket.allreduce(scale=1.0)                      # one flat fp32 all-reduce (NCCL on GPUs)
        norm = float(tot[0]) if normaliser is None else float(tot[2])
        self.optimizer.step(grad_scale=1.0 / max(norm, 1.0))
        self.repack(rebuild=False)
        self.iteration += 1
        return float(tot[0]), float(tot[1])

    def step(self, rays, gt, ray_ids=None, **kw):
        """One iteration on this rank's rays: one sub-batch, normalised by the global number of kept rays."""
        out = self.accumulate(rays, gt, ray_ids=ray_ids, first=True, **kw)
        n, loss = self.apply(out["n_rays"], out["loss_photo"])
        out["mse"] = loss / max(3.0 * n, 1.0)
        return out

    def check_schedule(self, iteration, upsamp_list=(), n_voxel_list=(), update_list=()):
        """TensorNeRF.check_schedule (modules/tensor_nerf.py:177-195) for this model: the sampler's occupancy update
        first (samplers/alphagrid.py:91-94, at the field's CURRENT resolution), then the field's upsampling
        (fields/tensor_base.py:234-243).  True = the optimiser (and its LambdaLR) was re-created (train.py:806-809)."""
        from .plugins import _n_to_reso
        if iteration in update_list:
            self.alpha_volume = self.scene.update_alpha_mask(self.meta["grid_size"])
        if iteration in upsamp_list:
            aabb = torch.as_tensor(self.meta["aabb"]).float().cpu()
            self.upsample(_n_to_reso(n_voxel_list[list(upsamp_list).index(iteration)], aabb), rebuild_occupancy=False)
            return True
        return False

    def fit(self, allrays, allrgbs, n_iters=None, upsamp_list=(), n_voxel_list=(), update_list=(), callback=None):
        """The optimiser loop of train.py:497-813 for model=tensorf on device-resident rays (N,6) / colours (N,3|4):
        per iteration, sub-batches of `num_rays` rays are accumulated until `lbatch_size` rays were seen; `num_rays`
        follows the valid-sample count (train.py:616-626: target_num_samples per sub-batch); the gradients are
        normalised by lbatch_size, clipped and applied; schedule events re-create the optimiser and reset the controller.
        Needs `params` (REFERENCE_PARAMS).  Returns the per-iteration history."""
        if self.hparams is None:
            raise _lib.NmfError("PlainTrainer.fit needs the `params` block (train.REFERENCE_PARAMS)")
        h = self.hparams
        n_iters = h["n_iters"] if n_iters is None else n_iters
        allrays, allrgbs = allrays.to(self.device), allrgbs.to(self.device)
        import torch.distributed as dist
        on = dist.is_available() and dist.is_initialized()
        rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
        # ray-sharded training: every rank draws its own ray ids (seed + rank) and runs the same number of
        # iterations; the loss normaliser is the global lbatch_size
        sampler = RayIdSampler(allrays.shape[0], h["batch_size"], self.device, seed=self.seed + 1000003 * rank)
        num_rays, prev = h["starting_batch_size"], None
        history = []
        for it in range(n_iters):
            lbatch = min(h["min_batch_size"] if num_rays < h["min_batch_size"] else num_rays, h["max_batch_size"])
            remaining, first, kept, loss, samples, subs = lbatch, True, 0, 0.0, 0, 0
            while remaining > 0:
                ln = min(num_rays, remaining)
                remaining -= ln
                ids = sampler.nextids(ln)
                rgba = allrgbs[ids]
                if rgba.shape[-1] == 4:                      # train.py:527-531, bg_col = white
                    rgba = rgba[:, :3] * rgba[:, 3:] + (1 - rgba[:, 3:])
                out = self.accumulate(allrays[ids], rgba, ray_ids=ids, first=first)
                first = False
                kept, loss, samples, subs = kept + out["n_rays"], loss + out["loss_photo"], samples + out["n_samples"], subs + 1
                ratio = out["n_rays"] / max(out["n_samples"], 1)                       # train.py:616-626
                prev = ratio if prev is None else min(0.1 * ratio + 0.9 * prev, ratio)
                num_rays = int(prev * h["target_num_samples"] + 1)
            self.apply(kept, loss, normaliser=lbatch)          # summed over the ranks inside
            rec = dict(iteration=it, lbatch_size=lbatch, sub_batches=subs, kept_rays=kept, n_samples=samples,
                       mse=loss / max(3.0 * kept, 1.0), next_num_rays=num_rays, lr_factor=self.optimizer.lr_factor(),
                       grid=list(self.meta["grid_size"]))
            if self.check_schedule(it, upsamp_list, n_voxel_list, update_list):
                num_rays, prev = h["starting_batch_size"], None                         # train.py:810-812
                self._on_reinit()
                rec["reinit"] = True
            history.append(rec)
            if callback is not None:
                callback(rec)
        return history


def benchmark_plain(grid=300, n_rays=4096, steps=20, iters=60, device="cuda:0"):
    """Times nmf_train_plain (CUDA events, `steps` steps after 3 warm-ups) and a short PlainTrainer loop (`iters` Adam
    iterations, wall clock around a synchronised loop) on the synthetic lego scene at grid^3.  Used by bench.py
    (`train_step` entry of the JSON line) and tools/train_bench.py."""
    import time
    from . import ops, synthetic
    from .scene import DeviceScene
    dev = torch.device(device)
    state, meta = synthetic.make_scene("lego", grid_size=grid, bg_resolution=32)
    state.update(synthetic.plain_mlp_state(0))
    sc = DeviceScene(state, meta["aabb"], meta["near_far"], meta["grid_size"], device=dev, model="plain")
    vol = sc.update_alpha_mask()
    H = W = 800
    focal = synthetic.focal_for(W)
    pose = synthetic.hemisphere_poses(4)[1]
    pix = torch.randperm(H * W, generator=torch.Generator().manual_seed(0))[:n_rays]
    rays = synthetic.camera_rays(pose, H, W, focal)[pix].contiguous().to(dev)
    gt = ops.render_rays(sc, rays, focal, chunk=n_rays, skip_eps=0.0, t_cut=0.0)[0]["rgb_map"].clone()
    out = train_plain(sc, rays, gt, focal=focal, seed=1)
    bufs = out["buffers"]
    for _ in range(3):
        train_plain(sc, rays, gt, focal=focal, seed=1, buffers=bufs, check_errors=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(steps):
        train_plain(sc, rays, gt, focal=focal, seed=1, buffers=bufs, check_errors=False)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / max(steps, 1)
    res = dict(what="nmf_train_plain forward+backward, model=tensorf (SURVEY 8f row 1)", grid=grid, rays=n_rays,
               n_samples=out["n_samples"], ms_per_step=ms, rays_per_s=n_rays / ms * 1e3,
               samples_per_s=out["n_samples"] / ms * 1e3, gpu_launches_per_step=7)
    if iters > 0:
        g = torch.Generator().manual_seed(2)
        st2 = {k: v.clone() for k, v in state.items()}
        for k in PLAIN_PARAM_KEYS:
            if ("app_rf" in k) or ("mlp" in k):
                st2[k] = st2[k] + 0.05 * st2[k].abs().mean() * torch.randn(st2[k].shape, generator=g)
        tr = PlainTrainer(st2, meta["aabb"], meta["near_far"], meta["grid_size"], alpha_volume=vol, device=dev)
        mse = [tr.step(rays, gt)["mse"] for _ in range(3)]
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        mse += [tr.step(rays, gt)["mse"] for _ in range(iters - 3)]
        torch.cuda.synchronize(dev)
        res.update(loop_iters=iters, loop_ms_per_iter=(time.perf_counter() - t0) / max(iters - 3, 1) * 1e3,
                   loop_mse_first=mse[0], loop_mse_last=mse[-1])
        res["optimizer"] = benchmark_optimizer(tr, steps)
    return res


def benchmark_optimizer(tr, steps=20):
    """The update half of an iteration at the trainer's resolution, CUDA-event timed: FusedAdam with the reference's
    settings (clip_grad_norm_ + weight decay + Adam: one norm pass over the flat gradient + one pass per parameter)
    next to torch.optim.Adam + clip_grad_norm_ on copies of the same tensors.  HBM-bound: algorithmic bytes per
    parameter element = 4 (norm) + 16 read + 12 written (p, g, m, v -> p, m, v)."""
    dev = tr.device
    from .distributed import FlatGradBucket
    flat = tr.flat_params.clone()                # same layout as the trainer: views of one flat buffer, groups contiguous
    ps, off = [], 0
    for p in tr.params.values():
        ps.append(torch.nn.Parameter(flat[off:off + p.numel()].view(p.shape)))
        off += p.numel()
    bucket = FlatGradBucket(ps)
    bucket.flat.normal_(generator=None)
    h = REFERENCE_PARAMS
    opt = FusedAdam([dict(params=ps, lr=2e-2)], betas=h["betas"], eps=h["eps"], weight_decay=h["weight_decay"],
                    clip_grad=h["clip_grad"], flat_grad=bucket.flat)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    for r, p in zip(ref, ps):
        r.grad = p.grad.detach().clone()
    topt = torch.optim.Adam(ref, lr=2e-2, betas=h["betas"], eps=h["eps"], weight_decay=h["weight_decay"])

    def timed(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / max(steps, 1)

    def torch_step():
        torch.nn.utils.clip_grad_norm_(ref, h["clip_grad"])
        topt.step()

    n = bucket.flat.numel()
    ms = timed(lambda: opt.step(grad_scale=1.0 / 4096))
    ms_torch = timed(torch_step)
    return dict(n_params=n, fused_ms=ms, fused_launches=opt.n_launches(), torch_ms=ms_torch, algorithmic_bytes=32 * n,
                fused_GBps=32 * n / ms / 1e6)


def benchmark_microfacet_forward(grid=300, n_rays=4096, steps=20, device="cuda:0"):
    """Times nmf_render_rays_train (TensorNeRF.forward(is_train=True) of microfacet_tensorf2: jittered steps, dynamic
    batch truncation at max_samples = 200000, one retrace level) for one training batch of the synthetic lego scene."""
    from . import ops, synthetic
    from .scene import DeviceScene
    dev = torch.device(device)
    state, meta = synthetic.make_scene("lego", grid_size=grid)
    sc = DeviceScene(state, meta["aabb"], meta["near_far"], meta["grid_size"], device=dev, model="microfacet")
    sc.update_alpha_mask()
    H = W = 800
    focal = synthetic.focal_for(W)
    pose = synthetic.hemisphere_poses(4)[1]
    pix = torch.randperm(H * W, generator=torch.Generator().manual_seed(0))[:n_rays]
    rays = synthetic.camera_rays(pose, H, W, focal)[pix].contiguous().to(dev)
    ims, st = ops.render_rays_train(sc, rays, focal, seed=1, max_samples=200000)
    bufs = st["buffers"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(steps):
        ops.render_rays_train(sc, rays, focal, seed=2 + i, max_samples=200000, buffers=bufs)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / max(steps, 1)
    return dict(what="nmf_render_rays_train: microfacet training forward (SURVEY 8f row 1, forward half)", grid=grid, rays=n_rays,
                kept_rays=st["n_kept"], n_samples=st["n_samples"], n_retrace=st["n_retrace"][0],
                n_bounce_rays=[st["n_bounce_rays0"][0], st["n_bounce_rays1"][0]], ms_per_step=ms,
                kept_rays_per_s=st["n_kept"] / ms * 1e3, note="includes the counter read-back (one host sync per step)")


# ------------------------------------------------------------------------------------------------------------
# model=microfacet_tensorf2 (BASELINE configs #3 / #4)
# ------------------------------------------------------------------------------------------------------------
# configs/model/microfacet_tensorf2.yaml:192-240 (`params:`), the values train.py reads for this model
MICROFACET_REFERENCE_PARAMS = dict(L1_weight_initial=8e-5, clip_grad=None, weight_decay=0.0, eps=1e-8, betas=(0.9, 0.99),
                                   starting_batch_size=100, min_batch_size=4096, max_batch_size=8000,
                                   target_num_samples=200000, n_iters=30000, batch_size=4096, lr_init=1.0, lr_final=1e-3,
                                   lr_delay_mult=0.1, lr_delay_steps=100, pred_lambda=3e-4, ori_lambda=0.1)


class MicrofacetTrainer(PlainTrainer):
    """The optimiser loop of train.py:497-813 for model=microfacet_tensorf2: every sub-batch is one nmf_train_microfacet
    call (forward + loss + reverse pass on the device, no autograd), the gradients of all sub-batches of an iteration
    accumulate in MicrofacetGradBuffers, are finished once (environment-map and stencil adjoints), all-reduced as ONE flat
    bucket when torch.distributed is initialised and applied by FusedAdam with the reference's
    optimiser groups (fields/tensoRF.py:298-313, models/microfacet.py get_optparam_groups, modules/integral_equirect.py
    get_optparam_groups); the model's own schedule (min_rough decay, detach_N, the adaptive re-trace budget,
    models/microfacet.py:112-121, 241-268) runs between iterations.  fixed_bg=True is train.py:267-284 (config #3 with
    backgrounds/forest.th): the map, its brightness and its scale are not trained (lr 0), mipbias still is."""
    PARAM_KEYS = MICROFACET_PARAM_KEYS
    MODEL = "microfacet"
    DEFAULT_PARAMS = MICROFACET_REFERENCE_PARAMS

    def __init__(self, state, aabb, near_far, grid_size, alpha_volume=None, device="cuda", lr_grid=2e-2, lr_net=1e-3,
                 lr_heads=1e-3, lr_brdf=1e-3, lr_bg=0.02, lr_mipbias=1e-4, lr_brightness=0.0, lr_mul=0.0, bg_betas=(0.9, 0.99),
                 mul_betas=(0.9, 0.9), fixed_bg=False, max_samples=200000, lambda_pred=3e-4, lambda_ori=0.1, seed=0, params=None,
                 min_rough_start=0.0, min_rough_decay=0.999, detach_N_iters=0, target_num_samples=(1000000,), **hp):
        self.lr_heads, self.lr_brdf = lr_heads, lr_brdf
        self.lr_bg, self.lr_mipbias, self.lr_brightness, self.lr_mul = lr_bg, lr_mipbias, lr_brightness, lr_mul
        if fixed_bg:
            self.lr_bg = self.lr_brightness = self.lr_mul = 0.0
        self.bg_betas, self.mul_betas = tuple(bg_betas), tuple(mul_betas)
        self.lambda_ori = lambda_ori
        self.min_rough, self.min_rough_decay = float(min_rough_start), float(min_rough_decay)
        self.detach_N, self.detach_N_iters = True, int(detach_N_iters)
        self.target_num_samples = list(target_num_samples)
        self.grads, self.ratio_list, self._subs = None, None, 0
        state = dict(state)
        for k, d in (("bg_module.mipbias", 1.0), ("bg_module.brightness", 0.0), ("bg_module.mul", 1.0)):
            state[k] = torch.as_tensor(state.get(k, d), dtype=torch.float32)
        super().__init__(state, aabb, near_far, grid_size, alpha_volume=alpha_volume, device=device, lr_grid=lr_grid, lr_net=lr_net,
                         max_samples=max_samples, lambda_pred=lambda_pred, seed=seed, params=params, **hp)
        self.start_max_retrace = list(self.scene.hp["max_retrace_rays"])

    def _group_defs(self):
        K = self.PARAM_KEYS
        sel = lambda f: [k for k in K if f(k)]
        return [(sel(self._is_grid), self.lr_grid, None),
                (["rf.basis_mat.weight"], self.lr_net, (0.9, 0.99)),
                (sel(lambda k: k.startswith("model.diffuse_module.")), self.lr_heads, None),
                (sel(lambda k: k.startswith("model.brdf.")), self.lr_brdf, None),
                (["bg_module.bg_mat"], self.lr_bg, self.bg_betas),
                (["bg_module.mipbias"], self.lr_mipbias, None),
                (["bg_module.brightness"], self.lr_brightness, None),
                (["bg_module.mul"], self.lr_mul, self.mul_betas)]

    ENV_SCALARS = ("bg_module.brightness", "bg_module.mul", "bg_module.mipbias")

    def _env_host(self, refresh=False):
        """(brightness, mul, mipbias) as python floats: they are kernel arguments by value, so each change needs them on the
        host -- ONE device-to-host copy per optimiser step instead of one per use."""
        if refresh or getattr(self, "_env_host_cache", None) is None:
            v = torch.stack([self.params[k].detach().reshape(()).float() for k in self.ENV_SCALARS]).tolist()
            self._env_host_cache = dict(zip(self.ENV_SCALARS, v))
        return self._env_host_cache

    def repack(self, rebuild=True):
        if rebuild:
            self._env_host_cache = None
        super().repack(rebuild)

    def _env_dev(self):
        """[brightness, mul, mipbias] as one fp32 device tensor (one small gather launch, no synchronisation)."""
        return torch.stack([self.params[k].detach().reshape(()).float() for k in self.ENV_SCALARS])

    def _refresh_scene(self, st):
        self._env_host_cache = None
        self.scene.refresh_microfacet(st, dev_scalars=self._env_dev())

    def _on_reinit(self):
        self.scene.update_hyper(max_retrace_rays=tuple(self.start_max_retrace))      # Microfacet.reset_counter
        self.ratio_list = None

    def upsample(self, grid_size, rebuild_occupancy=True):
        super().upsample(grid_size, rebuild_occupancy)
        self.grads = None

    def update_n_samples(self, n_samples1):
        """Microfacet.update_n_samples (models/microfacet.py:241-268, train.py:627): the re-trace budget follows
        target_num_samples * min(recent re-traced rays per secondary sample), capped by max_brdf_rays[0]."""
        hp = self.scene.hp
        cur = list(hp["max_retrace_rays"])
        if len(cur) != 1:
            return
        ratio = (cur[0] / n_samples1) if n_samples1 > 0 else 1e-3
        self.ratio_list = [ratio, 1e-3] if self.ratio_list is None else ([ratio] + self.ratio_list)[:20]
        new = min(int(self.target_num_samples[0] * min(self.ratio_list) + 1), int(hp["max_brdf_rays"][0]))
        if new != cur[0]:
            self.scene.update_hyper(max_retrace_rays=(new,))

    def accumulate(self, rays, gt, ray_ids=None, first=True, ray_id0=None):
        """One sub-batch (train.py:509-712): gradients accumulate on the device (MicrofacetGradBuffers).  ray_id0: global id
        of rays[0] for the keyed jitter (default: a fresh id range per call); ray-sharded ranks that pass the global offset
        of their slice draw the numbers a single-GPU pass over the whole batch would."""
        if self.grads is None:
            self.grads = MicrofacetGradBuffers(self.scene)
        self.grads.scene = self.scene
        if first:
            self.grads.zero_()
            self._subs = 0
        id0 = ((self.seed * 7919 + self._calls) << 20) if ray_id0 is None else int(ray_id0)
        out = train_microfacet(self.scene, rays, gt, seed=self.seed + self._calls, ray_id0=id0, max_samples=self.max_samples,
                               min_rough=self.min_rough, lambda_pred=self.lambda_pred, lambda_ori=self.lambda_ori,
                               detach_N=self.detach_N, grads=self.grads, zero_grads=False, buffers=self.buffers)
        self._calls += 1
        self._subs += 1
        self.buffers = out["buffers"]
        ns = out["n_samples"]
        out["n_samples_all"] = list(ns)
        out["n_samples"] = ns[0]
        if len(ns) > 1 and self.scene.c.max_retrace > 0:
            self.update_n_samples(ns[1])
        return out

    def apply(self, n_rays_local, loss_local=0.0, normaliser=None):
        self.finish_into_bucket()
        return super().apply(n_rays_local, loss_local, normaliser)

    def step(self, rays, gt, ray_ids=None, ray_id0=None, **kw):
        """One iteration with ONE sub-batch and ONE host synchronisation, at its end: the step, the finishing passes, the
        gradient hand-over, the all-reduce, FusedAdam and the scene re-pack are queued back to back, so the host's launch work
        overlaps the device.  What the host used to wait for stays on the device: the loss normaliser (global number of kept
        rays) and the overflow flag go to the optimiser through NmfAdam.control, the environment scalars through
        NmfScene.env_dyn.  If a device-side list overflowed (on any rank) the update was a no-op; the iteration is repeated
        with larger buffers."""
        import torch.distributed as dist
        from . import ops
        on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.grads is None:
            self.grads = MicrofacetGradBuffers(self.scene)
        self.grads.scene = self.scene
        id0 = ((self.seed * 7919 + self._calls) << 20) if ray_id0 is None else int(ray_id0)
        while True:
            self.grads.zero_()
            self._subs = 1
            out = train_microfacet(self.scene, rays, gt, seed=self.seed + self._calls, ray_id0=id0, max_samples=self.max_samples,
                                   min_rough=self.min_rough, lambda_pred=self.lambda_pred, lambda_ori=self.lambda_ori,
                                   detach_N=self.detach_N, grads=self.grads, zero_grads=False, buffers=self.buffers,
                                   check_errors=False, **kw)
            buffers = self.buffers = out["buffers"]
            self.finish_into_bucket()
            tot = torch.stack([buffers.n_kept[0].double(), buffers.loss3[0], buffers.counters["error"][0].double()])
            if on:
                dist.all_reduce(tot)                          # global kept rays, loss, overflow flag
            self.bucket.allreduce(scale=1.0)                  # one flat fp32 all-reduce (NCCL on GPUs)
            control = torch.stack([1.0 / tot[0].clamp(min=1.0), (tot[2] != 0).double()]).float()
            self.optimizer.step(grad_scale=1.0, control=control)
            self.repack(rebuild=False)
            overflow = False
            try:
                train_microfacet_readback(out)                # the iteration's one synchronisation
            except _lib.NmfOverflow:
                overflow = True
                if buffers.cap_scale >= 32:
                    raise
                self.buffers = ops.RenderBuffers(self.scene, buffers.n_rays, buffers.n_rays, ops.TRAIN_KEYS,
                                                 cap_scale=buffers.cap_scale * 2, train=True)
            n_glob, loss_glob, err_glob = tot.tolist()
            if overflow or err_glob != 0:                     # the update was skipped on every rank: take the step count back
                self.optimizer.t -= 1
                continue
            break
        self._calls += 1
        self.iteration += 1
        ns = out["n_samples"]
        out["n_samples_all"] = list(ns)
        out["n_samples"] = ns[0]
        if len(ns) > 1 and self.scene.c.max_retrace > 0:
            self.update_n_samples(ns[1])
        out["mse"] = loss_glob / max(3.0 * n_glob, 1.0)
        return out

    def finish_into_bucket(self):
        """After the last sub-batch of an iteration: the two whole-image finishing passes, then this rank's gradient of every
        parameter (plus the density L1 term) is written into the flat bucket the all-reduce and FusedAdam work on."""
        import torch.distributed as dist
        p = self.params
        self.grads.finish(p["bg_module.bg_mat"].data, None, None, scalars_dev=self._env_dev())
        self.grads.copy_into({k: q.grad for k, q in p.items()})
        if self.l1_weight > 0:            # train.py:675-678 adds the density L1 term to EVERY sub-batch's loss
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            self.l1_sum.zero_()
            for k, q in p.items():
                if ".density_rf." in k:
                    l1_reg(q.data, self.l1_weight * self._subs / world, q.grad, self.l1_sum)

    def check_schedule(self, iteration, upsamp_list=(), n_voxel_list=(), update_list=()):
        """TensorNeRF.check_schedule (modules/tensor_nerf.py:177-195): the model's schedule first
        (models/microfacet.py:112-121), then the sampler's occupancy update and the field's upsampling."""
        if iteration % 10 == 0:
            self.min_rough *= self.min_rough_decay
        if iteration > self.detach_N_iters:
            self.detach_N = False
        return super().check_schedule(iteration, upsamp_list, n_voxel_list, update_list)


def benchmark_microfacet_train(grid=300, n_rays=4096, steps=20, device="cuda:0", max_retrace=1000, detach_N=False, mlp="f16",
                               env="synthetic"):
    """Times nmf_train_microfacet (forward + loss + reverse pass of microfacet_tensorf2, BASELINE config #3's shape:
    G = 300, 4096-ray batches truncated at max_samples = 200000, one re-traced level) with CUDA events, phase by phase
    from the kernel launch list when `profile`.  Synthetic scene; gt = the eval render of the same scene."""
    from . import ops, synthetic
    from .scene import DeviceScene
    dev = torch.device(device)
    state, meta = synthetic.make_scene("materials", grid_size=grid)
    sc = DeviceScene(state, meta["aabb"], meta["near_far"], meta["grid_size"], device=dev, model="microfacet", mlp=mlp,
                     max_retrace_rays=(max_retrace,))
    sc.update_alpha_mask()
    H = W = 800
    focal = synthetic.focal_for(W)
    pose = synthetic.hemisphere_poses(4)[1]
    pix = torch.randperm(H * W, generator=torch.Generator().manual_seed(0))[:n_rays]
    rays = synthetic.camera_rays(pose, H, W, focal)[pix].contiguous().to(dev)
    gt = ops.render_rays(sc, rays, focal, chunk=n_rays, skip_eps=0.0, t_cut=0.0)[0]["rgb_map"].clone()
    out = train_microfacet(sc, rays, gt, focal=focal, seed=1, max_samples=200000, detach_N=detach_N)
    bufs, grads = out["buffers"], out["grads"]
    for i in range(3):
        train_microfacet(sc, rays, gt, focal=focal, seed=2 + i, max_samples=200000, detach_N=detach_N, grads=grads, buffers=bufs,
                         check_errors=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(steps):
        train_microfacet(sc, rays, gt, focal=focal, seed=5 + i, max_samples=200000, detach_N=detach_N, grads=grads, buffers=bufs,
                         check_errors=False)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / max(steps, 1)
    fwd = benchmark_microfacet_forward(grid, n_rays, steps, device) if max_retrace == 1000 else None
    c = out["counters"]
    return dict(what="nmf_train_microfacet: forward + loss + reverse pass, microfacet_tensorf2 (SURVEY 8f row 1, config #3 shape)",
                grid=grid, rays=n_rays, kept_rays=out["n_rays"], n_samples=out["n_samples"], max_retrace=max_retrace,
                n_retrace=c["n_retrace"][0], n_bounce_rays=[c["n_bounce_rays0"][0], c["n_bounce_rays1"][0]], detach_N=bool(detach_N),
                mlp=mlp, ms_per_step=ms, kept_rays_per_s=out["n_rays"] / ms * 1e3,
                forward_only_ms=None if fwd is None else fwd["ms_per_step"])


def benchmark_sharded_train(grid=300, n_rays=4096, steps=10, device="cuda:0", scene_name="ship"):
    """BASELINE config #4 (ship 800x800, ray-batch sharded, NCCL gradient all-reduce): every rank runs one
    MicrofacetTrainer iteration on ITS OWN 4096 rays (weak scaling of the batch: the global batch is world x 4096) --
    nmf_train_microfacet, finish, ONE flat fp32 all-reduce, FusedAdam, re-pack -- and the phases are timed with CUDA
    events (max over ranks; a 4-byte collective in front of the gradient all-reduce takes the arrival skew of the ranks, which
    render different views, into the forward + backward interval).  Needs torch.distributed initialised (NCCL); world = 1 works too (no collective)."""
    import torch.distributed as dist
    from . import ops, synthetic
    dev = torch.device(device)
    on = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    state, meta = synthetic.make_scene(scene_name, grid_size=grid)
    tr = MicrofacetTrainer(state, meta["aabb"], meta["near_far"], meta["grid_size"], device=dev, max_samples=200000, seed=7,
                           params=dict(MICROFACET_REFERENCE_PARAMS))
    tr.alpha_volume = tr.scene.update_alpha_mask()
    H = W = 800
    focal = synthetic.focal_for(W)
    pose = synthetic.hemisphere_poses(8)[1 + rank % 6]
    pix = torch.randperm(H * W, generator=torch.Generator().manual_seed(rank))[:n_rays]
    rays = synthetic.camera_rays(pose, H, W, focal)[pix].contiguous().to(dev)
    gt = ops.render_rays(tr.scene, rays, focal, chunk=n_rays, skip_eps=0.0, t_cut=0.0)[0]["rgb_map"].clone()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    acc = dict(step=0.0, allreduce=0.0, update=0.0, total=0.0)
    kept = 0
    skew = torch.zeros(1, device=dev)
    for it in range(steps + 3):
        e = [ev() for _ in range(5)]
        if on:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e[0].record()
        out = tr.accumulate(rays, gt, first=True)
        tr.finish_into_bucket()
        if on and world > 1:
            dist.all_reduce(skew)         # 4-byte collective: absorbs the ranks' arrival skew (they render different views),
        e[1].record()                     # so that the next interval is the gradient all-reduce itself, not the wait for the slowest rank
        tr.bucket.allreduce(scale=1.0)
        e[2].record()
        tr.optimizer.step(grad_scale=1.0 / (world * n_rays))
        tr.repack(rebuild=False)
        e[3].record()
        torch.cuda.synchronize(dev)
        if it >= 3:
            acc["step"] += e[0].elapsed_time(e[1]); acc["allreduce"] += e[1].elapsed_time(e[2])
            acc["update"] += e[2].elapsed_time(e[3]); acc["total"] += e[0].elapsed_time(e[3])
            kept += out["n_rays"]
    t = torch.tensor([acc[k] / steps for k in ("step", "allreduce", "update", "total")] + [float(kept) / steps], device=dev, dtype=torch.float64)
    if on and world > 1:
        tm = t.clone()
        dist.all_reduce(tm[:4], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[4:], op=dist.ReduceOp.SUM)
        t[:4] = tm[:4]
    nbytes = tr.bucket.flat.numel() * 4
    ar_ms = float(t[1])
    return dict(what="ray-sharded training iteration of microfacet_tensorf2 (config #4 shape): nmf_train_microfacet per rank + one "
                     "flat fp32 all-reduce + FusedAdam + re-pack; CUDA events, max over ranks",
                world=world, grid=grid, rays_per_rank=n_rays, kept_rays_global=float(t[4]), ms_fwd_bwd=float(t[0]), ms_allreduce=ar_ms,
                ms_update_repack=float(t[2]), ms_total=float(t[3]), allreduce_bytes=nbytes,
                allreduce_busbw_GBps=(2.0 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9) if world > 1 and ar_ms > 0 else None,
                allreduce_share=ar_ms / max(float(t[3]), 1e-9), kept_rays_per_s=float(t[4]) / (float(t[3]) * 1e-3))
