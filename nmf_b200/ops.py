"""Thin Python wrappers over the C ABI (include/nmf_b200.h): PyTorch tensors in, PyTorch tensors out.

PyTorch is used for device memory and streams only; every function here launches hand-written sm_100a kernels
from libnmf_b200.so on torch's current CUDA stream.  There is no fallback: CPU tensors are rejected.
"""
import ctypes as C

import torch

from . import _lib

IMAGE_SHAPES = dict(rgb_map=(3,), acc_map=(), depth=(), world_normal=(3,), normal=(3,), termination_xyz=(4,),
                    surf_width=(), cross_section=(3,), diffuse=(3,), tint=(3,), roughness=(3,), spec=(3,), albedo=(3,))
PLAIN_KEYS = ["rgb_map", "acc_map", "depth", "world_normal", "normal", "termination_xyz", "surf_width", "cross_section"]
DEFAULT_SKIP_EPS = 2e-5
DEFAULT_T_CUT = 1e-5


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _f32(t, dev):
    if not t.is_cuda:
        raise _lib.NmfError("nmf_b200 ops take CUDA tensors (there is no CPU path)")
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def env_lookup(scene, dirs, mip):
    """IntegralEquirect.forward (modules/integral_equirect.py:409-504)."""
    d = _f32(dirs.reshape(-1, 3), scene.device)
    m = _f32(mip.reshape(-1), scene.device)
    out = torch.empty_like(d)
    if d.shape[0]:
        _lib.check(_lib.lib().nmf_env_lookup(scene.ref(), _p(d), _p(m), d.shape[0], _p(out), _stream()), "nmf_env_lookup")
    return out


class EnvMapGrad:
    """Reverse pass of IntegralEquirect lookups w.r.t. the map (modules/integral_equirect.py:263-273, 409-504 under
    autograd): `scatter` once per batch of lookups (accumulates), `finish` once per optimiser step.
    nmf_env_lookup_bwd_scatter / nmf_env_lookup_bwd_finish (csrc/nmf_env_bwd.cu)."""

    def __init__(self, scene):
        self.scene = scene
        self.h, self.w = int(scene.c.env_h), int(scene.c.env_w)
        self.gsat = torch.zeros(self.h * self.w * 4 + 8, device=scene.device)
        self.d_mipbias = torch.zeros(1, device=scene.device)

    def zero(self):
        self.gsat.zero_()
        self.d_mipbias.zero_()

    def scatter(self, dirs, mip, g):
        d = _f32(dirs.reshape(-1, 3), self.scene.device)
        m = _f32(mip.reshape(-1), self.scene.device)
        u = _f32(g.reshape(-1, 3), self.scene.device)
        if not (d.shape[0] == m.shape[0] == u.shape[0]):
            raise _lib.NmfError("EnvMapGrad.scatter: dirs / mip / g disagree on the number of lookups")
        with torch.cuda.device(d.device):
            _lib.check(_lib.lib().nmf_env_lookup_bwd_scatter(self.scene.ref(), _p(d), _p(m), _p(u), d.shape[0], _p(self.gsat),
                                                             _stream()), "nmf_env_lookup_bwd_scatter")
            _lib.check(_lib.lib().nmf_env_lookup_bwd_mipbias(self.scene.ref(), _p(d), _p(m), _p(u), d.shape[0], _p(self.d_mipbias),
                                                             _stream()), "nmf_env_lookup_bwd_mipbias")

    def finish(self, bg_mat, brightness, mul):
        """-> (d bg_mat (1,3,h,w), d brightness, d mul, d mipbias); consumes (and re-zeroes) the accumulated scatter image."""
        bg = _f32(bg_mat.reshape(3, self.h, self.w), self.scene.device)
        brightness, mul = (float(v.detach()) if torch.is_tensor(v) else float(v) for v in (brightness, mul))
        d_bg = torch.zeros_like(bg)
        d_sc = torch.zeros(2, device=bg.device)
        with torch.cuda.device(bg.device):
            _lib.check(_lib.lib().nmf_env_lookup_bwd_finish(_p(self.gsat), self.h, self.w, _p(bg), brightness, mul,
                                                            _p(d_bg), _p(d_sc[0:1]), _p(d_sc[1:2]), _stream()),
                       "nmf_env_lookup_bwd_finish")
        d_mb = self.d_mipbias[0].clone()
        self.zero()
        return d_bg.reshape(1, 3, self.h, self.w), d_sc[0], d_sc[1], d_mb


class NormalsGrad:
    """Reverse pass of TensorBase.compute_normals w.r.t. the density factors (fields/tensor_base.py:107-129 ->
    modules/grid_sample_Cinf.py:109-281 under autograd): `scatter` once per batch of samples (accumulates), `finish` once per
    optimiser step.  nmf_vm_normals_bwd_scatter / nmf_vm_normals_bwd_finish (csrc/nmf_normals_bwd.cu)."""

    def __init__(self, scene):
        from .scene import derivative_stencils
        self.scene = scene
        if "dpack0" not in scene.keep:
            raise _lib.NmfError("NormalsGrad: the scene holds no derivative planes (model without normals)")
        self.gpack = [torch.zeros_like(scene.keep[f"dpack{p}"]) for p in range(3)]
        self.glpack = [torch.zeros_like(scene.keep[f"lpack{p}"]) for p in range(3)]
        self.c = _lib.NmfNormalGrads()
        for p in range(3):
            self.c.gpack[p] = self.gpack[p].data_ptr()
            self.c.glpack[p] = self.glpack[p].data_ptr()
        kx, ky = derivative_stencils()
        self.kx = kx.reshape(25).to(device=scene.device, dtype=torch.float32).contiguous()
        self.ky = ky.reshape(25).to(device=scene.device, dtype=torch.float32).contiguous()

    def zero(self):
        for t in self.gpack + self.glpack:
            t.zero_()

    def scatter(self, xyz, d_normals):
        x = _f32(xyz, self.scene.device)
        g = _f32(d_normals.reshape(-1, 3), self.scene.device)
        if x.dim() != 2 or x.shape[1] < 3 or x.shape[0] != g.shape[0]:
            raise _lib.NmfError("NormalsGrad.scatter: xyz (n, >=3) and d_normals (n, 3) expected")
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().nmf_vm_normals_bwd_scatter(self.scene.ref(), _p(x), x.shape[0], x.shape[1], _p(g), C.byref(self.c),
                                                             _stream()), "nmf_vm_normals_bwd_scatter")

    def finish(self):
        """-> ([d app_plane.p (1,16,H,W)], [d app_line.p (1,16,N,1)]) of rf.density_rf, in the reference's parameter layout;
        consumes (and re-zeroes) the accumulated images."""
        d_plane = [torch.zeros(g.shape[0], g.shape[1], 16, device=g.device) for g in self.gpack]
        d_line = [torch.zeros(g.shape[0], 16, device=g.device) for g in self.glpack]
        pp = (C.c_void_p * 3)(*[t.data_ptr() for t in d_plane])
        lp = (C.c_void_p * 3)(*[t.data_ptr() for t in d_line])
        with torch.cuda.device(self.kx.device):
            _lib.check(_lib.lib().nmf_vm_normals_bwd_finish(self.scene.ref(), C.byref(self.c), _p(self.kx), _p(self.ky), pp, lp,
                                                            _stream()), "nmf_vm_normals_bwd_finish")
        self.zero()
        return ([t.permute(2, 0, 1)[None].contiguous() for t in d_plane],
                [t.t()[None, :, :, None].contiguous() for t in d_line])


def vm_density(scene, xyz, activate=True):
    x = _f32(xyz, scene.device)
    out = torch.empty(x.shape[0], device=x.device)
    if x.shape[0]:
        _lib.check(_lib.lib().nmf_vm_density(scene.ref(), _p(x), x.shape[0], x.shape[1], int(activate), _p(out), _stream()),
                   "nmf_vm_density")
    return out


def vm_appfeature(scene, xyz):
    x = _f32(xyz, scene.device)
    out = torch.empty(x.shape[0], 24, device=x.device)
    if x.shape[0]:
        _lib.check(_lib.lib().nmf_vm_appfeature(scene.ref(), _p(x), x.shape[0], x.shape[1], _p(out), _stream()),
                   "nmf_vm_appfeature")
    return out


def vm_normals(scene, xyz):
    x = _f32(xyz, scene.device)
    out = torch.empty(x.shape[0], 3, device=x.device)
    if x.shape[0]:
        _lib.check(_lib.lib().nmf_vm_normals(scene.ref(), _p(x), x.shape[0], x.shape[1], _p(out), _stream()),
                   "nmf_vm_normals")
    return out


def sample_rays(scene, rays, override_near=None):
    """AlphaGridSampler.sample, eval mode: (ray_valid (B,S) bool, z_vals (B,S), n_valid (B) int32)."""
    r = _f32(rays[:, :6], scene.device)
    B, S = r.shape[0], scene.n_steps
    valid = torch.empty(B, S, dtype=torch.uint8, device=r.device)
    z = torch.empty(B, S, device=r.device)
    nv = torch.empty(B, dtype=torch.int32, device=r.device)
    if B:
        _lib.check(_lib.lib().nmf_sample_rays(scene.ref(), _p(r), B, -1.0 if override_near is None else float(override_near),
                                              _p(valid), _p(z), _p(nv), _stream()), "nmf_sample_rays")
    return valid.bool(), z, nv


def ggx_sample(u, V, N, r):
    dev = V.device
    u, V, N, r = _f32(u, dev), _f32(V, dev), _f32(N, dev), _f32(r.reshape(-1), dev)
    n = V.shape[0]
    L, lp, hl, dl = torch.empty_like(V), torch.empty(n, device=dev), torch.empty_like(V), torch.empty_like(V)
    if n:
        _lib.check(_lib.lib().nmf_ggx_sample(_p(u), _p(V), _p(N), _p(r), n, _p(L), _p(lp), _p(hl), _p(dl), _stream()),
                   "nmf_ggx_sample")
    return L, lp, hl, dl


def brdf_mlp(scene, feat, half_local, diff_local, rough):
    dev = scene.device
    f, h, d, r = _f32(feat, dev), _f32(half_local, dev), _f32(diff_local, dev), _f32(rough.reshape(-1), dev)
    out = torch.empty(f.shape[0], 3, device=dev)
    if f.shape[0]:
        _lib.check(_lib.lib().nmf_brdf_mlp(scene.ref(), _p(f), _p(h), _p(d), _p(r), f.shape[0], _p(out), _stream()),
                   "nmf_brdf_mlp")
    return out


def material_heads(scene, feat, with_r2=False):
    """RandHydraMLPDiffuse.forward (render_modules.py:519-574): (albedo, tint, f0 (n,3), r1 (n)[, r2 (n)])."""
    f = _f32(feat, scene.device)
    n = f.shape[0]
    a, t, f0 = (torch.empty(n, 3, device=f.device) for _ in range(3))
    r1 = torch.empty(n, device=f.device)
    r2 = torch.empty(n, device=f.device) if with_r2 else None
    if n:
        with torch.cuda.device(f.device):
            _lib.check(_lib.lib().nmf_material_heads(scene.ref(), _p(f), n, _p(a), _p(t), _p(f0), _p(r1), _p(r2), _stream()),
                       "nmf_material_heads")
    return (a, t, f0, r1, r2) if with_r2 else (a, t, f0, r1)


def material_heads_bwd(scene, feat, g_albedo, g_f0, g_rough, d_head_w=None, d_head_b=None):
    """Reverse pass of RandHydraMLPDiffuse.forward (render_modules.py:519-574 under autograd): -> (d_head_w (11,24),
    d_head_b (11), d_feat (n,24)); d_head_w / d_head_b accumulate into the given buffers (rows: diffuse 3, tint 3, f0 3,
    roughness 2)."""
    f = _f32(feat, scene.device)
    n = f.shape[0]
    ga, gf = _f32(g_albedo.reshape(-1, 3), scene.device), _f32(g_f0.reshape(-1, 3), scene.device)
    gr = _f32(g_rough.reshape(-1), scene.device)
    if f.dim() != 2 or f.shape[1] != 24 or not (ga.shape[0] == gf.shape[0] == gr.shape[0] == n):
        raise _lib.NmfError("material_heads_bwd: feat (n,24), g_albedo (n,3), g_f0 (n,3), g_rough (n) expected")
    dw = torch.zeros(11, 24, device=f.device) if d_head_w is None else d_head_w
    db = torch.zeros(11, device=f.device) if d_head_b is None else d_head_b
    d_feat = torch.empty_like(f)
    if n:
        with torch.cuda.device(f.device):
            _lib.check(_lib.lib().nmf_material_heads_bwd(scene.ref(), _p(f), _p(ga), _p(gf), _p(gr), n, _p(dw), _p(db), _p(d_feat),
                                                         _stream()), "nmf_material_heads_bwd")
    return dw, db, d_feat


def dense_alpha(scene, grid_size):
    gx, gy, gz = (int(g) for g in grid_size)
    out = torch.empty(gz, gy, gx, device=scene.device)
    # samplers/alphagrid.py:230-236: the lattice coordinates are torch.linspace tensors made on the host
    lins = torch.cat([torch.linspace(0, 1, g) for g in (gx, gy, gz)]).to(scene.device)
    _lib.check(_lib.lib().nmf_dense_alpha(scene.ref(), gx, gy, gz, _p(lins), _p(out), _stream()), "nmf_dense_alpha")
    return out


def upsample_bilinear(src, size):
    """F.interpolate(src, size, mode="bilinear", align_corners=True) for a (1,C,H,W) factor (fields/tensoRF.py:208-227)."""
    x = _f32(src, src.device)
    if x.dim() != 4 or x.shape[0] != 1:
        raise _lib.NmfError("upsample_bilinear takes a (1,C,H,W) tensor")
    _, Cn, H, W = x.shape
    H2, W2 = int(size[0]), int(size[1])
    out = torch.empty(1, Cn, H2, W2, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().nmf_upsample_bilinear(_p(x), Cn, H, W, _p(out), H2, W2, _stream()), "nmf_upsample_bilinear")
    return out


def generate_rays(c2w, H, W, fx, fy=None, cx=None, cy=None, pixel_ids=None, device="cuda", out=None):
    """(n,6) rays of one view generated on the device (dataLoader/ray_utils.py:23-89, blender.py:108-110,146).
    c2w: 3x4 / 4x4 camera-to-world in the OpenCV convention; pixel_ids: optional int32 device tensor (render order)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.NmfError("nmf_b200 ops take CUDA tensors (there is no CPU path)")
    m = torch.as_tensor(c2w, dtype=torch.float32).cpu()[:3, :4].contiguous()
    n = H * W if pixel_ids is None else int(pixel_ids.shape[0])
    if pixel_ids is not None and (pixel_ids.dtype != torch.int32 or not pixel_ids.is_cuda):
        raise _lib.NmfError("pixel_ids must be an int32 CUDA tensor")
    rays = torch.empty(n, 6, device=dev) if out is None else out
    fy = fx if fy is None else fy
    cx = W / 2 if cx is None else cx
    cy = H / 2 if cy is None else cy
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nmf_generate_rays(C.c_void_p(m.data_ptr()), H, W, float(fx), float(fy), float(cx), float(cy),
                                                _p(pixel_ids), n, _p(rays), _stream()), "nmf_generate_rays")
    return rays[:n]


def image_sq_error(rgb, gt, pixel_ids=None):
    """renderer.py:399-401 on the device: fp64 sum of squared errors of the 8-bit quantised render (a (1,) CUDA tensor)."""
    r, g = _f32(rgb.reshape(-1, 3), rgb.device), _f32(gt.reshape(-1, 3), rgb.device)
    out = torch.zeros(1, dtype=torch.float64, device=r.device)
    _lib.check(_lib.lib().nmf_image_sq_error(_p(r), _p(g), _p(pixel_ids), r.shape[0], _p(out), _stream()), "nmf_image_sq_error")
    return out


class RenderBuffers:
    """Output images, counters and scratch for batches of up to `n_rays` rays (grow-only cache per scene)."""

    def __init__(self, scene, n_rays, chunk, keys, cap_scale=1.0, train=False, min_ws_bytes=0):
        dev = scene.device
        self.train = train
        self.n_rays, self.chunk, self.cap_scale, self.keys = n_rays, chunk, float(cap_scale), list(keys)
        self.n_chunks = (n_rays + chunk - 1) // chunk
        self.images = {}
        self.c_images = _lib.NmfImages()
        for k in keys:
            dt = torch.int64 if k == "surf_width" else torch.float32
            self.images[k] = torch.empty((n_rays,) + IMAGE_SHAPES[k], dtype=dt, device=dev)
            setattr(self.c_images, k, self.images[k].data_ptr())
        # every counter (and the training step's loss sums / kept counts) is a view of ONE flat 4-byte-word buffer, so that
        # the host reads them all with a single device-to-host copy (read_counters) instead of one copy per counter
        self.counters = {}
        self.c_counters = _lib.NmfCounters()
        sizes = [("loss3", 6), ("n_kept", 2)]          # loss3: three doubles, first = 8-byte aligned
        for k in _lib.COUNTER_FIELDS:
            sizes.append((k, 2 if k == "n_shaded" else (1 if k == "error" else (4 * self.n_chunks if k == "stat4" else self.n_chunks))))
        self.counter_slices, off = {}, 0
        for k, n in sizes:
            self.counter_slices[k] = (off, n)
            off += (n + 3) // 4 * 4
        self.counter_flat = torch.zeros(off, dtype=torch.int32, device=dev)
        for k in _lib.COUNTER_FIELDS:
            o, n = self.counter_slices[k]
            v = self.counter_flat[o:o + n]
            self.counters[k] = v.view(torch.float32) if k == "stat4" else v
            setattr(self.c_counters, k, self.counters[k].data_ptr())
        self.loss3 = self.counter_flat[0:6].view(torch.float64)
        self.n_kept = self.counter_flat[self.counter_slices["n_kept"][0]:][:2]
        if train:     # one forward call per batch (chunk = n_rays) + the train-mode distance rows
            nbytes = _lib.lib().nmf_render_train_workspace_bytes(scene.ref(), n_rays, self.cap_scale)
            self.whole_valid = torch.zeros(n_rays, dtype=torch.uint8, device=dev)
        else:
            nbytes = _lib.lib().nmf_workspace_bytes_scaled(scene.ref(), n_rays, chunk, self.cap_scale)
        if nbytes == 0:
            raise _lib.NmfError("nmf_workspace_bytes: bad arguments")
        nbytes = max(int(nbytes), int(min_ws_bytes))
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        off = (-self.workspace.data_ptr()) % 256
        self.ws_ptr = self.workspace.data_ptr() + off
        self.ws_bytes = nbytes


def image_keys(scene):
    return list(IMAGE_SHAPES) if scene.hp["model"] == "microfacet" else list(PLAIN_KEYS)


def render_rays(scene, rays, focal, chunk=4096, seed=0, ray_id0=0, skip_eps=DEFAULT_SKIP_EPS, t_cut=DEFAULT_T_CUT,
                buffers=None, check_errors=True):
    """All chunks of `rays` (n,6) in one asynchronous call (nmf_render_rays).  Returns (images, stats):
    images as in TensorNeRF.forward (eval), stats = dict(n_samples=[per-chunk [M0, M1]], counters...)."""
    r = _f32(rays[:, :6], scene.device)
    n = r.shape[0]
    if buffers is None or buffers.n_rays < n or buffers.chunk != chunk:
        buffers = RenderBuffers(scene, n, chunk, image_keys(scene))
    while True:
        rp = _lib.NmfRender(n_rays=n, chunk=chunk, focal=float(focal), seed=int(seed), ray_id0=int(ray_id0),
                            skip_eps=float(skip_eps), t_cut=float(t_cut), white_bg=1, cap_scale=buffers.cap_scale)
        st = _lib.lib().nmf_render_rays(scene.ref(), C.byref(rp), _p(r), C.byref(buffers.c_images), C.byref(buffers.c_counters),
                                        C.c_void_p(buffers.ws_ptr), buffers.ws_bytes, _stream())
        _lib.check(st, "nmf_render_rays")
        images = {k: v[:n] for k, v in buffers.images.items()}
        stats = dict(buffers=buffers)
        if not check_errors:
            return images, stats
        try:
            stats.update(read_counters(buffers, n, chunk))
            return images, stats
        except _lib.NmfOverflow:
            # a scratch list was too small for this scene (results would be incomplete): grow and render again
            if buffers.cap_scale >= 16:
                raise
            buffers = RenderBuffers(scene, max(n, buffers.n_rays), chunk, buffers.keys, cap_scale=buffers.cap_scale * 2)


TRAIN_KEYS = ["rgb_map", "acc_map"]


def render_rays_train(scene, rays, focal, seed=0, ray_id0=0, max_samples=-1, min_rough=0.0, skip_eps=0.0, t_cut=0.0,
                      buffers=None):
    """TensorNeRF.forward(is_train=True, draw_debug=False) of the microfacet model for ONE ray batch
    (nmf_render_rays_train; forward only).  Returns (images, stats): images = rgb_map / acc_map rows of the kept rays,
    stats = dict(whole_valid (n) bool, n_kept, n_samples=[M0, M1], statistics=A19 dict, counters...)."""
    r = _f32(rays[:, :6], scene.device)
    n = r.shape[0]
    if n == 0:
        raise _lib.NmfError("render_rays_train: empty ray batch")
    if buffers is None or not getattr(buffers, "train", False) or buffers.n_rays != n:
        buffers = RenderBuffers(scene, n, n, TRAIN_KEYS, cap_scale=2.0, train=True)
    while True:
        rp = _lib.NmfRender(n_rays=n, chunk=n, focal=float(focal), seed=int(seed), ray_id0=int(ray_id0),
                            skip_eps=float(skip_eps), t_cut=float(t_cut), white_bg=1, cap_scale=buffers.cap_scale)
        tr = _lib.NmfRenderTrain(max_samples=int(max_samples), min_rough=float(min_rough),
                                 whole_valid=buffers.whole_valid.data_ptr(), n_kept=buffers.n_kept.data_ptr())
        st = _lib.lib().nmf_render_rays_train(scene.ref(), C.byref(rp), C.byref(tr), _p(r), C.byref(buffers.c_images),
                                              C.byref(buffers.c_counters), C.c_void_p(buffers.ws_ptr), buffers.ws_bytes,
                                              _stream())
        _lib.check(st, "nmf_render_rays_train")
        try:
            stats = read_counters(buffers, n, n)
        except _lib.NmfOverflow:
            if buffers.cap_scale >= 32:
                raise
            buffers = RenderBuffers(scene, n, n, TRAIN_KEYS, cap_scale=buffers.cap_scale * 2, train=True)
            continue
        kept, m0 = host_n_kept(buffers)
        stats.update(buffers=buffers, whole_valid=buffers.whole_valid[:n].bool(), n_kept=kept,
                     n_samples=stats["n_samples"][0], statistics=stats["statistics"][0])
        assert stats["n_samples"][0] == m0, "kept-sample count of the truncation and of the march disagree"
        return {k: v[:kept] for k, v in buffers.images.items()}, stats


def read_counters(buffers, n, chunk):
    """Synchronises; raises when a device-side list overflowed (results would be incomplete)."""
    nc = (n + chunk - 1) // chunk
    host = buffers.counter_flat.cpu()                  # the one synchronising copy
    buffers.host_counters = host
    c = {}
    for k in _lib.COUNTER_FIELDS:
        o, m = buffers.counter_slices[k]
        c[k] = host[o:o + m].view(torch.float32) if k == "stat4" else host[o:o + m]
    err = int(c["error"][0])
    if err:
        msgs = [m for bit, m in _lib.DEV_ERRORS.items() if err & bit]
        raise _lib.NmfOverflow("nmf_render_rays: " + "; ".join(msgs))
    out = {k: c[k][:nc].tolist() for k in ("n_samples0", "n_samples1", "n_cand", "n_bounce_rays0", "n_bounce_rays1", "n_retrace")}
    out["n_shaded"] = c["n_shaded"].tolist()
    out["n_samples"] = [[a, b] for a, b in zip(out["n_samples0"], out["n_samples1"])]
    out["statistics"] = chunk_statistics(c["stat4"][:4 * nc].reshape(nc, 4), out["n_samples0"])
    return out


def host_n_kept(buffers):
    """(kept rays, kept samples) of the last train-mode call, from the host copy read_counters took."""
    o, m = buffers.counter_slices["n_kept"]
    return buffers.host_counters[o:o + m].tolist()


def host_loss3(buffers):
    """(photometric sum, sum of acc, orientation sum) of the last nmf_train_microfacet call, from read_counters' host copy."""
    return buffers.host_counters[0:6].view(torch.float64).tolist()


def chunk_statistics(stat4, n_samples0, envmap_reg=None):
    """A19 -- the regulariser inputs of TensorNeRF.forward (modules/tensor_nerf.py:567-649), one dict per chunk (= per
    forward call of the reference), from the per-chunk sums the finish kernel leaves in NmfCounters.stat4."""
    res = []
    for (ori, diff, tint, acc), m in zip(stat4.tolist(), n_samples0):
        d = dict(ori_loss=ori, diffuse_reg=diff / 3.0, brdf_reg=max(tint / (3.0 * m), 0.0) if m > 0 else 0.0,
                 prediction_loss=2.0 * acc, distortion_loss=0.0)
        if envmap_reg is not None:
            d["envmap_reg"] = envmap_reg
        res.append(d)
    return res


def profile_enable(on=True):
    _lib.check(_lib.lib().nmf_profile_enable(int(on)), "nmf_profile_enable")


def profile_read():
    """Elapsed milliseconds of the phases of the last nmf_render_rays call (synchronise the stream first)."""
    L = _lib.lib()
    arr = (C.c_float * _lib.N_PHASES)()
    _lib.check(L.nmf_profile_read(arr, _lib.N_PHASES), "nmf_profile_read")
    return {L.nmf_profile_phase_name(i).decode(): float(arr[i]) for i in range(_lib.N_PHASES)}


def gather_peak(set_bytes=78 << 20, taps=64, threads=148 * 2048, reps=5, group=1):
    """Measured ceiling of independent random gathers over a resident set of `set_bytes` (nmf_bench_gather, CUDA events,
    best of `reps`): `group` lanes read one random segment of 16 * group bytes; GB/s = threads * taps * 16 / time.
    78 MB = the factor set of G = 300 (L2-resident)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    n = set_bytes // 16
    buf = torch.rand(n, 4, device=dev)
    threads = (threads + 255) // 256 * 256
    sink = torch.empty(threads, 4, device=dev)
    L = _lib.lib()
    best = None
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.nmf_bench_gather(_p(buf), n, taps, int(group), threads, _p(sink), _stream()), "nmf_bench_gather")
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if i >= 2 and (best is None or ms < best):
            best = ms
    return threads * taps * 16 / (best * 1e-3) / 1e9
