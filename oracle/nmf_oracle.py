"""TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the NMF per-ray render hot path.

This module is NOT part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
or as the timed CPU baseline.  Nothing under ``nmf_b200/`` imports it.

It restates, function by function, what half-potato/nmf computes for one chunk of rays in eval
mode with ``model=microfacet_tensorf2 field=tensorf`` (and the ``model=tensorf`` plumbing config).
All ``file:line`` citations are relative to the reference tree (``/root/reference``):

    sample_rays        samplers/alphagrid.py:131-207, 23-30, 47-50, 278-370
    vm_features        fields/tensoRF.py:161-205, 392-405 ; fields/tensor_base.py:66-93
    vm_normals         fields/tensor_base.py:107-129 ; modules/grid_sample_Cinf.py:109-281
    composite weights  modules/tensor_nerf.py:19-35
    material heads     modules/render_modules.py:519-574
    irradiance         modules/integral_equirect.py:324-360 ; modules/sh.py:97-142
    bounce counts      modules/pt_selectors.py:5-60
    ggx                brdf_samplers/base.py:11-20 ; brdf_samplers/ggx.py:61-268
    ish / brdf         modules/ish.py:94-105 ; modules/sh.py:251-308 ; modules/brdf.py:177-261
    shading            models/microfacet.py:271-673
    environment        modules/integral_equirect.py:18-173, 373-504
    render             modules/tensor_nerf.py:210-674 ; renderer.py:56-106

PARITY PIN.  The reference has no tests, golden vectors or fixtures for this path (SURVEY.md
section 4), so the restatement is pinned against *outputs of the reference itself*: with
``TorchRNG`` (oracle/keyed_rng.py) it consumes the torch global generator with the reference's
tensor shapes in the reference's call order, and ``oracle/make_golden.py`` asserts that it then
reproduces the reference ``TensorNeRF.forward`` on identical rays (see tests/golden/README.md
for the recorded agreement).  With ``KeyedRNG`` every random number is a pure function of
(ray, step, bounce-ray) keys; the CUDA kernels implement the same function, which is what makes
GPU-vs-oracle parity independent of sample ordering.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import keyed_rng as KR

EPS = float(torch.finfo(torch.float32).eps)
MAT_MODE = ((0, 1), (0, 2), (1, 2))   # fields/tensoRF.py:40
VEC_MODE = (2, 1, 0)                  # fields/tensoRF.py:41


def unit(v):
    """mutils.py:8-12"""
    return v / (v ** 2).sum(dim=-1, keepdim=True).clip(min=EPS).sqrt()


# --------------------------------------------------------------------------------------------
# scene container
# --------------------------------------------------------------------------------------------
DEFAULT_HP = dict(
    distance_scale=25.0, density_shift=-4.0, step_ratio=0.5,                # configs/field/tensorf.yaml
    rays_per_ray=128, max_brdf_rays=(650000, 450000), max_retrace_rays=(1000,), anoise=0.25,
    diffuse_bias=-0.619, diffuse_mul=1.5, roughness_bias=-1.0, tint_bias=0.0, f0_bias=0.0,
    brdf_bias=0.0,                                                           # microfacet_tensorf2.yaml
    alpha_mask_thres=1e-3, model="microfacet",
)


class Scene:
    """Weights (reference state_dict key names) + the hyper-parameters the path reads."""

    def __init__(self, state, aabb, near_far, grid_size, alpha_volume=None, requires_grad=False, **hp):
        """requires_grad=True keeps every weight as an autograd leaf (``self.params``: reference state_dict key ->
        leaf) so that ``render_chunk(..., is_train=True)`` can be back-propagated like the reference's training
        forward (gradient oracle for SURVEY 8f row 1)."""
        self.hp = dict(DEFAULT_HP)
        self.hp.update(hp)
        self.requires_grad = requires_grad
        self.params = {}

        def leaf(key, dtype=torch.float32):
            t = state[key].detach().to(dtype).contiguous().clone()
            if requires_grad:
                t.requires_grad_(True)
                self.params[key] = t
            return t

        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        self.aabb = f32(torch.as_tensor(aabb))
        self.near_far = (float(near_far[0]), float(near_far[1]))
        self.grid_size = [int(g) for g in grid_size]
        self.d_plane = [leaf(f"rf.density_rf.app_plane.{i}") for i in range(3)]
        self.d_line = [leaf(f"rf.density_rf.app_line.{i}") for i in range(3)]
        self.a_plane = [leaf(f"rf.app_rf.app_plane.{i}") for i in range(3)]
        self.a_line = [leaf(f"rf.app_rf.app_line.{i}") for i in range(3)]
        self.basis = leaf("rf.basis_mat.weight")
        if self.hp["model"] == "microfacet":
            g = lambda k: leaf(k) if "brdf_sampler" not in k else f32(state[k])
            self.heads = {h: (g(f"model.diffuse_module.{h}_mlp.0.weight"), g(f"model.diffuse_module.{h}_mlp.0.bias"))
                          for h in ("diffuse", "tint", "f0", "roughness")}
            self.brdf = [(g(f"model.brdf.mlp.{i}.weight"), g(f"model.brdf.mlp.{i}.bias")) for i in (0, 2, 4)]
            self.sobol = g("model.brdf_sampler.angs")
        else:
            g = lambda k: leaf(k)
            self.view_mlp = [(g(f"model.diffuse_module.mlp.{i}.weight"), g(f"model.diffuse_module.mlp.{i}.bias"))
                             for i in (0, 2, 4)]
        self.bg_mat = leaf("bg_module.bg_mat")
        self.mipbias = leaf("bg_module.mipbias", torch.float64)            # 0-dim float64 parameters
        self.brightness = leaf("bg_module.brightness", torch.float64)
        self.mul = leaf("bg_module.mul", torch.float64)
        # fields/tensor_base.py:56-62, 219-232
        self.aabb_size = self.aabb[1] - self.aabb[0]
        self.inv_aabb_size = 2.0 / self.aabb_size
        gs = torch.as_tensor(self.grid_size, dtype=torch.long)
        self.units = self.aabb_size / (gs - 1)
        self.stepsize = torch.min(self.units) * self.hp["step_ratio"]
        diag = torch.sqrt(torch.sum(torch.square(self.aabb_size)))
        self.n_samples = int((diag / self.stepsize).item()) + 1
        self.alpha_volume = None if alpha_volume is None else f32(alpha_volume).reshape(1, 1, *alpha_volume.shape[-3:])
        self._deriv = None
        self._env = None
        self._sh = None


# --------------------------------------------------------------------------------------------
# A1 / A2  sample generation, AABB clip, occupancy culling          samplers/alphagrid.py
# --------------------------------------------------------------------------------------------
def occupancy_lookup(sc, xyz):
    """alphagrid.py:23-30,47-50 -- trilinear lookup of the 0/1 volume; sample kept iff value > 0."""
    inv = 1.0 / sc.aabb_size * 2
    c = (xyz[..., :3] - sc.aabb[0]) * inv - 1
    vals = F.grid_sample(sc.alpha_volume, c.view(1, -1, 1, 1, 3), align_corners=True).view(-1)
    return vals > 0


def sample_rays(sc, rays, focal, override_near=None, is_train=False, rng=None, ray_keys=None, max_samples=-1):
    """AlphaGridSampler.sample (alphagrid.py:131-207, 278-370).

    Returns xyzs (M,4), ray_valid (B,S) bool, z_vals (B,S), dists (B,S); in train mode also whole_valid (B) bool
    (the rays kept by the dynamic batch truncation, :353-364) as a fifth value.
    Train mode (cumrand = True, :167-173): z = t_min + cumsum(U * stepsize + stepsize / 2)."""
    o, d = rays[:, :3], rays[:, 3:6]
    near, far = sc.near_far
    if override_near is not None:
        near = override_near
    S = sc.n_samples
    vec = torch.where(d == 0, torch.full_like(d, 1e-6), d)
    rate_a = (sc.aabb[1] - o) / vec
    rate_b = (sc.aabb[0] - o) / vec
    t_min = torch.minimum(rate_a, rate_b).amax(-1).clamp(min=near, max=far)
    if is_train:
        steps = rng.jitter(rays.shape[0], S, ray_keys) * sc.stepsize + sc.stepsize / 2       # :169-172
        z = t_min[..., None] + torch.cumsum(steps, dim=1)
    else:
        k = torch.arange(S)[None].float()
        z = t_min[..., None] + sc.stepsize * k
    pts = o[..., None, :] + d[..., None, :] * z[..., None]
    outside = ((sc.aabb[0] > pts) | (pts > sc.aabb[1])).any(dim=-1)
    pts = torch.cat([pts, z.unsqueeze(-1) / focal], dim=-1)
    valid = ~outside
    if sc.alpha_volume is not None:
        occ = occupancy_lookup(sc, pts[valid])
        inval = ~valid
        inval[valid] |= ~occ
        valid = ~inval
    dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), dim=-1)
    if not is_train:
        return pts[valid], valid, z, dists
    whole = torch.ones(rays.shape[0], dtype=torch.bool)
    if max_samples > 0 and valid.sum() > max_samples:                                        # :353-364
        whole = torch.cumsum(valid.sum(dim=1), dim=0) < max_samples
        valid, pts, z, dists = valid[whole], pts[whole], z[whole], dists[whole]
    return pts[valid], valid, z, dists, whole


# --------------------------------------------------------------------------------------------
# A3 / A5  VM-decomposed field                                       fields/tensoRF.py
# --------------------------------------------------------------------------------------------
def normalize_coord(sc, xyz):
    """tensor_base.py:66-69"""
    return (xyz[..., :3] - sc.aabb[0]) * sc.inv_aabb_size - 1


def _vm_grids(xn):
    xn = xn.detach()                                   # tensoRF.py:182-183: the field is not differentiated w.r.t. positions
    planes = torch.stack([xn[..., list(m)] for m in MAT_MODE]).view(3, -1, 1, 2)
    lines = torch.stack([xn[..., v] for v in VEC_MODE])
    lines = torch.stack((torch.zeros_like(lines), lines), dim=-1).view(3, -1, 1, 2)
    return planes, lines


def _gs(img, grid):
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def vm_products(planes, lines, xn):
    """tensoRF.py:181-205 -- list over the 3 plane/line pairs of (C, M) products."""
    gp, gl = _vm_grids(xn)
    out = []
    for i in range(3):
        pc = _gs(planes[i], gp[[i]]).view(-1, xn.shape[0])
        lc = _gs(lines[i], gl[[i]]).view(-1, xn.shape[0])
        out.append(pc * lc)
    return out


def density_feature(sc, xyz):
    """tensoRF.py:392-400 (dbasis=False): sum over the 48 products."""
    if xyz.shape[0] == 0:
        return torch.empty(0)
    return sum(vm_products(sc.d_plane, sc.d_line, normalize_coord(sc, xyz))).sum(dim=0)


def feature2density(sc, f):
    """tensor_base.py:83-85 (softplus)"""
    return F.softplus(f.clamp(-15, 1e3) + sc.hp["density_shift"])


def app_feature(sc, xyz):
    """tensoRF.py:402-405"""
    coefs = torch.cat(vm_products(sc.a_plane, sc.a_line, normalize_coord(sc, xyz)), dim=0).T
    return coefs @ sc.basis.T


# --------------------------------------------------------------------------------------------
# A4  normals: smoothed central-difference planes                    modules/grid_sample_Cinf.py
# --------------------------------------------------------------------------------------------
def derivative_stencils():
    """The two 5x5 stencils GridSampler2D.backward builds for smoothing >= 1
    (grid_sample_Cinf.py:24-29, 49-63, 118-121, 218-236).  Returns (Kx, Ky), each (1,1,5,5)."""
    f_blur = torch.tensor([0.0, 1.0, 0.0])
    f_edge = -1 * torch.tensor([1, 0.0, -1]) / 2
    dy = (f_blur[None, :] * f_edge[:, None]).reshape(1, 1, 3, 3)
    dx = dy.permute(0, 1, 3, 2)
    n = torch.arange(0, 3) - (3 - 1.0) / 2.0
    g1 = torch.exp(-(n ** 2) / (2 * 1.0 * 1.0))
    smooth = torch.outer(g1, g1)
    smooth = smooth / smooth.sum()
    comb = lambda k: -F.conv2d(smooth.reshape(1, 1, 3, 3), k.reshape(1, 1, 3, 3), stride=1, padding=2)
    return comb(dx), comb(dy)


def derivative_planes(sc):
    """Per density plane: (Pdx, Pdy); per density line: Ldy  (grid_sample_Cinf.py:237-242)."""
    if sc._deriv is None:
        kx, ky = derivative_stencils()
        conv = lambda img, k: F.conv2d(img.permute(1, 0, 2, 3), k, stride=1, padding=(2, 2)).permute(1, 0, 2, 3)
        sc._deriv = dict(pdx=[conv(p, kx) for p in sc.d_plane], pdy=[conv(p, ky) for p in sc.d_plane],
                         ldy=[conv(l, ky) for l in sc.d_line])
    return sc._deriv


def vm_normals(sc, xyz):
    """tensor_base.py:107-129: normalize(-d(density feature)/d(xyz)) with the Cinf backward."""
    xn = normalize_coord(sc, xyz)
    dv = derivative_planes(sc)
    gp, gl = _vm_grids(xn)
    g = torch.zeros(xyz.shape[0], 3)
    for i in range(3):
        M = xn.shape[0]
        pc = _gs(sc.d_plane[i], gp[[i]]).view(-1, M)
        lc = _gs(sc.d_line[i], gl[[i]]).view(-1, M)
        dpx = _gs(dv["pdx"][i], gp[[i]]).view(-1, M)
        dpy = _gs(dv["pdy"][i], gp[[i]]).view(-1, M)
        dly = _gs(dv["ldy"][i], gl[[i]]).view(-1, M)
        g[:, MAT_MODE[i][0]] += (lc * dpx).sum(dim=0)
        g[:, MAT_MODE[i][1]] += (lc * dpy).sum(dim=0)
        g[:, VEC_MODE[i]] += (pc * dly).sum(dim=0)
    g = g * sc.inv_aabb_size
    return unit(-g)


# --------------------------------------------------------------------------------------------
# A6  compositing weights                                            modules/tensor_nerf.py:19-35
# --------------------------------------------------------------------------------------------
def composite_weights(sigma, dist):
    alpha = 1.0 - torch.exp(-sigma * dist)
    T = torch.cumprod(torch.cat([torch.ones(alpha.shape[0], 1), 1.0 - alpha + 1e-10], dim=-1), dim=-1)
    return alpha * T[:, :-1]


# --------------------------------------------------------------------------------------------
# A7  material heads                                                 modules/render_modules.py:553-560
# --------------------------------------------------------------------------------------------
def material_heads(sc, feat):
    hp = sc.hp
    lin = lambda h: feat @ sc.heads[h][0].T + sc.heads[h][1]
    albedo = torch.sigmoid(hp["diffuse_mul"] * lin("diffuse") + hp["diffuse_bias"]).clip(min=0, max=1)
    r = (torch.sigmoid(lin("roughness") + hp["roughness_bias"]) / 2).clip(min=1e-2, max=1)
    tint = torch.sigmoid(lin("tint") + hp["tint_bias"])
    f0 = torch.sigmoid(lin("f0") + hp["f0_bias"])
    return albedo, tint, f0, r[:, 0:1]


# --------------------------------------------------------------------------------------------
# SH bases                                                           modules/sh.py
# --------------------------------------------------------------------------------------------
_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]


def sh9(dirs):
    """sh.py:97-121 (basis_dim = 9)"""
    out = torch.zeros((*dirs.shape[:-1], 9), dtype=dirs.dtype)
    x, y, z = dirs.unbind(-1)
    out[..., 0] = 0.28209479177387814
    out[..., 1] = 0.4886025119029199 * y
    out[..., 2] = 0.4886025119029199 * z
    out[..., 3] = 0.4886025119029199 * x
    out[..., 4] = _C2[0] * (x * y)
    out[..., 5] = _C2[1] * (y * z)
    out[..., 6] = _C2[2] * (3 * (z * z) - 1)
    out[..., 7] = _C2[3] * (x * z)
    out[..., 8] = _C2[4] * (x * x - y * y)
    return out


def ish18(v, rough):
    """ListISH(degs=[0,1,2,4]) (ish.py:102-105 -> sh.py:251-308).  kappa = 1/(rough+1e-3); the
    degree-4 terms carry no attenuation and degree-2 term 3 uses x*y -- both as in the reference."""
    kappa = (1 / (rough + 1e-3)).reshape(-1)
    x, y, z = v.T[0], v.T[1], v.T[2]
    xx, yy, zz = x * x, y * y, z * z
    x4, y4, z4 = x ** 4, y ** 4, z ** 4
    al = lambda l: torch.exp(-l * (l + 1) / 2 / (kappa + 1e-8))
    s0, s1, s2 = al(0), al(1), al(2)
    vals = [
        s0 * 0.28209479177387814 * torch.ones_like(x),
        -s1 * 0.488603 * x, s1 * 0.488603 * z, -s1 * 0.488603 * y,
        s2 * 1.092548 * y * x, -s2 * 1.092548 * y * z, s2 * 0.315392 * (3 * zz - 1), -s2 * 1.092548 * x * y,
        s2 * 0.546274 * (xx - yy),
        2.50334 * x * y * (xx - yy), -1.77013 * y * z * (-3 * xx + yy), 0.946175 * x * y * (7 * zz - 1),
        0.669047 * y * z * (7 * zz - 3), 3.70251 * z4 - 3.17358 * zz + 0.317358,
        0.669047 * x * z * (7 * zz - 3), (0.473087 * xx - 0.473087 * yy) * (7 * zz - 1),
        1.77013 * x * z * (xx - 3 * yy), 0.625836 * x4 - 3.755016 * xx * yy + 0.625836 * y4,
    ]
    return torch.stack(vals, dim=-1)


# --------------------------------------------------------------------------------------------
# A15  environment map: summed-area table lookups                    modules/integral_equirect.py
# --------------------------------------------------------------------------------------------
def env_tables(sc):
    """activation (exp) and SAT, integral_equirect.py:263-273, 431-433.  torch's CPU cumsum
    accumulates in float64 and rounds every prefix to float32."""
    stale = sc._env is not None and sc.requires_grad and torch.is_grad_enabled() and not sc._env[1].requires_grad
    if sc._env is None or stale:                      # a table cached under no_grad cannot carry the gradient
        x = sc.brightness + sc.mul * sc.bg_mat
        act = torch.exp(x.clip(max=20))
        sat = torch.cumsum(torch.cumsum(act / 1000, dim=2), dim=3)
        sc._env = (act, sat)
    return sc._env


def _box(bl, br, tl, tr, size, sat):
    """integral_equirect.py:18-39"""
    c = lambda p: p.clip(min=-1, max=1)
    t = lambda p: F.grid_sample(sat, c(p), mode="bilinear", align_corners=True, padding_mode="zeros")
    return (t(tr) + t(bl) - t(tl) - t(br)).reshape(3, -1).T / size


def _box_wrap_lr(bl, br, tl, tr, size, sat):
    """integral_equirect.py:42-93"""
    out = _box(bl, br, tl, tr, size, sat)
    ex = tr[..., 0] > 1
    if ex.any():
        sel = lambda p: p[ex].reshape(1, 1, -1, 2).clone()
        b1, t1, t2, b2 = sel(bl), sel(tl), sel(tr), sel(br)
        b1[..., 0] = -1.0
        t1[..., 0] = -1.0
        t2[..., 0] = t2[..., 0] - 2
        b2[..., 0] = b2[..., 0] - 2
        out[ex.reshape(-1)] += _box(b1, b2, t1, t2, size[ex.reshape(-1)], sat)
    ex = bl[..., 0] < -1
    if ex.any():
        sel = lambda p: p[ex].reshape(1, 1, -1, 2).clone()
        b1, t1, t2, b2 = sel(bl), sel(tl), sel(tr), sel(br)
        b1[..., 0] = b1[..., 0] + 2
        t1[..., 0] = t1[..., 0] + 2
        t2[..., 0] = 1.0
        b2[..., 0] = 1.0
        out[ex.reshape(-1)] += _box(b1, b2, t1, t2, size[ex.reshape(-1)], sat)
    return out


def _box_wrap(bl, br, tl, tr, size, sat):
    """integral_equirect.py:96-173 -- pole overhang: mirrored box shifted by half a turn."""
    out = _box_wrap_lr(bl, br, tl, tr, size, sat)
    ex = tl[..., 1] > 1
    if ex.any():
        sel = lambda p: p[ex].reshape(1, 1, -1, 2).clone()
        t1, t2, b1, b2 = sel(tl), sel(tr), sel(bl), sel(br)
        rot = torch.where(t1[..., 0] > 0, -1, 1)
        over = (t1[..., 1] - 1).clip(max=0.5, min=0)
        t1[..., 1] = 1.0
        t1[..., 0] = t1[..., 0] + rot
        t2[..., 1] = 1.0
        t2[..., 0] = t2[..., 0] + rot
        b1[..., 1] = 1.0 - over
        b1[..., 0] = b1[..., 0] + rot
        b2[..., 1] = 1.0 - over
        b2[..., 0] = b2[..., 0] + rot
        out[ex.reshape(-1)] += _box_wrap_lr(b1, b2, t1, t2, size[ex.reshape(-1)], sat)
    ex = bl[..., 1] < -1
    if ex.any():
        sel = lambda p: p[ex].reshape(1, 1, -1, 2).clone()
        t1, t2, b1, b2 = sel(tl), sel(tr), sel(bl), sel(br)
        rot = torch.where(t1[..., 0] > 0, -1, 1)
        over = (-1 - b1[..., 1]).clip(max=0.5, min=0)
        b1[..., 1] = -1.0
        b1[..., 0] = b1[..., 0] + rot
        b2[..., 1] = -1.0
        b2[..., 0] = b2[..., 0] + rot
        t1[..., 1] = -1.0 + over
        t1[..., 0] = t1[..., 0] + rot
        t2[..., 1] = -1.0 + over
        t2[..., 0] = t2[..., 0] + rot
        out[ex.reshape(-1)] += _box_wrap_lr(b1, b2, t1, t2, size[ex.reshape(-1)], sat)
    return out


def env_mip_levels(sc, u, sa):
    """sa2mip, integral_equirect.py:373-397 (mipnoise = 0)."""
    h, w = sc.bg_mat.shape[-2], sc.bg_mat.shape[-1]
    sa = sa.reshape(-1)
    cos = (1 - u[:, 2] ** 2).clip(min=EPS).sqrt()
    d = h * w / (2 * math.pi ** 2 * cos).clip(min=EPS)
    area = ((d / 2).log() + sa).exp()
    hh = (area.clip(min=EPS).sqrt() * cos).clip(min=EPS)
    ww = area / hh
    lw = ww.log() / math.log(2) + sc.mipbias
    lh = hh.log() / math.log(2) + sc.mipbias
    return lw.clip(0, 7).float(), lh.clip(0, 7).float()


class _Atan2Damped(torch.autograd.Function):
    """modules/safemath.py:8-32: atan2 whose backward divides by (x^2 + y^2 + 1e-5), the gradient the reference
    trains with (forward identical to torch.atan2)."""

    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x, y)
        return torch.atan2(x, y)

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        den = x ** 2 + y ** 2 + 1e-5
        return g * y / den, g * -x / den


def env_lookup(sc, dirs, sa, rng=None):
    """IntegralEquirect.forward, integral_equirect.py:409-504."""
    if dirs.shape[0] == 0:
        return torch.empty(0, 3)
    if rng is not None:
        rng.mip_noise(sa.reshape(-1))
    act, sat = env_tables(sc)
    h, w = sc.bg_mat.shape[-2], sc.bg_mat.shape[-1]
    lw, lh = env_mip_levels(sc, dirs, sa)
    sw = 2 ** lw / h / 2
    sh = 2 ** lh / h
    offset = torch.stack([sw, sh], dim=-1).reshape(1, 1, -1, 2)
    size = (offset / 2 * torch.tensor([w, h]).reshape(1, 1, 1, 2)).prod(dim=-1).reshape(-1, 1)
    a, b, c = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    norm2d = torch.sqrt(a ** 2 + b ** 2)
    phi = _Atan2Damped.apply(b, a)
    theta = _Atan2Damped.apply(c, norm2d)
    coords = torch.cat([(phi % (2 * math.pi) - math.pi) / math.pi, -theta / math.pi * 2], dim=1)
    x = coords.reshape(1, 1, -1, 2)
    bl = x - offset / 2
    tr = x + offset / 2
    br = x + torch.stack([sw, -sh], dim=-1).reshape(1, 1, -1, 2) / 2
    tl = x + torch.stack([-sw, sh], dim=-1).reshape(1, 1, -1, 2) / 2
    vals = _box_wrap(bl, br, tl, tr, size, sat) * 1000
    cutoff = 1 - 2 / h * 3
    top_row = act[..., 0, :].mean(dim=-1)
    bot_row = act[..., -1, :].mean(dim=-1)
    vals[coords[:, 1] > cutoff] = bot_row
    vals[coords[:, 1] < -cutoff] = top_row
    return vals


def sh_irradiance_coeffs(sc, rng=None, G=100, mipval=-5.0):
    """get_spherical_harmonics, integral_equirect.py:324-360 -> conv_coeffs / pi, shape (9,3)."""
    if sc._sh is not None and rng is None:
        return sc._sh
    _t = torch.linspace(0, np.pi, G // 2)
    _p = torch.linspace(0, 2 * np.pi, G)
    theta, phi = torch.meshgrid(_t, _p, indexing="ij")
    dirs = torch.stack([torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi), torch.cos(theta)],
                       dim=-1).reshape(-1, 3)
    n = dirs.shape[0]
    bg = env_lookup(sc, dirs, mipval * torch.ones(n, 1), rng)
    ev = sh9(dirs)
    coeffs = 2 * np.pi ** 2 * (bg.reshape(n, 1, 3) * ev.reshape(n, -1, 1) * torch.sin(theta.reshape(n, 1, 1))).mean(dim=0)
    al2 = torch.tensor([math.pi] + [2 * math.pi / 3] * 3 + [math.pi / 4] * 5)   # sh.py:149-157
    out = al2.reshape(-1, 1) * coeffs / np.pi
    sc._sh = out
    return out


class EnvScene:
    """Environment-only scene (an IntegralEquirect state_dict: bg_mat, mipbias, brightness, mul) for env_lookup /
    sh_irradiance_coeffs -- e.g. backgrounds/forest.th, which is 1024 x 2048 and carries non-trivial scalars."""

    def __init__(self, sd, **hp):
        self.hp = dict(DEFAULT_HP)
        self.hp.update(hp)
        self.requires_grad, self.params = False, {}
        self.bg_mat = torch.as_tensor(sd["bg_mat"]).detach().float().reshape(1, 3, *sd["bg_mat"].shape[-2:])
        f64 = lambda k, d: torch.as_tensor(sd.get(k, d)).detach().to(torch.float64)
        self.mipbias, self.brightness, self.mul = f64("mipbias", 1.0), f64("brightness", 0.0), f64("mul", 1.0)
        self._env = self._sh = None


def env_lookup_state(sd, dirs, sa):
    """IntegralEquirect(state_dict).forward(dirs, sa)"""
    return env_lookup(EnvScene(sd), dirs, sa.reshape(-1, 1) if sa.dim() == 1 else sa)


def sh_conv_state(sd, G=100):
    """IntegralEquirect(state_dict).get_spherical_harmonics(G)[1] (clamped-cosine-convolved coefficients / pi, :359), in
    the REFERENCE's basis convention: modules/sh.py:67-73 has all-positive degree-2 constants, sh9 here carries a minus on
    the yz / xz terms (same sign in projection and evaluation: the irradiance E(n) is identical)."""
    sign = torch.tensor([1, 1, 1, 1, 1, -1, 1, -1, 1.0]).reshape(-1, 1)
    return sh_irradiance_coeffs(EnvScene(sd), G=G) * sign


# --------------------------------------------------------------------------------------------
# A9  bounce counts                                                  modules/pt_selectors.py:5-60
# --------------------------------------------------------------------------------------------
def bounce_counts(sc, weights, valid, recur, rng, sample_keys, dense_keys):
    """Returns pt_limit.floor() per valid sample (float tensor (M,)) -- ray j of sample i is
    active iff j < floor(pt_limit_i); m = clip(max floor, 0, 400)."""
    if recur == 0:
        w = weights[valid]
        lim = w * sc.hp["rays_per_ray"] + rng.bounce_jitter(w, sample_keys) - 0.5
    else:
        budget = sc.hp["max_brdf_rays"][recur]
        w = weights + 1e-3 * rng.bounce_jitter(weights, dense_keys)
        N = budget - valid.sum()
        if N > 0:
            lim = w / (w.sum().clip(min=1e-3)) * N + 1
        else:
            lim = w / (w.sum().clip(min=1e-3)) * budget + 0.5
        lim = lim[valid]
    return lim.floor()


# --------------------------------------------------------------------------------------------
# A10 / A11  GGX VNDF sampling                                       brdf_samplers/ggx.py
# --------------------------------------------------------------------------------------------
def ggx_sample(u1, u2, V, N, r, ray_mask):
    """GGXSampler.sample (ggx.py:61-226).  V,N (n,3); r (n,1); u (n,m); ray_mask (n,m) bool.
    Returns L (R,3), col_basis (R,3,3) [columns t,b,n], log pdf (R)."""
    n = N.shape[0]
    m = ray_mask.shape[1]
    z_up = torch.tensor([0.0, 0.0, 1.0]).reshape(1, 3).expand(n, 3)
    x_up = torch.tensor([-1.0, 0.0, 0.0]).reshape(1, 3).expand(n, 3)
    up = torch.where(N[:, 2:3].abs() < 0.999, z_up, x_up)
    t = unit(torch.linalg.cross(up, N))
    b = unit(torch.linalg.cross(N, t))
    rows = torch.stack([t, b, N], dim=1).reshape(n, 3, 3)
    V_l = torch.matmul(rows, V.unsqueeze(-1)).squeeze(-1)
    rc = r.squeeze(-1)
    Vs = unit(torch.stack([rc * V_l[..., 0], rc * V_l[..., 1], V_l[..., 2]], dim=-1)).unsqueeze(1)
    T1 = torch.where(Vs[..., 2:3] < 0.999, unit(torch.linalg.cross(Vs, z_up.unsqueeze(1), dim=-1)), x_up.unsqueeze(1))
    T2 = unit(torch.linalg.cross(T1, Vs, dim=-1))
    z = Vs[..., 2].reshape(-1, 1)
    a = (1 / (1 + z.detach()).clip(min=1e-8)).clip(max=1e4)                                  # ggx.py:116
    am = a.expand(u1.shape)[ray_mask]
    rm = rc.reshape(-1, 1).expand(u1.shape)[ray_mask]
    zm = z.expand(u1.shape)[ray_mask]
    u1m, u2m = u1[ray_mask], u2[ray_mask]
    T1m = T1.expand(-1, m, 3)[ray_mask]
    T2m = T2.expand(-1, m, 3)[ray_mask]
    Vsm = Vs.expand(-1, m, 3)[ray_mask]
    cols = rows.permute(0, 2, 1).reshape(n, 1, 3, 3).expand(n, m, 3, 3)[ray_mask]
    rad = torch.sqrt(u1m)
    phi = torch.where(u2m < am, u2m / am * math.pi, (u2m - am) / (1 - am) * math.pi + math.pi)
    tmod = 100 * np.pi
    P1 = (rad * torch.cos(phi % tmod)).unsqueeze(-1)
    P2 = (rad * torch.sin(phi % tmod) * torch.where(u2m < am, torch.tensor(1.0), zm)).unsqueeze(-1)
    Ns = P1 * T1m + P2 * T2m + (1 - P1 * P1 - P2 * P2).clip(min=EPS).sqrt() * Vsm
    H_l = unit(torch.stack([Ns[..., 0] * rm, Ns[..., 1] * rm, Ns[..., 2]], dim=-1))
    H = torch.matmul(cols, H_l.unsqueeze(-1)).squeeze(-1)
    Vo = V.unsqueeze(1).expand(-1, m, 3)[ray_mask]
    eN = N.unsqueeze(1).expand(-1, m, 3)[ray_mask]
    L = unit(2.0 * (Vo * H).sum(dim=-1, keepdim=True) * H - Vo)
    sign = torch.where((L * eN).sum(dim=-1, keepdim=True) > 0, 1, -1)
    L = L * sign
    L_l = torch.matmul(cols.permute(0, 2, 1), L.unsqueeze(-1)).squeeze(-1)
    Vo_l = torch.matmul(cols.permute(0, 2, 1), Vo.unsqueeze(-1)).squeeze(-1)
    with torch.no_grad():                                                                      # ggx.py:218
        logpdf = ggx_pdf(L_l, Vo_l, H_l, rm).clip(min=EPS).log().reshape(-1)
    return L, cols, logpdf


def ggx_pdf(L_l, V_l, H_l, r):
    """compute_prob (ggx.py:228-268), isotropic (r2 = r1)."""
    r2 = r.reshape(-1).clip(min=EPS)
    r1 = (r + r2).reshape(-1).clip(min=EPS) / 2
    lam = (-1 + (1 + ((L_l[:, 0] * r1) ** 2 + (L_l[:, 1] * r2) ** 2) / (L_l[:, 2] ** 2).clip(min=1e-6)).clip(min=EPS).sqrt()) / 2
    invG = 1 + lam
    invD = math.pi * r1 * r2 * (H_l[:, 0] ** 2 / r1 ** 2 + H_l[:, 1] ** 2 / r2 ** 2 + H_l[:, 2] ** 2) ** 2
    logD = -(invG * invD).clip(min=EPS).log() - (4 * V_l[..., 2]).clip(min=EPS).log()
    prob = logD.exp().reshape(-1, 1)
    return torch.where(L_l[:, 2:3] > 0, prob, torch.zeros_like(prob))


# --------------------------------------------------------------------------------------------
# A13  BRDF MLP                                                      modules/brdf.py:177-261
# --------------------------------------------------------------------------------------------
def brdf_mlp(sc, feat, half_l, diff_l, rough):
    x = torch.cat([feat, ish18(half_l, rough), half_l, ish18(diff_l, rough), diff_l], dim=-1)
    (w0, b0), (w1, b1), (w2, b2) = sc.brdf
    h = torch.relu(x @ w0.T + b0)
    h = torch.relu(h @ w1.T + b1)
    out = h @ w2.T + b2
    return torch.sigmoid(out[..., :3] + sc.hp["brdf_bias"])


def row_mask_sum(mat, mask):
    """modules/row_mask_sum.py:15-22"""
    idx = torch.where(mask)[0]
    out = torch.zeros((mask.shape[0], mat.shape[1]), dtype=mat.dtype)
    out.scatter_add_(0, idx[:, None].expand(-1, mat.shape[1]), mat)
    return out


def srgb(img, noclip=False):
    """modules/tonemap.py:38-49"""
    limit = 0.0031308
    out = torch.where(img > limit, 1.055 * (img.clip(min=limit) ** (1.0 / 2.4)) - 0.055, 12.92 * img)
    return out if noclip else out.clip(0, 1)


# --------------------------------------------------------------------------------------------
# A5 .. A16  per-sample shading                                      models/microfacet.py:271-673
# --------------------------------------------------------------------------------------------
def shade_microfacet(sc, xyz, feat, view, normals, weights, valid, recur, rng, ray_keys, trace, aux, is_train=False,
                     detach_N=True):
    """Microfacet.forward (models/microfacet.py:271-673).  The detach / no_grad points of the reference are kept so
    that autograd through this function is the reference's gradient (they change nothing in the forward values)."""
    M = xyz.shape[0]
    rows, steps = torch.where(valid)
    skeys = KR.sample_keys(ray_keys[rows.numpy()], steps.numpy()) if rng.keyed else None
    noise = rng.app_noise(feat, skeys)
    nfeat = feat + noise * sc.hp["anoise"]
    albedo, tint, f0, r1 = material_heads(sc, feat)
    rng.head_noise(albedo, torch.empty(M, 2))
    with torch.no_grad():                                                                      # microfacet.py:305-315
        conv = sh_irradiance_coeffs(sc, rng if not rng.keyed else None)
        E = (conv.reshape(1, -1, 3) * sh9(normals).reshape(M, -1, 1)).sum(dim=1)
    diffuse = albedo * E

    dense_keys = None
    if rng.keyed and recur > 0:
        S = valid.shape[1]
        dense_keys = KR.sample_keys(np.repeat(ray_keys, S), np.tile(np.arange(S), valid.shape[0]))
    with torch.no_grad():                                                                      # pt_selectors.py:5
        kf = bounce_counts(sc, weights, valid, recur, rng, skeys, dense_keys)
    m = int(kf.max().clip(min=0, max=400).int()) if M > 0 else 0
    ray_mask_all = torch.arange(m).reshape(1, -1) < kf.reshape(-1, 1)
    bmask = ray_mask_all.sum(dim=-1) > 0
    ray_mask = ray_mask_all[bmask]
    aux["bounce_count"] = ray_mask_all.sum(dim=-1)

    reflect = torch.zeros_like(diffuse)
    brdf_rgb = torch.zeros_like(diffuse)
    spec = torch.zeros_like(diffuse)
    if bmask.any() and ray_mask.any():
        ri, rj = torch.where(ray_mask)
        bN = normals[bmask]
        if detach_N:                                                                           # microfacet.py:352-353
            bN = bN.detach()
        bV = -view[bmask]
        bN = bN * (bV * bN).sum(dim=-1, keepdim=True).sign()
        rr = r1[bmask]
        if is_train:
            rr = rr.clip(min=sc.hp.get("min_rough", 0.0))                                        # microfacet.py:361-363
        nb = bN.shape[0]
        bkeys = skeys[bmask.numpy()] if rng.keyed else None
        off = rng.sobol_offset(nb, bkeys) * 0.25
        angs = (sc.sobol.reshape(1, -1, 2)[:, :m, :].expand(nb, m, 2) + off) % 1.0
        L, cols, lpdf = ggx_sample(angs[..., 0], angs[..., 1], bV, bN, rr, ray_mask)
        eV = bV[ri]
        eN = bN[ri]
        ea = rr.expand(ray_mask.shape)[ri, rj]
        efeat = nfeat[bmask][ri]
        exyz = xyz[bmask][:, :3][ri]
        H = unit((eV + L) / 2)
        to_local = cols.permute(0, 2, 1)
        diff_l = torch.matmul(to_local, L.unsqueeze(-1)).squeeze(-1)
        half_l = torch.matmul(to_local, H.unsqueeze(-1)).squeeze(-1)
        pdf = lpdf.exp().reshape(-1, 1)
        count = ray_mask.sum(dim=1)
        mip = -torch.log(count[ri].clip(min=1)) - lpdf
        brays = torch.cat([exyz + L * 5e-3, L], dim=-1)
        bw = brdf_mlp(sc, efeat, half_l.detach(), diff_l.detach(), ea.detach())                  # microfacet.py:461-472
        ray_count = (count + 1e-8)[..., None]
        rkeys = KR.bounce_ray_keys(bkeys[ri.numpy()], rj.numpy()) if rng.keyed else None
        R = brays.shape[0]
        incoming = torch.zeros(R, 3)
        retr = sc.hp["max_retrace_rays"]
        if len(retr) > recur:
            n_re = min(R, retr[recur])
            with torch.no_grad():                                                              # microfacet.py:479
                per_sample = weights[valid][bmask].reshape(-1, 1) / ray_count
                per_ray = bw.max(dim=-1, keepdim=True).values * ((eV * eN).sum(dim=-1, keepdim=True) > 0) * pdf
                cc = per_ray.reshape(-1) * per_sample.expand(ray_mask.shape)[ri, rj]
                cc = cc / cc.sum() * n_re
                cc = cc + rng.tie_break(cc, rkeys)
                order = cc.argsort()
            cut = max(order.shape[0] - n_re, 0)
            re_idx, no_idx = order[cut:], order[:cut]
            aux["retrace_score"] = cc
            aux["retrace_idx"] = re_idx
            if len(re_idx) > 0:
                incoming[re_idx] = trace(brays[re_idx], mip[re_idx], True, None if rkeys is None else rkeys[re_idx.numpy()])
            if len(no_idx) > 0:
                incoming[no_idx] = trace(brays[no_idx], mip[no_idx], False, None)
        else:
            incoming = trace(brays, mip, False, None)
        ecount = ray_count.reshape(-1, 1).expand(ray_mask.shape)[ray_mask].reshape(-1, 1).clip(min=1)
        brdf_rgb[bmask] = row_mask_sum(bw / ecount, ray_mask)
        spec[bmask] = row_mask_sum(incoming / ecount, ray_mask)
        R0 = f0[bmask][ri]
        ediff = diffuse[bmask][ri]
        cost = (-eV * H).sum(dim=-1, keepdim=True).abs()
        fres = R0 + (1 - R0) * (1 - cost).clip(min=0, max=1) ** 5
        comb = fres * incoming * bw + (1 - fres) * ediff
        reflect[bmask] = row_mask_sum(comb / ecount, ray_mask)
        aux.update(L=L, logpdf=lpdf, brdf=bw, mip=mip, incoming=incoming, ray_owner=torch.where(bmask)[0][ri], ray_j=rj)
    cost = (-view * normals).sum(dim=-1, keepdim=True).abs()
    fres = f0 + (1 - f0) * (1 - cost).clip(min=0, max=1) ** 5
    debug = dict(diffuse=(1 - fres) * diffuse, tint=fres * brdf_rgb, roughness=r1, spec=spec, albedo=albedo)
    aux.update(albedo=albedo, f0=f0, rough=r1, E=E, app_noise=noise)
    return reflect, debug


def shade_plain(sc, feat, view):
    """model=tensorf: MLPRender_Fea(viewpe=2, feape=2), render_modules.py:201-235."""
    def pe(p, nf):
        bands = (2 ** torch.arange(nf).float())
        pts = (p[..., None] * bands).reshape(p.shape[:-1] + (nf * p.shape[-1],))
        return torch.cat([torch.sin(pts), torch.cos(pts)], dim=-1)
    x = torch.cat([feat, view, pe(feat, 2), pe(view, 2)], dim=-1)
    (w0, b0), (w1, b1), (w2, b2) = sc.view_mlp
    h = torch.relu(x @ w0.T + b0)
    h = torch.relu(h @ w1.T + b1)
    return torch.sigmoid(h @ w2.T + b2), {}


# --------------------------------------------------------------------------------------------
# TensorNeRF.forward (eval)                                          modules/tensor_nerf.py:210-674
# --------------------------------------------------------------------------------------------
def render_chunk(sc, rays, focal, rng, ray_keys=None, recur=0, start_mip=None, override_near=None,
                 white_bg=True, tonemap=True, draw_debug=True, want_aux=False, is_train=False, detach_N=True,
                 max_samples=-1):
    """One TensorNeRF.forward call (modules/tensor_nerf.py:210-674).  Returns (images, statistics[, aux]).
    is_train=True is the training forward: jittered steps, dynamic batch truncation (statistics["whole_valid"]), no
    debug maps (pass draw_debug=False like train.py:556-564); with a Scene built with requires_grad=True the result can
    be back-propagated (detach_N mirrors Microfacet.detach_N, which train.py turns off after the first iterations)."""
    if sc.requires_grad and recur == 0:
        sc._deriv = sc._env = sc._sh = None          # cached tables belong to the previous autograd graph
    aux = {}
    whole = torch.ones(rays.shape[0], dtype=torch.bool)
    if is_train:
        xyz, valid, z, dists, whole = sample_rays(sc, rays, focal, override_near, True, rng, ray_keys, max_samples)
        rays = rays[whole]
        if ray_keys is not None:
            ray_keys = ray_keys[whole.numpy()]
    else:
        xyz, valid, z, dists = sample_rays(sc, rays, focal, override_near)
    B = rays.shape[0]
    M = xyz.shape[0]
    S = valid.shape[1]
    n_samples = [M]
    view = rays[:, 3:6].view(-1, 1, 3).expand(B, S, 3)
    sigma = torch.zeros(B, S)
    normals = torch.zeros(M, 3)
    if M > 0:
        sigma[valid] = feature2density(sc, density_feature(sc, xyz))
    weights = composite_weights(sigma, dists * sc.hp["distance_scale"])
    pw = weights[valid]

    def trace(brays, mip, retrace, keys):
        if retrace:
            im, st = render_chunk(sc, brays, focal, rng, keys, recur + 1, mip.reshape(-1),
                                  3 * sc.stepsize, white_bg=False, tonemap=False, draw_debug=False,
                                  is_train=is_train, detach_N=detach_N, max_samples=-1)   # tensor_nerf.py:291-317
            # (dynamic_batch_size=False in the recursion, tensor_nerf.py:299: re-traced rays are never truncated)
            n_samples.extend(st["n_samples"])
            return im["rgb_map"]
        return env_lookup(sc, brays[..., 3:6], mip.reshape(-1), rng)

    if M > 0:
        feat = app_feature(sc, xyz)
        if sc.hp["model"] == "microfacet":
            normals = vm_normals(sc, xyz)
            rgb, debug = shade_microfacet(sc, xyz, feat, view[valid], normals, weights, valid, recur, rng,
                                          ray_keys, trace, aux, is_train=is_train, detach_N=detach_N)
        else:
            rgb, debug = shade_plain(sc, feat, view[valid])
    else:
        keys = dict(diffuse=3, roughness=1, tint=3, spec=3, albedo=3) if sc.hp["model"] == "microfacet" else {}
        debug = {k: torch.empty(0, v) for k, v in keys.items()}
        rgb = torch.empty(0, 3)
    acc = torch.sum(weights, 1)
    ew = pw[..., None]
    rgb_map = row_mask_sum(ew * rgb, valid)
    images = {}
    stats = dict(recur=recur, whole_valid=whole, n_samples=n_samples)
    if not white_bg:
        mipv = -100 * torch.ones(B, 1) if start_mip is None else start_mip
        bg = env_lookup(sc, view[:, 0, :], mipv, rng).reshape(-1, 3)
        if tonemap:
            bg = srgb(bg, noclip=True)
    else:
        bg = torch.tensor([1.0, 1.0, 1.0]).reshape(1, 3)
    if draw_debug:
        images["depth"] = torch.sum(weights * z, 1)
        wn = row_mask_sum(normals * pw[..., None], valid)
        images["world_normal"] = acc[..., None] * wn + (1 - acc[..., None])
        images["normal"] = acc[..., None] * torch.zeros(B, 3) + (1 - acc[..., None])
        inds = weights.max(dim=1).indices.clip(min=0)
        full = torch.zeros(B, S, 4)
        full[valid] = xyz
        images["termination_xyz"] = full[range(B), inds]
        images["surf_width"] = valid.sum(dim=1)
        below = normalize_coord(sc, xyz)[..., 2] < 0
        images["cross_section"] = row_mask_sum(below[..., None] * ew * rgb.clip(min=0, max=1), valid)
        for k, v in debug.items():
            images[k] = row_mask_sum(v * ew, valid) + (1 - acc[..., None]) * bg
    elif recur == 0:
        stats.update(training_statistics(sc, pw, view[valid], normals, debug))
        for k, v in debug.items():
            images[k] = v
    if tonemap:
        rgb_map = srgb(rgb_map)
    images["rgb_map"] = rgb_map + (1 - acc[..., None]) * bg
    images["acc_map"] = acc
    if want_aux:
        aux.update(xyz=xyz, valid=valid, z=z, dists=dists, sigma=sigma, weights=weights, normals=normals,
                   feat=feat if M > 0 else torch.empty(0, sc.basis.shape[0]), rgb=rgb)
        return images, stats, aux
    return images, stats


def training_statistics(sc, aweight, view, world_normal, debug):
    """A19: the regulariser inputs TensorNeRF.forward returns when it is not drawing debug maps
    (modules/tensor_nerf.py:567-649; microfacet_tensorf2: no normal module => pred_norms = 0, geonorm_iters = -1,
    align_pred_norms = True, distortion loss hard-coded to 0)."""
    ndv2 = (-view.reshape(-1, 3) * world_normal.reshape(-1, 3)).sum(dim=-1)                     # :573-577
    st = dict(ori_loss=(aweight * ndv2.clamp(max=0) ** 2).sum())                                 # :583
    st["distortion_loss"] = torch.tensor(0.0)                                                    # :596
    st["prediction_loss"] = (aweight * (2 * (1 - torch.zeros_like(aweight)))).sum()              # :598-602
    mean_color = torch.exp((sc.brightness + sc.mul * sc.bg_mat).clip(max=20)).reshape(-1, 3).mean(dim=0)
    st["envmap_reg"] = (mean_color.mean() - 0.05).clip(min=0).float()                            # :606-608
    st["brdf_reg"] = debug["tint"].mean().clip(min=0) if "tint" in debug and debug["tint"].numel() else torch.tensor(0.0)
    st["diffuse_reg"] = ((aweight.detach().reshape(-1, 1) * debug["diffuse"]).sum() / 3) if "diffuse" in debug \
        else torch.tensor(0.0)                                                                   # :639-641
    return st


def render_rays(sc, rays, focal, rng, chunk=4096, seed=0, ray_id0=0, keys=None):
    """renderer.chunk_renderer (renderer.py:56-106) in eval / render2completion mode."""
    outs, ns = {}, []
    for c0 in range(0, rays.shape[0], chunk):
        r = rays[c0:c0 + chunk]
        rk = KR.primary_ray_keys(seed, np.arange(ray_id0 + c0, ray_id0 + c0 + r.shape[0])) if rng.keyed else None
        im, st = render_chunk(sc, r, focal, rng, rk)
        for k, v in im.items():
            if keys is None or k in keys:
                outs.setdefault(k, []).append(v)
        ns.append(st["n_samples"])
    return {k: torch.cat(v, 0) for k, v in outs.items()}, ns


# --------------------------------------------------------------------------------------------
# A21  occupancy rebuild                                             samplers/alphagrid.py:209-276
# --------------------------------------------------------------------------------------------
def build_alpha_volume(sc, grid_size=None, use_existing_mask=False):
    """updateAlphaMask: returns the (Gz,Gy,Gx) 0/1 float volume."""
    gs = sc.grid_size if grid_size is None else [int(g) for g in grid_size]
    lin = [torch.linspace(0, 1, g) for g in gs]
    samples = torch.stack(torch.meshgrid(*lin, indexing="ij"), -1)
    dense = sc.aabb[0] * (1 - samples) + sc.aabb[1] * samples
    alpha = torch.zeros_like(dense[..., 0])
    for i in range(gs[0]):
        p = dense[i].view(-1, 3)
        if use_existing_mask and sc.alpha_volume is not None:
            keep = occupancy_lookup(sc, p)
        else:
            keep = torch.ones(p.shape[0], dtype=torch.bool)
        sig = torch.zeros(p.shape[0])
        if keep.any():
            sig[keep] = feature2density(sc, density_feature(sc, p[keep]))
        alpha[i] = (1 - torch.exp(-sig * sc.stepsize)).view(gs[1], gs[2])
    alpha = alpha.clamp(0, 1).transpose(0, 2).contiguous()[None, None]
    alpha = F.max_pool3d(alpha, kernel_size=3, padding=1, stride=1).view(gs[::-1])
    thr = sc.hp["alpha_mask_thres"]
    return (alpha >= thr).float()
