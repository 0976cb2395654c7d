"""Device-side scene: packs a reference-format state_dict into the layouts the kernels gather from.

Everything here runs once per weight update (or once per checkpoint load), not per ray: it is host
orchestration in PyTorch (plumbing), the per-ray work is in csrc/.  Layouts are described in DESIGN.md
("HBM layout") and in include/nmf_b200.h (struct NmfScene).

Reference behaviour restated here (file:line relative to the reference tree):
  * stepsize / nSamples                    fields/tensor_base.py:219-232
  * smoothed-difference derivative planes  modules/grid_sample_Cinf.py:24-29,49-63,118-121,218-242
  * summed-area table of exp(bg)/1000      modules/integral_equirect.py:263-273,431-433
  * SH irradiance coefficients             modules/integral_equirect.py:324-360, modules/sh.py:97-157
"""
import ctypes as C
import math

import torch
import torch.nn.functional as F

from . import _lib

MAT_MODE = ((0, 1), (0, 2), (1, 2))   # fields/tensoRF.py:40
VEC_MODE = (2, 1, 0)                  # fields/tensoRF.py:41

APP_STRIDE = 24        # floats per appearance texel (= include/nmf_b200.h NMF_APP_STRIDE)

DEFAULT_HP = dict(
    distance_scale=25.0, density_shift=-4.0, step_ratio=0.5,                 # configs/field/tensorf.yaml
    rays_per_ray=128, max_brdf_rays=(650000, 450000), max_retrace_rays=(1000,), anoise=0.25,
    diffuse_bias=-0.619, diffuse_mul=1.5, roughness_bias=-1.0, tint_bias=0.0, f0_bias=0.0,
    brdf_bias=0.0,                                                            # configs/model/microfacet_tensorf2.yaml
    alpha_mask_thres=1e-3, model="microfacet",
)


def derivative_stencils():
    """The two 5x5 stencils GridSampler2D.backward builds for smoothing=1: a 3x3 Gaussian (std 1, normalised)
    convolved with a central difference (grid_sample_Cinf.py:24-29, 49-63, 218-236).  Returns (Kx, Ky)."""
    blur = torch.tensor([0.0, 1.0, 0.0])
    edge = -1 * torch.tensor([1, 0.0, -1]) / 2
    dy = (blur[None, :] * edge[:, None]).reshape(1, 1, 3, 3)
    dx = dy.permute(0, 1, 3, 2)
    n = torch.arange(0, 3) - (3 - 1.0) / 2.0
    g1 = torch.exp(-(n ** 2) / 2.0)
    smooth = torch.outer(g1, g1)
    smooth = (smooth / smooth.sum()).reshape(1, 1, 3, 3)
    comb = lambda k: -F.conv2d(smooth, k, stride=1, padding=2)
    return comb(dx), comb(dy)


def step_size_and_count(aabb, grid_size, step_ratio):
    aabb_size = aabb[1] - aabb[0]
    gs = torch.as_tensor(list(grid_size), dtype=torch.long, device=aabb.device)
    units = aabb_size / (gs - 1)
    stepsize = torch.min(units) * step_ratio
    diag = torch.sqrt(torch.sum(torch.square(aabb_size)))
    return stepsize, int((diag / stepsize).item()) + 1


def pack_bits(vol):
    """(D,H,W) bool -> (int32 words, pitch): bit (z*H + y)*pitch + x."""
    D, H, W = vol.shape
    pitch = (W + 31) // 32 * 32
    v = torch.zeros(D, H, pitch, dtype=torch.int64, device=vol.device)
    v[..., :W] = vol.to(torch.int64)
    shifts = torch.arange(32, device=vol.device, dtype=torch.int64)
    words = (v.view(D, H, pitch // 32, 32) << shifts).sum(dim=-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    return words.contiguous().view(-1), pitch


def round_tf32(x):
    """round-to-nearest (ties away, like cvt.rna.tf32.f32) of fp32 to the 10-bit TF32 mantissa"""
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def build_sat(bg_mat, brightness, mul):
    """exp activation and summed-area table (integral_equirect.py:263-273, 431-433).  The reference runs two fp32
    cumsums; on CPU ATen accumulates each in double and rounds every prefix to fp32, which is reproduced here
    on any device so that the table does not depend on the scan order of a device cumsum."""
    x = (brightness + mul * bg_mat).to(torch.float32)
    act = torch.exp(x.clip(max=20))
    c1 = torch.cumsum((act / 1000).double(), dim=2).float()
    sat = torch.cumsum(c1.double(), dim=3).float()
    return act, sat


class DeviceScene:
    """Owns the packed device tensors and the NmfScene struct that points at them."""

    def __init__(self, state, aabb, near_far, grid_size, alpha_volume=None, device="cuda", sh_conv=None, **hp):
        self.hp = dict(DEFAULT_HP)
        self.hp.update(hp)
        dev = torch.device(device)
        self.device = dev
        f32 = lambda t: torch.as_tensor(t).detach().to(device=dev, dtype=torch.float32).contiguous()
        self.keep = {}
        s = _lib.NmfScene()
        aabb = f32(aabb)
        aabb_size = aabb[1] - aabb[0]
        inv2 = 2.0 / aabb_size
        self.grid_size = [int(g) for g in grid_size]
        stepsize, n_steps = step_size_and_count(aabb, self.grid_size, self.hp["step_ratio"])
        self.stepsize = float(stepsize)
        self.n_steps = n_steps
        for i in range(3):
            s.aabb0[i] = float(aabb[0, i])
            s.aabb1[i] = float(aabb[1, i])
            s.inv_aabb2[i] = float(inv2[i])
        s.stepsize = self.stepsize
        s.near, s.far = float(near_far[0]), float(near_far[1])
        s.distance_scale = float(self.hp["distance_scale"])
        s.density_shift = float(self.hp["density_shift"])
        s.n_steps = n_steps
        self.aabb = aabb
        self.near_far = (float(near_far[0]), float(near_far[1]))

        self.c = s
        self._pack_factors(state)

        model = self.hp["model"]          # "microfacet" | "plain" | "field" (factors only: plugin-slot field queries)
        s.model = 0 if model == "microfacet" else 1
        if model == "microfacet":
            self._pack_shading(state)
        elif model == "plain":
            self._pack_plain_mlp(state)
        self.c = s
        self.update_hyper()

        self.set_alpha_volume(alpha_volume)
        if "bg_module.bg_mat" in state:
            self._set_env(state, sh_conv)

    def update_hyper(self, **hp):
        """(Re)sets the scalar hyper-parameters NmfScene carries -- the calibrated biases (models/microfacet.py:79-96) and
        the ray budgets the adaptive controller moves (models/microfacet.py:241-268) -- without re-packing any tensor.
        Scratch buffers sized for another max_retrace_rays must be re-created by the caller."""
        self.hp.update(hp)
        s = self.c
        for k in ("diffuse_mul", "diffuse_bias", "tint_bias", "f0_bias", "roughness_bias", "brdf_bias", "anoise"):
            setattr(s, k, float(self.hp[k]))
        s.rays_per_ray = int(self.hp["rays_per_ray"])
        s.max_brdf_rays1 = int(self.hp["max_brdf_rays"][1]) if len(self.hp["max_brdf_rays"]) > 1 else 0
        s.max_retrace = int(self.hp["max_retrace_rays"][0]) if len(self.hp["max_retrace_rays"]) > 0 else 0

    def _pack_shading(self, state, heads=True, brdf=True):
        """Material heads (render_modules.py:519-574) and the BRDF MLP (modules/brdf.py:177-261) + the Sobol table."""
        dev, s = self.device, self.c
        g = lambda k: torch.as_tensor(state[k]).detach().to(device=dev, dtype=torch.float32).contiguous()
        if dev.type == "cuda" and not getattr(self, "_torch_pack", False):
            return self._pack_shading_cuda(state, g, heads, brdf)
        if heads:
            hw = torch.cat([g(f"model.diffuse_module.{h}_mlp.0.weight") for h in ("diffuse", "tint", "f0", "roughness")])
            hb = torch.cat([g(f"model.diffuse_module.{h}_mlp.0.bias") for h in ("diffuse", "tint", "f0", "roughness")])
            assert tuple(hw.shape) == (11, 24), hw.shape
            self._ptr(s, "head_w", hw.contiguous())
            self._ptr(s, "head_b", hb.contiguous())
        if brdf:
            for i, li in enumerate((0, 2, 4)):
                w, b = g(f"model.brdf.mlp.{li}.weight"), g(f"model.brdf.mlp.{li}.bias")
                self._ptr(s, f"brdf_w{i}t", w.t().contiguous())
                self._ptr(s, f"brdf_b{i}", b)
            assert tuple(self.keep["brdf_w0t"].shape) == (66, 64) and tuple(self.keep["brdf_w2t"].shape) == (64, 4)
            # tensor-core operands (csrc/nmf_mlp_tc.cuh): W (rows out, K in) -> fp16 [K/8][rows][8], K zero-padded to 80,
            # bias in column 66 (it multiplies the constant-1 input); layer 3 padded from 4 to 16 output rows
            for i, rows in ((0, 64), (1, 64), (2, 16)):
                w = self.keep[f"brdf_w{i}t"].t()                                  # (out, K)
                wp = torch.zeros(rows, 80, device=dev)
                wp[:w.shape[0], :w.shape[1]] = w
                wp[:w.shape[0], 66] = self.keep[f"brdf_b{i}"]
                self._ptr(s, f"brdf_w{i}u", wp.half().view(rows, 10, 8).permute(1, 0, 2).contiguous())
                # the same tile in BF16 for the reverse pass (csrc/nmf_mlp_tc_bwd.cuh)
                self._ptr(s, f"brdf_w{i}b", wp.bfloat16().view(rows, 10, 8).permute(1, 0, 2).contiguous())
            if self.hp.get("mlp", "f16") not in ("f16", "fp32"):
                raise _lib.NmfError("mlp must be 'f16' (tcgen05, fp16 operands / fp32 accumulate) or 'fp32' (SIMT)")
            s.mlp_mode = 0 if self.hp.get("mlp", "f16") == "f16" else 1
            if "model.brdf_sampler.angs" in state:
                sob = g("model.brdf_sampler.angs")
                assert sob.shape[0] >= 400 and sob.shape[1] == 2
                self._ptr(s, "sobol", sob)

    def _pack_shading_cuda(self, state, g, heads, brdf):
        """_pack_shading as ONE nmf_pack_shading launch into persistent buffers (the NmfScene pointers stay valid across the
        optimiser steps of a training run)."""
        from .ops import _stream
        dev, s = self.device, self.c
        pk = _lib.NmfShadingPack()
        hold = []                                   # keeps the (possibly converted) sources alive until the launch is queued

        def out(name, shape, dtype=torch.float32):
            t = self.keep.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                t = torch.zeros(*shape, device=dev, dtype=dtype)
                self._ptr(s, name, t)
            return t
        if heads:
            hw, hb = out("head_w", (11, 24)), out("head_b", (11,))
            pk.head_w_out, pk.head_b_out = hw.data_ptr(), hb.data_ptr()
            for i, (h, rows) in enumerate((("diffuse", 3), ("tint", 3), ("f0", 3), ("roughness", 2))):
                w, b = g(f"model.diffuse_module.{h}_mlp.0.weight"), g(f"model.diffuse_module.{h}_mlp.0.bias")
                assert tuple(w.shape) == (rows, 24), w.shape
                hold += [w, b]
                pk.head_w[i], pk.head_b[i], pk.head_rows[i] = w.data_ptr(), b.data_ptr(), rows
        if brdf:
            for i, (li, o, k, rows) in enumerate(((0, 64, 66, 64), (2, 64, 64, 64), (4, 4, 64, 16))):
                w, b = g(f"model.brdf.mlp.{li}.weight"), g(f"model.brdf.mlp.{li}.bias")
                assert tuple(w.shape) == (o, k), w.shape
                hold += [w, b]
                pk.w[i], pk.b[i], pk.n_out[i], pk.n_in[i] = w.data_ptr(), b.data_ptr(), o, k
                pk.wt[i] = out(f"brdf_w{i}t", (k, o)).data_ptr()
                pk.bo[i] = out(f"brdf_b{i}", (o,)).data_ptr()
                pk.w16[i] = out(f"brdf_w{i}u", (10, rows, 8), torch.float16).data_ptr()
                pk.wbf[i] = out(f"brdf_w{i}b", (10, rows, 8), torch.bfloat16).data_ptr()
            if self.hp.get("mlp", "f16") not in ("f16", "fp32"):
                raise _lib.NmfError("mlp must be 'f16' (tcgen05, fp16 operands / fp32 accumulate) or 'fp32' (SIMT)")
            s.mlp_mode = 0 if self.hp.get("mlp", "f16") == "f16" else 1
            if "model.brdf_sampler.angs" in state:
                sob = g("model.brdf_sampler.angs")
                assert sob.shape[0] >= 400 and sob.shape[1] == 2
                if "sobol" not in self.keep or self.keep["sobol"].data_ptr() != sob.data_ptr():
                    self._ptr(s, "sobol", sob)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nmf_pack_shading(C.byref(pk), _stream()), "nmf_pack_shading")

    @classmethod
    def shading_only(cls, state, device="cuda", heads=True, brdf=True, **hp):
        """A scene that only carries the material heads and / or the BRDF MLP (the diffuse_module / brdf plugins used on
        their own: nmf_material_heads, nmf_brdf_mlp).  `state`: reference keys (model.diffuse_module.*, model.brdf.*)."""
        self = cls.__new__(cls)
        self.hp = dict(DEFAULT_HP, model="shading")
        self.hp.update(hp)
        self.device = torch.device(device)
        self.keep = {}
        self.c = _lib.NmfScene()
        self._pack_shading(state, heads=heads, brdf=brdf)
        for k in ("diffuse_mul", "diffuse_bias", "tint_bias", "f0_bias", "roughness_bias", "brdf_bias", "anoise"):
            setattr(self.c, k, float(self.hp[k]))
        return self

    def _put(self, name, t, arr=None, index=None):
        """Keeps `t` under `name`; an existing buffer of the same shape is overwritten in place (its device pointer, which
        NmfScene holds, stays valid), otherwise the pointer is (re)set."""
        old = self.keep.get(name)
        if old is not None and old.shape == t.shape and old.dtype == t.dtype:
            old.copy_(t)
            return old
        t = t.contiguous()
        self.keep[name] = t
        if arr is not None:
            arr[index] = t.data_ptr()
        else:
            setattr(self.c, name, t.data_ptr())
        return t

    def _pack_factors_cuda(self, state, derivatives=True):
        """_pack_factors on a CUDA device: nmf_pack_factor (csrc/nmf_repack.cu) writes the channel-last buffers and the
        smoothed-difference planes straight from the reference-layout parameters, in place when the shapes are unchanged."""
        from .ops import _p, _stream
        s, dev = self.c, self.device
        L = _lib.lib()
        f32 = lambda t: torch.as_tensor(t).detach().to(device=dev, dtype=torch.float32).contiguous()
        if "kx25" not in self.keep:
            kx, ky = derivative_stencils()
            self.keep["kx25"], self.keep["ky25"] = f32(kx.reshape(25)), f32(ky.reshape(25))

        def buf(name, shape, arr, p):
            old = self.keep.get(f"{name}{p}")
            if old is None or tuple(old.shape) != tuple(shape):
                old = torch.zeros(*shape, device=dev, dtype=torch.float32)       # (padding channels stay zero)
                self.keep[f"{name}{p}"] = old
            arr[p] = old.data_ptr()
            return old
        with torch.cuda.device(dev):
            for p in range(3):
                dp, dl = f32(state[f"rf.density_rf.app_plane.{p}"]), f32(state[f"rf.density_rf.app_line.{p}"])
                ap, al = f32(state[f"rf.app_rf.app_plane.{p}"]), f32(state[f"rf.app_rf.app_line.{p}"])
                if dp.shape[1] != 16 or ap.shape[1] != 24:
                    raise _lib.NmfError("kernels are compiled for density_n_comp=16, appearance_n_comp=24")
                H, W, N = dp.shape[2], dp.shape[3], dl.shape[2]
                s.plane_w[p], s.plane_h[p], s.line_n[p] = W, H, N
                dval, lval = buf("dval", (H, W, 16), s.dval, p), buf("lval", (N, 16), s.lval, p)
                aval, alval = buf("aval", (H, W, APP_STRIDE), s.aval, p), buf("alval", (N, APP_STRIDE), s.alval, p)
                if derivatives:
                    dpack, lpack = buf("dpack", (H, W, 48), s.dpack, p), buf("lpack", (N, 4, 8), s.lpack, p)
                else:
                    dpack = lpack = None
                    for name, arr in (("dpack", s.dpack), ("lpack", s.lpack)):
                        self.keep.pop(f"{name}{p}", None)
                        arr[p] = None
                kx, ky = _p(self.keep["kx25"]), _p(self.keep["ky25"])
                for src, C, h, w, val, pack in ((dp, 16, H, W, dval, dpack), (dl, 16, N, 1, lval, lpack),
                                                (ap, 24, H, W, aval, None), (al, 24, N, 1, alval, None)):
                    _lib.check(L.nmf_pack_factor(_p(src), C, h, w, kx, ky, _p(val), _p(pack), _stream()), "nmf_pack_factor")
        basis = f32(state["rf.basis_mat.weight"])
        if tuple(basis.shape) != (24, 72):
            raise _lib.NmfError("kernels are compiled for app_dim=24")
        self._put("basis_t", basis.t())

    def _pack_factors(self, state, derivatives=True):
        """Reference-layout factors (1,C,H,W) / (1,C,N,1) -> the channel-last buffers of DESIGN.md section 2.  On a CUDA
        device the hand-written re-pack kernels do it (this runs once per optimiser step in training); the torch version
        below serves CPU-side containers (the host checks of tests/hostcheck) and is the kernels' reference in the tests."""
        if self.device.type == "cuda" and not getattr(self, "_torch_pack", False):
            return self._pack_factors_cuda(state, derivatives)
        s, dev = self.c, self.device
        f32 = lambda t: torch.as_tensor(t).detach().to(device=dev, dtype=torch.float32).contiguous()
        if derivatives:
            kx, ky = derivative_stencils()
            kx, ky = kx.to(dev), ky.to(dev)
            conv = lambda img, k: F.conv2d(img.permute(1, 0, 2, 3), k, stride=1, padding=(2, 2)).permute(1, 0, 2, 3)
        for p in range(3):
            dp = f32(state[f"rf.density_rf.app_plane.{p}"])
            dl = f32(state[f"rf.density_rf.app_line.{p}"])
            ap = f32(state[f"rf.app_rf.app_plane.{p}"])
            al = f32(state[f"rf.app_rf.app_line.{p}"])
            if dp.shape[1] != 16 or ap.shape[1] != 24:
                raise _lib.NmfError("kernels are compiled for density_n_comp=16, appearance_n_comp=24")
            H, W = dp.shape[2], dp.shape[3]
            N = dl.shape[2]
            s.plane_w[p], s.plane_h[p], s.line_n[p] = W, H, N
            dval = dp[0].permute(1, 2, 0)                                                # (H,W,16)
            lval = dl[0, :, :, 0].permute(1, 0)                                          # (N,16)
            packed = [("dval", dval, s.dval), ("lval", lval, s.lval),
                      ("aval", F.pad(ap[0].permute(1, 2, 0), (0, APP_STRIDE - 24)), s.aval),          # (H,W,APP_STRIDE)
                      ("alval", F.pad(al[0, :, :, 0].permute(1, 0), (0, APP_STRIDE - 24)), s.alval)]  # (N,APP_STRIDE)
            if derivatives:
                pdx, pdy = conv(dp, kx)[0].permute(1, 2, 0), conv(dp, ky)[0].permute(1, 2, 0)
                ldy = conv(dl, ky)[0, :, :, 0].permute(1, 0)
                packed += [("dpack", torch.cat([dval, pdx, pdy], dim=2), s.dpack),      # (H,W,48): [val16 | dx16 | dy16]
                           ("lpack", torch.stack([lval.reshape(N, 4, 4), ldy.reshape(N, 4, 4)], dim=2).reshape(N, 4, 8), s.lpack)]
            else:
                for name, arr in (("dpack", s.dpack), ("lpack", s.lpack)):              # stale derivative planes: drop them
                    self.keep.pop(f"{name}{p}", None)
                    arr[p] = None
            for name, t, arr in packed:
                self._put(f"{name}{p}", t, arr, p)
        basis = f32(state["rf.basis_mat.weight"])
        if tuple(basis.shape) != (24, 72):
            raise _lib.NmfError("kernels are compiled for app_dim=24")
        self._put("basis_t", basis.t())

    def _pack_plain_mlp(self, state):
        f32 = lambda t: torch.as_tensor(t).detach().to(device=self.device, dtype=torch.float32).contiguous()
        for i, li in enumerate((0, 2, 4)):
            w, b = f32(state[f"model.diffuse_module.mlp.{li}.weight"]), f32(state[f"model.diffuse_module.mlp.{li}.bias"])
            self._put(f"plain_w{i}t", w.t())
            self._put(f"plain_b{i}", b)
            if i < 2:
                self._put(f"plain_w{i}", w)              # (out, in) as stored: read by the training backward
        assert tuple(self.keep["plain_w0t"].shape) == (135, 128) and tuple(self.keep["plain_w2t"].shape) == (128, 3)

    def refresh_plain(self, state):
        """After an optimiser step of model=tensorf: re-packs the factors and the view MLP IN PLACE (same buffers, same
        NmfScene pointers; no derivative planes -- this model has no normals -- and the occupancy is left alone, as the
        reference only rebuilds it on its schedule, samplers/alphagrid.py:188-199)."""
        if self.hp["model"] != "plain":
            raise _lib.NmfError("refresh_plain: model=tensorf scenes only")
        self._pack_factors(state, derivatives=False)
        self._pack_plain_mlp(state)

    def refresh_microfacet(self, state, env_scalars=None, dev_scalars=None):
        """After an optimiser step of model=microfacet_tensorf2: re-packs the factors (with their smoothed-difference
        planes), the material heads / BRDF MLP operands and the environment tables (SAT, pole rows, SH irradiance) from the
        updated parameters; the occupancy is left alone (the reference rebuilds it on its schedule only)."""
        if self.hp["model"] != "microfacet":
            raise _lib.NmfError("refresh_microfacet: microfacet scenes only")
        self._pack_factors(state, derivatives=True)
        keep_sobol = self.keep.get("sobol")
        self._pack_shading(state)
        if "sobol" not in self.keep and keep_sobol is not None:
            self._ptr(self.c, "sobol", keep_sobol)
        if "bg_module.bg_mat" in state:
            if dev_scalars is not None:       # [brightness, mul, mipbias] stay on the device: no host round trip at all
                return self._set_env(state, None, dev_scalars=dev_scalars)
            if env_scalars is not None:       # host copies of brightness / mul / mipbias, fetched AFTER the packing kernels
                state = dict(state)           # above are queued (the fetch synchronises)
                state.update(env_scalars())
            self._set_env(state, None)

    @classmethod
    def env_only(cls, bg_mat, mipbias=1.0, brightness=0.0, mul=1.0, device="cuda"):
        """A scene that only carries the environment tables (for the IntegralEquirect plugin used on its own)."""
        self = cls.__new__(cls)
        self.hp = dict(DEFAULT_HP, model="env")
        self.device = torch.device(device)
        self.keep = {}
        self.c = _lib.NmfScene()
        self._set_env({"bg_module.bg_mat": bg_mat, "bg_module.mipbias": mipbias, "bg_module.brightness": brightness,
                       "bg_module.mul": mul}, None)
        return self

    def _set_env(self, state, sh_conv, dev_scalars=None):
        s, dev, model = self.c, self.device, self.hp["model"]
        f32 = lambda t: torch.as_tensor(t).detach().to(device=dev, dtype=torch.float32).contiguous()
        bg = f32(state["bg_module.bg_mat"])
        to64 = lambda k, d: torch.as_tensor(state.get(k, d)).detach().to(device=dev, dtype=torch.float64)

        def hostf(k, d):          # python floats pass through without a device round trip (the trainer caches them per step)
            v = state.get(k, d)
            return float(v.detach()) if torch.is_tensor(v) else float(v)
        eh, ew = bg.shape[-2], bg.shape[-1]
        cuda_pack = dev.type == "cuda" and not getattr(self, "_torch_pack", False)
        if dev_scalars is not None and cuda_pack:
            return self._set_env_device(bg, dev_scalars, sh_conv)
        s.env_dyn = None                              # host mode: the kernels read env_mipbias / env_top / env_bot
        brightness, mul = (hostf("bg_module.brightness", 0.0), hostf("bg_module.mul", 1.0)) if cuda_pack else \
            (to64("bg_module.brightness", 0.0), to64("bg_module.mul", 1.0))
        if cuda_pack:
            from .ops import _p, _stream
            sat4 = self.keep.get("env_sat")
            if sat4 is None or tuple(sat4.shape) != (eh, ew, 4):
                sat4 = torch.empty(eh, ew, 4, device=dev, dtype=torch.float32)
                self.keep["env_c1"] = torch.empty(3, eh, ew, device=dev, dtype=torch.float32)
                self.keep["env_pole"] = torch.zeros(6, device=dev, dtype=torch.float64)
            act = None
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().nmf_env_build_sat(_p(bg), eh, ew, float(brightness), float(mul), _p(self.keep["env_c1"]), None,
                                                        _p(sat4), _p(self.keep["env_pole"]), _stream()), "nmf_env_build_sat")
            # paired table for the forward lookups (NmfScene.env_sat2): only while it fits the L2 next to the factor set
            if eh * ew <= 512 * 1024:
                sat8 = self.keep.get("env_sat2")
                if sat8 is None or tuple(sat8.shape) != (eh, ew, 8):
                    sat8 = torch.empty(eh, ew, 8, device=dev, dtype=torch.float32)
                with torch.cuda.device(dev):
                    _lib.check(_lib.lib().nmf_env_pair_sat(_p(sat4), eh, ew, _p(sat8), _stream()), "nmf_env_pair_sat")
                self._ptr(s, "env_sat2", sat8)
            else:
                self.keep.pop("env_sat2", None)
                s.env_sat2 = None
            pole = (self.keep["env_pole"] / ew).float().cpu()          # ONE device-to-host copy for the six pole means
            top, bot = pole[:3], pole[3:]
        else:
            self.keep.pop("env_sat2", None)
            s.env_sat2 = None
            act, sat = build_sat(bg, brightness, mul)
            sat4 = torch.zeros(eh, ew, 4, device=dev, dtype=torch.float32)
            sat4[..., :3] = sat[0].permute(1, 2, 0)
            sat4 = sat4.contiguous()
            top, bot = act[0, :, 0, :].mean(dim=-1), act[0, :, -1, :].mean(dim=-1)
        self._ptr(s, "env_sat", sat4)
        s.env_h, s.env_w = eh, ew
        s.env_mipbias = hostf("bg_module.mipbias", 1.0)
        for i in range(3):
            s.env_top[i] = float(top[i])
            s.env_bot[i] = float(bot[i])
        self.env_act = act
        if model == "microfacet":
            if sh_conv is None:
                sh_conv = self.sh_irradiance()
            self._ptr(s, "sh_conv", f32(sh_conv))

    def _set_env_device(self, bg, dev_scalars, sh_conv=None):
        """_set_env without any host round trip (training): dev_scalars = fp32 device tensor [brightness, mul, mipbias]; the
        tables are built by nmf_env_build_sat_dev, the three quantities the kernels otherwise take by value (mipbias, the pole-row
        means) are left in device memory (NmfScene.env_dyn)."""
        from .ops import _p, _stream
        s, dev = self.c, self.device
        eh, ew = bg.shape[-2], bg.shape[-1]
        sat4 = self.keep.get("env_sat")
        if sat4 is None or tuple(sat4.shape) != (eh, ew, 4) or "env_c1" not in self.keep:
            sat4 = torch.empty(eh, ew, 4, device=dev, dtype=torch.float32)
            self.keep["env_c1"] = torch.empty(3, eh, ew, device=dev, dtype=torch.float32)
            self.keep["env_pole"] = torch.zeros(6, device=dev, dtype=torch.float64)
        dyn = self.keep.get("env_dyn")
        if dyn is None:
            dyn = torch.zeros(8, device=dev, dtype=torch.float32)
        sc = dev_scalars.detach().to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nmf_env_build_sat_dev(_p(bg), eh, ew, _p(sc), _p(self.keep["env_c1"]), None, _p(sat4),
                                                        _p(self.keep["env_pole"]), _p(dyn), _stream()), "nmf_env_build_sat_dev")
            if eh * ew <= 512 * 1024:
                sat8 = self.keep.get("env_sat2")
                if sat8 is None or tuple(sat8.shape) != (eh, ew, 8):
                    sat8 = torch.empty(eh, ew, 8, device=dev, dtype=torch.float32)
                _lib.check(_lib.lib().nmf_env_pair_sat(_p(sat4), eh, ew, _p(sat8), _stream()), "nmf_env_pair_sat")
                self._ptr(s, "env_sat2", sat8)
            else:
                self.keep.pop("env_sat2", None)
                s.env_sat2 = None
        self._ptr(s, "env_sat", sat4)
        self._ptr(s, "env_dyn", dyn)
        s.env_h, s.env_w = eh, ew
        self.env_act = None
        if self.hp["model"] == "microfacet":
            if sh_conv is None:
                sh_conv = self.sh_irradiance()
            old = self.keep.get("sh_conv")
            if old is not None and old.shape == sh_conv.shape and old.dtype == torch.float32:
                old.copy_(sh_conv)
            else:
                self._ptr(s, "sh_conv", sh_conv.float().contiguous())

    def _ptr(self, s, name, t):
        self.keep[name] = t
        setattr(s, name, t.data_ptr())

    def set_alpha_volume(self, alpha_volume):
        """alpha_volume: (Gz,Gy,Gx) 0/1 (AlphaGridMask.alpha_volume, samplers/alphagrid.py:6-21) or None."""
        s = self.c
        if alpha_volume is None:
            s.has_occ = 0
            s.ow = s.oh = s.od = s.opitch = 0
            s.occ_vox = s.occ_cell = s.occ_coarse = None
            s.ocw = s.och = s.ocd = 0
            self.alpha_volume = None
            return
        vol = torch.as_tensor(alpha_volume).to(self.device)
        vol = vol.reshape(vol.shape[-3:]) > 0
        D, H, W = vol.shape
        pad = F.pad(vol.float()[None, None], (0, 1, 0, 1, 0, 1))
        cell = F.max_pool3d(pad, kernel_size=2, stride=1)[0, 0] > 0
        vox_bits, pitch = pack_bits(vol)
        cell_bits, _ = pack_bits(cell)
        self.keep["occ_vox"], self.keep["occ_cell"] = vox_bits, cell_bits
        s.occ_vox, s.occ_cell = vox_bits.data_ptr(), cell_bits.data_ptr()
        s.ow, s.oh, s.od, s.opitch, s.has_occ = W, H, D, pitch, 1
        # conservative coarse field (8 fine cells per axis, dilated by one voxel): coarse cell c is set iff a voxel in
        # [8c-1, 8c+9] is set on every axis; flat bit order, small enough for shared memory or left out
        cdim = [(n + 7) // 8 for n in (W, H, D)]
        s.occ_coarse, s.ocw, s.och, s.ocd = None, 0, 0, 0
        if (cdim[0] * cdim[1] * cdim[2] + 31) // 32 <= 2048:
            padr = [8 * c + 2 - n for c, n in zip(cdim, (W, H, D))]
            padded = F.pad(vol.float()[None, None], (1, padr[0], 1, padr[1], 1, padr[2]))
            coarse = (F.max_pool3d(padded, kernel_size=11, stride=8)[0, 0] > 0).reshape(-1)
            assert coarse.numel() == cdim[0] * cdim[1] * cdim[2]
            nw = (coarse.numel() + 31) // 32
            bits = torch.zeros(nw * 32, dtype=torch.int64, device=vol.device)
            bits[:coarse.numel()] = coarse.to(torch.int64)
            words = (bits.view(nw, 32) << torch.arange(32, device=vol.device, dtype=torch.int64)).sum(dim=1)
            words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).contiguous()
            self.keep["occ_coarse"] = words
            s.occ_coarse = words.data_ptr()
            s.ocw, s.och, s.ocd = cdim
            size = (self.aabb[1] - self.aabb[0]).cpu()
            for i, n in enumerate((W, H, D)):
                s.occ_scale[i] = float((n - 1) / float(size[i]))
        self.alpha_volume = vol

    def update_alpha_mask(self, grid_size=None):
        """AlphaGridSampler.updateAlphaMask (samplers/alphagrid.py:249-276): dense alpha on the lattice (CUDA,
        nmf_dense_alpha), 3^3 max-pool dilation, threshold -> 0/1 volume (Gz,Gy,Gx); installs it as the occupancy."""
        from . import ops
        gs = self.grid_size if grid_size is None else [int(g) for g in grid_size]
        alpha = ops.dense_alpha(self, gs)
        if self.device.type != "cuda" or getattr(self, "_torch_pack", False):
            alpha = F.max_pool3d(alpha.clamp(0, 1)[None, None], kernel_size=3, padding=1, stride=1)[0, 0]
            vol = (alpha >= self.hp["alpha_mask_thres"]).float()
            self.set_alpha_volume(vol)
            return vol
        # pool + threshold + the three bit-fields in hand-written kernels (csrc/nmf_repack.cu)
        from .ops import _p, _stream
        gx, gy, gz = gs
        dev, s = self.device, self.c
        pitch = (gx + 31) // 32 * 32
        n_words = gz * gy * (pitch // 32)
        vox = torch.empty(n_words, dtype=torch.int32, device=dev)
        cell = torch.empty(n_words, dtype=torch.int32, device=dev)
        vol = torch.empty(gz, gy, gx, device=dev)
        cdim = [(n + 7) // 8 for n in (gx, gy, gz)]
        cwords = (cdim[0] * cdim[1] * cdim[2] + 31) // 32
        coarse = torch.zeros(cwords, dtype=torch.int32, device=dev) if cwords <= 2048 else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nmf_occupancy_from_alpha(_p(alpha), gx, gy, gz, float(self.hp["alpha_mask_thres"]), pitch, _p(vox),
                                                           _p(cell), _p(coarse), _p(vol), _stream()), "nmf_occupancy_from_alpha")
        self.keep["occ_vox"], self.keep["occ_cell"] = vox, cell
        s.occ_vox, s.occ_cell = vox.data_ptr(), cell.data_ptr()
        s.ow, s.oh, s.od, s.opitch, s.has_occ = gx, gy, gz, pitch, 1
        s.occ_coarse, s.ocw, s.och, s.ocd = None, 0, 0, 0
        if coarse is not None:
            self.keep["occ_coarse"] = coarse
            s.occ_coarse = coarse.data_ptr()
            s.ocw, s.och, s.ocd = cdim
            size = (self.aabb[1] - self.aabb[0]).cpu()
            for i, n in enumerate((gx, gy, gz)):
                s.occ_scale[i] = float((n - 1) / float(size[i]))
        self.alpha_volume = vol > 0
        return vol

    def sh_irradiance(self, G=100, mipval=-5.0):
        """get_spherical_harmonics(100) (integral_equirect.py:324-360) convolved with the clamped-cosine lobe
        (sh.py:149-157), divided by pi (models/microfacet.py:304-316): (9,3).  Uses the CUDA env lookup."""
        from . import ops
        dev = self.device
        cache = self.keep.get("_sh_quadrature")
        if cache is None or cache[0] != (G, mipval):
            # the quadrature is constant: directions, mip level and the (n, 9) weight matrix = SH basis * sin(theta) *
            # 2 pi^2 / n * the clamped-cosine lobe / pi are built once per scene; every later call is one lookup + one GEMM
            _t = torch.linspace(0, math.pi, G // 2)
            _p = torch.linspace(0, 2 * math.pi, G)
            theta, phi = torch.meshgrid(_t, _p, indexing="ij")
            dirs = torch.stack([torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi), torch.cos(theta)],
                               dim=-1).reshape(-1, 3).to(dev)
            n = dirs.shape[0]
            x, y, z = dirs.unbind(-1)
            c2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
            ev = torch.stack([torch.full_like(x, 0.28209479177387814), 0.4886025119029199 * y, 0.4886025119029199 * z,
                              0.4886025119029199 * x, c2[0] * (x * y), c2[1] * (y * z), c2[2] * (3 * (z * z) - 1),
                              c2[3] * (x * z), c2[4] * (x * x - y * y)], dim=-1)
            st = torch.sin(theta).reshape(n, 1).to(dev)
            al2 = torch.tensor([math.pi] + [2 * math.pi / 3] * 3 + [math.pi / 4] * 5, device=dev)
            wq = (ev * st * (2 * math.pi ** 2 / n)) * (al2 / math.pi).reshape(1, 9)
            cache = ((G, mipval), dirs.contiguous(), torch.full((n,), mipval, device=dev), wq.t().contiguous())
            self.keep["_sh_quadrature"] = cache
        _, dirs, mip, wq_t = cache
        bg = ops.env_lookup(self, dirs, mip)
        return wq_t @ bg.reshape(-1, 3)

    def ref(self):
        return C.byref(self.c)
