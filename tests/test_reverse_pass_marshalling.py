"""CPU check of the HOST side of the reverse-pass operators (ops.EnvMapGrad, ops.NormalsGrad, ops.material_heads_bwd and the
plugin methods over them): the bodies of tests/test_gpu_zz_env_bwd.py are run with the C-ABI entry points replaced -- in this
test process only -- by the host restatements of the same per-element math (tests/hostcheck).  What this pins without a GPU:
argument order and types of every call, buffer layouts and sizes (gsat + pole slots, gradient images, channel-last outputs and
their permutation back to the reference's parameter layout), accumulation across batches / optimiser steps, `.grad` handling in
the plugins, and the tolerances of the GPU tests on their exact inputs.  The kernels themselves are gated by the GPU tests.
TEST INFRASTRUCTURE: the product has no such routing (nmf_b200/_lib.py raises without the CUDA library)."""
import contextlib
import ctypes as C

import pytest
import torch

import conftest
import test_gpu_zz_env_bwd as G


class _HostAbi:
    """the five reverse-pass entry points of include/nmf_b200.h, served by libnmf_hostcheck.so"""

    def __init__(self, hc):
        self.hc = hc
        hc.hc_env_mipbias_grad.restype = C.c_double

    def nmf_env_lookup_bwd_scatter(self, sp, dirs, mip, g, n, gsat, stream):
        s, base = sp._obj, gsat.value
        poles = base + s.env_h * s.env_w * 16
        self.hc.hc_env_bwd_map(s.env_h, s.env_w, C.c_float(s.env_mipbias), dirs, mip, g, n, C.c_void_p(base), C.c_void_p(poles),
                               C.c_void_p(poles + 16))
        return 0

    def nmf_env_lookup_bwd_mipbias(self, sp, dirs, mip, g, n, out, stream):
        C.cast(out, C.POINTER(C.c_float))[0] += self.hc.hc_env_mipbias_grad(sp, dirs, mip, g, n)
        return 0

    def nmf_env_lookup_bwd_finish(self, gsat, h, w, bg, brightness, mul, d_bg, d_br, d_mul, stream):
        poles = gsat.value + h * w * 16
        self.hc.hc_env_map_grad_finish(gsat, h, w, C.c_void_p(poles), C.c_void_p(poles + 16), bg, C.c_float(brightness), C.c_float(mul),
                                       d_bg)
        g = torch.frombuffer((C.c_float * (3 * h * w)).from_address(d_bg.value), dtype=torch.float32)
        b = torch.frombuffer((C.c_float * (3 * h * w)).from_address(bg.value), dtype=torch.float32)
        C.cast(d_br, C.POINTER(C.c_float))[0] += float((g / mul).double().sum())         # = sum d act * act (clip open)
        C.cast(d_mul, C.POINTER(C.c_float))[0] += float((g / mul * b).double().sum())
        return 0

    def nmf_vm_normals_bwd_scatter(self, sp, xyz, n, stride, d_normals, imgs, stream):
        self.hc.hc_normals_bwd(sp, xyz, stride, d_normals, n, imgs._obj.gpack, imgs._obj.glpack)
        return 0

    def nmf_vm_normals_bwd_finish(self, sp, imgs, kx, ky, d_plane, d_line, stream):
        s, im = sp._obj, imgs._obj
        for p in range(3):
            self.hc.hc_normal_grad_finish(C.c_void_p(im.gpack[p]), s.plane_h[p], s.plane_w[p], C.c_void_p(im.glpack[p]), s.line_n[p],
                                          kx, ky, C.c_void_p(d_plane[p]), C.c_void_p(d_line[p]))
        return 0

    def nmf_material_heads_bwd(self, sp, feat, g_albedo, g_f0, g_rough, n, d_w, d_b, d_feat, stream):
        s = sp._obj
        self.hc.hc_heads_bwd(feat, C.c_void_p(s.head_w), C.c_void_p(s.head_b), C.c_float(s.diffuse_mul), C.c_float(s.diffuse_bias),
                             C.c_float(s.f0_bias), C.c_float(s.roughness_bias), g_albedo, g_f0, g_rough, n, d_w, d_b, d_feat)
        return 0


@pytest.fixture
def host_abi(hostcheck, monkeypatch):
    from nmf_b200 import _lib, ops
    from oracle import nmf_oracle as O
    fake = _HostAbi(hostcheck)
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(ops, "_f32", lambda t, dev: t.detach().to(dtype=torch.float32).contiguous())
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(G, "device_scene", lambda fix, dev, **kw: conftest.device_scene(
        fix, "cpu", sh_conv=O.sh_irradiance_coeffs(conftest.oracle_scene(fix)), **kw))
    return "cpu"


@pytest.mark.parametrize("name", ["microfacet_g40", "microfacet_g56_ship", "fullsize_512x1024"])
def test_env_map_gradient_host_side(host_abi, name):
    G.test_env_map_gradient_on_device(host_abi, name)


def test_env_linearity_and_plugin_host_side(host_abi):
    G.test_env_map_gradient_is_linear_and_skips_zero_upstream(host_abi)
    G.test_plugin_accumulates_into_parameter_grads(host_abi)


@pytest.mark.parametrize("name", ["microfacet_g40", "microfacet_g56_ship", "microfacet_noncubic"])
def test_normals_gradient_host_side(host_abi, name):
    G.test_normals_gradient_on_device(host_abi, name)


@pytest.mark.parametrize("n", [1, 255, 5000])
def test_material_heads_gradient_host_side(host_abi, n):
    G.test_material_heads_gradient_on_device(host_abi, n)
