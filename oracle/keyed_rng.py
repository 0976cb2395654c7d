"""TEST INFRASTRUCTURE (oracle side) -- random sources for the CPU restatement.

The reference draws every random number from torch's global generator, with tensor shapes that
depend on data (valid-sample count M, bounce counts ...).  Two sources are provided:

* ``TorchRNG``  -- draws from the torch global generator with the reference's shapes, in the
  reference's call order.  With the same ``torch.manual_seed`` the restatement therefore consumes
  bit-identical random tensors as the reference; this is what pins the oracle to the reference.
* ``KeyedRNG``  -- a counter-based generator: every random number is a pure function of a 64-bit
  key that identifies *what it is for* (ray, step, bounce ray, stream).  The CUDA kernels implement
  the same function (nmf_b200/csrc/nmf_rng.cuh), so the oracle and the GPU consume identical
  uniforms without any shape-dependent stream alignment.  The hash is the splitmix64 finaliser.

Stream ids (must match nmf_rng.cuh):
    0       seed of the appearance-feature noise (24 normals per sample: noise24)  (models/microfacet.py:297)
    32      bounce-count jitter U                                  (modules/pt_selectors.py:10,12)
    33, 34  per-sample Sobol offset (u, v)                         (brdf_samplers/base.py:17)
    35      retrace tie-break U (keyed by bounce-ray key)          (models/microfacet.py:506)
    1000+j  derives the key of bounce ray j from its sample key
"""
import numpy as np
import torch

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)

STREAM_NOISE0 = 0
STREAM_BOUNCE = 32
STREAM_OFF_U = 33
STREAM_OFF_V = 34
STREAM_TIE = 35
STREAM_JITTER = 36       # train-mode step jitter, keyed by (ray key, step)             (samplers/alphagrid.py:167-173)
STREAM_NOISE_B = 64      # second uniform of the Box-Muller pair for noise dim d is stream 64 + d
STREAM_RAY0 = 1000


def _u64(x):
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    return np.asarray(x).astype(np.uint64)


def mix64(a, b):
    """key' = splitmix64_finalise(a + GOLDEN * (b + 1)); numpy uint64 arithmetic wraps mod 2^64."""
    a = _u64(a)
    b = _u64(b)
    with np.errstate(over="ignore"):
        z = a + GOLDEN * (b + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * M1
        z = (z ^ (z >> np.uint64(27))) * M2
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(key, stream):
    """U[0,1) with 24 random bits (exactly representable in fp32)."""
    z = mix64(key, stream)
    return torch.from_numpy(((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)))


def normal(key, stream_a, stream_b):
    """Box-Muller from two keyed uniforms; u1 is shifted into (0,1]."""
    za = mix64(key, stream_a) >> np.uint64(40)
    u1 = torch.from_numpy((za.astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24))
    u2 = uniform(key, stream_b)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(np.float32(2 * np.pi) * u2)


def fmix32(x):
    """murmur3 32-bit finaliser (numpy uint32 arithmetic wraps mod 2^32)."""
    x = np.asarray(x).astype(np.uint32)
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint32(16))
        x = x * np.uint32(0x85EBCA6B)
        x = x ^ (x >> np.uint32(13))
        x = x * np.uint32(0xC2B2AE35)
        x = x ^ (x >> np.uint32(16))
    return x


def noise24(sample_keys_):
    """(n, 24) standard normals per sample key: one splitmix64 mix seeds two 32-bit counters, feature pair p uses both
    outputs of a Box-Muller transform of (fmix32(lo + c1 (p+1)), fmix32(hi + c2 (p+1)))  (nmf_math.cuh: nmf_noise_pair)."""
    seed = mix64(sample_keys_, STREAM_NOISE0)
    lo = (seed & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (seed >> np.uint64(32)).astype(np.uint32)
    cols = []
    with np.errstate(over="ignore"):
        for p in range(12):
            a = fmix32(lo + np.uint32(0x9E3779B9) * np.uint32(p + 1))
            b = fmix32(hi + np.uint32(0x85EBCA6B) * np.uint32(p + 1))
            u1 = torch.from_numpy(((a >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24))
            u2 = torch.from_numpy((b >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))
            r = torch.sqrt(-2.0 * torch.log(u1))
            ang = np.float32(2 * np.pi) * u2
            cols += [r * torch.cos(ang), r * torch.sin(ang)]
    return torch.stack(cols, dim=1)


def primary_ray_keys(seed, ray_ids):
    return mix64(np.uint64(seed), _u64(ray_ids))


def sample_keys(ray_keys, steps):
    return mix64(ray_keys, steps)


def bounce_ray_keys(sample_keys_, j):
    return mix64(sample_keys_, _u64(j) + np.uint64(STREAM_RAY0))


class TorchRNG:
    """Reference-identical random streams (torch global generator, reference shapes and order)."""
    keyed = False

    def jitter(self, B, S, ray_keys):
        return torch.rand((B, S))                       # alphagrid.py:169

    def app_noise(self, feat, skeys):
        return torch.randn_like(feat)

    def head_noise(self, diffuse, r):
        # render_modules.py:556,558 -- drawn even though std == 0
        torch.randn_like(diffuse)
        torch.randn_like(r)

    def mip_noise(self, sa):
        # integral_equirect.py:392,395 -- drawn even though mipnoise == 0
        torch.rand_like(sa)
        torch.rand_like(sa)

    def bounce_jitter(self, like, keys):
        return torch.rand_like(like)

    def sobol_offset(self, n, skeys):
        return torch.rand(n, 1, 2)

    def tie_break(self, like, bkeys):
        return torch.rand_like(like)


class KeyedRNG:
    keyed = True

    def jitter(self, B, S, ray_keys):
        keys = sample_keys(np.repeat(_u64(ray_keys), S), np.tile(np.arange(S), B))
        return uniform(keys, STREAM_JITTER).reshape(B, S)

    def app_noise(self, feat, skeys):
        assert feat.shape[1] == 24
        return noise24(skeys).reshape(feat.shape)

    def head_noise(self, diffuse, r):
        pass

    def mip_noise(self, sa):
        pass

    def bounce_jitter(self, like, keys):
        return uniform(keys, STREAM_BOUNCE).reshape(like.shape)

    def sobol_offset(self, n, skeys):
        return torch.stack([uniform(skeys, STREAM_OFF_U), uniform(skeys, STREAM_OFF_V)], -1).reshape(n, 1, 2)

    def tie_break(self, like, bkeys):
        return uniform(bkeys, STREAM_TIE).reshape(like.shape)
