// Per-element math of the NMF render path, shared by every kernel (and compiled for the host by
// tests/hostcheck to check it against the oracle without a GPU).  Each function cites the reference
// code it restates (file:line relative to the reference tree).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define NMF_HD __host__ __device__ __forceinline__
#else
#define NMF_HD inline
#endif

// The index path (sample positions, AABB and occupancy tests) must reproduce ATen's fp32 op order bit
// for bit, so it never contracts a mul and an add into an FMA.
#ifdef __CUDA_ARCH__
#define NMF_MUL(a, b) __fmul_rn((a), (b))
#define NMF_ADD(a, b) __fadd_rn((a), (b))
#define NMF_SUB(a, b) __fsub_rn((a), (b))
#define NMF_DIV(a, b) __fdiv_rn((a), (b))
#else
#define NMF_MUL(a, b) ((a) * (b))   // host build uses -ffp-contract=off
#define NMF_ADD(a, b) ((a) + (b))
#define NMF_SUB(a, b) ((a) - (b))
#define NMF_DIV(a, b) ((a) / (b))
#endif

#define NMF_EPS 1.1920929e-07f
#define NMF_PI 3.14159265358979323846f

// ------------------------------------------------------------------------------------------------
// keyed random numbers (oracle/keyed_rng.py): splitmix64 finaliser over (key, stream)
// ------------------------------------------------------------------------------------------------
#define NMF_STREAM_NOISE0 0u
#define NMF_STREAM_BOUNCE 32u
#define NMF_STREAM_OFF_U 33u
#define NMF_STREAM_OFF_V 34u
#define NMF_STREAM_TIE 35u
#define NMF_STREAM_NOISE_B 64u
#define NMF_STREAM_RAY0 1000u

NMF_HD uint64_t nmf_mix64(uint64_t a, uint64_t b) {
  uint64_t z = a + 0x9E3779B97F4A7C15ull * (b + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
NMF_HD float nmf_uniform(uint64_t key, uint32_t stream) {
  return (float)(uint32_t)(nmf_mix64(key, stream) >> 40) * 5.9604644775390625e-8f;
}
NMF_HD float nmf_normal(uint64_t key, uint32_t sa, uint32_t sb) {
  float u1 = ((float)(uint32_t)(nmf_mix64(key, sa) >> 40) + 1.0f) * 5.9604644775390625e-8f;
  float u2 = nmf_uniform(key, sb);
#ifdef __CUDA_ARCH__
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);   // cos(2 pi u2) without cosf's range-reduction slow path
#else
  return sqrtf(-2.0f * logf(u1)) * cosf(6.2831855f * u2);
#endif
}

// Appearance-feature noise (models/microfacet.py:297: 24 standard normals per shaded sample).  One 64-bit mix per sample
// seeds two 32-bit counters; pair p (features 2p, 2p+1) hashes them with the murmur3 finaliser and uses BOTH outputs of
// a Box-Muller transform -- a third of the integer work of 24 independent splitmix64 streams (oracle/keyed_rng.py).
NMF_HD uint32_t nmf_fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
NMF_HD uint64_t nmf_noise_seed(uint64_t sample_key) { return nmf_mix64(sample_key, NMF_STREAM_NOISE0); }
NMF_HD void nmf_noise_pair(uint64_t seed, uint32_t p, float* n0, float* n1) {
  const uint32_t a = nmf_fmix32((uint32_t)seed + 0x9E3779B9u * (p + 1u));
  const uint32_t b = nmf_fmix32((uint32_t)(seed >> 32) + 0x85EBCA6Bu * (p + 1u));
  const float u1 = ((float)(a >> 8) + 1.0f) * 5.9604644775390625e-8f;      // (0, 1]
  const float u2 = (float)(b >> 8) * 5.9604644775390625e-8f;               // [0, 1)
#ifdef __CUDA_ARCH__
  // SFU versions (MUFU.LG2 / MUFU.SIN / MUFU.COS): absolute error ~1e-6 on a feature noise of scale `anoise`, far below the
  // parity tolerance of anything downstream; the accurate libm versions cost ~45 instructions per pair, 12 pairs per sample
  const float r = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(6.2831855f * u2, &sn, &cs);
#else
  const float r = sqrtf(-2.0f * logf(u1));
  const float sn = sinf(6.2831855f * u2), cs = cosf(6.2831855f * u2);
#endif
  *n0 = r * cs;
  *n1 = r * sn;
}

// ------------------------------------------------------------------------------------------------
// small vector helpers
// ------------------------------------------------------------------------------------------------
struct nmf_v3 { float x, y, z; };
NMF_HD nmf_v3 nmf_mk3(float x, float y, float z) { nmf_v3 r; r.x = x; r.y = y; r.z = z; return r; }
NMF_HD float nmf_dot(nmf_v3 a, nmf_v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
NMF_HD nmf_v3 nmf_cross(nmf_v3 a, nmf_v3 b) {
  return nmf_mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
NMF_HD nmf_v3 nmf_scale(nmf_v3 a, float s) { return nmf_mk3(a.x * s, a.y * s, a.z * s); }
NMF_HD nmf_v3 nmf_add3(nmf_v3 a, nmf_v3 b) { return nmf_mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
// mutils.py:8-12  v / sqrt(clip(sum v^2, eps))
NMF_HD nmf_v3 nmf_unit(nmf_v3 v) {
  float n = sqrtf(fmaxf(v.x * v.x + v.y * v.y + v.z * v.z, NMF_EPS));
  return nmf_mk3(v.x / n, v.y / n, v.z / n);
}
NMF_HD float nmf_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
NMF_HD float nmf_clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ------------------------------------------------------------------------------------------------
// A1 sample generation: samplers/alphagrid.py:145-207
// ------------------------------------------------------------------------------------------------
NMF_HD float nmf_ray_tmin(const float* o, const float* d, const float* aabb0, const float* aabb1, float near_, float far_) {
  float m = -INFINITY;
  for (int a = 0; a < 3; ++a) {
    float vec = (d[a] == 0.0f) ? 1e-6f : d[a];                 // alphagrid.py:149
    float ra = NMF_DIV(NMF_SUB(aabb1[a], o[a]), vec);          // :150
    float rb = NMF_DIV(NMF_SUB(aabb0[a], o[a]), vec);          // :151
    m = fmaxf(m, fminf(ra, rb));                               // :152
  }
  return fminf(fmaxf(m, near_), far_);
}
// z_k = t_min + stepsize * k  (alphagrid.py:190-193);  p = o + d * z (:195)
NMF_HD float nmf_step_z(float tmin, float stepsize, int k) { return NMF_ADD(tmin, NMF_MUL(stepsize, (float)k)); }
NMF_HD void nmf_step_pos(const float* o, const float* d, float z, float* p) {
  p[0] = NMF_ADD(o[0], NMF_MUL(d[0], z));
  p[1] = NMF_ADD(o[1], NMF_MUL(d[1], z));
  p[2] = NMF_ADD(o[2], NMF_MUL(d[2], z));
}
NMF_HD bool nmf_inside(const float* p, const float* aabb0, const float* aabb1) {     // alphagrid.py:197
  return !((aabb0[0] > p[0]) | (p[0] > aabb1[0]) | (aabb0[1] > p[1]) | (p[1] > aabb1[1]) | (aabb0[2] > p[2]) | (p[2] > aabb1[2]));
}
// normalised coordinate (alphagrid.py:47-50, tensor_base.py:66-69): (p - aabb0) * inv - 1
NMF_HD float nmf_norm_coord(float p, float a0, float inv2) { return NMF_SUB(NMF_MUL(NMF_SUB(p, a0), inv2), 1.0f); }
// grid_sample unnormalise with align_corners=True: ((x + 1) / 2) * (size - 1)
NMF_HD float nmf_unnorm(float x, int size) { return NMF_MUL(NMF_MUL(NMF_ADD(x, 1.0f), 0.5f), (float)(size - 1)); }

// ------------------------------------------------------------------------------------------------
// A2 occupancy: samplers/alphagrid.py:23-30 -- trilinear lookup of the 0/1 volume, keep iff > 0.
// value > 0  <=>  some in-range corner voxel is set and all three of its 1-D weights are > 0.
// ------------------------------------------------------------------------------------------------
NMF_HD bool nmf_bit(const uint32_t* bits, int w, int h, int d, int pitch, int x, int y, int z) {
  if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h || (unsigned)z >= (unsigned)d) return false;
  size_t i = ((size_t)z * h + y) * (size_t)pitch + x;
  return (bits[i >> 5] >> (i & 31)) & 1u;
}
NMF_HD bool nmf_occupied(const uint32_t* vox, const uint32_t* cell, int w, int h, int d, int pitch, float cx, float cy, float cz) {
  float ix = nmf_unnorm(cx, w), iy = nmf_unnorm(cy, h), iz = nmf_unnorm(cz, d);
  float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
  int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
  float tx = ix - fx0, ty = iy - fy0, tz = iz - fz0;   // exact
  if (tx > 0.0f && ty > 0.0f && tz > 0.0f && x0 >= 0 && y0 >= 0 && z0 >= 0) {
    // all eight weights are positive: the answer is the OR of the in-range corners
    return nmf_bit(cell, w, h, d, pitch, x0, y0, z0);
  }
  // a sample exactly on a lattice plane: the far corner along that axis has weight 0
  for (int c = 0; c < 8; ++c) {
    int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
    float wx = dx ? tx : (fx0 + 1.0f) - ix;
    float wy = dy ? ty : (fy0 + 1.0f) - iy;
    float wz = dz ? tz : (fz0 + 1.0f) - iz;
    if (wx * wy * wz > 0.0f && nmf_bit(vox, w, h, d, pitch, x0 + dx, y0 + dy, z0 + dz)) return true;
  }
  return false;
}

// The common case of nmf_occupied split in two, so that a kernel can issue the loads of several steps before it
// consumes any of them: nmf_occ_fast returns true when all eight trilinear weights are positive (one bit of the
// cell field answers); *word = index of the 32-bit word to load (or -1: the cell is out of range => not occupied).
NMF_HD bool nmf_occ_fast(int w, int h, int d, int pitch, float cx, float cy, float cz, long long* word, int* shift) {
  float ix = nmf_unnorm(cx, w), iy = nmf_unnorm(cy, h), iz = nmf_unnorm(cz, d);
  float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
  int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
  float tx = ix - fx0, ty = iy - fy0, tz = iz - fz0;   // exact
  if (!(tx > 0.0f && ty > 0.0f && tz > 0.0f && x0 >= 0 && y0 >= 0 && z0 >= 0)) return false;
  if (x0 >= w || y0 >= h || z0 >= d) { *word = -1; *shift = 0; return true; }
  size_t i = ((size_t)z0 * h + y0) * (size_t)pitch + x0;
  *word = (long long)(i >> 5);
  *shift = (int)(i & 31);
  return true;
}

// ------------------------------------------------------------------------------------------------
// A3 bilinear taps of F.grid_sample(align_corners=True, zeros padding): fields/tensoRF.py:181-205
// ------------------------------------------------------------------------------------------------
struct NmfLerp { int i0, i1; float w0, w1; };   // value = w0 * v[i0] + w1 * v[i1]; out-of-range taps get weight 0
NMF_HD NmfLerp nmf_lerp_setup(float xn, int size) {
  NmfLerp L;
  float ix = nmf_unnorm(xn, size);
  float f = floorf(ix);
  float t = ix - f;
  int i0 = (int)f;
  L.w1 = t;
  L.w0 = 1.0f - t;
  L.i0 = i0;
  L.i1 = i0 + 1;
  if ((unsigned)L.i0 >= (unsigned)size) { L.w0 = 0.0f; L.i0 = 0; }
  if ((unsigned)L.i1 >= (unsigned)size) { L.w1 = 0.0f; L.i1 = 0; }
  return L;
}
// feature -> density: tensor_base.py:83-85  softplus(clamp(f, -15, 1e3) + shift), torch softplus threshold 20
NMF_HD float nmf_feature2density(float f, float shift) {
  float x = nmf_clampf(f, -15.0f, 1000.0f) + shift;
  return x > 20.0f ? x : log1pf(expf(x));
}

// ------------------------------------------------------------------------------------------------
// SH bases: modules/sh.py:97-121 (9 terms) and the roughness-attenuated list basis of
// modules/ish.py:94-105 -> modules/sh.py:251-308 (degrees 0,1,2,4 = 18 terms)
// ------------------------------------------------------------------------------------------------
NMF_HD void nmf_sh9(nmf_v3 d, float* o) {
  float x = d.x, y = d.y, z = d.z;
  o[0] = 0.28209479177387814f;
  o[1] = 0.4886025119029199f * y;
  o[2] = 0.4886025119029199f * z;
  o[3] = 0.4886025119029199f * x;
  o[4] = 1.0925484305920792f * (x * y);
  o[5] = -1.0925484305920792f * (y * z);
  o[6] = 0.31539156525252005f * (3.0f * (z * z) - 1.0f);
  o[7] = -1.0925484305920792f * (x * z);
  o[8] = 0.5462742152960396f * (x * x - y * y);
}
// roughness attenuation of degrees 1 and 2: exp(-l(l+1)/2/kappa), kappa = 1/(rough + 1e-3)  (sh.py:268-275)
NMF_HD void nmf_ish_scales(float rough, float* s1, float* s2) {
  float kappa = 1.0f / (rough + 1e-3f);
  float kk = kappa + 1e-8f;
  *s1 = expf(-1.0f / kk);
  *s2 = expf(-3.0f / kk);
}
NMF_HD void nmf_ish18_s(nmf_v3 v, float s1, float s2, float* o) {
  float x = v.x, y = v.y, z = v.z;
  float xx = x * x, yy = y * y, zz = z * z;
  float x4 = xx * xx, y4 = yy * yy, z4 = zz * zz;
  float s0 = 1.0f;                      // exp(-0)
  o[0] = s0 * 0.28209479177387814f;
  o[1] = -s1 * 0.488603f * x;
  o[2] = s1 * 0.488603f * z;
  o[3] = -s1 * 0.488603f * y;
  o[4] = s2 * 1.092548f * y * x;
  o[5] = -s2 * 1.092548f * y * z;
  o[6] = s2 * 0.315392f * (3.0f * zz - 1.0f);
  o[7] = -s2 * 1.092548f * x * y;       // sic: the reference uses x*y here (sh.py:283-308)
  o[8] = s2 * 0.546274f * (xx - yy);
  o[9] = 2.50334f * x * y * (xx - yy);
  o[10] = -1.77013f * y * z * (-3.0f * xx + yy);
  o[11] = 0.946175f * x * y * (7.0f * zz - 1.0f);
  o[12] = 0.669047f * y * z * (7.0f * zz - 3.0f);
  o[13] = 3.70251f * z4 - 3.17358f * zz + 0.317358f;
  o[14] = 0.669047f * x * z * (7.0f * zz - 3.0f);
  o[15] = (0.473087f * xx - 0.473087f * yy) * (7.0f * zz - 1.0f);
  o[16] = 1.77013f * x * z * (xx - 3.0f * yy);
  o[17] = 0.625836f * x4 - 3.755016f * xx * yy + 0.625836f * y4;
}
NMF_HD void nmf_ish18(nmf_v3 v, float rough, float* o) {
  float s1, s2;
  nmf_ish_scales(rough, &s1, &s2);
  nmf_ish18_s(v, s1, s2, o);
}

// ------------------------------------------------------------------------------------------------
// A10/A11 GGX VNDF importance sampling: brdf_samplers/ggx.py:61-226 (sample), :228-268 (pdf)
// V: unit vector to the viewer, N: normal already flipped to V's side (models/microfacet.py:354-356)
// ------------------------------------------------------------------------------------------------
struct NmfGGX {
  nmf_v3 L;        // world-space outgoing direction
  nmf_v3 half_l;   // normalize((V+L)/2) in the local frame (models/microfacet.py:388-400)
  nmf_v3 diff_l;   // L in the local frame
  nmf_v3 H;        // world-space normalize((V+L)/2)
  float logpdf;
};
NMF_HD float nmf_ggx_pdf(nmf_v3 L_l, nmf_v3 V_l, nmf_v3 H_l, float r) {
  float r2 = fmaxf(r, NMF_EPS);
  float r1 = fmaxf(r + r2, NMF_EPS) / 2.0f;
  float lx = L_l.x * r1, ly = L_l.y * r2;
  float lam = (-1.0f + sqrtf(fmaxf(1.0f + (lx * lx + ly * ly) / fmaxf(L_l.z * L_l.z, 1e-6f), NMF_EPS))) / 2.0f;
  float invG = 1.0f + lam;
  float q = H_l.x * H_l.x / (r1 * r1) + H_l.y * H_l.y / (r2 * r2) + H_l.z * H_l.z;
  float invD = NMF_PI * r1 * r2 * (q * q);
  float logD = -logf(fmaxf(invG * invD, NMF_EPS)) - logf(fmaxf(4.0f * V_l.z, NMF_EPS));
  return L_l.z > 0.0f ? expf(logD) : 0.0f;
}
// The part of the sampler that only depends on the shaded sample (V, N, roughness), not on the bounce ray:
// tangent frame, view vector in the local / stretched frame, the orthonormal basis of the VNDF disk (ggx.py:83-140)
struct NmfGGXFrame { nmf_v3 t, b, V_l, Vs, T1, T2; float a; };
NMF_HD NmfGGXFrame nmf_ggx_frame(nmf_v3 V, nmf_v3 N, float r) {
  NmfGGXFrame f;
  nmf_v3 up = (fabsf(N.z) < 0.999f) ? nmf_mk3(0.f, 0.f, 1.f) : nmf_mk3(-1.f, 0.f, 0.f);
  f.t = nmf_unit(nmf_cross(up, N));
  f.b = nmf_unit(nmf_cross(N, f.t));
  f.V_l = nmf_mk3(nmf_dot(f.t, V), nmf_dot(f.b, V), nmf_dot(N, V));
  f.Vs = nmf_unit(nmf_mk3(r * f.V_l.x, r * f.V_l.y, f.V_l.z));
  f.T1 = (f.Vs.z < 0.999f) ? nmf_unit(nmf_cross(f.Vs, nmf_mk3(0.f, 0.f, 1.f))) : nmf_mk3(-1.f, 0.f, 0.f);
  f.T2 = nmf_unit(nmf_cross(f.T1, f.Vs));
  f.a = fminf(1.0f / fmaxf(1.0f + f.Vs.z, 1e-8f), 1e4f);
  return f;
}
NMF_HD NmfGGX nmf_ggx_sample_f(const NmfGGXFrame& f, float u1, float u2, nmf_v3 V, nmf_v3 N, float r) {
  NmfGGX g;
  const nmf_v3 t = f.t, b = f.b, V_l = f.V_l, Vs = f.Vs, T1 = f.T1, T2 = f.T2;
  const float a = f.a;
  float rad = sqrtf(u1);
  bool lo = u2 < a;
  float phi = lo ? (u2 / a * NMF_PI) : ((u2 - a) / (1.0f - a) * NMF_PI + NMF_PI);
  float pm = fmodf(phi, 100.0f * NMF_PI);
  float P1 = rad * cosf(pm);
  float P2 = rad * sinf(pm) * (lo ? 1.0f : Vs.z);
  float c = sqrtf(fmaxf(1.0f - P1 * P1 - P2 * P2, NMF_EPS));
  nmf_v3 Ns = nmf_add3(nmf_add3(nmf_scale(T1, P1), nmf_scale(T2, P2)), nmf_scale(Vs, c));
  nmf_v3 H_l = nmf_unit(nmf_mk3(Ns.x * r, Ns.y * r, Ns.z));
  nmf_v3 H = nmf_add3(nmf_add3(nmf_scale(t, H_l.x), nmf_scale(b, H_l.y)), nmf_scale(N, H_l.z));
  float vh = nmf_dot(V, H);
  nmf_v3 L = nmf_unit(nmf_mk3(2.0f * vh * H.x - V.x, 2.0f * vh * H.y - V.y, 2.0f * vh * H.z - V.z));
  if (!(nmf_dot(L, N) > 0.0f)) L = nmf_scale(L, -1.0f);
  nmf_v3 L_l = nmf_mk3(nmf_dot(t, L), nmf_dot(b, L), nmf_dot(N, L));
  g.logpdf = logf(fmaxf(nmf_ggx_pdf(L_l, V_l, H_l, r), NMF_EPS));
  g.L = L;
  g.diff_l = L_l;
  nmf_v3 H2 = nmf_unit(nmf_mk3((V.x + L.x) / 2.0f, (V.y + L.y) / 2.0f, (V.z + L.z) / 2.0f));
  g.H = H2;
  g.half_l = nmf_mk3(nmf_dot(t, H2), nmf_dot(b, H2), nmf_dot(N, H2));
  return g;
}
NMF_HD NmfGGX nmf_ggx_sample(float u1, float u2, nmf_v3 V, nmf_v3 N, float r) {
  return nmf_ggx_sample_f(nmf_ggx_frame(V, N, r), u1, u2, V, N, r);
}
// (u, v) of bounce ray j: brdf_samplers/base.py:11-20  (sobol[j] + 0.25 * U) mod 1
NMF_HD float nmf_wrap01(float x) { return x - floorf(x); }

// ------------------------------------------------------------------------------------------------
// A15 environment: modules/integral_equirect.py:373-504 (sa2mip, forward) and :18-173 (box integrals)
// `Tap` is a functor  void operator()(float px, float py, float sign, float* acc3)  that adds
// sign * bilinear(SAT, clip(p, -1, 1)) to acc3.
// ------------------------------------------------------------------------------------------------
struct NmfEnvBox { float cx, cy, sw, sh, size; };
NMF_HD NmfEnvBox nmf_env_box(nmf_v3 u, float sa, int h, int w, float mipbias) {
  NmfEnvBox bx;
  float cosv = sqrtf(fmaxf(1.0f - u.z * u.z, NMF_EPS));                                   // :376
  float d = (float)(h * w) / fmaxf((float)(2.0 * 3.14159265358979323846 * 3.14159265358979323846) * cosv, NMF_EPS);
  float area = expf(logf(d / 2.0f) + sa);                                                  // :381-383
  float hh = fmaxf(sqrtf(fmaxf(area, NMF_EPS)) * cosv, NMF_EPS);
  float ww = area / hh;
  const float ln2 = 0.6931471805599453f;
  float lw = nmf_clampf(logf(ww) / ln2 + mipbias, 0.0f, 7.0f);
  float lh = nmf_clampf(logf(hh) / ln2 + mipbias, 0.0f, 7.0f);
  bx.sw = exp2f(lw) / (float)h / 2.0f;                                                     // :463-464
  bx.sh = exp2f(lh) / (float)h;
  bx.size = (bx.sw / 2.0f * (float)w) * (bx.sh / 2.0f * (float)h);
  float norm2d = sqrtf(u.x * u.x + u.y * u.y);
  float phi = atan2f(u.y, u.x);
  float theta = atan2f(u.z, norm2d);
  const float twopi = 6.2831855f;
  float pm = phi - twopi * floorf(phi / twopi);                                            // torch.remainder
  bx.cx = (pm - 3.1415927f) / 3.1415927f;
  bx.cy = -theta / 3.1415927f * 2.0f;
  return bx;
}
// One axis-aligned box (x0,y0)-(x1,y1): integral_equirect.py:18-39   (tr + bl - tl - br) / size
template <class Tap>
NMF_HD void nmf_env_box1(Tap& tap, float x0, float y0, float x1, float y1, float inv_size, float* out) {
  float acc[3] = {0.f, 0.f, 0.f};
  tap(x1, y1, 1.0f, acc);
  tap(x0, y0, 1.0f, acc);
  tap(x0, y1, -1.0f, acc);
  tap(x1, y0, -1.0f, acc);
  out[0] += acc[0] * inv_size;
  out[1] += acc[1] * inv_size;
  out[2] += acc[2] * inv_size;
}
// The box, its pole overhangs (integral_equirect.py:96-173: mirrored across the pole and shifted by half a turn) and
// the left/right wrap-around pieces of each (:42-93): up to 9 boxes, summed in the reference's order.  Written as
// rolled loops around ONE box evaluation: most lookups are a single box, and nine inlined copies of the 16-tap body
// cost more in instruction fetch than the loop costs in control flow.
template <class Tap>
NMF_HD void nmf_env_integrate(Tap& tap, const NmfEnvBox& bx, float* out) {
  const float hx = bx.sw / 2.0f, hy = bx.sh / 2.0f;
  const float bx0 = bx.cx - hx, by0 = bx.cy - hy, bx1 = bx.cx + hx, by1 = bx.cy + hy;
  const float inv_size = 1.0f / bx.size;
  const float rot = bx0 > 0.0f ? -1.0f : 1.0f;
  out[0] = out[1] = out[2] = 0.0f;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int part = 0; part < 3; ++part) {
    float x0 = bx0, x1 = bx1, y0 = by0, y1 = by1;
    if (part == 1) {
      if (!(by1 > 1.0f)) continue;
      const float over = nmf_clampf(by1 - 1.0f, 0.0f, 0.5f);
      x0 = bx0 + rot; x1 = bx1 + rot; y0 = 1.0f - over; y1 = 1.0f;
    } else if (part == 2) {
      if (!(by0 < -1.0f)) continue;
      const float over = nmf_clampf(-1.0f - by0, 0.0f, 0.5f);
      x0 = bx0 + rot; x1 = bx1 + rot; y0 = -1.0f; y1 = -1.0f + over;
    }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int wrap = 0; wrap < 3; ++wrap) {
      float u0 = x0, u1 = x1;
      if (wrap == 1) {
        if (!(x1 > 1.0f)) continue;
        u0 = -1.0f; u1 = x1 - 2.0f;
      } else if (wrap == 2) {
        if (!(x0 < -1.0f)) continue;
        u0 = x0 + 2.0f; u1 = 1.0f;
      }
      nmf_env_box1(tap, u0, y0, u1, y1, inv_size, out);
    }
  }
}
// bilinear tap of the channel-last SAT ([h][w][4]) at clip(p, -1, 1), align_corners=True
struct NmfSatTap {
  const float* sat; int h, w;
  NMF_HD void operator()(float px, float py, float sign, float* acc) const {
    px = nmf_clampf(px, -1.0f, 1.0f);
    py = nmf_clampf(py, -1.0f, 1.0f);
    float ix = (px + 1.0f) * 0.5f * (float)(w - 1), iy = (py + 1.0f) * 0.5f * (float)(h - 1);
    float fx = floorf(ix), fy = floorf(iy);
    int x0 = (int)fx, y0 = (int)fy;
    float tx = ix - fx, ty = iy - fy;
    int x1 = x0 + 1 < w ? x0 + 1 : x0, y1 = y0 + 1 < h ? y0 + 1 : y0;   // the clamped tap has weight 0
    float w00 = (1.0f - tx) * (1.0f - ty), w10 = tx * (1.0f - ty), w01 = (1.0f - tx) * ty, w11 = tx * ty;
#ifdef __CUDA_ARCH__
    float4 a = __ldg((const float4*)sat + (size_t)y0 * w + x0), b = __ldg((const float4*)sat + (size_t)y0 * w + x1);
    float4 c = __ldg((const float4*)sat + (size_t)y1 * w + x0), d = __ldg((const float4*)sat + (size_t)y1 * w + x1);
    acc[0] += sign * (a.x * w00 + b.x * w10 + c.x * w01 + d.x * w11);
    acc[1] += sign * (a.y * w00 + b.y * w10 + c.y * w01 + d.y * w11);
    acc[2] += sign * (a.z * w00 + b.z * w10 + c.z * w01 + d.z * w11);
#else
    const float* a = sat + ((size_t)y0 * w + x0) * 4; const float* b = sat + ((size_t)y0 * w + x1) * 4;
    const float* c = sat + ((size_t)y1 * w + x0) * 4; const float* d = sat + ((size_t)y1 * w + x1) * 4;
    for (int k = 0; k < 3; ++k) acc[k] += sign * (a[k] * w00 + b[k] * w10 + c[k] * w01 + d[k] * w11);
#endif
  }
};
// The same tap from the PAIRED table ([h][w][8]: texel x and its right neighbour in one 32-byte record): one 256-bit load
// per row.  Same values, same arithmetic order as NmfSatTap -- only the number of load instructions (= L1 wavefronts per
// lane) is halved.
struct NmfSatTap2 {
  const float* sat8; int h, w;
  NMF_HD void operator()(float px, float py, float sign, float* acc) const {
    px = nmf_clampf(px, -1.0f, 1.0f);
    py = nmf_clampf(py, -1.0f, 1.0f);
    float ix = (px + 1.0f) * 0.5f * (float)(w - 1), iy = (py + 1.0f) * 0.5f * (float)(h - 1);
    float fx = floorf(ix), fy = floorf(iy);
    int x0 = (int)fx, y0 = (int)fy;
    float tx = ix - fx, ty = iy - fy;
    int y1 = y0 + 1 < h ? y0 + 1 : y0;                                   // the clamped tap has weight 0
    float w00 = (1.0f - tx) * (1.0f - ty), w10 = tx * (1.0f - ty), w01 = (1.0f - tx) * ty, w11 = tx * ty;
    const float* r0 = sat8 + ((size_t)y0 * w + x0) * 8;
    const float* r1 = sat8 + ((size_t)y1 * w + x0) * 8;
#ifdef __CUDA_ARCH__
    float a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2, c3, d0, d1, d2, d3;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "l"(r0));
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(c0), "=f"(c1), "=f"(c2), "=f"(c3), "=f"(d0), "=f"(d1), "=f"(d2), "=f"(d3) : "l"(r1));
    acc[0] += sign * (a0 * w00 + b0 * w10 + c0 * w01 + d0 * w11);
    acc[1] += sign * (a1 * w00 + b1 * w10 + c1 * w01 + d1 * w11);
    acc[2] += sign * (a2 * w00 + b2 * w10 + c2 * w01 + d2 * w11);
#else
    for (int k = 0; k < 3; ++k) acc[k] += sign * (r0[k] * w00 + r0[4 + k] * w10 + r1[k] * w01 + r1[4 + k] * w11);
#endif
  }
};
NMF_HD void nmf_env_lookup1_pair(const float* sat8, int h, int w, float mipbias, const float* top, const float* bot,
                                 nmf_v3 dir, float sa, float* rgb) {
  NmfEnvBox bx = nmf_env_box(dir, sa, h, w, mipbias);
  NmfSatTap2 tap; tap.sat8 = sat8; tap.h = h; tap.w = w;
  nmf_env_integrate(tap, bx, rgb);
  rgb[0] *= 1000.0f; rgb[1] *= 1000.0f; rgb[2] *= 1000.0f;
  float cutoff = 1.0f - 2.0f / (float)h * 3.0f;
  if (bx.cy > cutoff) { rgb[0] = bot[0]; rgb[1] = bot[1]; rgb[2] = bot[2]; }
  if (bx.cy < -cutoff) { rgb[0] = top[0]; rgb[1] = top[1]; rgb[2] = top[2]; }
}
// IntegralEquirect.forward for one direction (integral_equirect.py:409-504)
NMF_HD void nmf_env_lookup1(const float* sat, int h, int w, float mipbias, const float* top, const float* bot,
                            nmf_v3 dir, float sa, float* rgb) {
  NmfEnvBox bx = nmf_env_box(dir, sa, h, w, mipbias);
  NmfSatTap tap; tap.sat = sat; tap.h = h; tap.w = w;
  nmf_env_integrate(tap, bx, rgb);
  rgb[0] *= 1000.0f; rgb[1] *= 1000.0f; rgb[2] *= 1000.0f;
  float cutoff = 1.0f - 2.0f / (float)h * 3.0f;
  if (bx.cy > cutoff) { rgb[0] = bot[0]; rgb[1] = bot[1]; rgb[2] = bot[2]; }
  if (bx.cy < -cutoff) { rgb[0] = top[0]; rgb[1] = top[1]; rgb[2] = top[2]; }
}

// modules/tonemap.py:38-49
NMF_HD float nmf_srgb(float x) {
  const float limit = 0.0031308f;
  return x > limit ? 1.055f * powf(fmaxf(x, limit), 1.0f / 2.4f) - 0.055f : 12.92f * x;
}
// Schlick Fresnel: models/microfacet.py:585-590, 639-641
NMF_HD float nmf_fresnel(float f0, float cost) {
  float m = nmf_clampf(1.0f - cost, 0.0f, 1.0f);
  float m2 = m * m;
  return f0 + (1.0f - f0) * (m2 * m2 * m);
}
