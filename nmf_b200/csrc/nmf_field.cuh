// VM-decomposed field queries (fields/tensoRF.py:161-205, 392-405; fields/tensor_base.py:66-129) on the
// channel-last layouts of NmfScene.  Written per *channel group* (4 channels = one 16-byte load per tap) so that
// the kernels can spread the groups of one sample over neighbouring lanes; the host check loops over the groups.
#pragma once
#include "../../include/nmf_b200.h"
#include "nmf_math.cuh"

// The three environment quantities that change with every optimiser step (mipbias and the pole-row means) are read from
// device memory when NmfScene.env_dyn is set, so that a training iteration does not need them on the host.  They are
// loaded ONCE per kernel into a local struct (forming pointers into the by-value kernel parameter cost 10-15 % on the
// environment kernels: the parameter block had to be addressable)
struct NmfEnvDyn { float mipbias, top[3], bot[3]; };
NMF_HD NmfEnvDyn nmf_env_dyn_load(const NmfScene& s) {
  NmfEnvDyn e;
  if (s.env_dyn) {
    e.mipbias = s.env_dyn[0];
    e.top[0] = s.env_dyn[1]; e.top[1] = s.env_dyn[2]; e.top[2] = s.env_dyn[3];
    e.bot[0] = s.env_dyn[4]; e.bot[1] = s.env_dyn[5]; e.bot[2] = s.env_dyn[6];
  } else {
    e.mipbias = s.env_mipbias;
    e.top[0] = s.env_top[0]; e.top[1] = s.env_top[1]; e.top[2] = s.env_top[2];
    e.bot[0] = s.env_bot[0]; e.bot[1] = s.env_bot[1]; e.bot[2] = s.env_bot[2];
  }
  return e;
}

#ifdef __CUDACC__
typedef float4 nmf_f4;
#else
struct nmf_f4 { float x, y, z, w; };
#endif
NMF_HD nmf_f4 nmf_ld4(const float* p) {
#ifdef __CUDA_ARCH__
  return __ldg((const float4*)p);     // read-only path, one 16-byte load
#else
  return *(const nmf_f4*)p;
#endif
}
#define NMF_LD4(p) nmf_ld4(p)
// two adjacent 16-byte groups (32-byte aligned) as ONE 256-bit load (LDG.E.256): half the load instructions, and with them
// half the L1 data-pipe wavefronts, of two nmf_ld4 calls -- the gather kernels are bound by that pipe
NMF_HD void nmf_ld8(const float* p, nmf_f4& a, nmf_f4& b) {
#ifdef __CUDA_ARCH__
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
  a = *(const nmf_f4*)p; b = *(const nmf_f4*)(p + 4);
#endif
}

NMF_HD nmf_f4 nmf_f4_zero() { nmf_f4 r; r.x = r.y = r.z = r.w = 0.f; return r; }
NMF_HD void nmf_f4_fma(nmf_f4& a, nmf_f4 v, float w) { a.x += v.x * w; a.y += v.y * w; a.z += v.z * w; a.w += v.w * w; }
NMF_HD float nmf_f4_dot(nmf_f4 a, nmf_f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// the same for records that are read exactly once (bounce-ray records): evict-first, so that they do not displace the
// summed-area table and the factor planes in L1 / L2
NMF_HD void nmf_ld8_stream(const float* p, nmf_f4& a, nmf_f4& b) {
#ifdef __CUDA_ARCH__
  asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
  a = *(const nmf_f4*)p; b = *(const nmf_f4*)(p + 4);
#endif
}
// two 16-byte groups to a 32-byte aligned address as ONE 256-bit store
NMF_HD void nmf_st8(float* p, nmf_f4 a, nmf_f4 b) {
#ifdef __CUDA_ARCH__
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
#else
  *(nmf_f4*)p = a; *(nmf_f4*)(p + 4) = b;
#endif
}
// (plain-cached variant of nmf_ld8 for data written earlier in the same launch sequence is not needed: records are
// produced by one kernel and consumed by the next, so the read-only path is safe)

// matMode / vecMode of fields/tensoRF.py:40-41
#define NMF_MAT0(p) ((p) == 2 ? 1 : 0)
#define NMF_MAT1(p) ((p) == 0 ? 1 : 2)
#define NMF_VEC(p) (2 - (p))

struct NmfTaps { NmfLerp px[3], py[3], pl[3]; };

NMF_HD void nmf_normalize_xyz(const NmfScene& s, const float* p, float* xn) {   // tensor_base.py:66-69
  xn[0] = nmf_norm_coord(p[0], s.aabb0[0], s.inv_aabb2[0]);
  xn[1] = nmf_norm_coord(p[1], s.aabb0[1], s.inv_aabb2[1]);
  xn[2] = nmf_norm_coord(p[2], s.aabb0[2], s.inv_aabb2[2]);
}
// Conservative coarse occupancy (NmfScene.occ_coarse, here already in shared memory): false => no voxel within one
// cell of the sample's 8-cell block is set, so the exact trilinear test (nmf_occupied) is false.  The cell index comes
// from one multiply (error ~1e-5 cells); the field is dilated by a whole voxel on each side, so a floor that lands one
// cell off is still covered.
NMF_HD bool nmf_occ_coarse(const NmfScene& s, const uint32_t* coarse, const float* p) {
  int x = (int)((p[0] - s.aabb0[0]) * s.occ_scale[0]), y = (int)((p[1] - s.aabb0[1]) * s.occ_scale[1]);
  int z = (int)((p[2] - s.aabb0[2]) * s.occ_scale[2]);
  x = x < 0 ? 0 : (x >= s.ow ? s.ow - 1 : x);
  y = y < 0 ? 0 : (y >= s.oh ? s.oh - 1 : y);
  z = z < 0 ? 0 : (z >= s.od ? s.od - 1 : z);
  const unsigned i = ((unsigned)(z >> 3) * (unsigned)s.och + (unsigned)(y >> 3)) * (unsigned)s.ocw + (unsigned)(x >> 3);
  return (coarse[i >> 5] >> (i & 31)) & 1u;
}

NMF_HD NmfTaps nmf_vm_taps(const NmfScene& s, const float* xn) {                // tensoRF.py:161-179
  // every factor that is indexed by axis a has grid[a] texels along it (plane p: width grid[mat0(p)], height
  // grid[mat1(p)]; line p: grid[vec(p)]; checked by the host before any launch), so the nine bilinear set-ups of
  // the three plane/line pairs are three set-ups, one per axis
  const NmfLerp lx = nmf_lerp_setup(xn[0], s.plane_w[0]);
  const NmfLerp ly = nmf_lerp_setup(xn[1], s.plane_h[0]);
  const NmfLerp lz = nmf_lerp_setup(xn[2], s.plane_h[1]);
  NmfTaps t;
  t.px[0] = lx; t.py[0] = ly; t.pl[0] = lz;      // plane (x, y), line z
  t.px[1] = lx; t.py[1] = lz; t.pl[1] = ly;      // plane (x, z), line y
  t.px[2] = ly; t.py[2] = lz; t.pl[2] = lx;      // plane (y, z), line x
  return t;
}
// bilinear tap of a channel-last plane: `stride` floats per texel, `off` float offset of this lane's 4 channels
NMF_HD nmf_f4 nmf_bilerp4(const float* plane, int w, int stride, int off, const NmfLerp& lx, const NmfLerp& ly) {
  const float* r0 = plane + (size_t)ly.i0 * w * stride + off;
  const float* r1 = plane + (size_t)ly.i1 * w * stride + off;
  nmf_f4 a = NMF_LD4(r0 + (size_t)lx.i0 * stride), b = NMF_LD4(r0 + (size_t)lx.i1 * stride);
  nmf_f4 c = NMF_LD4(r1 + (size_t)lx.i0 * stride), d = NMF_LD4(r1 + (size_t)lx.i1 * stride);
  nmf_f4 o = nmf_f4_zero();
  nmf_f4_fma(o, a, ly.w0 * lx.w0);
  nmf_f4_fma(o, b, ly.w0 * lx.w1);
  nmf_f4_fma(o, c, ly.w1 * lx.w0);
  nmf_f4_fma(o, d, ly.w1 * lx.w1);
  return o;
}
NMF_HD nmf_f4 nmf_lerp4(const float* line, int stride, int off, const NmfLerp& l) {
  nmf_f4 a = NMF_LD4(line + (size_t)l.i0 * stride + off), b = NMF_LD4(line + (size_t)l.i1 * stride + off);
  nmf_f4 o = nmf_f4_zero();
  nmf_f4_fma(o, a, l.w0);
  nmf_f4_fma(o, b, l.w1);
  return o;
}

// density feature, channels 4g..4g+3 of all three plane/line pairs (tensoRF.py:392-400, dbasis = False)
NMF_HD float nmf_density_group(const NmfScene& s, const NmfTaps& t, int g) {
  float acc = 0.f;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    nmf_f4 pv = nmf_bilerp4(s.dval[p], s.plane_w[p], 16, 4 * g, t.px[p], t.py[p]);
    nmf_f4 lv = nmf_lerp4(s.lval[p], 16, 4 * g, t.pl[p]);
    acc += nmf_f4_dot(pv, lv);
  }
  return acc;
}
// appearance coefficients, channels 4g..4g+3 (g < 6) of plane p (tensoRF.py:402-405 before basis_mat)
NMF_HD nmf_f4 nmf_app_group(const NmfScene& s, const NmfTaps& t, int p, int g) {
  nmf_f4 pv = nmf_bilerp4(s.aval[p], s.plane_w[p], NMF_APP_STRIDE, 4 * g, t.px[p], t.py[p]);
  nmf_f4 lv = nmf_lerp4(s.alval[p], NMF_APP_STRIDE, 4 * g, t.pl[p]);
  nmf_f4 o; o.x = pv.x * lv.x; o.y = pv.y * lv.y; o.z = pv.z * lv.z; o.w = pv.w * lv.w;
  return o;
}
// x-taps of a bilinear lookup as a PAIR of adjacent texels (base, base + 1) with weights (wa, wb), so that the two
// texels of a row are one contiguous span of memory.  Out-of-range taps keep weight 0 (zeros padding); at the right
// border (i0 = size - 1, weight-0 i1) the pair slides to (size - 2, size - 1).
struct NmfPair { int base; float wa, wb; };
NMF_HD NmfPair nmf_pair_setup(const NmfLerp& l, int size) {
  NmfPair q;
  q.base = l.i0 < size - 2 ? l.i0 : size - 2;
  q.wa = (l.i0 == q.base ? l.w0 : 0.f) + (l.i1 == q.base ? l.w1 : 0.f);
  q.wb = (l.i0 == q.base + 1 ? l.w0 : 0.f) + (l.i1 == q.base + 1 ? l.w1 : 0.f);
  return q;
}
// gradient of the density feature wrt normalised coordinates (tensor_base.py:107-129 with the smoothed-difference
// backward of grid_sample_Cinf.py:218-281), the share of lane `l` of an 8-lane group; summing over l = 0..7 gives
// d(feature)/d(xn).  dpack is [h][w][val16 | dx16 | dy16]: the two x-taps of a row are 96 contiguous floats that the
// 8 lanes read as three 128-byte pieces (lane l takes bytes 16 l .. 16 l + 15 of a piece):
//   piece 0 = tex0.val | tex0.dx     piece 1 = tex0.dy | tex1.val     piece 2 = tex1.dx | tex1.dy
// so lane (g = l & 3, hi = l >> 2) sees channel group g of (val, dy, dx) [hi = 0] or (dx, val, dy) [hi = 1].
NMF_HD void nmf_normal_lane(const NmfScene& s, const NmfTaps& t, int l, float* grad) {
  const int g = l & 3;
  const bool hi = l >= 4;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const int w = s.plane_w[p];
    const NmfLerp& ly = t.py[p];
    const NmfLerp& ll = t.pl[p];
    const NmfPair px = nmf_pair_setup(t.px[p], w);
    // line value and smoothed derivative of channel group g (both taps)
    const float* l0 = s.lpack[p] + (size_t)ll.i0 * 32 + 8 * g;
    const float* l1 = s.lpack[p] + (size_t)ll.i1 * 32 + 8 * g;
    nmf_f4 lv = nmf_f4_zero(), ld = nmf_f4_zero();
    {
      nmf_f4 v0, d0, v1, d1;                       // (val4 | dy4) of a line texel are 32 contiguous, 32-byte aligned bytes
      nmf_ld8(l0, v0, d0);
      nmf_ld8(l1, v1, d1);
      nmf_f4_fma(lv, v0, ll.w0); nmf_f4_fma(ld, d0, ll.w0);
      nmf_f4_fma(lv, v1, ll.w1); nmf_f4_fma(ld, d1, ll.w1);
    }
    const float* r0 = s.dpack[p] + ((size_t)ly.i0 * w + px.base) * 48 + 4 * l;
    const float* r1 = s.dpack[p] + ((size_t)ly.i1 * w + px.base) * 48 + 4 * l;
    nmf_f4 a0 = nmf_f4_zero(), a1 = nmf_f4_zero(), a2 = nmf_f4_zero();
    nmf_f4_fma(a0, NMF_LD4(r0), ly.w0); nmf_f4_fma(a1, NMF_LD4(r0 + 32), ly.w0); nmf_f4_fma(a2, NMF_LD4(r0 + 64), ly.w0);
    nmf_f4_fma(a0, NMF_LD4(r1), ly.w1); nmf_f4_fma(a1, NMF_LD4(r1 + 32), ly.w1); nmf_f4_fma(a2, NMF_LD4(r1 + 64), ly.w1);
    const float e0 = px.wa * nmf_f4_dot(a0, hi ? lv : ld);                 // tex0: dx [hi] or val
    const float e1 = (hi ? px.wb : px.wa) * nmf_f4_dot(a1, hi ? ld : lv);  // tex1.val [hi] or tex0.dy
    const float e2 = px.wb * nmf_f4_dot(a2, lv);                           // tex1: dy [hi] or dx
    grad[NMF_VEC(p)] += hi ? e1 : e0;
    grad[NMF_MAT0(p)] += hi ? e0 : e2;
    grad[NMF_MAT1(p)] += hi ? e2 : e1;
  }
}
// n = normalize(-(g * invaabbSize))  (tensor_base.py:126-128)
NMF_HD nmf_v3 nmf_normal_from_grad(const NmfScene& s, const float* grad) {
  return nmf_unit(nmf_mk3(-(grad[0] * s.inv_aabb2[0]), -(grad[1] * s.inv_aabb2[1]), -(grad[2] * s.inv_aabb2[2])));
}
