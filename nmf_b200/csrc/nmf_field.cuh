// VM-decomposed field queries (fields/tensoRF.py:161-205, 392-405; fields/tensor_base.py:66-129) on the
// channel-last layouts of NmfScene.  Written per *channel group* (4 channels = one 16-byte load per tap) so that
// the kernels can spread the groups of one sample over neighbouring lanes; the host check loops over the groups.
#pragma once
#include "../../include/nmf_b200.h"
#include "nmf_math.cuh"

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

NMF_HD nmf_f4 nmf_f4_zero() { nmf_f4 r; r.x = r.y = r.z = r.w = 0.f; return r; }
NMF_HD void nmf_f4_fma(nmf_f4& a, nmf_f4 v, float w) { a.x += v.x * w; a.y += v.y * w; a.z += v.z * w; a.w += v.w * w; }
NMF_HD float nmf_f4_dot(nmf_f4 a, nmf_f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

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
NMF_HD NmfTaps nmf_vm_taps(const NmfScene& s, const float* xn) {                // tensoRF.py:161-179
  NmfTaps t;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    t.px[p] = nmf_lerp_setup(xn[NMF_MAT0(p)], s.plane_w[p]);
    t.py[p] = nmf_lerp_setup(xn[NMF_MAT1(p)], s.plane_h[p]);
    t.pl[p] = nmf_lerp_setup(xn[NMF_VEC(p)], s.line_n[p]);
  }
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
  nmf_f4 pv = nmf_bilerp4(s.aval[p], s.plane_w[p], 24, 4 * g, t.px[p], t.py[p]);
  nmf_f4 lv = nmf_lerp4(s.alval[p], 24, 4 * g, t.pl[p]);
  nmf_f4 o; o.x = pv.x * lv.x; o.y = pv.y * lv.y; o.z = pv.z * lv.z; o.w = pv.w * lv.w;
  return o;
}
// gradient of the density feature wrt normalised coordinates, restricted to channel group g and to the plane rows
// y-tap `half` (0: row i0, 1: row i1) / the matching line tap: tensor_base.py:107-129 with the smoothed-difference
// backward of grid_sample_Cinf.py:218-281.  Summing over g in 0..3 and half in 0..1 gives d(feature)/d(xn).
NMF_HD void nmf_normal_group(const NmfScene& s, const NmfTaps& t, int g, int half, float* grad) {
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const int w = s.plane_w[p];
    const NmfLerp& lx = t.px[p];
    const NmfLerp& ly = t.py[p];
    const NmfLerp& ll = t.pl[p];
    int yi = half ? ly.i1 : ly.i0;
    float wy = half ? ly.w1 : ly.w0;
    const float* r = s.dpack[p] + ((size_t)yi * w) * 48 + 12 * g;
    const float* ta = r + (size_t)lx.i0 * 48;
    const float* tb = r + (size_t)lx.i1 * 48;
    float wa = wy * lx.w0, wb = wy * lx.w1;
    nmf_f4 val = nmf_f4_zero(), dx = nmf_f4_zero(), dy = nmf_f4_zero();
    nmf_f4_fma(val, NMF_LD4(ta), wa); nmf_f4_fma(dx, NMF_LD4(ta + 4), wa); nmf_f4_fma(dy, NMF_LD4(ta + 8), wa);
    nmf_f4_fma(val, NMF_LD4(tb), wb); nmf_f4_fma(dx, NMF_LD4(tb + 4), wb); nmf_f4_fma(dy, NMF_LD4(tb + 8), wb);
    // full line value / derivative (both taps): cheap, and keeps the two halves symmetric
    const float* l0 = s.lpack[p] + (size_t)ll.i0 * 32 + 8 * g;
    const float* l1 = s.lpack[p] + (size_t)ll.i1 * 32 + 8 * g;
    nmf_f4 lv = nmf_f4_zero(), ld = nmf_f4_zero();
    nmf_f4_fma(lv, NMF_LD4(l0), ll.w0); nmf_f4_fma(ld, NMF_LD4(l0 + 4), ll.w0);
    nmf_f4_fma(lv, NMF_LD4(l1), ll.w1); nmf_f4_fma(ld, NMF_LD4(l1 + 4), ll.w1);
    grad[NMF_MAT0(p)] += nmf_f4_dot(lv, dx);
    grad[NMF_MAT1(p)] += nmf_f4_dot(lv, dy);
    grad[NMF_VEC(p)] += nmf_f4_dot(val, ld);
  }
}
// n = normalize(-(g * invaabbSize))  (tensor_base.py:126-128)
NMF_HD nmf_v3 nmf_normal_from_grad(const NmfScene& s, const float* grad) {
  return nmf_unit(nmf_mk3(-(grad[0] * s.inv_aabb2[0]), -(grad[1] * s.inv_aabb2[1]), -(grad[2] * s.inv_aabb2[2])));
}
