// Per-element backward math of the MICROFACET model (remainder): every piece of the
// reverse pass stated per bounce ray / per sample / per texel, compiled for the host (tests/hostcheck) and checked against
// the oracle's autograd (= the reference's gradient, oracle/check_train.py) in the CPU suite -- piece by piece and composed
// into the whole two-level training reverse pass (hc_train_microfacet, hc_train_microfacet_retrace).  The first kernels that
// include it: csrc/nmf_env_bwd.cu (map / mipbias gradient), csrc/nmf_normals_bwd.cu, csrc/nmf_shade_bwd.cu (material heads).
//
//   nmf_ggx_sample_dr     d L / d roughness and d H / d roughness of the GGX VNDF sample (brdf_samplers/ggx.py:61-226):
//                         forward-mode (dual-number) restatement of nmf_ggx_frame + nmf_ggx_sample_f.  The reference
//                         detaches `a = 1 / (1 + Vs.z)` (ggx.py:116) and evaluates the pdf under no_grad (ggx.py:218),
//                         so the mip level of the environment lookup carries NO roughness gradient.
//   nmf_fresnel_mix_bwd   comb = F * L_in * brdf + (1 - F) * diffuse, F = R0 + (1 - R0) (1 - |v.h|)^5
//                         (models/microfacet.py:565-613)
//   nmf_heads_bwd         the sigmoid material heads on the 24-d feature with their clip gates
//                         (modules/render_modules.py:553-560)
//   nmf_env_lookup1_bwd_map  gradient of an environment lookup w.r.t. the map: scatter into a SAT-shaped image with the
//                         forward's own box walk (modules/integral_equirect.py:409-504); the adjoint of the double
//                         cumsum and the activation chain are whole-map passes (stated in the test)
//   nmf_env_lookup1_d     directional derivative of a lookup along a tangent of the direction (d L / d roughness), forward-mode
//   nmf_brdf_row_fwd_bwd  the 66-64-64-4 BRDF MLP, one row;  nmf_bounce_sample_bwd  the composed reverse pass of a bounce sample
//   nmf_normal_vec_bwd / nmf_normal_bwd  the normal path that opens with detach_N off (scatter into dpack / lpack-shaped images)
// tests/hostcheck hc_train_microfacet composes all of it into the training reverse pass of one shading level; its gradients
// equal the oracle's for every parameter (tests/test_hostmath.py::test_train_microfacet_host_gradients).
#pragma once
#include "nmf_train.cuh"

// ---- minimal forward-mode AD: value + derivative w.r.t. ONE scalar (the sample's roughness) ----
struct NmfDual { float v, d; };
NMF_HD NmfDual nmf_dk(float c) { NmfDual r; r.v = c; r.d = 0.f; return r; }
NMF_HD NmfDual nmf_dmk(float v, float d) { NmfDual r; r.v = v; r.d = d; return r; }
NMF_HD NmfDual operator+(NmfDual a, NmfDual b) { return nmf_dmk(a.v + b.v, a.d + b.d); }
NMF_HD NmfDual operator-(NmfDual a, NmfDual b) { return nmf_dmk(a.v - b.v, a.d - b.d); }
NMF_HD NmfDual operator*(NmfDual a, NmfDual b) { return nmf_dmk(a.v * b.v, a.d * b.v + a.v * b.d); }
NMF_HD NmfDual operator/(NmfDual a, NmfDual b) { return nmf_dmk(a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)); }
NMF_HD NmfDual operator*(float a, NmfDual b) { return nmf_dmk(a * b.v, a * b.d); }
NMF_HD NmfDual operator*(NmfDual a, float b) { return nmf_dmk(a.v * b, a.d * b); }
// sqrt(max(x, floor)): torch.clip passes no gradient below the floor
NMF_HD NmfDual nmf_dsqrt_floor(NmfDual x, float floor_) {
  if (x.v > floor_) { const float s = sqrtf(x.v); return nmf_dmk(s, x.d / (2.0f * s)); }
  return nmf_dk(sqrtf(floor_));
}
struct NmfDual3 { NmfDual x, y, z; };
NMF_HD NmfDual3 nmf_d3(NmfDual x, NmfDual y, NmfDual z) { NmfDual3 r; r.x = x; r.y = y; r.z = z; return r; }
NMF_HD NmfDual3 nmf_d3k(nmf_v3 v) { return nmf_d3(nmf_dk(v.x), nmf_dk(v.y), nmf_dk(v.z)); }
NMF_HD NmfDual nmf_ddot(NmfDual3 a, NmfDual3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
NMF_HD NmfDual3 nmf_dcross(NmfDual3 a, NmfDual3 b) {
  return nmf_d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
NMF_HD NmfDual3 nmf_dscale(NmfDual3 a, NmfDual s) { return nmf_d3(a.x * s, a.y * s, a.z * s); }
NMF_HD NmfDual3 nmf_dadd(NmfDual3 a, NmfDual3 b) { return nmf_d3(a.x + b.x, a.y + b.y, a.z + b.z); }
NMF_HD NmfDual3 nmf_dunit(NmfDual3 v) {                      // v / sqrt(max(|v|^2, eps))  (nmf_unit)
  const NmfDual n = nmf_dsqrt_floor(nmf_ddot(v, v), NMF_EPS);
  return nmf_d3(v.x / n, v.y / n, v.z / n);
}

struct NmfGGXdr {
  nmf_v3 L, dL;        // outgoing direction and d L / d roughness
  nmf_v3 H, dH;        // normalize((V + L) / 2) (models/microfacet.py:388) and its derivative
};
// V: unit vector to the viewer, N: normal flipped to V's side, r: roughness of the sample (r2 = r1).  N and r carry ONE
// tangent: seed r = (r, 1), N constant for d / d roughness (N is detached while Microfacet.detach_N is on,
// microfacet.py:352-353); seed N = (N, e_c), r constant for the c-th column of d / d N once detach_N is off.
NMF_HD NmfGGXdr nmf_ggx_sample_dual3(float u1, float u2, NmfDual3 Vd, NmfDual3 N, NmfDual r) {
  const NmfDual3 up = nmf_d3k((fabsf(N.z.v) < 0.999f) ? nmf_mk3(0.f, 0.f, 1.f) : nmf_mk3(-1.f, 0.f, 0.f));
  const NmfDual3 t = nmf_dunit(nmf_dcross(up, N));
  const NmfDual3 b = nmf_dunit(nmf_dcross(N, t));
  const NmfDual3 V_l = nmf_d3(nmf_ddot(t, Vd), nmf_ddot(b, Vd), nmf_ddot(N, Vd));
  const NmfDual3 Vs = nmf_dunit(nmf_d3(r * V_l.x, r * V_l.y, V_l.z));
  const NmfDual3 zup = nmf_d3k(nmf_mk3(0.f, 0.f, 1.f));
  const NmfDual3 T1 = (Vs.z.v < 0.999f) ? nmf_dunit(nmf_dcross(Vs, zup)) : nmf_d3k(nmf_mk3(-1.f, 0.f, 0.f));
  const NmfDual3 T2 = nmf_dunit(nmf_dcross(T1, Vs));
  const float a = fminf(1.0f / fmaxf(1.0f + Vs.z.v, 1e-8f), 1e4f);          // detached (ggx.py:116)
  const float rad = sqrtf(u1);
  const bool lo = u2 < a;
  const float phi = lo ? (u2 / a * NMF_PI) : ((u2 - a) / (1.0f - a) * NMF_PI + NMF_PI);
  const float pm = fmodf(phi, 100.0f * NMF_PI);
  const NmfDual P1 = nmf_dk(rad * cosf(pm));
  const NmfDual P2 = lo ? nmf_dk(rad * sinf(pm)) : (rad * sinf(pm)) * Vs.z;
  const NmfDual c = nmf_dsqrt_floor(nmf_dk(1.0f) - P1 * P1 - P2 * P2, NMF_EPS);
  const NmfDual3 Ns = nmf_dadd(nmf_dadd(nmf_dscale(T1, P1), nmf_dscale(T2, P2)), nmf_dscale(Vs, c));
  const NmfDual3 H_l = nmf_dunit(nmf_d3(Ns.x * r, Ns.y * r, Ns.z));
  const NmfDual3 Hs = nmf_dadd(nmf_dadd(nmf_dscale(t, H_l.x), nmf_dscale(b, H_l.y)), nmf_dscale(N, H_l.z));
  const NmfDual vh = nmf_ddot(Vd, Hs);
  NmfDual3 L = nmf_dunit(nmf_d3(2.0f * vh * Hs.x - Vd.x, 2.0f * vh * Hs.y - Vd.y, 2.0f * vh * Hs.z - Vd.z));
  if (!(L.x.v * N.x.v + L.y.v * N.y.v + L.z.v * N.z.v > 0.0f)) L = nmf_dscale(L, nmf_dk(-1.0f));
  const NmfDual3 H2 = nmf_dunit(nmf_d3((Vd.x + L.x) * 0.5f, (Vd.y + L.y) * 0.5f, (Vd.z + L.z) * 0.5f));
  NmfGGXdr o;
  o.L = nmf_mk3(L.x.v, L.y.v, L.z.v);
  o.dL = nmf_mk3(L.x.d, L.y.d, L.z.d);
  o.H = nmf_mk3(H2.x.v, H2.y.v, H2.z.v);
  o.dH = nmf_mk3(H2.x.d, H2.y.d, H2.z.d);
  return o;
}
NMF_HD NmfGGXdr nmf_ggx_sample_dual(float u1, float u2, nmf_v3 V, NmfDual3 N, NmfDual r) {
  return nmf_ggx_sample_dual3(u1, u2, nmf_d3k(V), N, r);
}
// c-th column of d / d V: the tangent a RE-TRACED ray's shading sees (its view vector is minus the parent's bounce direction,
// which moves with the parent's roughness / normal; positions are detached in the field, tensoRF.py:182-183)
NMF_HD NmfGGXdr nmf_ggx_sample_dV(float u1, float u2, nmf_v3 V, nmf_v3 N, float r, int c) {
  NmfDual3 Vd = nmf_d3k(V);
  if (c == 0) Vd.x.d = 1.0f; else if (c == 1) Vd.y.d = 1.0f; else Vd.z.d = 1.0f;
  return nmf_ggx_sample_dual3(u1, u2, Vd, nmf_d3k(N), nmf_dk(r));
}
NMF_HD NmfGGXdr nmf_ggx_sample_dr(float u1, float u2, nmf_v3 V, nmf_v3 N, float r) {
  return nmf_ggx_sample_dual(u1, u2, V, nmf_d3k(N), nmf_dmk(r, 1.0f));
}
// c-th column of d / d N (c = 0, 1, 2)
NMF_HD NmfGGXdr nmf_ggx_sample_dN(float u1, float u2, nmf_v3 V, nmf_v3 N, float r, int c) {
  NmfDual3 Nd = nmf_d3k(N);
  if (c == 0) Nd.x.d = 1.0f; else if (c == 1) Nd.y.d = 1.0f; else Nd.z.d = 1.0f;
  return nmf_ggx_sample_dual(u1, u2, V, Nd, nmf_dk(r));
}

// comb_c = F_c * inc_c * bw_c + (1 - F_c) * diff_c,  F_c = R0_c + (1 - R0_c) * m^5,  m = clip(1 - cost, 0, 1),
// cost = |v . h|  (models/microfacet.py:584-600).  g = d loss / d comb.  Returns d loss / d cost.
NMF_HD float nmf_fresnel_mix_bwd(const float* R0, float cost, const float* inc, const float* bw, const float* diff, const float* g,
                                 float* dR0, float* dinc, float* dbw, float* ddiff) {
  const float m = nmf_clampf(1.0f - cost, 0.0f, 1.0f);
  const float m2 = m * m, m5 = m2 * m2 * m, m4 = m2 * m2;
  const bool open = (1.0f - cost) >= 0.0f && (1.0f - cost) <= 1.0f;       // torch.clip passes the gradient on the closed interval
  float dcost = 0.f;
  for (int c = 0; c < 3; ++c) {
    const float F = R0[c] + (1.0f - R0[c]) * m5;
    const float dF = g[c] * (inc[c] * bw[c] - diff[c]);
    dR0[c] = dF * (1.0f - m5);
    dinc[c] = g[c] * F * bw[c];
    dbw[c] = g[c] * F * inc[c];
    ddiff[c] = g[c] * (1.0f - F);
    if (open) dcost -= dF * (1.0f - R0[c]) * 5.0f * m4;
  }
  return dcost;
}

// Material heads (render_modules.py:553-560) on one sample: albedo_c = clip(sigmoid(mul * lin_c + bias_d), 0, 1),
// f0_c = sigmoid(lin_{6+c} + bias_f), r = clip(sigmoid(lin_9 + bias_r) / 2, 1e-2, 1) with lin = W feat + b
// (W rows: diffuse 0..2, tint 3..5, f0 6..8, roughness 9..10).  Upstream: g_albedo[3], g_f0[3], g_rough.
// Accumulates dW (11 x 24), db (11) and writes dfeat (24).  The tint head and r2 feed nothing on this path.
// d loss / d lin (the 11 pre-activations) of one sample; shared by nmf_heads_bwd and k_heads_bwd (csrc/nmf_shade_bwd.cu)
NMF_HD void nmf_heads_dlin(const float* feat, const float* W, const float* b, float diffuse_mul, float diffuse_bias, float f0_bias,
                           float roughness_bias, const float* g_albedo, const float* g_f0, float g_rough, float* dlin) {
  for (int h = 0; h < 11; ++h) dlin[h] = 0.f;
  float lin[11];
  for (int h = 0; h < 11; ++h) {
    float v = b[h];
    for (int k = 0; k < 24; ++k) v += W[h * 24 + k] * feat[k];
    lin[h] = v;
  }
  for (int c = 0; c < 3; ++c) {
    const float sa = nmf_sigmoid(diffuse_mul * lin[c] + diffuse_bias);
    dlin[c] = (sa >= 0.f && sa <= 1.f) ? g_albedo[c] * sa * (1.0f - sa) * diffuse_mul : 0.f;
    const float sf = nmf_sigmoid(lin[6 + c] + f0_bias);
    dlin[6 + c] = g_f0[c] * sf * (1.0f - sf);
  }
  const float sr = nmf_sigmoid(lin[9] + roughness_bias);
  dlin[9] = (sr / 2.0f >= 1e-2f && sr / 2.0f <= 1.0f) ? g_rough * 0.5f * sr * (1.0f - sr) : 0.f;
}
NMF_HD void nmf_heads_bwd(const float* feat, const float* W, const float* b, float diffuse_mul, float diffuse_bias, float f0_bias,
                          float roughness_bias, const float* g_albedo, const float* g_f0, float g_rough, float* dW, float* db,
                          float* dfeat) {
  float dlin[11];
  nmf_heads_dlin(feat, W, b, diffuse_mul, diffuse_bias, f0_bias, roughness_bias, g_albedo, g_f0, g_rough, dlin);
  for (int k = 0; k < 24; ++k) dfeat[k] = 0.f;
  for (int h = 0; h < 11; ++h) {
    if (dlin[h] == 0.f) continue;
    db[h] += dlin[h];
    for (int k = 0; k < 24; ++k) {
      dW[h * 24 + k] += dlin[h] * feat[k];
      dfeat[k] += dlin[h] * W[h * 24 + k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Environment lookup, gradient w.r.t. the map (modules/integral_equirect.py:263-273, 409-504).
//   forward:  act = exp(min(brightness + mul * bg_mat, 20)),  SAT = cumsum_y cumsum_x (act / 1000),
//             rgb = 1000 * sum_boxes sign * bilinear(SAT, corner) / size        (pole rows: the mean of act's first / last row)
//   backward: (1) every lookup scatters g * sign * w_bilinear / size into a SAT-shaped gradient image -- the SAME box walk
//                 as the forward (nmf_env_integrate) with a scattering tap; pole lookups add g to the row-mean gradient;
//             (2) d act = reverse cumsum over x and y of that image (the 1/1000 and the 1000 cancel), + row-mean terms;
//             (3) d bg_mat = d act * act * mul where the clip is open, d brightness = sum d act * act, d mul = sum d act * act * bg.
// Step (1) is per bounce ray (atomics into an 8 MB image on the device), steps (2) / (3) are two prefix-sum passes and an
// elementwise pass over the 512 x 1024 map per optimiser step.
// ------------------------------------------------------------------------------------------------
struct NmfSatScatterTap {
  float* gsat; int h, w;          // channel-last gradient image [h][w][4] (zeroed by the caller)
  float gs[3];                    // upstream gradient * 1 / size of the lookup's box
  NMF_HD void operator()(float px, float py, float sign, float* acc) const {
    (void)acc;
    px = nmf_clampf(px, -1.0f, 1.0f);
    py = nmf_clampf(py, -1.0f, 1.0f);
    float ix = (px + 1.0f) * 0.5f * (float)(w - 1), iy = (py + 1.0f) * 0.5f * (float)(h - 1);
    float fx = floorf(ix), fy = floorf(iy);
    int x0 = (int)fx, y0 = (int)fy;
    float tx = ix - fx, ty = iy - fy;
    int x1 = x0 + 1 < w ? x0 + 1 : x0, y1 = y0 + 1 < h ? y0 + 1 : y0;
    const float wgt[4] = {(1.0f - tx) * (1.0f - ty), tx * (1.0f - ty), (1.0f - tx) * ty, tx * ty};
    const size_t idx[4] = {(size_t)y0 * w + x0, (size_t)y0 * w + x1, (size_t)y1 * w + x0, (size_t)y1 * w + x1};
    for (int q = 0; q < 4; ++q) {
      if (wgt[q] == 0.f) continue;
      for (int k = 0; k < 3; ++k) {
#ifdef __CUDA_ARCH__
        atomicAdd(gsat + idx[q] * 4 + k, sign * wgt[q] * gs[k]);
#else
        gsat[idx[q] * 4 + k] += sign * wgt[q] * gs[k];
#endif
      }
    }
  }
};
// one lookup: g = d loss / d rgb.  g_top / g_bot (3 floats each) collect the pole-row terms.
NMF_HD void nmf_env_lookup1_bwd_map(float* gsat, int h, int w, float mipbias, nmf_v3 dir, float sa, const float* g, float* g_top,
                                    float* g_bot) {
  const NmfEnvBox bx = nmf_env_box(dir, sa, h, w, mipbias);
  const float cutoff = 1.0f - 2.0f / (float)h * 3.0f;
  if (bx.cy < -cutoff) { for (int k = 0; k < 3; ++k) NMF_ATOMIC_ADD(g_top + k, g[k]); return; }
  if (bx.cy > cutoff) { for (int k = 0; k < 3; ++k) NMF_ATOMIC_ADD(g_bot + k, g[k]); return; }
  NmfSatScatterTap tap;
  tap.gsat = gsat; tap.h = h; tap.w = w;
  const float inv_size = 1.0f / bx.size;
  for (int k = 0; k < 3; ++k) tap.gs[k] = g[k] * inv_size;
  float unused[3];
  nmf_env_integrate(tap, bx, unused);
}

// ------------------------------------------------------------------------------------------------
// Environment lookup, DIRECTIONAL derivative along a tangent of the direction (the bounce direction moves with the
// roughness: tangent = d L / d roughness from nmf_ggx_sample_dr; the mip level `sa` carries no gradient, see above).
// Forward-mode restatement of nmf_env_box + nmf_env_integrate + NmfSatTap with the reference's gradient conventions:
// _Atan2Damped (modules/safemath.py:8-32: backward divides by x^2 + y^2 + 1e-5), torch.clip gates (closed interval),
// grid_sample's bilinear coordinate gradient, constants where the reference assigns constants (wrap / pole pieces).
// ------------------------------------------------------------------------------------------------
NMF_HD NmfDual nmf_dclamp(NmfDual x, float lo, float hi) {
  if (x.v < lo) return nmf_dk(lo);
  if (x.v > hi) return nmf_dk(hi);
  return x;
}
NMF_HD NmfDual nmf_dmax_floor(NmfDual x, float floor_) { return x.v >= floor_ ? x : nmf_dk(floor_); }
NMF_HD NmfDual nmf_dlog(NmfDual x) { return nmf_dmk(logf(x.v), x.d / x.v); }
NMF_HD NmfDual nmf_dexp2(NmfDual x) { const float e = exp2f(x.v); return nmf_dmk(e, 0.6931471805599453f * e * x.d); }
NMF_HD NmfDual nmf_datan2_damped(NmfDual y, NmfDual x) {      // atan2(y, x); d = (x dy - y dx) / (x^2 + y^2 + 1e-5)
  return nmf_dmk(atan2f(y.v, x.v), (x.v * y.d - y.v * x.d) / (x.v * x.v + y.v * y.v + 1e-5f));
}
struct NmfEnvBoxD { NmfDual cx, cy, sw, sh, size; };
NMF_HD NmfEnvBoxD nmf_env_box_d(NmfDual3 u, float sa, int h, int w, NmfDual mipbias) {   // the tangent may sit on the bias too
  NmfEnvBoxD bx;
  const NmfDual cosv = nmf_dsqrt_floor(nmf_dk(1.0f) - u.z * u.z, NMF_EPS);
  const NmfDual d = nmf_dk((float)(h * w)) / nmf_dmax_floor((float)(2.0 * 3.14159265358979323846 * 3.14159265358979323846) * cosv, NMF_EPS);
  const float es = expf(sa);
  const NmfDual area = nmf_dmk(expf(logf(d.v / 2.0f) + sa), 0.5f * es * d.d);
  const NmfDual hh = nmf_dmax_floor(nmf_dsqrt_floor(area, NMF_EPS) * cosv, NMF_EPS);
  const NmfDual ww = area / hh;
  const float ln2 = 0.6931471805599453f;
  const NmfDual lw = nmf_dclamp(nmf_dlog(ww) * (1.0f / ln2) + mipbias, 0.0f, 7.0f);
  const NmfDual lh = nmf_dclamp(nmf_dlog(hh) * (1.0f / ln2) + mipbias, 0.0f, 7.0f);
  bx.sw = nmf_dexp2(lw) * (1.0f / (float)h / 2.0f);
  bx.sh = nmf_dexp2(lh) * (1.0f / (float)h);
  bx.size = (bx.sw * (0.5f * (float)w)) * (bx.sh * (0.5f * (float)h));
  const NmfDual n2 = u.x * u.x + u.y * u.y;
  const float n2s = sqrtf(n2.v);
  const NmfDual norm2d = nmf_dmk(n2s, n2s > 0.f ? n2.d / (2.0f * n2s) : 0.f);
  const NmfDual phi = nmf_datan2_damped(u.y, u.x);
  const NmfDual theta = nmf_datan2_damped(u.z, norm2d);
  const float twopi = 6.2831855f;
  const float pm = phi.v - twopi * floorf(phi.v / twopi);
  bx.cx = nmf_dmk((pm - 3.1415927f) / 3.1415927f, phi.d / 3.1415927f);
  bx.cy = nmf_dmk(-theta.v / 3.1415927f * 2.0f, -theta.d / 3.1415927f * 2.0f);
  return bx;
}
// bilinear tap of the channel-last SAT at clip(p, -1, 1) with its derivative along (px.d, py.d)
NMF_HD void nmf_sat_tap_d(const float* sat, int h, int w, NmfDual px, NmfDual py, float sign, NmfDual* acc) {
  px = nmf_dclamp(px, -1.0f, 1.0f);
  py = nmf_dclamp(py, -1.0f, 1.0f);
  const NmfDual ix = (px + nmf_dk(1.0f)) * (0.5f * (float)(w - 1)), iy = (py + nmf_dk(1.0f)) * (0.5f * (float)(h - 1));
  const float fx = floorf(ix.v), fy = floorf(iy.v);
  const int x0 = (int)fx, y0 = (int)fy;
  const NmfDual tx = nmf_dmk(ix.v - fx, ix.d), ty = nmf_dmk(iy.v - fy, iy.d);
  const int x1 = x0 + 1 < w ? x0 + 1 : x0, y1 = y0 + 1 < h ? y0 + 1 : y0;
  const NmfDual one = nmf_dk(1.0f);
  const NmfDual w00 = (one - tx) * (one - ty), w10 = tx * (one - ty), w01 = (one - tx) * ty, w11 = tx * ty;
  const float* a = sat + ((size_t)y0 * w + x0) * 4; const float* b = sat + ((size_t)y0 * w + x1) * 4;
  const float* c = sat + ((size_t)y1 * w + x0) * 4; const float* d = sat + ((size_t)y1 * w + x1) * 4;
  for (int k = 0; k < 3; ++k) acc[k] = acc[k] + sign * (w00 * a[k] + w10 * b[k] + w01 * c[k] + w11 * d[k]);
}
NMF_HD void nmf_env_box1_d(const float* sat, int h, int w, NmfDual x0, NmfDual y0, NmfDual x1, NmfDual y1, NmfDual inv_size,
                           NmfDual* out) {
  NmfDual acc[3] = {nmf_dk(0.f), nmf_dk(0.f), nmf_dk(0.f)};
  nmf_sat_tap_d(sat, h, w, x1, y1, 1.0f, acc);
  nmf_sat_tap_d(sat, h, w, x0, y0, 1.0f, acc);
  nmf_sat_tap_d(sat, h, w, x0, y1, -1.0f, acc);
  nmf_sat_tap_d(sat, h, w, x1, y0, -1.0f, acc);
  for (int k = 0; k < 3; ++k) out[k] = out[k] + acc[k] * inv_size;
}
// rgb[k] = value, drgb[k] = derivative along the tangent carried by `dir`
NMF_HD void nmf_env_lookup1_dm(const float* sat, int h, int w, NmfDual mipbias, const float* top, const float* bot, NmfDual3 dir,
                               float sa, float* rgb, float* drgb) {
  const NmfEnvBoxD bx = nmf_env_box_d(dir, sa, h, w, mipbias);
  const float cutoff = 1.0f - 2.0f / (float)h * 3.0f;
  if (bx.cy.v > cutoff) { for (int k = 0; k < 3; ++k) { rgb[k] = bot[k]; drgb[k] = 0.f; } return; }
  if (bx.cy.v < -cutoff) { for (int k = 0; k < 3; ++k) { rgb[k] = top[k]; drgb[k] = 0.f; } return; }
  const NmfDual hx = bx.sw * 0.5f, hy = bx.sh * 0.5f;
  const NmfDual bx0 = bx.cx - hx, by0 = bx.cy - hy, bx1 = bx.cx + hx, by1 = bx.cy + hy;
  const NmfDual inv_size = nmf_dk(1.0f) / bx.size;
  const NmfDual rot = nmf_dk(bx0.v > 0.0f ? -1.0f : 1.0f);
  NmfDual out[3] = {nmf_dk(0.f), nmf_dk(0.f), nmf_dk(0.f)};
  for (int part = 0; part < 3; ++part) {
    NmfDual x0 = bx0, x1 = bx1, y0 = by0, y1 = by1;
    if (part == 1) {
      if (!(by1.v > 1.0f)) continue;
      const NmfDual over = nmf_dclamp(by1 - nmf_dk(1.0f), 0.0f, 0.5f);
      x0 = bx0 + rot; x1 = bx1 + rot; y0 = nmf_dk(1.0f) - over; y1 = nmf_dk(1.0f);
    } else if (part == 2) {
      if (!(by0.v < -1.0f)) continue;
      const NmfDual over = nmf_dclamp(nmf_dk(-1.0f) - by0, 0.0f, 0.5f);
      x0 = bx0 + rot; x1 = bx1 + rot; y0 = nmf_dk(-1.0f); y1 = nmf_dk(-1.0f) + over;
    }
    for (int wrap = 0; wrap < 3; ++wrap) {
      NmfDual u0 = x0, u1 = x1;
      if (wrap == 1) {
        if (!(x1.v > 1.0f)) continue;
        u0 = nmf_dk(-1.0f); u1 = x1 - nmf_dk(2.0f);
      } else if (wrap == 2) {
        if (!(x0.v < -1.0f)) continue;
        u0 = x0 + nmf_dk(2.0f); u1 = nmf_dk(1.0f);
      }
      nmf_env_box1_d(sat, h, w, u0, y0, u1, y1, inv_size, out);
    }
  }
  for (int k = 0; k < 3; ++k) { rgb[k] = out[k].v * 1000.0f; drgb[k] = out[k].d * 1000.0f; }
}
NMF_HD void nmf_env_lookup1_d(const float* sat, int h, int w, float mipbias, const float* top, const float* bot, NmfDual3 dir,
                              float sa, float* rgb, float* drgb) {
  nmf_env_lookup1_dm(sat, h, w, nmf_dk(mipbias), top, bot, dir, sa, rgb, drgb);
}
// d rgb / d mipbias of one lookup (IntegralEquirect.mipbias is a parameter: the box size moves with it, sa2mip
// integral_equirect.py:373-397): the same forward-mode pass with the unit tangent on the bias and none on the direction.
NMF_HD void nmf_env_lookup1_dmipbias(const float* sat, int h, int w, float mipbias, const float* top, const float* bot, nmf_v3 dir,
                                     float sa, float* rgb, float* drgb) {
  nmf_env_lookup1_dm(sat, h, w, nmf_dmk(mipbias, 1.0f), top, bot, nmf_d3k(dir), sa, rgb, drgb);
}

// ------------------------------------------------------------------------------------------------
// BRDF MLP (modules/brdf.py:177-261), one row: x = [feat(24) | ISH(h)(18) | h(3) | ISH(d)(18) | d(3)] -> 64 ReLU -> 64 ReLU -> 4,
// bw = sigmoid(out[:3] + bias).  Weights transposed as in NmfScene (w0t [66][64], w1t [64][64], w2t [64][4]).
// Backward of g = d loss / d bw: accumulates the weight / bias gradients (same layouts) and returns d x[0..23]
// (the encodings and the roughness reach the MLP detached: models/microfacet.py:461-472).
// ------------------------------------------------------------------------------------------------
NMF_HD void nmf_brdf_input(const float* feat, nmf_v3 half_l, nmf_v3 diff_l, float rough, float* x) {
  for (int i = 0; i < 24; ++i) x[i] = feat[i];
  nmf_ish18(half_l, rough, x + 24);
  x[42] = half_l.x; x[43] = half_l.y; x[44] = half_l.z;
  nmf_ish18(diff_l, rough, x + 45);
  x[63] = diff_l.x; x[64] = diff_l.y; x[65] = diff_l.z;
}
NMF_HD void nmf_brdf_row_fwd_bwd(const float* x, const float* w0t, const float* b0, const float* w1t, const float* b1, const float* w2t,
                                 const float* b2, float brdf_bias, const float* g, float* bw, float* dw0t, float* db0, float* dw1t,
                                 float* db1, float* dw2t, float* db2, float* dfeat) {
  float h1[64], h2[64], o[4];
  for (int j = 0; j < 64; ++j) { float v = b0[j]; for (int k = 0; k < 66; ++k) v += x[k] * w0t[k * 64 + j]; h1[j] = fmaxf(v, 0.f); }
  for (int j = 0; j < 64; ++j) { float v = b1[j]; for (int k = 0; k < 64; ++k) v += h1[k] * w1t[k * 64 + j]; h2[j] = fmaxf(v, 0.f); }
  for (int j = 0; j < 4; ++j) { float v = b2[j]; for (int k = 0; k < 64; ++k) v += h2[k] * w2t[k * 4 + j]; o[j] = v; }
  float dout[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < 3; ++c) { bw[c] = nmf_sigmoid(o[c] + brdf_bias); if (g) dout[c] = g[c] * bw[c] * (1.0f - bw[c]); }
  if (!g) return;
  float dh2[64], dh1[64];
  for (int k = 0; k < 64; ++k) {
    float v = 0.f;
    for (int j = 0; j < 3; ++j) { NMF_ATOMIC_ADD(dw2t + k * 4 + j, h2[k] * dout[j]); v += w2t[k * 4 + j] * dout[j]; }
    dh2[k] = h2[k] > 0.f ? v : 0.f;
  }
  for (int j = 0; j < 3; ++j) NMF_ATOMIC_ADD(db2 + j, dout[j]);
  for (int k = 0; k < 64; ++k) {
    float v = 0.f;
    for (int j = 0; j < 64; ++j) { NMF_ATOMIC_ADD(dw1t + k * 64 + j, h1[k] * dh2[j]); v += w1t[k * 64 + j] * dh2[j]; }
    dh1[k] = h1[k] > 0.f ? v : 0.f;
  }
  for (int j = 0; j < 64; ++j) NMF_ATOMIC_ADD(db1 + j, dh2[j]);
  for (int k = 0; k < 66; ++k) {
    float v = 0.f;
    for (int j = 0; j < 64; ++j) { NMF_ATOMIC_ADD(dw0t + k * 64 + j, x[k] * dh1[j]); v += w0t[k * 64 + j] * dh1[j]; }
    if (k < 24) dfeat[k] += v;
  }
  for (int j = 0; j < 64; ++j) NMF_ATOMIC_ADD(db0 + j, dh1[j]);
}

// ------------------------------------------------------------------------------------------------
// One bounce sample with m environment rays (no re-trace), models/microfacet.py:352-613: the reverse pass composed from
// the pieces above.  reflect = mean_j [ F_j inc_j bw_j + (1 - F_j) diffuse ];  g = d loss / d reflect.
// Outputs: d R0 (3), d diffuse (3), d roughness (through L: Fresnel angle and environment direction), d feat (24, the
// noisy feature the BRDF MLP sees), BRDF weight gradients, environment-map scatter.
// ------------------------------------------------------------------------------------------------
struct NmfBrdfGrads { float* w0t; float* b0; float* w1t; float* b1; float* w2t; float* b2; };
NMF_HD void nmf_bounce_sample_bwd(const NmfScene& s, const float* nfeat, nmf_v3 V, nmf_v3 N, const float* R0, const float* diffuse,
                                  float rough, const float* u, int m, const float* g, float* dR0, float* ddiffuse, float* drough,
                                  float* dfeat, NmfBrdfGrads bg, float* gsat, float* g_top, float* g_bot, float* dN = nullptr) {
  const NmfEnvDyn ed = nmf_env_dyn_load(s);     // mipbias and pole means: by value or from NmfScene.env_dyn
  const float inv_m = 1.0f / (float)m;
  float gm[3] = {g[0] * inv_m, g[1] * inv_m, g[2] * inv_m};
  for (int c = 0; c < 3; ++c) { dR0[c] = 0.f; ddiffuse[c] = 0.f; }
  for (int k = 0; k < 24; ++k) dfeat[k] = 0.f;
  float dr = 0.f;
  if (dN) dN[0] = dN[1] = dN[2] = 0.f;
  for (int j = 0; j < m; ++j) {
    const float u1 = u[2 * j], u2 = u[2 * j + 1];
    const NmfGGX fw = nmf_ggx_sample(u1, u2, V, N, rough);             // half_l / diff_l / logpdf as the forward draws them
    const NmfGGXdr dg = nmf_ggx_sample_dr(u1, u2, V, N, rough);
    const float mip = -logf((float)m) - fw.logpdf;                     // models/microfacet.py:445-448 (no gradient)
    float x[66], bw[3], inc[3], dinc_dr[3];
    nmf_brdf_input(nfeat, fw.half_l, fw.diff_l, rough, x);
    const NmfDual3 Ld = nmf_d3(nmf_dmk(dg.L.x, dg.dL.x), nmf_dmk(dg.L.y, dg.dL.y), nmf_dmk(dg.L.z, dg.dL.z));
    nmf_env_lookup1_d(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot, Ld, mip, inc, dinc_dr);
    // forward value of the BRDF weight first (g = NULL), then the mix backward, then the MLP backward with d bw
    nmf_brdf_row_fwd_bwd(x, s.brdf_w0t, s.brdf_b0, s.brdf_w1t, s.brdf_b1, s.brdf_w2t, s.brdf_b2, s.brdf_bias, nullptr, bw, nullptr,
                         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    const float vh = nmf_dot(V, dg.H);
    const float cost = fabsf(vh);
    float a_R0[3], a_inc[3], a_bw[3], a_diff[3];
    const float dcost = nmf_fresnel_mix_bwd(R0, cost, inc, bw, diffuse, gm, a_R0, a_inc, a_bw, a_diff);
    for (int c = 0; c < 3; ++c) { dR0[c] += a_R0[c]; ddiffuse[c] += a_diff[c]; dr += a_inc[c] * dinc_dr[c]; }
    const float svh = vh > 0.f ? 1.0f : (vh < 0.f ? -1.0f : 0.f);
    dr += dcost * svh * nmf_dot(V, dg.dH);
    if (dN) {                                  // detach_N off: the same two paths (environment direction, Fresnel angle) per column of N
      for (int c = 0; c < 3; ++c) {
        const NmfGGXdr dn = nmf_ggx_sample_dN(u1, u2, V, N, rough, c);
        const NmfDual3 Ln = nmf_d3(nmf_dmk(dn.L.x, dn.dL.x), nmf_dmk(dn.L.y, dn.dL.y), nmf_dmk(dn.L.z, dn.dL.z));
        float inc_n[3], dinc_dn[3];
        nmf_env_lookup1_d(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot, Ln, mip, inc_n, dinc_dn);
        float acc = dcost * svh * nmf_dot(V, dn.dH);
        for (int k = 0; k < 3; ++k) acc += a_inc[k] * dinc_dn[k];
        dN[c] += acc;
      }
    }
    float bw2[3];
    nmf_brdf_row_fwd_bwd(x, s.brdf_w0t, s.brdf_b0, s.brdf_w1t, s.brdf_b1, s.brdf_w2t, s.brdf_b2, s.brdf_bias, a_bw, bw2, bg.w0t, bg.b0,
                         bg.w1t, bg.b1, bg.w2t, bg.b2, dfeat);
    nmf_env_lookup1_bwd_map(gsat, s.env_h, s.env_w, ed.mipbias, dg.L, mip, a_inc, g_top, g_bot);
  }
  *drough = dr;
}

// ------------------------------------------------------------------------------------------------
// Normals (fields/tensor_base.py:107-129 with the smoothed-difference planes of modules/grid_sample_Cinf.py:218-281), the
// path that opens when detach_N goes off:  grad[mat0] += sum_c line_c * plane_dx_c,  grad[mat1] += sum_c line_c * plane_dy_c,
// grad[vec] += sum_c plane_c * line_dy_c  (all bilinear taps),  n = normalize(-grad * invaabbSize).
// nmf_normal_vec_bwd: d n -> d grad.   nmf_normal_bwd: d grad -> scatter into gradient images laid out like dpack
// ([h][w][val16 | dx16 | dy16]) and lpack ([n][4][val4 | dy4]); the dx / dy parts still have to go through the adjoint of
// the 5x5 stencil convolution (a whole-plane pass per step) to reach the density planes / lines.
// ------------------------------------------------------------------------------------------------
NMF_HD void nmf_normal_vec_bwd(const NmfScene& s, const float* grad, const float* dn, float* dgrad) {
  float v[3], n2 = 0.f;
  for (int c = 0; c < 3; ++c) { v[c] = -(grad[c] * s.inv_aabb2[c]); n2 += v[c] * v[c]; }
  if (!(n2 > NMF_EPS)) { dgrad[0] = dgrad[1] = dgrad[2] = 0.f; return; }
  const float len = sqrtf(n2);
  float nd = 0.f;
  for (int c = 0; c < 3; ++c) nd += (v[c] / len) * dn[c];
  for (int c = 0; c < 3; ++c) dgrad[c] = -s.inv_aabb2[c] * (dn[c] - (v[c] / len) * nd) / len;
}
NMF_HD void nmf_normal_bwd(const NmfScene& s, const NmfTaps& t, const float* dgrad, float* const* gpack, float* const* glpack) {
  for (int p = 0; p < 3; ++p) {
    const int w = s.plane_w[p];
    const NmfLerp& lx = t.px[p]; const NmfLerp& ly = t.py[p]; const NmfLerp& ll = t.pl[p];
    const float g0 = dgrad[NMF_MAT0(p)], g1 = dgrad[NMF_MAT1(p)], gv = dgrad[NMF_VEC(p)];
    const size_t tex[4] = {(size_t)ly.i0 * w + lx.i0, (size_t)ly.i0 * w + lx.i1, (size_t)ly.i1 * w + lx.i0, (size_t)ly.i1 * w + lx.i1};
    const float tw[4] = {ly.w0 * lx.w0, ly.w0 * lx.w1, ly.w1 * lx.w0, ly.w1 * lx.w1};
    for (int c = 0; c < 16; ++c) {
      const int lo = (c >> 2) * 8 + (c & 3);              // lpack: [texel][group][val4 | dy4]
      float pc = 0.f, dpx = 0.f, dpy = 0.f;
      for (int q = 0; q < 4; ++q) {
        if (tw[q] == 0.f) continue;
        const float* e = s.dpack[p] + tex[q] * 48;
        pc += tw[q] * e[c]; dpx += tw[q] * e[16 + c]; dpy += tw[q] * e[32 + c];
      }
      const float* a0 = s.lpack[p] + (size_t)ll.i0 * 32;
      const float* a1 = s.lpack[p] + (size_t)ll.i1 * 32;
      const float lc = (ll.w0 != 0.f ? ll.w0 * a0[lo] : 0.f) + (ll.w1 != 0.f ? ll.w1 * a1[lo] : 0.f);
      const float dly = (ll.w0 != 0.f ? ll.w0 * a0[lo + 4] : 0.f) + (ll.w1 != 0.f ? ll.w1 * a1[lo + 4] : 0.f);
      const float d_lc = g0 * dpx + g1 * dpy, d_dpx = g0 * lc, d_dpy = g1 * lc, d_pc = gv * dly, d_dly = gv * pc;
      for (int q = 0; q < 4; ++q) {
        if (tw[q] == 0.f) continue;
        float* e = gpack[p] + tex[q] * 48;
        NMF_ATOMIC_ADD(e + c, tw[q] * d_pc); NMF_ATOMIC_ADD(e + 16 + c, tw[q] * d_dpx); NMF_ATOMIC_ADD(e + 32 + c, tw[q] * d_dpy);
      }
      if (ll.w0 != 0.f) { NMF_ATOMIC_ADD(glpack[p] + (size_t)ll.i0 * 32 + lo, ll.w0 * d_lc); NMF_ATOMIC_ADD(glpack[p] + (size_t)ll.i0 * 32 + lo + 4, ll.w0 * d_dly); }
      if (ll.w1 != 0.f) { NMF_ATOMIC_ADD(glpack[p] + (size_t)ll.i1 * 32 + lo, ll.w1 * d_lc); NMF_ATOMIC_ADD(glpack[p] + (size_t)ll.i1 * 32 + lo + 4, ll.w1 * d_dly); }
    }
  }
}

// The same scatter with 16-byte accumulates (device: red.global.add.v4.f32): channel groups of 4 are contiguous in both
// gradient images, so a sample issues 192 vector atomics instead of 768 scalar ones.  Used by k_mf_sample_bwd; equal to
// nmf_normal_bwd up to the order of the additions (the GPU gradient-parity tests cover it).
NMF_HD nmf_f4 nmf_f4_axpby(float a, nmf_f4 x, float b, nmf_f4 y) {
  nmf_f4 o; o.x = a * x.x + b * y.x; o.y = a * x.y + b * y.y; o.z = a * x.z + b * y.z; o.w = a * x.w + b * y.w; return o;
}
NMF_HD void nmf_normal_bwd4(const NmfScene& s, const NmfTaps& t, const float* dgrad, float* const* gpack, float* const* glpack) {
  for (int p = 0; p < 3; ++p) {
    const int w = s.plane_w[p];
    const NmfLerp& lx = t.px[p]; const NmfLerp& ly = t.py[p]; const NmfLerp& ll = t.pl[p];
    const float g0 = dgrad[NMF_MAT0(p)], g1 = dgrad[NMF_MAT1(p)], gv = dgrad[NMF_VEC(p)];
    const size_t tex[4] = {(size_t)ly.i0 * w + lx.i0, (size_t)ly.i0 * w + lx.i1, (size_t)ly.i1 * w + lx.i0, (size_t)ly.i1 * w + lx.i1};
    const float tw[4] = {ly.w0 * lx.w0, ly.w0 * lx.w1, ly.w1 * lx.w0, ly.w1 * lx.w1};
    for (int g = 0; g < 4; ++g) {
      nmf_f4 pc = nmf_f4_zero(), dpx = nmf_f4_zero(), dpy = nmf_f4_zero();
      for (int q = 0; q < 4; ++q) {
        if (tw[q] == 0.f) continue;
        const float* e = s.dpack[p] + tex[q] * 48 + 4 * g;
        nmf_f4_fma(pc, NMF_LD4(e), tw[q]); nmf_f4_fma(dpx, NMF_LD4(e + 16), tw[q]); nmf_f4_fma(dpy, NMF_LD4(e + 32), tw[q]);
      }
      const float* a0 = s.lpack[p] + (size_t)ll.i0 * 32 + 8 * g;
      const float* a1 = s.lpack[p] + (size_t)ll.i1 * 32 + 8 * g;
      nmf_f4 lc = nmf_f4_zero(), dly = nmf_f4_zero();
      if (ll.w0 != 0.f) { nmf_f4_fma(lc, NMF_LD4(a0), ll.w0); nmf_f4_fma(dly, NMF_LD4(a0 + 4), ll.w0); }
      if (ll.w1 != 0.f) { nmf_f4_fma(lc, NMF_LD4(a1), ll.w1); nmf_f4_fma(dly, NMF_LD4(a1 + 4), ll.w1); }
      const nmf_f4 d_lc = nmf_f4_axpby(g0, dpx, g1, dpy);
      for (int q = 0; q < 4; ++q) {
        float* e = gpack[p] + tex[q] * 48 + 4 * g;
        nmf_acc4(e, dly, tw[q] * gv);            // d value plane  = gv * line'
        nmf_acc4(e + 16, lc, tw[q] * g0);        // d dx plane     = g0 * line
        nmf_acc4(e + 32, lc, tw[q] * g1);        // d dy plane     = g1 * line
      }
      float* b0 = glpack[p] + (size_t)ll.i0 * 32 + 8 * g;
      float* b1 = glpack[p] + (size_t)ll.i1 * 32 + 8 * g;
      nmf_acc4(b0, d_lc, ll.w0); nmf_acc4(b0 + 4, pc, ll.w0 * gv);
      nmf_acc4(b1, d_lc, ll.w1); nmf_acc4(b1 + 4, pc, ll.w1 * gv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Finishing pass of the normal path: the gradient images scattered by nmf_normal_bwd hold d loss / d (value | dx | dy) per
// texel; the dx / dy planes are cross-correlations of the density plane with the 5x5 smoothed-difference stencils
// (grid_sample_Cinf.py:218-242, zero padding 2), so the plane gradient is  val + K_x^T * g_dx + K_y^T * g_dy  -- one texel
// and one channel per call (a whole-plane, HBM-streaming pass per optimiser step: 25 taps x 2 images).  Output layout =
// NmfPlainGrads.d_plane / d_line (channel-last), so the reverse-pass kernels can add it to the compositing path's gradient.
// ------------------------------------------------------------------------------------------------
NMF_HD float nmf_plane_grad_finish(const float* gpack, int h, int w, const float* kx25, const float* ky25, int y, int x, int c) {
  float acc = gpack[((size_t)y * w + x) * 48 + c];
  for (int i = 0; i < 5; ++i) {
    const int yy = y - i + 2;
    if (yy < 0 || yy >= h) continue;
    for (int j = 0; j < 5; ++j) {
      const int xx = x - j + 2;
      if (xx < 0 || xx >= w) continue;
      const float* e = gpack + ((size_t)yy * w + xx) * 48;
      acc += kx25[i * 5 + j] * e[16 + c] + ky25[i * 5 + j] * e[32 + c];
    }
  }
  return acc;
}
// lines are (N, 1) images: only the centre column of the stencil meets data
NMF_HD float nmf_line_grad_finish(const float* glpack, int n, const float* ky25, int i, int c) {
  const int lo = (c >> 2) * 8 + (c & 3);
  float acc = glpack[(size_t)i * 32 + lo];
  for (int r = 0; r < 5; ++r) {
    const int ii = i - r + 2;
    if (ii < 0 || ii >= n) continue;
    acc += ky25[r * 5 + 2] * glpack[(size_t)ii * 32 + lo + 4];
  }
  return acc;
}

// Reverse pass of TensorBase.compute_normals (fields/tensor_base.py:107-129) for ONE sample: recomputes the smoothed-difference
// gradient of the density feature at p (the forward's own taps), sends d n through the normalisation and scatters into the
// gradient images.  Shared by k_normals_bwd_scatter (csrc/nmf_normals_bwd.cu) and tests/hostcheck.
NMF_HD void nmf_normals_bwd_sample(const NmfScene& s, const float* p, const float* dn, float* const* gpack, float* const* glpack) {
  if (dn[0] == 0.f && dn[1] == 0.f && dn[2] == 0.f) return;
  float xn[3];
  nmf_normalize_xyz(s, p, xn);
  const NmfTaps t = nmf_vm_taps(s, xn);
  float grad[3] = {0.f, 0.f, 0.f}, dgrad[3];
  for (int l = 0; l < 8; ++l) nmf_normal_lane(s, t, l, grad);
  nmf_normal_vec_bwd(s, grad, dn, dgrad);
  if (dgrad[0] == 0.f && dgrad[1] == 0.f && dgrad[2] == 0.f) return;
  nmf_normal_bwd(s, t, dgrad, gpack, glpack);
}

// Finishing pass of the environment-map gradient: reverse prefix sums of the scattered SAT gradient over x and y (the adjoint
// of integral_equirect.py:271-273's double cumsum), the pole-row means, then the exp activation with its clip at 20
// (integral_equirect.py:263-270).  Sequential reference of what is two scan passes + one elementwise pass on the device.
// gsat: [h][w][4] (overwritten with d act), bg: (3, h, w) parameter, out: (3, h, w) gradient (accumulated).
inline void nmf_env_map_grad_finish(float* gsat, int h, int w, const float* g_top, const float* g_bot, const float* bg, float brightness,
                                    float mul, float* out) {
  for (int y = 0; y < h; ++y)
    for (int x = w - 2; x >= 0; --x)
      for (int k = 0; k < 3; ++k) gsat[((size_t)y * w + x) * 4 + k] += gsat[((size_t)y * w + x + 1) * 4 + k];
  for (int y = h - 2; y >= 0; --y)
    for (int x = 0; x < w; ++x)
      for (int k = 0; k < 3; ++k) gsat[((size_t)y * w + x) * 4 + k] += gsat[((size_t)(y + 1) * w + x) * 4 + k];
  for (int k = 0; k < 3; ++k)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        float d = gsat[((size_t)y * w + x) * 4 + k];
        if (y == 0) d += g_top[k] / (float)w;
        if (y == h - 1) d += g_bot[k] / (float)w;
        const float pre = brightness + mul * bg[((size_t)k * h + y) * w + x];
        if (pre <= 20.0f) out[((size_t)k * h + y) * w + x] += d * expf(pre) * mul;
      }
}

// ------------------------------------------------------------------------------------------------
// Tangent of one bounce sample's reflected radiance along a tangent of its VIEW vector (forward mode): what a re-traced ray's
// samples contribute to d incoming / d (parent roughness | parent normal).  The bounce directions, the Fresnel angle and the
// environment directions move with V; the BRDF weight does not (its encodings are detached, microfacet.py:461-472), nor do
// the heads, the irradiance or the sample position.
// ------------------------------------------------------------------------------------------------
// One bounce ray of the above: d comb / d (tangent of V) with the BRDF weight `bw` given (it does not move with V);
// comb = F inc bw + (1 - F) diffuse.  Shared by the host restatement below and k_mf_tangent1 (csrc/nmf_mf_train.cu).
NMF_HD void nmf_bounce_ray_tangent(const NmfScene& s, NmfDual3 V, nmf_v3 N, const float* R0, const float* diffuse, float rough,
                                   float u1, float u2, float mip, const float* bw, float* comb, float* dcomb) {
  const NmfEnvDyn ed = nmf_env_dyn_load(s);     // mipbias and pole means: by value or from NmfScene.env_dyn
  const NmfGGXdr dg = nmf_ggx_sample_dual3(u1, u2, V, nmf_d3k(N), nmf_dk(rough));
  float inc[3], dinc[3];
  const NmfDual3 Ld = nmf_d3(nmf_dmk(dg.L.x, dg.dL.x), nmf_dmk(dg.L.y, dg.dL.y), nmf_dmk(dg.L.z, dg.dL.z));
  nmf_env_lookup1_d(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot, Ld, mip, inc, dinc);
  const NmfDual3 Hd = nmf_d3(nmf_dmk(dg.H.x, dg.dH.x), nmf_dmk(dg.H.y, dg.dH.y), nmf_dmk(dg.H.z, dg.dH.z));
  NmfDual vh = nmf_ddot(V, Hd);
  if (vh.v < 0.f) vh = nmf_dmk(-vh.v, -vh.d);                          // |v.h|
  const NmfDual mm = nmf_dclamp(nmf_dk(1.0f) - vh, 0.0f, 1.0f);
  const NmfDual m2 = mm * mm, m5 = m2 * m2 * mm;
  for (int c = 0; c < 3; ++c) {
    const NmfDual F = nmf_dk(R0[c]) + (1.0f - R0[c]) * m5;
    const NmfDual v = F * nmf_dmk(inc[c], dinc[c]) * bw[c] + (nmf_dk(1.0f) - F) * diffuse[c];
    comb[c] = v.v; dcomb[c] = v.d;
  }
}
NMF_HD void nmf_bounce_sample_tangent(const NmfScene& s, const float* nfeat, NmfDual3 V, nmf_v3 N, const float* R0, const float* diffuse,
                                      float rough, const float* u, int m, float* refl, float* drefl) {
  const nmf_v3 Vv = nmf_mk3(V.x.v, V.y.v, V.z.v);
  float acc[3] = {0.f, 0.f, 0.f}, dacc[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < m; ++j) {
    const float u1 = u[2 * j], u2 = u[2 * j + 1];
    const NmfGGX fw = nmf_ggx_sample(u1, u2, Vv, N, rough);
    const float mip = -logf((float)m) - fw.logpdf;
    float x[66], bw[3], comb[3], dcomb[3];
    nmf_brdf_input(nfeat, fw.half_l, fw.diff_l, rough, x);
    nmf_brdf_row_fwd_bwd(x, s.brdf_w0t, s.brdf_b0, s.brdf_w1t, s.brdf_b1, s.brdf_w2t, s.brdf_b2, s.brdf_bias, nullptr, bw, nullptr,
                         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    nmf_bounce_ray_tangent(s, V, N, R0, diffuse, rough, u1, u2, mip, bw, comb, dcomb);
    for (int c = 0; c < 3; ++c) { acc[c] += comb[c]; dacc[c] += dcomb[c]; }
  }
  for (int c = 0; c < 3; ++c) { refl[c] = acc[c] / (float)m; drefl[c] = dacc[c] / (float)m; }
}
