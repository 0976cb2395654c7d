// TEST INFRASTRUCTURE -- compiles the per-element math of the kernels (nmf_b200/csrc/nmf_math.cuh,
// nmf_field.cuh) for the host so that `pytest -m "not gpu"` can check it against the oracle without a GPU.
// It is not part of the product and nothing under nmf_b200/ loads it.
//   g++ -O2 -ffp-contract=off -shared -fPIC -o libnmf_hostcheck.so hostcheck.cpp
#include <vector>

#include "../../nmf_b200/csrc/nmf_train.cuh"
#include "../../nmf_b200/csrc/nmf_microfacet_bwd.cuh"

extern "C" {

void hc_mix64(const uint64_t* a, const uint64_t* b, int n, uint64_t* out) {
  for (int i = 0; i < n; ++i) out[i] = nmf_mix64(a[i], b[i]);
}
void hc_uniform(const uint64_t* key, uint32_t stream, int n, float* out) {
  for (int i = 0; i < n; ++i) out[i] = nmf_uniform(key[i], stream);
}
void hc_normal(const uint64_t* key, uint32_t sa, uint32_t sb, int n, float* out) {
  for (int i = 0; i < n; ++i) out[i] = nmf_normal(key[i], sa, sb);
}

void hc_noise24(const uint64_t* key, int n, float* out) {
  for (int i = 0; i < n; ++i) {
    const uint64_t seed = nmf_noise_seed(key[i]);
    for (uint32_t p = 0; p < 12; ++p) nmf_noise_pair(seed, p, out + 24 * i + 2 * p, out + 24 * i + 2 * p + 1);
  }
}

// AlphaGridSampler.sample (eval): dense validity + z
void hc_sample_rays(const NmfScene* s, const float* rays, int n, float near_override, uint8_t* valid, float* z) {
  const int S = s->n_steps;
  float near_ = near_override >= 0.f ? near_override : s->near;
  for (int r = 0; r < n; ++r) {
    const float* o = rays + 6 * r;
    const float* d = o + 3;
    float tmin = nmf_ray_tmin(o, d, s->aabb0, s->aabb1, near_, s->far);
    for (int k = 0; k < S; ++k) {
      float zk = nmf_step_z(tmin, s->stepsize, k);
      float p[3];
      nmf_step_pos(o, d, zk, p);
      bool ok = nmf_inside(p, s->aabb0, s->aabb1);
      if (ok && s->has_occ) {
        float xn[3];
        nmf_normalize_xyz(*s, p, xn);
        ok = nmf_occupied(s->occ_vox, s->occ_cell, s->ow, s->oh, s->od, s->opitch, xn[0], xn[1], xn[2]);
      }
      valid[(size_t)r * S + k] = ok;
      z[(size_t)r * S + k] = zk;
    }
  }
}

void hc_vm_density(const NmfScene* s, const float* xyz, int n, int stride, int activate, float* out) {
  for (int i = 0; i < n; ++i) {
    float xn[3];
    nmf_normalize_xyz(*s, xyz + (size_t)i * stride, xn);
    NmfTaps t = nmf_vm_taps(*s, xn);
    float f = 0.f;
    for (int g = 0; g < 4; ++g) f += nmf_density_group(*s, t, g);
    out[i] = activate ? nmf_feature2density(f, s->density_shift) : f;
  }
}
void hc_vm_appfeature(const NmfScene* s, const float* xyz, int n, int stride, float* out) {
  for (int i = 0; i < n; ++i) {
    float xn[3];
    nmf_normalize_xyz(*s, xyz + (size_t)i * stride, xn);
    NmfTaps t = nmf_vm_taps(*s, xn);
    float coef[72];
    for (int p = 0; p < 3; ++p)
      for (int g = 0; g < 6; ++g) {
        nmf_f4 c = nmf_app_group(*s, t, p, g);
        float* q = coef + p * 24 + 4 * g;
        q[0] = c.x; q[1] = c.y; q[2] = c.z; q[3] = c.w;
      }
    for (int o = 0; o < 24; ++o) {
      float acc = 0.f;
      for (int j = 0; j < 72; ++j) acc += s->basis_t[j * 24 + o] * coef[j];
      out[(size_t)i * 24 + o] = acc;
    }
  }
}
void hc_vm_normals(const NmfScene* s, const float* xyz, int n, int stride, float* out) {
  for (int i = 0; i < n; ++i) {
    float xn[3];
    nmf_normalize_xyz(*s, xyz + (size_t)i * stride, xn);
    NmfTaps t = nmf_vm_taps(*s, xn);
    float grad[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < 8; ++l) nmf_normal_lane(*s, t, l, grad);
    nmf_v3 nn = nmf_normal_from_grad(*s, grad);
    out[3 * i] = nn.x; out[3 * i + 1] = nn.y; out[3 * i + 2] = nn.z;
  }
}
void hc_env_lookup(const NmfScene* s, const float* dirs, const float* mip, int n, float* out) {
  for (int i = 0; i < n; ++i)
    nmf_env_lookup1(s->env_sat, s->env_h, s->env_w, s->env_mipbias, s->env_top, s->env_bot,
                    nmf_mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), mip[i], out + 3 * i);
}
void hc_ggx(const float* u, const float* V, const float* N, const float* r, int n, float* L, float* logpdf,
            float* half_l, float* diff_l) {
  for (int i = 0; i < n; ++i) {
    NmfGGX g = nmf_ggx_sample(u[2 * i], u[2 * i + 1], nmf_mk3(V[3 * i], V[3 * i + 1], V[3 * i + 2]),
                              nmf_mk3(N[3 * i], N[3 * i + 1], N[3 * i + 2]), r[i]);
    L[3 * i] = g.L.x; L[3 * i + 1] = g.L.y; L[3 * i + 2] = g.L.z;
    logpdf[i] = g.logpdf;
    half_l[3 * i] = g.half_l.x; half_l[3 * i + 1] = g.half_l.y; half_l[3 * i + 2] = g.half_l.z;
    diff_l[3 * i] = g.diff_l.x; diff_l[3 * i + 1] = g.diff_l.y; diff_l[3 * i + 2] = g.diff_l.z;
  }
}
void hc_ish18(const float* v, const float* rough, int n, float* out) {
  for (int i = 0; i < n; ++i) nmf_ish18(nmf_mk3(v[3 * i], v[3 * i + 1], v[3 * i + 2]), rough[i], out + 18 * i);
}
void hc_sh9(const float* v, int n, float* out) {
  for (int i = 0; i < n; ++i) nmf_sh9(nmf_mk3(v[3 * i], v[3 * i + 1], v[3 * i + 2]), out + 9 * i);
}
void hc_srgb(const float* x, int n, float* out) {
  for (int i = 0; i < n; ++i) out[i] = nmf_srgb(x[i]);
}

// ---- training slice (nmf_b200/csrc/nmf_train.cuh) ----
// AlphaGridSampler.sample(is_train=True): z = t_min + cumsum(jittered steps); the fp64 sum is exact (see k_train_sample)
void hc_sample_rays_train(const NmfScene* s, const float* rays, int n, float near_override, uint64_t seed, uint64_t ray_id0,
                          const uint64_t* ray_ids, uint8_t* valid, float* z) {
  const int S = s->n_steps;
  const float near_ = near_override >= 0.f ? near_override : s->near;
  for (int r = 0; r < n; ++r) {
    const float* o = rays + 6 * r;
    const float* d = o + 3;
    const float tmin = nmf_ray_tmin(o, d, s->aabb0, s->aabb1, near_, s->far);
    const uint64_t key = nmf_primary_key(seed, ray_ids ? ray_ids[r] : ray_id0 + (uint64_t)r);
    double run = 0.0;
    for (int k = 0; k < S; ++k) {
      run += (double)nmf_jitter_step(key, k, s->stepsize);
      const float zk = NMF_ADD(tmin, (float)run);
      float p[3];
      nmf_step_pos(o, d, zk, p);
      bool ok = nmf_inside(p, s->aabb0, s->aabb1);
      if (ok && s->has_occ) {
        float xn[3];
        nmf_normalize_xyz(*s, p, xn);
        ok = nmf_occupied(s->occ_vox, s->occ_cell, s->ow, s->oh, s->od, s->opitch, xn[0], xn[1], xn[2]);
      }
      valid[(size_t)r * S + k] = ok;
      z[(size_t)r * S + k] = zk;
    }
  }
}

// nmf_train_plain restated sequentially from the same per-element functions (the kernels' warp / tile choreography is
// what `-m gpu` covers).  `g` buffers must be zeroed by the caller.
void hc_train_plain(const NmfScene* s, const NmfTrain* tp, const float* rays, const float* gt, const NmfPlainGrads* g,
                    float* rgb_map, float* acc_map, uint8_t* whole, double* loss, int* n_kept) {
  const int n = tp->n_rays, S = s->n_steps;
  std::vector<uint8_t> valid((size_t)n * S);
  std::vector<float> z((size_t)n * S);
  hc_sample_rays_train(s, rays, n, -1.0f, tp->seed, tp->ray_id0, tp->ray_ids, valid.data(), z.data());
  std::vector<int> nv(n, 0);
  long long total = 0;
  for (int r = 0; r < n; ++r) { for (int k = 0; k < S; ++k) nv[r] += valid[(size_t)r * S + k]; total += nv[r]; }
  const bool trunc = tp->max_samples > 0 && total > tp->max_samples;
  long long run = 0;
  int kept = 0, M = 0;
  for (int r = 0; r < n; ++r) {
    run += nv[r];
    whole[r] = !trunc || run < tp->max_samples;
    if (whole[r]) { kept = r + 1; M = (int)run; }
  }
  n_kept[0] = kept; n_kept[1] = M;
  loss[0] = loss[1] = 0.0;
  const float bg[3] = {tp->white_bg ? 1.f : 0.f, tp->white_bg ? 1.f : 0.f, tp->white_bg ? 1.f : 0.f};
  struct Smp { float z, dist, f, alpha, T, w, rgb[3], x[135], h1[128], h2[128], dw; NmfTaps t; };
  for (int r = 0; r < n; ++r) {
    for (int c = 0; c < 3; ++c) rgb_map[3 * r + c] = 0.f;
    acc_map[r] = 0.f;
    if (r >= kept) continue;
    const float* o = rays + 6 * r;
    const float* d = o + 3;
    std::vector<Smp> sm;
    float T = 1.0f, acc = 0.f, lin[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < S; ++k) {
      if (!valid[(size_t)r * S + k]) continue;
      Smp q;
      q.z = z[(size_t)r * S + k];
      q.dist = (k + 1 < S ? NMF_SUB(z[(size_t)r * S + k + 1], q.z) : 0.f) * s->distance_scale;
      float p[3], xn[3];
      nmf_step_pos(o, d, q.z, p);
      nmf_normalize_xyz(*s, p, xn);
      q.t = nmf_vm_taps(*s, xn);
      q.f = 0.f;
      for (int gi = 0; gi < 4; ++gi) q.f += nmf_density_group(*s, q.t, gi);
      q.alpha = 1.0f - expf(-nmf_feature2density(q.f, s->density_shift) * q.dist);
      q.T = T;
      q.w = q.alpha * T;
      T *= 1.0f - q.alpha + 1e-10f;
      acc += q.w;
      // appearance + MLP forward
      float coef[72], feat[24];
      nmf_app_coef(*s, q.t, coef);
      for (int oo = 0; oo < 24; ++oo) {
        float a = 0.f;
        for (int j = 0; j < 72; ++j) a += s->basis_t[j * 24 + oo] * coef[j];
        feat[oo] = a;
      }
      nmf_plain_encode(feat, d, q.x, 1);
      for (int j = 0; j < 128; ++j) {
        float a = s->plain_b0[j];
        for (int i = 0; i < 135; ++i) a += q.x[i] * s->plain_w0t[i * 128 + j];
        q.h1[j] = fmaxf(a, 0.f);
      }
      for (int j = 0; j < 128; ++j) {
        float a = s->plain_b1[j];
        for (int i = 0; i < 128; ++i) a += q.h1[i] * s->plain_w1t[i * 128 + j];
        q.h2[j] = fmaxf(a, 0.f);
      }
      for (int c = 0; c < 3; ++c) {
        float a = s->plain_b2[c];
        for (int i = 0; i < 128; ++i) a += q.h2[i] * s->plain_w2t[i * 3 + c];
        q.rgb[c] = nmf_sigmoid(a);
        lin[c] += q.w * q.rgb[c];
      }
      sm.push_back(q);
    }
    float gl[3], ga;
    loss[0] += nmf_train_loss_ray(lin, acc, bg, gt + 3 * r, tp->lambda_pred, rgb_map + 3 * r, gl, &ga);
    loss[1] += acc;
    acc_map[r] = acc;
    // backward: MLP, encoding, basis, appearance factors
    for (Smp& q : sm) {
      float dpre[3], dh2[128], dh1[128], dx[135], dfeat[24], coef[72], dcoef[72];
      q.dw = ga;
      for (int c = 0; c < 3; ++c) {
        q.dw += gl[c] * q.rgb[c];
        dpre[c] = q.w * gl[c] * q.rgb[c] * (1.0f - q.rgb[c]);
        g->b2[c] += dpre[c];
      }
      for (int k = 0; k < 128; ++k) {
        float a = 0.f;
        for (int c = 0; c < 3; ++c) { g->w2t[k * 3 + c] += q.h2[k] * dpre[c]; a += s->plain_w2t[k * 3 + c] * dpre[c]; }
        dh2[k] = q.h2[k] > 0.f ? a : 0.f;
      }
      for (int k = 0; k < 128; ++k) {
        float a = 0.f;
        for (int j = 0; j < 128; ++j) { g->w1t[k * 128 + j] += q.h1[k] * dh2[j]; a += s->plain_w1[j * 128 + k] * dh2[j]; }
        dh1[k] = q.h1[k] > 0.f ? a : 0.f;
      }
      for (int j = 0; j < 128; ++j) { g->b1[j] += dh2[j]; g->b0[j] += dh1[j]; }
      for (int i = 0; i < 135; ++i) {
        float a = 0.f;
        for (int j = 0; j < 128; ++j) { g->w0t[i * 128 + j] += q.x[i] * dh1[j]; a += s->plain_w0[j * 135 + i] * dh1[j]; }
        dx[i] = a;
      }
      for (int oo = 0; oo < 24; ++oo)
        dfeat[oo] = nmf_plain_encode_bwd(q.x, 1, oo, dx[oo], dx[27 + 2 * oo], dx[28 + 2 * oo], dx[75 + 2 * oo], dx[76 + 2 * oo]);
      nmf_app_coef(*s, q.t, coef);
      for (int j = 0; j < 72; ++j) {
        float a = 0.f;
        for (int oo = 0; oo < 24; ++oo) { g->basis_t[j * 24 + oo] += coef[j] * dfeat[oo]; a += s->basis_t[j * 24 + oo] * dfeat[oo]; }
        dcoef[j] = a;
      }
      nmf_app_bwd(*s, q.t, dcoef, g->a_plane, g->a_line);
    }
    // backward: compositing and density factors
    float suffix = 0.f;
    for (int i = (int)sm.size() - 1; i >= 0; --i) {
      const Smp& q = sm[i];
      const float dsigma = nmf_composite_bwd(q.dw, q.T, q.alpha, q.dist, suffix);
      suffix += q.dw * q.w;
      const float df = dsigma * nmf_feature2density_grad(q.f, s->density_shift);
      if (df != 0.f) nmf_density_bwd(*s, q.t, df, g->d_plane, g->d_line);
    }
  }
}

// Training forward + backward of the MICROFACET model on the host for the case without re-trace (max_retrace_rays = ())
// and detach_N = True: TensorNeRF.forward(is_train=True) -> Microfacet.forward -> photometric loss, then the reverse pass
// composed from csrc/nmf_microfacet_bwd.cuh and the compositing / VM-factor backward of the model=tensorf slice.
// Reference for the reverse-pass kernels of DESIGN.md section 9 (checked against the oracle's autograd of render_chunk).
void hc_train_microfacet(const NmfScene* s, const NmfTrain* tp, const float* rays, const float* gt, const NmfPlainGrads* g,
                         float* d_head_w, float* d_head_b, float* dw0t, float* db0, float* dw1t, float* db1, float* dw2t, float* db2,
                         float* gsat, float* g_top, float* g_bot, float* rgb_map, float* acc_map, double* loss, int* n_samples,
                         int detach_N, float* const* gpack, float* const* glpack, float lambda_ori) {
  const int n = tp->n_rays, S = s->n_steps;
  std::vector<uint8_t> valid((size_t)n * S);
  std::vector<float> z((size_t)n * S);
  hc_sample_rays_train(s, rays, n, -1.0f, tp->seed, tp->ray_id0, tp->ray_ids, valid.data(), z.data());
  loss[0] = loss[1] = loss[2] = 0.0;
  *n_samples = 0;
  // dynamic batch truncation (alphagrid.py:353-364): rays kept = the prefix whose inclusive sample count stays below max_samples
  int kept = n;
  if (tp->max_samples > 0) {
    long long total = 0, run = 0;
    std::vector<int> nv(n, 0);
    for (int r = 0; r < n; ++r) { for (int k = 0; k < S; ++k) nv[r] += valid[(size_t)r * S + k]; total += nv[r]; }
    if (total > tp->max_samples) {
      kept = 0;
      for (int r = 0; r < n; ++r) { run += nv[r]; if (run < tp->max_samples) kept = r + 1; }
    }
  }
  n_samples[1] = kept;
  const float bg[3] = {1.f, 1.f, 1.f};
  NmfBrdfGrads bgr{dw0t, db0, dw1t, db1, dw2t, db2};
  struct Smp { float z, dist, f, alpha, T, w, dw; NmfTaps t; float feat[24], nfeat[24], albedo[3], f0[3], rough, E[3], refl[3];
               nmf_v3 V, Nf; int count; uint64_t skey; float ngrad[3], sgn, vn; };
  for (int r = 0; r < n; ++r) {
    const float* o = rays + 6 * r;
    const float* d = o + 3;
    const uint64_t rkey = nmf_primary_key(tp->seed, tp->ray_ids ? tp->ray_ids[r] : tp->ray_id0 + (uint64_t)r);
    if (r >= kept) { for (int c = 0; c < 3; ++c) rgb_map[3 * r + c] = 0.f; acc_map[r] = 0.f; continue; }
    std::vector<Smp> sm;
    float T = 1.0f, acc = 0.f, lin[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < S; ++k) {
      if (!valid[(size_t)r * S + k]) continue;
      Smp q;
      q.z = z[(size_t)r * S + k];
      q.dist = (k + 1 < S ? NMF_SUB(z[(size_t)r * S + k + 1], q.z) : 0.f) * s->distance_scale;
      float p[3], xn[3];
      nmf_step_pos(o, d, q.z, p);
      nmf_normalize_xyz(*s, p, xn);
      q.t = nmf_vm_taps(*s, xn);
      q.f = 0.f;
      for (int gi = 0; gi < 4; ++gi) q.f += nmf_density_group(*s, q.t, gi);
      q.alpha = 1.0f - expf(-nmf_feature2density(q.f, s->density_shift) * q.dist);
      q.T = T;
      q.w = q.alpha * T;
      T *= 1.0f - q.alpha + 1e-10f;
      acc += q.w;
      // appearance feature, normal, heads, irradiance
      float coef[72];
      nmf_app_coef(*s, q.t, coef);
      for (int oo = 0; oo < 24; ++oo) {
        float a = 0.f;
        for (int j = 0; j < 72; ++j) a += s->basis_t[j * 24 + oo] * coef[j];
        q.feat[oo] = a;
      }
      float grad[3] = {0.f, 0.f, 0.f};
      for (int l = 0; l < 8; ++l) nmf_normal_lane(*s, q.t, l, grad);
      const nmf_v3 nrm = nmf_normal_from_grad(*s, grad);
      float lin11[11];
      for (int h = 0; h < 11; ++h) {
        float v = s->head_b[h];
        for (int i = 0; i < 24; ++i) v += s->head_w[h * 24 + i] * q.feat[i];
        lin11[h] = v;
      }
      float sh[9];
      nmf_sh9(nrm, sh);
      for (int c = 0; c < 3; ++c) {
        q.albedo[c] = nmf_clampf(nmf_sigmoid(s->diffuse_mul * lin11[c] + s->diffuse_bias), 0.f, 1.f);
        q.f0[c] = nmf_sigmoid(lin11[6 + c] + s->f0_bias);
        float e = 0.f;
        for (int i = 0; i < 9; ++i) e += s->sh_conv[i * 3 + c] * sh[i];
        q.E[c] = e;
        q.refl[c] = 0.f;
      }
      q.rough = nmf_clampf(nmf_sigmoid(lin11[9] + s->roughness_bias) / 2.0f, 1e-2f, 1.0f);
      q.V = nmf_mk3(-d[0], -d[1], -d[2]);
      const float vn = nmf_dot(q.V, nrm);
      const float sgn = vn > 0.f ? 1.f : (vn < 0.f ? -1.f : 0.f);
      q.Nf = nmf_mk3(nrm.x * sgn, nrm.y * sgn, nrm.z * sgn);
      q.sgn = sgn;
      q.vn = vn;
      for (int c = 0; c < 3; ++c) q.ngrad[c] = grad[c];
      if (vn < 0.f) loss[2] += (double)(q.w * vn * vn);                    // ori_loss (tensor_nerf.py:573-583)
      q.skey = nmf_mix64(rkey, (uint64_t)k);
      const float kf = floorf(q.w * (float)s->rays_per_ray + nmf_uniform(q.skey, NMF_STREAM_BOUNCE) - 0.5f);     // pt_selectors.py:20-40
      q.count = (int)nmf_clampf(kf, 0.f, (float)NMF_MAX_BOUNCE);
      const uint64_t nseed = nmf_noise_seed(q.skey);
      for (uint32_t pp = 0; pp < 12; ++pp) {
        float n0, n1;
        nmf_noise_pair(nseed, pp, &n0, &n1);
        q.nfeat[2 * pp] = q.feat[2 * pp] + s->anoise * n0;
        q.nfeat[2 * pp + 1] = q.feat[2 * pp + 1] + s->anoise * n1;
      }
      // forward shading of the bounce rays (the reverse pass below recomputes them)
      if (q.count > 0) {
        const float offu = 0.25f * nmf_uniform(q.skey, NMF_STREAM_OFF_U), offv = 0.25f * nmf_uniform(q.skey, NMF_STREAM_OFF_V);
        for (int j = 0; j < q.count; ++j) {
          const float u1 = nmf_wrap01(s->sobol[2 * j] + offu), u2 = nmf_wrap01(s->sobol[2 * j + 1] + offv);
          const NmfGGX fw = nmf_ggx_sample(u1, u2, q.V, q.Nf, q.rough);
          const float mip = -logf((float)q.count) - fw.logpdf;
          float x[66], bw[3], inc[3];
          nmf_brdf_input(q.nfeat, fw.half_l, fw.diff_l, q.rough, x);
          nmf_brdf_row_fwd_bwd(x, s->brdf_w0t, s->brdf_b0, s->brdf_w1t, s->brdf_b1, s->brdf_w2t, s->brdf_b2, s->brdf_bias, nullptr, bw,
                               nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
          nmf_env_lookup1(s->env_sat, s->env_h, s->env_w, s->env_mipbias, s->env_top, s->env_bot, fw.L, mip, inc);
          const float cost = fabsf(nmf_dot(q.V, fw.H));
          for (int c = 0; c < 3; ++c) {
            const float F = nmf_fresnel(q.f0[c], cost);
            q.refl[c] += (F * inc[c] * bw[c] + (1.0f - F) * q.albedo[c] * q.E[c]) / (float)q.count;
          }
        }
      }
      for (int c = 0; c < 3; ++c) lin[c] += q.w * q.refl[c];
      sm.push_back(q);
    }
    *n_samples += (int)sm.size();
    float gl[3], ga;
    loss[0] += nmf_train_loss_ray(lin, acc, bg, gt + 3 * r, tp->lambda_pred, rgb_map + 3 * r, gl, &ga);
    loss[1] += acc;
    acc_map[r] = acc;
    // ---- reverse pass ----
    for (Smp& q : sm) {
      q.dw = ga;
      for (int c = 0; c < 3; ++c) q.dw += gl[c] * q.refl[c];
      if (lambda_ori != 0.f && q.vn < 0.f) {       // ori_lambda * sum w min(v.n, 0)^2: to the weight and, through the normal, to the density factors
        q.dw += lambda_ori * q.vn * q.vn;
        const float k2 = lambda_ori * q.w * 2.0f * q.vn;
        float dn[3] = {k2 * q.V.x, k2 * q.V.y, k2 * q.V.z}, dgrad[3];
        nmf_normal_vec_bwd(*s, q.ngrad, dn, dgrad);
        nmf_normal_bwd(*s, q.t, dgrad, gpack, glpack);
      }
      if (q.count == 0) continue;
      const float offu = 0.25f * nmf_uniform(q.skey, NMF_STREAM_OFF_U), offv = 0.25f * nmf_uniform(q.skey, NMF_STREAM_OFF_V);
      std::vector<float> u(2 * (size_t)q.count);
      for (int j = 0; j < q.count; ++j) { u[2 * j] = nmf_wrap01(s->sobol[2 * j] + offu); u[2 * j + 1] = nmf_wrap01(s->sobol[2 * j + 1] + offv); }
      float diffuse[3], gre[3], dR0[3], ddiff[3], drough, dnfeat[24], dfeat_h[24], g_alb[3];
      for (int c = 0; c < 3; ++c) { diffuse[c] = q.albedo[c] * q.E[c]; gre[c] = q.w * gl[c]; }
      float dNf[3];
      nmf_bounce_sample_bwd(*s, q.nfeat, q.V, q.Nf, q.f0, diffuse, q.rough, u.data(), q.count, gre, dR0, ddiff, &drough, dnfeat, bgr,
                            gsat, g_top, g_bot, detach_N ? nullptr : dNf);
      if (!detach_N) {                             // normal path: d Nf -> d n (sign flip) -> d grad -> derivative-plane scatter
        float dn[3] = {q.sgn * dNf[0], q.sgn * dNf[1], q.sgn * dNf[2]}, dgrad[3];
        nmf_normal_vec_bwd(*s, q.ngrad, dn, dgrad);
        nmf_normal_bwd(*s, q.t, dgrad, gpack, glpack);
      }
      for (int c = 0; c < 3; ++c) g_alb[c] = ddiff[c] * q.E[c];                       // diffuse = albedo * E, E under no_grad
      nmf_heads_bwd(q.feat, s->head_w, s->head_b, s->diffuse_mul, s->diffuse_bias, s->f0_bias, s->roughness_bias, g_alb, dR0, drough,
                    d_head_w, d_head_b, dfeat_h);
      float coef[72], dcoef[72];
      nmf_app_coef(*s, q.t, coef);
      for (int j = 0; j < 72; ++j) {
        float a = 0.f;
        for (int oo = 0; oo < 24; ++oo) {
          const float df = dnfeat[oo] + dfeat_h[oo];
          g->basis_t[j * 24 + oo] += coef[j] * df;
          a += s->basis_t[j * 24 + oo] * df;
        }
        dcoef[j] = a;
      }
      nmf_app_bwd(*s, q.t, dcoef, g->a_plane, g->a_line);
    }
    float suffix = 0.f;
    for (int i = (int)sm.size() - 1; i >= 0; --i) {
      const Smp& q = sm[i];
      const float dsigma = nmf_composite_bwd(q.dw, q.T, q.alpha, q.dist, suffix);
      suffix += q.dw * q.w;
      const float df = dsigma * nmf_feature2density_grad(q.f, s->density_shift);
      if (df != 0.f) nmf_density_bwd(*s, q.t, df, g->d_plane, g->d_line);
    }
  }
}

// optimiser step (nmf_adam_step / nmf_l1_reg on the host)
void hc_adam(float* p, const float* g, float* m, float* v, int n, float lr, float b1, float b2, float eps, float wd, int step,
             float grad_scale, float max_norm) {
  NmfAdamScalars h;
  h.one_minus_b1 = (float)(1.0 - (double)b1); h.b2 = b2; h.one_minus_b2 = (float)(1.0 - (double)b2); h.eps = eps;
  h.weight_decay = wd;
  h.step_size = (float)((double)lr / (1.0 - pow((double)b1, (double)step)));
  h.bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)step));
  h.grad_scale = grad_scale; h.max_norm = max_norm;
  double sq = 0.0;
  for (int i = 0; i < n; ++i) sq += (double)g[i] * g[i];
  const float gmul = grad_scale * nmf_clip_coef(sq, grad_scale, max_norm);
  for (int i = 0; i < n; ++i) nmf_adam_elem(p + i, g[i], m + i, v + i, h, gmul);
}
double hc_l1(const float* p, int n, float coef, float* g) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) { s += fabs((double)p[i]); g[i] += nmf_l1_grad(p[i], coef); }
  return s;
}

// microfacet backward, first stage (csrc/nmf_microfacet_bwd.cuh)
void hc_ggx_dr(const float* u, const float* V, const float* N, const float* r, int n, float* L, float* dL, float* H, float* dH) {
  for (int i = 0; i < n; ++i) {
    const NmfGGXdr g = nmf_ggx_sample_dr(u[2 * i], u[2 * i + 1], nmf_mk3(V[3 * i], V[3 * i + 1], V[3 * i + 2]),
                                         nmf_mk3(N[3 * i], N[3 * i + 1], N[3 * i + 2]), r[i]);
    L[3 * i] = g.L.x; L[3 * i + 1] = g.L.y; L[3 * i + 2] = g.L.z;
    dL[3 * i] = g.dL.x; dL[3 * i + 1] = g.dL.y; dL[3 * i + 2] = g.dL.z;
    H[3 * i] = g.H.x; H[3 * i + 1] = g.H.y; H[3 * i + 2] = g.H.z;
    dH[3 * i] = g.dH.x; dH[3 * i + 1] = g.dH.y; dH[3 * i + 2] = g.dH.z;
  }
}
void hc_ggx_dN(const float* u, const float* V, const float* N, const float* r, int n, int c, float* dL, float* dH) {
  for (int i = 0; i < n; ++i) {
    const NmfGGXdr g = nmf_ggx_sample_dN(u[2 * i], u[2 * i + 1], nmf_mk3(V[3 * i], V[3 * i + 1], V[3 * i + 2]),
                                         nmf_mk3(N[3 * i], N[3 * i + 1], N[3 * i + 2]), r[i], c);
    dL[3 * i] = g.dL.x; dL[3 * i + 1] = g.dL.y; dL[3 * i + 2] = g.dL.z;
    dH[3 * i] = g.dH.x; dH[3 * i + 1] = g.dH.y; dH[3 * i + 2] = g.dH.z;
  }
}
void hc_ggx_dV(const float* u, const float* V, const float* N, const float* r, int n, int c, float* dL, float* dH) {
  for (int i = 0; i < n; ++i) {
    const NmfGGXdr g = nmf_ggx_sample_dV(u[2 * i], u[2 * i + 1], nmf_mk3(V[3 * i], V[3 * i + 1], V[3 * i + 2]),
                                         nmf_mk3(N[3 * i], N[3 * i + 1], N[3 * i + 2]), r[i], c);
    dL[3 * i] = g.dL.x; dL[3 * i + 1] = g.dL.y; dL[3 * i + 2] = g.dL.z;
    dH[3 * i] = g.dH.x; dH[3 * i + 1] = g.dH.y; dH[3 * i + 2] = g.dH.z;
  }
}
void hc_fresnel_mix_bwd(const float* R0, const float* cost, const float* inc, const float* bw, const float* diff, const float* g, int n,
                        float* dR0, float* dinc, float* dbw, float* ddiff, float* dcost) {
  for (int i = 0; i < n; ++i)
    dcost[i] = nmf_fresnel_mix_bwd(R0 + 3 * i, cost[i], inc + 3 * i, bw + 3 * i, diff + 3 * i, g + 3 * i, dR0 + 3 * i, dinc + 3 * i,
                                   dbw + 3 * i, ddiff + 3 * i);
}
void hc_heads_bwd(const float* feat, const float* W, const float* b, float diffuse_mul, float diffuse_bias, float f0_bias,
                  float roughness_bias, const float* g_albedo, const float* g_f0, const float* g_rough, int n, float* dW, float* db,
                  float* dfeat) {
  for (int i = 0; i < n; ++i)
    nmf_heads_bwd(feat + 24 * i, W, b, diffuse_mul, diffuse_bias, f0_bias, roughness_bias, g_albedo + 3 * i, g_f0 + 3 * i, g_rough[i],
                  dW, db, dfeat + 24 * i);
}

void hc_env_bwd_map(int h, int w, float mipbias, const float* dirs, const float* sa, const float* g, int n, float* gsat, float* g_top,
                    float* g_bot) {
  for (int i = 0; i < n; ++i)
    nmf_env_lookup1_bwd_map(gsat, h, w, mipbias, nmf_mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), sa[i], g + 3 * i, g_top, g_bot);
}

double hc_env_mipbias_grad(const NmfScene* s, const float* dirs, const float* mip, const float* g, int n) {
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    float rgb[3], d[3];
    nmf_env_lookup1_dmipbias(s->env_sat, s->env_h, s->env_w, s->env_mipbias, s->env_top, s->env_bot,
                             nmf_mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), mip[i], rgb, d);
    acc += (double)(g[3 * i] * d[0] + g[3 * i + 1] * d[1] + g[3 * i + 2] * d[2]);
  }
  return acc;
}

void hc_env_lookup_d(const NmfScene* s, const float* dirs, const float* tangent, const float* mip, int n, float* rgb, float* drgb) {
  for (int i = 0; i < n; ++i) {
    const NmfDual3 d = nmf_d3(nmf_dmk(dirs[3 * i], tangent[3 * i]), nmf_dmk(dirs[3 * i + 1], tangent[3 * i + 1]),
                              nmf_dmk(dirs[3 * i + 2], tangent[3 * i + 2]));
    nmf_env_lookup1_d(s->env_sat, s->env_h, s->env_w, s->env_mipbias, s->env_top, s->env_bot, d, mip[i], rgb + 3 * i, drgb + 3 * i);
  }
}

// n bounce samples with m rays each: the composed reverse pass of one shading level without re-trace
void hc_bounce_samples_bwd(const NmfScene* s, const float* nfeat, const float* V, const float* N, const float* R0, const float* diffuse,
                           const float* rough, const float* u, int n, int m, const float* g, float* dR0, float* ddiffuse, float* drough,
                           float* dfeat, float* dw0t, float* db0, float* dw1t, float* db1, float* dw2t, float* db2, float* gsat,
                           float* g_top, float* g_bot) {
  NmfBrdfGrads bg{dw0t, db0, dw1t, db1, dw2t, db2};
  for (int i = 0; i < n; ++i)
    nmf_bounce_sample_bwd(*s, nfeat + 24 * i, nmf_mk3(V[3 * i], V[3 * i + 1], V[3 * i + 2]), nmf_mk3(N[3 * i], N[3 * i + 1], N[3 * i + 2]),
                          R0 + 3 * i, diffuse + 3 * i, rough[i], u + (size_t)2 * m * i, m, g + 3 * i, dR0 + 3 * i, ddiffuse + 3 * i,
                          drough + i, dfeat + 24 * i, bg, gsat, g_top, g_bot);
}

// ori_loss = sum w * min(v.n, 0)^2 at given points: value and the normal-path scatter (tensor_nerf.py:573-583)
double hc_ori_loss_bwd(const NmfScene* s, const float* xyz, const float* V, const float* w, int n, float* const* gpack,
                       float* const* glpack) {
  double total = 0.0;
  for (int i = 0; i < n; ++i) {
    float xn[3];
    nmf_normalize_xyz(*s, xyz + 4 * (size_t)i, xn);
    const NmfTaps t = nmf_vm_taps(*s, xn);
    float grad[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < 8; ++l) nmf_normal_lane(*s, t, l, grad);
    const nmf_v3 nn = nmf_normal_from_grad(*s, grad);
    const float vn = V[3 * i] * nn.x + V[3 * i + 1] * nn.y + V[3 * i + 2] * nn.z;
    if (!(vn < 0.f)) continue;
    total += (double)(w[i] * vn * vn);
    const float k2 = w[i] * 2.0f * vn;
    float dn[3] = {k2 * V[3 * i], k2 * V[3 * i + 1], k2 * V[3 * i + 2]}, dgrad[3];
    nmf_normal_vec_bwd(*s, grad, dn, dgrad);
    nmf_normal_bwd(*s, t, dgrad, gpack, glpack);
  }
  return total;
}

void hc_normals_bwd(const NmfScene* s, const float* xyz, int stride, const float* dn, int n, float* const* gpack, float* const* glpack) {
  for (int i = 0; i < n; ++i) nmf_normals_bwd_sample(*s, xyz + (size_t)stride * i, dn + 3 * (size_t)i, gpack, glpack);
}

void hc_normal_grad_finish(const float* gpack, int h, int w, const float* glpack, int n, const float* kx25, const float* ky25,
                           float* d_plane, float* d_line) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      for (int c = 0; c < 16; ++c) d_plane[((size_t)y * w + x) * 16 + c] += nmf_plane_grad_finish(gpack, h, w, kx25, ky25, y, x, c);
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < 16; ++c) d_line[(size_t)i * 16 + c] += nmf_line_grad_finish(glpack, n, ky25, i, c);
}

void hc_env_map_grad_finish(float* gsat, int h, int w, const float* g_top, const float* g_bot, const float* bg, float brightness,
                            float mul, float* out) {
  nmf_env_map_grad_finish(gsat, h, w, g_top, g_bot, bg, brightness, mul, out);
}

void hc_bounce_samples_tangent(const NmfScene* s, const float* nfeat, const float* V, const float* dV, const float* N, const float* R0,
                               const float* diffuse, const float* rough, const float* u, int n, int m, float* refl, float* drefl) {
  for (int i = 0; i < n; ++i) {
    const NmfDual3 Vd = nmf_d3(nmf_dmk(V[3 * i], dV[3 * i]), nmf_dmk(V[3 * i + 1], dV[3 * i + 1]), nmf_dmk(V[3 * i + 2], dV[3 * i + 2]));
    nmf_bounce_sample_tangent(*s, nfeat + 24 * i, Vd, nmf_mk3(N[3 * i], N[3 * i + 1], N[3 * i + 2]), R0 + 3 * i, diffuse + 3 * i, rough[i],
                              u + (size_t)2 * m * i, m, refl + 3 * i, drefl + 3 * i);
  }
}

void hc_upsample(const float* src, int C, int H, int W, float* dst, int H2, int W2) {
  for (int c = 0; c < C; ++c)
    for (int y = 0; y < H2; ++y)
      for (int x = 0; x < W2; ++x) {
        int x0, x1, y0, y1;
        float wx0, wx1, hy0, hy1;
        nmf_resize_tap(x, W, W2, &x0, &x1, &wx0, &wx1);
        nmf_resize_tap(y, H, H2, &y0, &y1, &hy0, &hy1);
        dst[((size_t)c * H2 + y) * W2 + x] = nmf_resize_pixel(src + (size_t)c * H * W, W, y0, y1, hy0, hy1, x0, x1, wx0, wx1);
      }
}
}
// Training forward + backward of the microfacet model with ONE re-traced level where EVERY bounce ray is re-traced
// (max_retrace_rays[0] >= number of bounce rays: no selection), on the host.
extern "C" void hc_train_microfacet_retrace(const NmfScene* s, const NmfTrain* tp, const float* rays, const float* gt, const NmfPlainGrads* g,
                                 float* d_head_w, float* d_head_b, float* dw0t, float* db0, float* dw1t, float* db1, float* dw2t, float* db2,
                                 float* gsat, float* g_top, float* g_bot, float* rgb_map, double* loss, int* n_samples, int detach_N,
                                 float* const* gpack, float* const* glpack) {
  const int n = tp->n_rays, S = s->n_steps;
  NmfBrdfGrads bgr{dw0t, db0, dw1t, db1, dw2t, db2};
  struct Smp { float z, dist, f, alpha, T, w, dw; NmfTaps t; float pos[3]; float feat[24], nfeat[24], albedo[3], f0[3], rough, E[3], refl[3];
               nmf_v3 V, Nf; int count; uint64_t skey; float ngrad[3], sgn; int ray0; };
  struct Sec { float o[3], d[3]; uint64_t key; float mip; std::vector<Smp> sm; float acc; double usum; float lin[3], bg[3], rgb[3]; };
  // ---- shared: march one ray (level 0: eval-key by (seed, id); level 1: key given) and shade heads; counts are set by the caller ----
  auto march = [&](const float* o, const float* d, uint64_t rkey, float near_, std::vector<Smp>& sm, float& acc, double* usum) {
    const float tmin = nmf_ray_tmin(o, d, s->aabb0, s->aabb1, near_, s->far);
    std::vector<float> z(S);
    std::vector<uint8_t> ok(S);
    double run = 0.0;
    for (int k = 0; k < S; ++k) {
      run += (double)nmf_jitter_step(rkey, k, s->stepsize);
      z[k] = NMF_ADD(tmin, (float)run);
      float p[3];
      nmf_step_pos(o, d, z[k], p);
      bool v = nmf_inside(p, s->aabb0, s->aabb1);
      if (v && s->has_occ) {
        float xn[3];
        nmf_normalize_xyz(*s, p, xn);
        v = nmf_occupied(s->occ_vox, s->occ_cell, s->ow, s->oh, s->od, s->opitch, xn[0], xn[1], xn[2]);
      }
      ok[k] = v;
      if (usum) *usum += (double)nmf_uniform(nmf_mix64(rkey, (uint64_t)k), NMF_STREAM_BOUNCE);
    }
    float T = 1.0f;
    acc = 0.f;
    for (int k = 0; k < S; ++k) {
      if (!ok[k]) continue;
      Smp q;
      q.z = z[k];
      q.dist = (k + 1 < S ? NMF_SUB(z[k + 1], q.z) : 0.f) * s->distance_scale;
      float xn[3];
      nmf_step_pos(o, d, q.z, q.pos);
      nmf_normalize_xyz(*s, q.pos, xn);
      q.t = nmf_vm_taps(*s, xn);
      q.f = 0.f;
      for (int gi = 0; gi < 4; ++gi) q.f += nmf_density_group(*s, q.t, gi);
      q.alpha = 1.0f - expf(-nmf_feature2density(q.f, s->density_shift) * q.dist);
      q.T = T;
      q.w = q.alpha * T;
      T *= 1.0f - q.alpha + 1e-10f;
      acc += q.w;
      float coef[72];
      nmf_app_coef(*s, q.t, coef);
      for (int oo = 0; oo < 24; ++oo) {
        float a = 0.f;
        for (int j = 0; j < 72; ++j) a += s->basis_t[j * 24 + oo] * coef[j];
        q.feat[oo] = a;
      }
      float grad[3] = {0.f, 0.f, 0.f};
      for (int l = 0; l < 8; ++l) nmf_normal_lane(*s, q.t, l, grad);
      const nmf_v3 nrm = nmf_normal_from_grad(*s, grad);
      float lin11[11];
      for (int h = 0; h < 11; ++h) {
        float v = s->head_b[h];
        for (int i = 0; i < 24; ++i) v += s->head_w[h * 24 + i] * q.feat[i];
        lin11[h] = v;
      }
      float sh[9];
      nmf_sh9(nrm, sh);
      for (int c = 0; c < 3; ++c) {
        q.albedo[c] = nmf_clampf(nmf_sigmoid(s->diffuse_mul * lin11[c] + s->diffuse_bias), 0.f, 1.f);
        q.f0[c] = nmf_sigmoid(lin11[6 + c] + s->f0_bias);
        float e = 0.f;
        for (int i = 0; i < 9; ++i) e += s->sh_conv[i * 3 + c] * sh[i];
        q.E[c] = e;
        q.refl[c] = 0.f;
      }
      q.rough = nmf_clampf(nmf_sigmoid(lin11[9] + s->roughness_bias) / 2.0f, 1e-2f, 1.0f);
      q.V = nmf_mk3(-d[0], -d[1], -d[2]);
      const float vn = nmf_dot(q.V, nrm);
      q.sgn = vn > 0.f ? 1.f : (vn < 0.f ? -1.f : 0.f);
      q.Nf = nmf_mk3(nrm.x * q.sgn, nrm.y * q.sgn, nrm.z * q.sgn);
      for (int c = 0; c < 3; ++c) q.ngrad[c] = grad[c];
      q.skey = nmf_mix64(rkey, (uint64_t)k);
      const uint64_t nseed = nmf_noise_seed(q.skey);
      for (uint32_t pp = 0; pp < 12; ++pp) {
        float n0, n1;
        nmf_noise_pair(nseed, pp, &n0, &n1);
        q.nfeat[2 * pp] = q.feat[2 * pp] + s->anoise * n0;
        q.nfeat[2 * pp + 1] = q.feat[2 * pp + 1] + s->anoise * n1;
      }
      q.count = 0; q.ray0 = -1; q.dw = 0.f;
      sm.push_back(q);
    }
  };
  auto sample_u = [&](const Smp& q, std::vector<float>& u) {
    const float offu = 0.25f * nmf_uniform(q.skey, NMF_STREAM_OFF_U), offv = 0.25f * nmf_uniform(q.skey, NMF_STREAM_OFF_V);
    u.resize(2 * (size_t)q.count);
    for (int j = 0; j < q.count; ++j) { u[2 * j] = nmf_wrap01(s->sobol[2 * j] + offu); u[2 * j + 1] = nmf_wrap01(s->sobol[2 * j + 1] + offv); }
  };
  // ---- level 0 ----
  std::vector<std::vector<Smp>> L0(n);
  std::vector<float> acc0(n);
  std::vector<Sec> secs;
  for (int r = 0; r < n; ++r) {
    const uint64_t rkey = nmf_primary_key(tp->seed, tp->ray_id0 + (uint64_t)r);
    march(rays + 6 * r, rays + 6 * r + 3, rkey, s->near, L0[r], acc0[r], nullptr);
    for (Smp& q : L0[r]) {
      const float kf = floorf(q.w * (float)s->rays_per_ray + nmf_uniform(q.skey, NMF_STREAM_BOUNCE) - 0.5f);
      q.count = (int)nmf_clampf(kf, 0.f, (float)NMF_MAX_BOUNCE);
      if (q.count == 0) continue;
      q.ray0 = (int)secs.size();
      std::vector<float> u;
      sample_u(q, u);
      for (int j = 0; j < q.count; ++j) {
        const NmfGGX fw = nmf_ggx_sample(u[2 * j], u[2 * j + 1], q.V, q.Nf, q.rough);
        Sec sc;
        sc.o[0] = q.pos[0] + fw.L.x * 5e-3f; sc.o[1] = q.pos[1] + fw.L.y * 5e-3f; sc.o[2] = q.pos[2] + fw.L.z * 5e-3f;
        sc.d[0] = fw.L.x; sc.d[1] = fw.L.y; sc.d[2] = fw.L.z;
        sc.key = nmf_mix64(q.skey, (uint64_t)j + NMF_STREAM_RAY0);
        sc.mip = -logf((float)q.count) - fw.logpdf;
        sc.usum = 0.0;
        secs.push_back(sc);
      }
    }
  }
  // ---- level 1: march, chunk totals, budgeted counts, shading ----
  long long n1 = 0;
  double wsum = 0.0;
  for (Sec& sc : secs) {
    march(sc.o, sc.d, sc.key, NMF_MUL(3.0f, s->stepsize), sc.sm, sc.acc, &sc.usum);
    n1 += (long long)sc.sm.size();
    wsum += (double)sc.acc + 1e-3 * sc.usum;
  }
  n_samples[0] = 0;
  for (int r = 0; r < n; ++r) n_samples[0] += (int)L0[r].size();
  n_samples[1] = (int)n1;
  const float wsumf = fmaxf((float)wsum, 1e-3f);
  const int budget = s->max_brdf_rays1;
  const int Nb = budget - (int)n1;
  auto shade_fwd = [&](Smp& q) {        // all rays to the environment
    for (int c = 0; c < 3; ++c) q.refl[c] = 0.f;
    if (q.count == 0) return;
    std::vector<float> u;
    sample_u(q, u);
    for (int j = 0; j < q.count; ++j) {
      const NmfGGX fw = nmf_ggx_sample(u[2 * j], u[2 * j + 1], q.V, q.Nf, q.rough);
      const float mip = -logf((float)q.count) - fw.logpdf;
      float x[66], bw[3], inc[3];
      nmf_brdf_input(q.nfeat, fw.half_l, fw.diff_l, q.rough, x);
      nmf_brdf_row_fwd_bwd(x, s->brdf_w0t, s->brdf_b0, s->brdf_w1t, s->brdf_b1, s->brdf_w2t, s->brdf_b2, s->brdf_bias, nullptr, bw,
                           nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
      nmf_env_lookup1(s->env_sat, s->env_h, s->env_w, s->env_mipbias, s->env_top, s->env_bot, fw.L, mip, inc);
      const float cost = fabsf(nmf_dot(q.V, fw.H));
      for (int c = 0; c < 3; ++c) {
        const float F = nmf_fresnel(q.f0[c], cost);
        q.refl[c] += (F * inc[c] * bw[c] + (1.0f - F) * q.albedo[c] * q.E[c]) / (float)q.count;
      }
    }
  };
  for (Sec& sc : secs) {
    for (int c = 0; c < 3; ++c) sc.lin[c] = 0.f;
    for (Smp& q : sc.sm) {
      const float U = nmf_uniform(q.skey, NMF_STREAM_BOUNCE);
      const float wj = q.w + 1e-3f * U;
      const float kf = Nb > 0 ? floorf(wj / wsumf * (float)Nb + 1.0f) : floorf(wj / wsumf * (float)budget + 0.5f);
      q.count = (int)nmf_clampf(kf, 0.f, (float)NMF_MAX_BOUNCE);
      shade_fwd(q);
      for (int c = 0; c < 3; ++c) sc.lin[c] += q.w * q.refl[c];
    }
    nmf_env_lookup1(s->env_sat, s->env_h, s->env_w, s->env_mipbias, s->env_top, s->env_bot, nmf_mk3(sc.d[0], sc.d[1], sc.d[2]), sc.mip, sc.bg);
    for (int c = 0; c < 3; ++c) sc.rgb[c] = sc.lin[c] + (1.0f - sc.acc) * sc.bg[c];
  }
  // ---- level 0 reflect + loss + reverse ----
  loss[0] = loss[1] = 0.0;
  const float bgw[3] = {1.f, 1.f, 1.f};
  // reverse of one sample's shading given d loss / d reflect (gre): level-independent part (heads, BRDF, features, normals)
  auto factors_bwd = [&](const Smp& q, const float* dnfeat, const float* dR0, const float* ddiff, float drough, const float* dNf) {
    float g_alb[3], dfeat_h[24];
    for (int c = 0; c < 3; ++c) g_alb[c] = ddiff[c] * q.E[c];
    nmf_heads_bwd(q.feat, s->head_w, s->head_b, s->diffuse_mul, s->diffuse_bias, s->f0_bias, s->roughness_bias, g_alb, dR0, drough,
                  d_head_w, d_head_b, dfeat_h);
    float coef[72], dcoef[72];
    nmf_app_coef(*s, q.t, coef);
    for (int j = 0; j < 72; ++j) {
      float a = 0.f;
      for (int oo = 0; oo < 24; ++oo) {
        const float df = dnfeat[oo] + dfeat_h[oo];
        g->basis_t[j * 24 + oo] += coef[j] * df;
        a += s->basis_t[j * 24 + oo] * df;
      }
      dcoef[j] = a;
    }
    nmf_app_bwd(*s, q.t, dcoef, g->a_plane, g->a_line);
    if (dNf) {
      float dn[3] = {q.sgn * dNf[0], q.sgn * dNf[1], q.sgn * dNf[2]}, dgrad[3];
      nmf_normal_vec_bwd(*s, q.ngrad, dn, dgrad);
      nmf_normal_bwd(*s, q.t, dgrad, gpack, glpack);
    }
  };
  auto composite_bwd = [&](std::vector<Smp>& sm) {
    float suffix = 0.f;
    for (int i = (int)sm.size() - 1; i >= 0; --i) {
      const Smp& q = sm[i];
      const float dsigma = nmf_composite_bwd(q.dw, q.T, q.alpha, q.dist, suffix);
      suffix += q.dw * q.w;
      const float df = dsigma * nmf_feature2density_grad(q.f, s->density_shift);
      if (df != 0.f) nmf_density_bwd(*s, q.t, df, g->d_plane, g->d_line);
    }
  };
  // tangent of a secondary ray's radiance along a tangent dL of its direction (V1 = -d1; background lookup)
  auto sec_tangent = [&](const Sec& sc, nmf_v3 dL, float* out) {
    out[0] = out[1] = out[2] = 0.f;
    const NmfDual3 Vd = nmf_d3(nmf_dmk(-sc.d[0], -dL.x), nmf_dmk(-sc.d[1], -dL.y), nmf_dmk(-sc.d[2], -dL.z));
    for (const Smp& q : sc.sm) {
      if (q.count == 0) continue;
      std::vector<float> u;
      sample_u(q, u);
      float diffuse[3], refl[3], drefl[3];
      for (int c = 0; c < 3; ++c) diffuse[c] = q.albedo[c] * q.E[c];
      nmf_bounce_sample_tangent(*s, q.nfeat, Vd, q.Nf, q.f0, diffuse, q.rough, u.data(), q.count, refl, drefl);
      for (int c = 0; c < 3; ++c) out[c] += q.w * drefl[c];
    }
    const NmfDual3 Dd = nmf_d3(nmf_dmk(sc.d[0], dL.x), nmf_dmk(sc.d[1], dL.y), nmf_dmk(sc.d[2], dL.z));
    float bgv[3], dbg[3];
    nmf_env_lookup1_d(s->env_sat, s->env_h, s->env_w, s->env_mipbias, s->env_top, s->env_bot, Dd, sc.mip, bgv, dbg);
    for (int c = 0; c < 3; ++c) out[c] += (1.0f - sc.acc) * dbg[c];
  };
  for (int r = 0; r < n; ++r) {
    float lin[3] = {0.f, 0.f, 0.f};
    std::vector<std::vector<float>> bws(L0[r].size());
    for (size_t si = 0; si < L0[r].size(); ++si) {
      Smp& q = L0[r][si];
      if (q.count == 0) continue;
      std::vector<float> u;
      sample_u(q, u);
      bws[si].resize(3 * (size_t)q.count);
      for (int j = 0; j < q.count; ++j) {
        const NmfGGX fw = nmf_ggx_sample(u[2 * j], u[2 * j + 1], q.V, q.Nf, q.rough);
        float x[66];
        nmf_brdf_input(q.nfeat, fw.half_l, fw.diff_l, q.rough, x);
        nmf_brdf_row_fwd_bwd(x, s->brdf_w0t, s->brdf_b0, s->brdf_w1t, s->brdf_b1, s->brdf_w2t, s->brdf_b2, s->brdf_bias, nullptr,
                             &bws[si][3 * j], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        const float cost = fabsf(nmf_dot(q.V, fw.H));
        const Sec& sc = secs[q.ray0 + j];
        for (int c = 0; c < 3; ++c) {
          const float F = nmf_fresnel(q.f0[c], cost);
          q.refl[c] += (F * sc.rgb[c] * bws[si][3 * j + c] + (1.0f - F) * q.albedo[c] * q.E[c]) / (float)q.count;
        }
      }
      for (int c = 0; c < 3; ++c) lin[c] += q.w * q.refl[c];
    }
    float gl[3], ga;
    loss[0] += nmf_train_loss_ray(lin, acc0[r], bgw, gt + 3 * r, tp->lambda_pred, rgb_map + 3 * r, gl, &ga);
    loss[1] += acc0[r];
    for (size_t si = 0; si < L0[r].size(); ++si) {
      Smp& q = L0[r][si];
      q.dw = ga;
      for (int c = 0; c < 3; ++c) q.dw += gl[c] * q.refl[c];
      if (q.count == 0) continue;
      std::vector<float> u;
      sample_u(q, u);
      float diffuse[3], gm[3], dR0[3] = {0.f, 0.f, 0.f}, ddiff[3] = {0.f, 0.f, 0.f}, dnfeat[24], dNf[3] = {0.f, 0.f, 0.f};
      float dr = 0.f;
      for (int k = 0; k < 24; ++k) dnfeat[k] = 0.f;
      for (int c = 0; c < 3; ++c) { diffuse[c] = q.albedo[c] * q.E[c]; gm[c] = q.w * gl[c] / (float)q.count; }
      for (int j = 0; j < q.count; ++j) {
        Sec& sc = secs[q.ray0 + j];
        const float u1 = u[2 * j], u2 = u[2 * j + 1];
        const NmfGGX fw = nmf_ggx_sample(u1, u2, q.V, q.Nf, q.rough);
        const NmfGGXdr dg = nmf_ggx_sample_dr(u1, u2, q.V, q.Nf, q.rough);
        const float vh = nmf_dot(q.V, dg.H), cost = fabsf(vh), svh = vh > 0.f ? 1.f : (vh < 0.f ? -1.f : 0.f);
        float a_R0[3], a_inc[3], a_bw[3], a_diff[3];
        const float dcost = nmf_fresnel_mix_bwd(q.f0, cost, sc.rgb, &bws[si][3 * j], diffuse, gm, a_R0, a_inc, a_bw, a_diff);
        for (int c = 0; c < 3; ++c) { dR0[c] += a_R0[c]; ddiff[c] += a_diff[c]; }
        float tang[3];
        sec_tangent(sc, dg.dL, tang);
        dr += dcost * svh * nmf_dot(q.V, dg.dH) + a_inc[0] * tang[0] + a_inc[1] * tang[1] + a_inc[2] * tang[2];
        if (!detach_N)
          for (int c = 0; c < 3; ++c) {
            const NmfGGXdr dn = nmf_ggx_sample_dN(u1, u2, q.V, q.Nf, q.rough, c);
            sec_tangent(sc, dn.dL, tang);
            dNf[c] += dcost * svh * nmf_dot(q.V, dn.dH) + a_inc[0] * tang[0] + a_inc[1] * tang[1] + a_inc[2] * tang[2];
          }
        float x[66], bw2[3];
        nmf_brdf_input(q.nfeat, fw.half_l, fw.diff_l, q.rough, x);
        nmf_brdf_row_fwd_bwd(x, s->brdf_w0t, s->brdf_b0, s->brdf_w1t, s->brdf_b1, s->brdf_w2t, s->brdf_b2, s->brdf_bias, a_bw, bw2, bgr.w0t,
                             bgr.b0, bgr.w1t, bgr.b1, bgr.w2t, bgr.b2, dnfeat);
        // ---- level-1 reverse with the linear upstream a_inc on this secondary ray's radiance ----
        float gbg = 0.f;
        for (int c = 0; c < 3; ++c) gbg += a_inc[c] * sc.bg[c];
        float gbg3[3] = {(1.0f - sc.acc) * a_inc[0], (1.0f - sc.acc) * a_inc[1], (1.0f - sc.acc) * a_inc[2]};
        nmf_env_lookup1_bwd_map(gsat, s->env_h, s->env_w, s->env_mipbias, nmf_mk3(sc.d[0], sc.d[1], sc.d[2]), sc.mip, gbg3, g_top, g_bot);
        for (Smp& q1 : sc.sm) {
          q1.dw = -gbg;
          for (int c = 0; c < 3; ++c) q1.dw += a_inc[c] * q1.refl[c];
          if (q1.count == 0) continue;
          std::vector<float> u1v;
          sample_u(q1, u1v);
          float diffuse1[3], gre1[3], dR01[3], ddiff1[3], drough1, dnfeat1[24], dNf1[3];
          for (int c = 0; c < 3; ++c) { diffuse1[c] = q1.albedo[c] * q1.E[c]; gre1[c] = q1.w * a_inc[c]; }
          nmf_bounce_sample_bwd(*s, q1.nfeat, q1.V, q1.Nf, q1.f0, diffuse1, q1.rough, u1v.data(), q1.count, gre1, dR01, ddiff1, &drough1,
                                dnfeat1, bgr, gsat, g_top, g_bot, detach_N ? nullptr : dNf1);
          factors_bwd(q1, dnfeat1, dR01, ddiff1, drough1, detach_N ? nullptr : dNf1);
        }
        composite_bwd(sc.sm);
      }
      factors_bwd(q, dnfeat, dR0, ddiff, dr, detach_N ? nullptr : dNf);
    }
    composite_bwd(L0[r]);
  }
}
