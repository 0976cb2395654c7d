// TEST INFRASTRUCTURE -- compiles the per-element math of the kernels (nmf_b200/csrc/nmf_math.cuh,
// nmf_field.cuh) for the host so that `pytest -m "not gpu"` can check it against the oracle without a GPU.
// It is not part of the product and nothing under nmf_b200/ loads it.
//   g++ -O2 -ffp-contract=off -shared -fPIC -o libnmf_hostcheck.so hostcheck.cpp
#include "../../nmf_b200/csrc/nmf_field.cuh"

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
}
