// Reverse pass of the material heads (DESIGN.md section 9, row "material heads"): what autograd does to
// RandHydraMLPDiffuse.forward (modules/render_modules.py:519-574, the sigmoid heads :553-560) w.r.t. the head weights, their
// biases and the 24-d appearance feature, for upstream gradients of albedo, f0 and the roughness.
//   k_heads_bwd   CTA = 288 threads over a tile of 256 samples.  Phase 1, thread per sample: nmf_heads_dlin (host-checked,
//                 csrc/nmf_microfacet_bwd.cuh) with the 11 x 24 weights in shared memory; d feat = W^T dlin is written straight
//                 out; dlin and the feature are parked in shared memory.  Phase 2, thread per weight (264) / bias (11): the
//                 tile's outer-product sum  dW[h][k] = sum_j dlin_j[h] feat_j[k]  from shared memory, one atomic per weight
//                 and tile (275 atomics per 256 samples instead of 275 per sample).
// Streams feat (96 B), the upstream (28 B) and d feat (96 B) per sample once: HBM-bound, 220 B per sample.
#include <cuda_runtime.h>
#include <stdint.h>
#include "nmf_microfacet_bwd.cuh"

#define CKL() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)
#define HB_TILE 256
#define HB_T 288

__global__ void __launch_bounds__(HB_T) k_heads_bwd(const NmfScene s, const float* __restrict__ feat, const float* __restrict__ g_albedo,
                                                    const float* __restrict__ g_f0, const float* __restrict__ g_rough, int n,
                                                    float* d_head_w, float* d_head_b, float* __restrict__ d_feat) {
  __shared__ float sW[11 * 24], sB[11];
  __shared__ float sF[HB_TILE][25];      // 25: the per-sample rows start in different banks
  __shared__ float sD[HB_TILE][11];
  const int t = threadIdx.x;
  for (int q = t; q < 11 * 24; q += HB_T) sW[q] = s.head_w[q];
  if (t < 11) sB[t] = s.head_b[t];
  __syncthreads();
  if (t < HB_TILE) {
    const size_t i = (size_t)blockIdx.x * HB_TILE + t;
    float f[24], dlin[11];
    if (i < (size_t)n) {
      for (int k = 0; k < 24; ++k) f[k] = feat[i * 24 + k];
      const float ga[3] = {g_albedo[i * 3], g_albedo[i * 3 + 1], g_albedo[i * 3 + 2]};
      const float gf[3] = {g_f0[i * 3], g_f0[i * 3 + 1], g_f0[i * 3 + 2]};
      nmf_heads_dlin(f, sW, sB, s.diffuse_mul, s.diffuse_bias, s.f0_bias, s.roughness_bias, ga, gf, g_rough[i], dlin);
      for (int k = 0; k < 24; ++k) {
        float a = 0.f;
        for (int h = 0; h < 11; ++h) a += dlin[h] * sW[h * 24 + k];
        d_feat[i * 24 + k] = a;
      }
    } else {
      for (int k = 0; k < 24; ++k) f[k] = 0.f;
      for (int h = 0; h < 11; ++h) dlin[h] = 0.f;
    }
    for (int k = 0; k < 24; ++k) sF[t][k] = f[k];
    for (int h = 0; h < 11; ++h) sD[t][h] = dlin[h];
  }
  __syncthreads();
  if (t < 11 * 24) {
    const int h = t / 24, k = t - 24 * h;
    float a = 0.f;
#pragma unroll 8
    for (int j = 0; j < HB_TILE; ++j) a += sD[j][h] * sF[j][k];
    if (a != 0.f) atomicAdd(d_head_w + t, a);
  } else if (t < 11 * 24 + 11) {
    const int h = t - 11 * 24;
    float a = 0.f;
    for (int j = 0; j < HB_TILE; ++j) a += sD[j][h];
    if (a != 0.f) atomicAdd(d_head_b + h, a);
  }
}

extern "C" int nmf_material_heads_bwd(const NmfScene* scene, const float* feat, const float* g_albedo, const float* g_f0,
                                      const float* g_rough, int n, float* d_head_w, float* d_head_b, float* d_feat, void* stream) {
  if (!scene || !scene->head_w || !scene->head_b || !feat || !g_albedo || !g_f0 || !g_rough || !d_head_w || !d_head_b || !d_feat || n < 0)
    return NMF_E_ARG;
  if (n == 0) return NMF_OK;
  k_heads_bwd<<<(n + HB_TILE - 1) / HB_TILE, HB_T, 0, (cudaStream_t)stream>>>(*scene, feat, g_albedo, g_f0, g_rough, n, d_head_w,
                                                                              d_head_b, d_feat);
  CKL();
  return NMF_OK;
}
