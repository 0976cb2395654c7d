// Environment-map reverse pass (DESIGN.md section 9, row "environment"): the first reverse-pass KERNELS of the microfacet model.
//   reference: IntegralEquirect.activation_fn / calc_sat / forward (modules/integral_equirect.py:263-273, 409-504) under autograd.
//   forward:   act = exp(min(brightness + mul * bg_mat, 20)),  SAT = cumsum_y cumsum_x (act / 1000),
//              rgb = 1000 * sum_boxes sign * bilinear(SAT, corner) / size     (pole rows: the mean of act's first / last row)
//   backward:  k_env_bwd_scatter   one thread per lookup: the forward's own box walk (nmf_env_integrate) with a scattering tap
//                                  (nmf_env_lookup1_bwd_map, host-checked against autograd in tests/test_hostmath.py);
//                                  fp32 atomics into an [h][w][4] image that stays L2-resident (8 MB at 512 x 1024)
//              k_env_bwd_scan_x    reverse inclusive prefix sum along x, one CTA per map row (adjoint of cumsum over x)
//              k_env_bwd_finish    reverse prefix sum along y (16 row segments per column, segment totals through shared memory)
//                                  fused with the pole-row means and the exp activation's chain rule: d bg_mat (accumulated),
//                                  d brightness, d mul; brightness / mul by value or from device memory (training)
// Both passes stream the 8 MB image once: HBM/L2-bound, 2 x (read + write) of h*w*16 bytes + the (3,h,w) parameter and gradient.
#include <cuda_runtime.h>
#include <stdint.h>
#include "nmf_microfacet_bwd.cuh"

#define FULL 0xffffffffu
#define CKL() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)
#define SCAN_T 256

__global__ void k_env_bwd_scatter(const NmfScene s, const float* __restrict__ dirs, const float* __restrict__ mip,
                                  const float* __restrict__ g, int n, float* gsat, float* g_top, float* g_bot) {
  const NmfEnvDyn ed = nmf_env_dyn_load(s);     // mipbias and pole means: by value or from NmfScene.env_dyn
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi[3] = {g[3 * i], g[3 * i + 1], g[3 * i + 2]};
  if (gi[0] == 0.f && gi[1] == 0.f && gi[2] == 0.f) return;
  nmf_env_lookup1_bwd_map(gsat, s.env_h, s.env_w, ed.mipbias, nmf_mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), mip[i], gi,
                          g_top, g_bot);
}

// One CTA per row: thread t owns the `per` consecutive texels [t*per, (t+1)*per); local suffix sums right to left, a suffix
// scan of the per-thread totals in shared memory, then the carry of everything to the right is added.
__global__ void __launch_bounds__(SCAN_T) k_env_bwd_scan_x(float4* __restrict__ gsat, int w) {
  __shared__ float tot[2][SCAN_T][3];
  float4* row = gsat + (size_t)blockIdx.x * w;
  const int t = threadIdx.x;
  const int per = (w + SCAN_T - 1) / SCAN_T;
  const int x0 = t * per, x1 = min(x0 + per, w);
  float run[3] = {0.f, 0.f, 0.f};
  for (int x = x1 - 1; x >= x0; --x) {
    float4 v = row[x];
    run[0] += v.x; run[1] += v.y; run[2] += v.z;
    row[x] = make_float4(run[0], run[1], run[2], 0.f);
  }
  for (int k = 0; k < 3; ++k) tot[0][t][k] = run[k];
  __syncthreads();
  int cur = 0;
  for (int off = 1; off < SCAN_T; off <<= 1) {          // inclusive suffix scan over threads (Hillis-Steele)
    for (int k = 0; k < 3; ++k) tot[cur ^ 1][t][k] = tot[cur][t][k] + (t + off < SCAN_T ? tot[cur][t + off][k] : 0.f);
    cur ^= 1;
    __syncthreads();
  }
  if (t + 1 < SCAN_T && x0 < x1) {
    const float c0 = tot[cur][t + 1][0], c1 = tot[cur][t + 1][1], c2 = tot[cur][t + 1][2];
    for (int x = x0; x < x1; ++x) {
      float4 v = row[x];
      row[x] = make_float4(v.x + c0, v.y + c1, v.z + c2, 0.f);
    }
  }
}

// Reverse prefix sum along y.  CTA = 16 columns x 16 row segments, a thread owns one column segment with all three channels
// (float4 rows): segment totals meet in shared memory, then the segment is walked bottom to top with the carry of everything
// below it, the pole-row means are added and d act -> d bg_mat / d brightness / d mul is applied where the clip at 20 is
// open.  (One thread per (column, channel) walking all rows took 524 us at 512 x 1024: 3072 threads on 148 SMs.)
#define ENVB_SEG 16
__global__ void __launch_bounds__(256) k_env_bwd_finish(const float* __restrict__ gsat, int h, int w, const float* __restrict__ g_top,
                                                        const float* __restrict__ g_bot, const float* __restrict__ bg, float brightness,
                                                        float mul, const float* __restrict__ scalars_dev, float* __restrict__ d_bg,
                                                        float* d_brightness, float* d_mul) {
  __shared__ float tot[ENVB_SEG][16][3];
  if (scalars_dev) { brightness = scalars_dev[0]; mul = scalars_dev[1]; }      // device-resident parameters (training)
  const int col = threadIdx.x & 15, seg = threadIdx.x >> 4;
  const int x = blockIdx.x * 16 + col;
  const int R = (h + ENVB_SEG - 1) / ENVB_SEG;
  const int y0 = seg * R, y1 = min(y0 + R, h);
  const float4* g4 = (const float4*)gsat;
  float run[3] = {0.f, 0.f, 0.f};
  if (x < w)
    for (int y = y1 - 1; y >= y0; --y) {
      const float4 v = g4[(size_t)y * w + x];
      run[0] += v.x; run[1] += v.y; run[2] += v.z;
    }
  for (int k = 0; k < 3; ++k) tot[seg][col][k] = run[k];
  __syncthreads();
  float sb = 0.f, sm = 0.f;
  if (x < w) {
    for (int k = 0; k < 3; ++k) run[k] = 0.f;
    for (int sgm = ENVB_SEG - 1; sgm > seg; --sgm)
      for (int k = 0; k < 3; ++k) run[k] += tot[sgm][col][k];
    const float top[3] = {g_top[0] / (float)w, g_top[1] / (float)w, g_top[2] / (float)w};
    const float bot[3] = {g_bot[0] / (float)w, g_bot[1] / (float)w, g_bot[2] / (float)w};
    for (int y = y1 - 1; y >= y0; --y) {
      const float4 v = g4[(size_t)y * w + x];
      run[0] += v.x; run[1] += v.y; run[2] += v.z;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float d = run[k];
        if (y == 0) d += top[k];
        if (y == h - 1) d += bot[k];
        const size_t o = ((size_t)k * h + y) * w + x;
        const float b = bg[o];
        const float pre = brightness + mul * b;
        if (pre <= 20.0f) {
          const float da = d * expf(pre);
          d_bg[o] += da * mul;
          sb += da;
          sm += da * b;
        }
      }
    }
  }
  if (d_brightness || d_mul) {
    for (int o = 16; o; o >>= 1) {
      sb += __shfl_xor_sync(FULL, sb, o);
      sm += __shfl_xor_sync(FULL, sm, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (d_brightness) atomicAdd(d_brightness, sb);
      if (d_mul) atomicAdd(d_mul, sm);
    }
  }
}

// d loss / d mipbias: per lookup the forward-mode pass nmf_env_lookup1_dmipbias (unit tangent on the bias: the box size moves,
// the taps' bilinear weights and 1 / size with it), dotted with the upstream; warp sum, one atomic per warp.
__global__ void k_env_bwd_mipbias(const NmfScene s, const float* __restrict__ dirs, const float* __restrict__ mip,
                                  const float* __restrict__ g, int n, float* d_mipbias) {
  const NmfEnvDyn ed = nmf_env_dyn_load(s);     // mipbias and pole means: by value or from NmfScene.env_dyn
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (i < n) {
    const float gi[3] = {g[3 * i], g[3 * i + 1], g[3 * i + 2]};
    if (gi[0] != 0.f || gi[1] != 0.f || gi[2] != 0.f) {
      float rgb[3], d[3];
      nmf_env_lookup1_dmipbias(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot,
                               nmf_mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), mip[i], rgb, d);
      acc = gi[0] * d[0] + gi[1] * d[1] + gi[2] * d[2];
    }
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
  if ((threadIdx.x & 31) == 0 && acc != 0.f) atomicAdd(d_mipbias, acc);
}

extern "C" int nmf_env_lookup_bwd_mipbias(const NmfScene* scene, const float* dirs, const float* mip, const float* g, int n,
                                          float* d_mipbias, void* stream) {
  if (!scene || !scene->env_sat || !dirs || !mip || !g || !d_mipbias || n < 0) return NMF_E_ARG;
  if (n == 0) return NMF_OK;
  k_env_bwd_mipbias<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*scene, dirs, mip, g, n, d_mipbias);
  CKL();
  return NMF_OK;
}

extern "C" int nmf_env_lookup_bwd_scatter(const NmfScene* scene, const float* dirs, const float* mip, const float* g, int n,
                                          float* gsat, void* stream) {
  if (!scene || !dirs || !mip || !g || !gsat || n < 0 || scene->env_h <= 0 || scene->env_w <= 0) return NMF_E_ARG;
  if (n == 0) return NMF_OK;
  float* poles = gsat + (size_t)scene->env_h * scene->env_w * 4;
  k_env_bwd_scatter<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*scene, dirs, mip, g, n, gsat, poles, poles + 4);
  CKL();
  return NMF_OK;
}

static int env_bwd_finish_impl(float* gsat, int h, int w, const float* bg_mat, float brightness, float mul, const float* scalars_dev,
                               float* d_bg_mat, float* d_brightness, float* d_mul, void* stream) {
  if (!gsat || !bg_mat || !d_bg_mat || h <= 0 || w <= 0) return NMF_E_ARG;
  const float* poles = gsat + (size_t)h * w * 4;
  k_env_bwd_scan_x<<<h, SCAN_T, 0, (cudaStream_t)stream>>>((float4*)gsat, w);
  CKL();
  k_env_bwd_finish<<<(w + 15) / 16, 256, 0, (cudaStream_t)stream>>>(gsat, h, w, poles, poles + 4, bg_mat, brightness, mul, scalars_dev,
                                                                       d_bg_mat, d_brightness, d_mul);
  CKL();
  return NMF_OK;
}
extern "C" int nmf_env_lookup_bwd_finish(float* gsat, int h, int w, const float* bg_mat, float brightness, float mul, float* d_bg_mat,
                                       float* d_brightness, float* d_mul, void* stream) {
  return env_bwd_finish_impl(gsat, h, w, bg_mat, brightness, mul, nullptr, d_bg_mat, d_brightness, d_mul, stream);
}
// the same with brightness / mul read from device memory (scalars_dev = { brightness, mul, ... }): no host copy of the parameters
extern "C" int nmf_env_lookup_bwd_finish_dev(float* gsat, int h, int w, const float* bg_mat, const float* scalars_dev, float* d_bg_mat,
                                           float* d_brightness, float* d_mul, void* stream) {
  if (!scalars_dev) return NMF_E_ARG;
  return env_bwd_finish_impl(gsat, h, w, bg_mat, 0.f, 1.f, scalars_dev, d_bg_mat, d_brightness, d_mul, stream);
}
