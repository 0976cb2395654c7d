// nmf_b200 -- scene re-pack kernels: what has to be rebuilt from the parameters after every optimiser step of a training
// run (once per ITERATION), hand-written instead of cuDNN / torch scans:
//   k_pack_plane / k_pack_line   reference-layout factors (1,C,H,W) / (1,C,N,1) -> the channel-last gather layouts of
//                                NmfScene, with the smoothed-difference planes of modules/grid_sample_Cinf.py:218-242 (a 5x5
//                                cross-correlation of the WHOLE plane, zero padding 2; the reference recomputes it inside every
//                                compute_normals call) -- the forward twin of k_normals_bwd_planes (csrc/nmf_normals_bwd.cu)
//   k_env_act_scan_y, k_env_scan_x   modules/integral_equirect.py:263-273, 431-433: act = exp(min(brightness + mul * bg, 20)),
//                                SAT = cumsum_x(cumsum_y(act / 1000)) with ATen's CPU semantics (fp64 accumulation, every prefix
//                                rounded to fp32 after EACH of the two scans), channel-last [h][w][4]; pole-row means
//   k_occ_pool_pack, k_occ_cells, k_occ_coarse   samplers/alphagrid.py:249-276: 3^3 max-pool + threshold of the dense alpha
//                                lattice -> voxel bit-field, per-cell OR of the 8 corners, conservative coarse field
// All of them stream their input once; bytes per element are stated at each kernel.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/nmf_b200.h"

#define FULLM 0xffffffffu
#define CKL() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

// thread per (texel, channel), channel fastest: 4 B read (+ 2 x 25 cached taps) / 4 or 16 B written per element
template <int C>
__global__ void __launch_bounds__(256) k_pack_plane(const float* __restrict__ src, int H, int W, const float* __restrict__ kx25,
                                                    const float* __restrict__ ky25, float* __restrict__ val, float* __restrict__ pack) {
  __shared__ float kx[25], ky[25];
  if (pack && threadIdx.x < 25) { kx[threadIdx.x] = kx25[threadIdx.x]; ky[threadIdx.x] = ky25[threadIdx.x]; }
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t texel = idx / C;
  const int c = (int)(idx - texel * C);
  if (texel >= (size_t)H * W) return;
  const int y = (int)(texel / W), x = (int)(texel - (size_t)y * W);
  const float* p = src + (size_t)c * H * W;
  const float v = p[(size_t)y * W + x];
  if (val) val[texel * (C == 24 ? NMF_APP_STRIDE : C) + c] = v;
  if (pack) {
    float dx = 0.f, dy = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int yy = y + i - 2;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int xx = x + j - 2;
        if (xx < 0 || xx >= W) continue;
        const float s = p[(size_t)yy * W + xx];
        dx = fmaf(kx[i * 5 + j], s, dx);
        dy = fmaf(ky[i * 5 + j], s, dy);
      }
    }
    float* o = pack + texel * (3 * C);
    o[c] = v; o[C + c] = dx; o[2 * C + c] = dy;
  }
}
// lines are (N, 1) images: only the centre column of the stencil meets data.  lpack: [n][4][val4 | dy4]
__global__ void k_pack_line(const float* __restrict__ src, int C, int N, const float* __restrict__ ky25, float* __restrict__ val,
                            float* __restrict__ pack) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = idx / C, c = idx - n * C;
  if (n >= N) return;
  const float* p = src + (size_t)c * N;
  const float v = p[n];
  if (val) val[(size_t)n * (C == 24 ? NMF_APP_STRIDE : C) + c] = v;
  if (pack) {
    float dy = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int nn = n + i - 2;
      if (nn >= 0 && nn < N) dy = fmaf(ky25[i * 5 + 2], p[nn], dy);
    }
    const int lo = (c >> 2) * 8 + (c & 3);
    pack[(size_t)n * 32 + lo] = v;
    pack[(size_t)n * 32 + lo + 4] = dy;
  }
}

extern "C" int nmf_pack_factor(const float* src, int C, int H, int W, const float* kx25, const float* ky25, float* val, float* pack,
                               void* stream) {
  if (!src || (!val && !pack) || H <= 0 || W <= 0 || (C != 16 && C != 24)) return NMF_E_ARG;
  if (pack && (C != 16 || !kx25 || !ky25)) return NMF_E_ARG;
  cudaStream_t cs = (cudaStream_t)stream;
  if (W == 1) {                     // a line (1,C,N,1)
    k_pack_line<<<(H * C + 255) / 256, 256, 0, cs>>>(src, C, H, ky25, val, pack);
  } else {
    const size_t n = (size_t)H * W * C;
    if (C == 16) k_pack_plane<16><<<(unsigned)((n + 255) / 256), 256, 0, cs>>>(src, H, W, kx25, ky25, val, pack);
    else k_pack_plane<24><<<(unsigned)((n + 255) / 256), 256, 0, cs>>>(src, H, W, kx25, ky25, val, pack);
  }
  CKL();
  return NMF_OK;
}

// ---- environment: activation + summed-area table ----
// pass 1: act and the running sum over rows in fp64, rounded to fp32 per prefix (torch.cumsum(act / 1000, dim=2) on the
// CPU).  CTA = 16 columns x 16 row segments of one channel: every thread sums its segment, the segment totals meet in
// shared memory, then the thread walks its rows again with the carry of the segments above (fp64 sums of <= 2048 fp32
// values: any association agrees with the sequential sum to 1e-16 relative, i.e. the same fp32 after rounding; a single
// thread per column walking all rows took 259 us at 512 x 1024).  c1: (3, h, w) fp32 scratch.  Also the pole-row sums.
#define ENV_SEG 16
__global__ void __launch_bounds__(256) k_env_act_scan_y(const float* __restrict__ bg, int h, int w, float brightness, float mul,
                                                        const float* __restrict__ scalars_dev, float* __restrict__ c1,
                                                        float* __restrict__ act_out, double* __restrict__ pole) {
  __shared__ double tot[ENV_SEG][16];
  if (scalars_dev) { brightness = scalars_dev[0]; mul = scalars_dev[1]; }      // device-resident parameters (training)
  const int col = threadIdx.x & 15, seg = threadIdx.x >> 4;
  const int cpb = (w + 15) / 16;                       // CTAs per channel
  const int k = blockIdx.x / cpb, x = (blockIdx.x - k * cpb) * 16 + col;
  const int R = (h + ENV_SEG - 1) / ENV_SEG;
  const int y0 = seg * R, y1 = min(y0 + R, h);
  auto act = [&](size_t o) { return expf(fminf(__fadd_rn(brightness, __fmul_rn(mul, bg[o])), 20.0f)); };   // two roundings, like the two torch ops
  double run = 0.0;
  if (x < w)
    for (int y = y0; y < y1; ++y) run += (double)(act(((size_t)k * h + y) * w + x) / 1000.0f);
  tot[seg][col] = run;
  __syncthreads();
  if (x >= w) return;
  run = 0.0;
  for (int sgm = 0; sgm < seg; ++sgm) run += tot[sgm][col];
  for (int y = y0; y < y1; ++y) {
    const size_t o = ((size_t)k * h + y) * w + x;
    const float a = act(o);
    if (act_out) act_out[o] = a;
    if (y == 0) atomicAdd(pole + k, (double)a);        // pole rows: mean over the row (integral_equirect.py:498-502)
    if (y == h - 1) atomicAdd(pole + 3 + k, (double)a);
    run += (double)(a / 1000.0f);
    c1[o] = (float)run;
  }
}
// pass 2: one warp per (channel, row): inclusive scan over x in fp64 (chunks of 32 with a carried total), rounded to fp32,
// written channel-last.  sat: [h][w][4]
__global__ void __launch_bounds__(256) k_env_scan_x(const float* __restrict__ c1, int h, int w, float* __restrict__ sat) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // k * h + y
  if (row >= 3 * h) return;
  const int k = row / h, y = row - k * h;
  const float* p = c1 + (size_t)row * w;
  double carry = 0.0;
  for (int x0 = 0; x0 < w; x0 += 32) {
    const int x = x0 + lane;
    double v = x < w ? (double)p[x] : 0.0;
    for (int off = 1; off < 32; off <<= 1) {
      const double u = __shfl_up_sync(FULLM, v, off);
      if (lane >= off) v += u;
    }
    v += carry;
    carry = __shfl_sync(FULLM, v, 31);
    if (x < w) sat[((size_t)y * w + x) * 4 + k] = (float)v;
  }
}
__global__ void k_env_pad(float* __restrict__ sat, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sat[i * 4 + 3] = 0.f;
}

// sat8[y][x] = { sat4[y][x], sat4[y][min(x + 1, w - 1)] }: both x-taps of a bilinear lookup in one aligned 32-byte record
__global__ void k_env_pair(const float4* __restrict__ sat4, int w, size_t n, float4* __restrict__ sat8) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % (size_t)w);
  sat8[2 * i] = sat4[i];
  sat8[2 * i + 1] = sat4[x + 1 < w ? i + 1 : i];
}
extern "C" int nmf_env_pair_sat(const float* sat4, int h, int w, float* sat8, void* stream) {
  if (!sat4 || !sat8 || h <= 0 || w <= 0) return NMF_E_ARG;
  const size_t n = (size_t)h * w;
  k_env_pair<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)sat4, w, n, (float4*)sat8);
  CKL();
  return NMF_OK;
}

// env_dyn = { mipbias, top rgb, bottom rgb }: what NmfScene.env_dyn points at
__global__ void k_env_dyn_finish(const double* __restrict__ pole, int w, const float* __restrict__ scalars_dev, float* __restrict__ env_dyn) {
  const int t = threadIdx.x;
  if (t == 0) env_dyn[0] = scalars_dev[2];
  if (t < 6) env_dyn[1 + t] = (float)(pole[t] / (double)w);
}
static int env_build_sat_impl(const float* bg_mat, int h, int w, float brightness, float mul, const float* scalars_dev, float* scratch_c1,
                              float* act, float* sat4, double* pole_sums, float* env_dyn, void* stream);
extern "C" int nmf_env_build_sat(const float* bg_mat, int h, int w, float brightness, float mul, float* scratch_c1, float* act,
                                 float* sat4, double* pole_sums, void* stream) {
  return env_build_sat_impl(bg_mat, h, w, brightness, mul, nullptr, scratch_c1, act, sat4, pole_sums, nullptr, stream);
}
extern "C" int nmf_env_build_sat_dev(const float* bg_mat, int h, int w, const float* scalars_dev, float* scratch_c1, float* act,
                                     float* sat4, double* pole_sums, float* env_dyn, void* stream) {
  if (!scalars_dev || !env_dyn) return NMF_E_ARG;
  return env_build_sat_impl(bg_mat, h, w, 0.f, 1.f, scalars_dev, scratch_c1, act, sat4, pole_sums, env_dyn, stream);
}
static int env_build_sat_impl(const float* bg_mat, int h, int w, float brightness, float mul, const float* scalars_dev, float* scratch_c1,
                              float* act, float* sat4, double* pole_sums, float* env_dyn, void* stream) {
  if (!bg_mat || !scratch_c1 || !sat4 || !pole_sums || h <= 0 || w <= 0) return NMF_E_ARG;
  cudaStream_t cs = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(pole_sums, 0, 6 * sizeof(double), cs);
  if (e != cudaSuccess) return (int)e;
  k_env_act_scan_y<<<3 * ((w + 15) / 16), 256, 0, cs>>>(bg_mat, h, w, brightness, mul, scalars_dev, scratch_c1, act, pole_sums);
  CKL();
  k_env_scan_x<<<(3 * h * 32 + 255) / 256, 256, 0, cs>>>(scratch_c1, h, w, sat4);
  CKL();
  const size_t n = (size_t)h * w;
  k_env_pad<<<(unsigned)((n + 255) / 256), 256, 0, cs>>>(sat4, n);
  CKL();
  if (env_dyn) {
    k_env_dyn_finish<<<1, 32, 0, cs>>>(pole_sums, w, scalars_dev, env_dyn);
    CKL();
  }
  return NMF_OK;
}

// ---- occupancy: 3^3 max-pool (padding 1) of clamp(alpha, 0, 1), threshold, bit-fields ----
// one thread per 32-voxel word of the voxel field: 27 x 32 cached reads of alpha per word.  vox: bit (z*gy + y)*pitch + x
__global__ void __launch_bounds__(256) k_occ_pool_pack(const float* __restrict__ alpha, int gx, int gy, int gz, int pitch, float thres,
                                                       uint32_t* __restrict__ vox, float* __restrict__ vol) {
  const size_t wi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int wpr = pitch >> 5;
  const size_t n_words = (size_t)gz * gy * wpr;
  if (wi >= n_words) return;
  const int xw = (int)(wi % wpr);
  const int y = (int)((wi / wpr) % gy), z = (int)(wi / ((size_t)wpr * gy));
  uint32_t bits = 0;
  for (int b = 0; b < 32; ++b) {
    const int x = xw * 32 + b;
    if (x >= gx) break;
    float m = 0.f;
    for (int dz = -1; dz <= 1; ++dz) {
      const int zz = z + dz;
      if (zz < 0 || zz >= gz) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= gy) continue;
        const float* r = alpha + ((size_t)zz * gy + yy) * gx;
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= gx) continue;
          m = fmaxf(m, fminf(fmaxf(r[xx], 0.f), 1.f));
        }
      }
    }
    const bool on = m >= thres;
    if (on) bits |= 1u << b;
    if (vol) vol[((size_t)z * gy + y) * gx + x] = on ? 1.f : 0.f;
  }
  vox[wi] = bits;
}
// cell (x, y, z) = OR of the voxels (x..x+1, y..y+1, z..z+1) that exist: word-parallel (shift by one bit + neighbour word)
__global__ void __launch_bounds__(256) k_occ_cells(const uint32_t* __restrict__ vox, int gy, int gz, int pitch, uint32_t* __restrict__ cell) {
  const size_t wi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int wpr = pitch >> 5;
  const size_t n_words = (size_t)gz * gy * wpr;
  if (wi >= n_words) return;
  const int xw = (int)(wi % wpr);
  const int y = (int)((wi / wpr) % gy), z = (int)(wi / ((size_t)wpr * gy));
  uint32_t acc = 0;
  for (int dz = 0; dz <= 1; ++dz) {
    if (z + dz >= gz) continue;
    for (int dy = 0; dy <= 1; ++dy) {
      if (y + dy >= gy) continue;
      const size_t base = ((size_t)(z + dz) * gy + (y + dy)) * wpr + xw;
      const uint32_t a = vox[base];
      const uint32_t nxt = xw + 1 < wpr ? vox[base + 1] : 0u;
      acc |= a | (a >> 1) | (nxt << 31);
    }
  }
  cell[wi] = acc;
}
// coarse cell c (8 fine cells per axis) is set iff a voxel with index in [8c-1, 8c+9] on every axis is set
__global__ void __launch_bounds__(256) k_occ_coarse(const uint32_t* __restrict__ vox, int gx, int gy, int gz, int pitch, int cw, int ch,
                                                    int cd, uint32_t* __restrict__ coarse) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = c < cw * ch * cd;
  bool on = false;
  if (in) {
    const int cx = c % cw, cy = (c / cw) % ch, cz = c / (cw * ch);
    const int x0 = max(8 * cx - 1, 0), x1 = min(8 * cx + 9, gx - 1);
    for (int z = max(8 * cz - 1, 0); z <= min(8 * cz + 9, gz - 1) && !on; ++z)
      for (int y = max(8 * cy - 1, 0); y <= min(8 * cy + 9, gy - 1) && !on; ++y) {
        const size_t row = ((size_t)z * gy + y) * (size_t)pitch;
        for (int x = x0; x <= x1; ++x) {
          const size_t i = row + x;
          if ((vox[i >> 5] >> (i & 31)) & 1u) { on = true; break; }
        }
      }
  }
  const unsigned m = __ballot_sync(FULLM, on);
  if ((threadIdx.x & 31) == 0 && (c >> 5) < (cw * ch * cd + 31) / 32) coarse[c >> 5] = m;
}

extern "C" int nmf_occupancy_from_alpha(const float* alpha, int gx, int gy, int gz, float thres, int pitch, uint32_t* vox,
                                        uint32_t* cell, uint32_t* coarse, float* volume, void* stream) {
  if (!alpha || !vox || !cell || gx < 2 || gy < 2 || gz < 2 || pitch < gx || (pitch & 31)) return NMF_E_ARG;
  cudaStream_t cs = (cudaStream_t)stream;
  const size_t n_words = (size_t)gz * gy * (pitch >> 5);
  k_occ_pool_pack<<<(unsigned)((n_words + 255) / 256), 256, 0, cs>>>(alpha, gx, gy, gz, pitch, thres, vox, volume);
  CKL();
  k_occ_cells<<<(unsigned)((n_words + 255) / 256), 256, 0, cs>>>(vox, gy, gz, pitch, cell);
  CKL();
  if (coarse) {
    const int cw = (gx + 7) / 8, ch = (gy + 7) / 8, cd = (gz + 7) / 8;
    const int n = cw * ch * cd;
    k_occ_coarse<<<((n + 31) / 32 * 32 + 255) / 256, 256, 0, cs>>>(vox, gx, gy, gz, pitch, cw, ch, cd, coarse);
    CKL();
  }
  return NMF_OK;
}

// ---- gradient hand-over: channel-last kernel layouts -> the reference's parameter layouts, all tensors in ONE launch ----
// job j: dst[c * n + i] = src[i * c_dim + c]  (c_dim = 1: a plain copy).  blockIdx.y = job; a CTA moves tiles of 128 rows
// through shared memory so that both the reads and the writes are coalesced.  (31 torch copy_ launches cost 0.38 ms of
// host time per training iteration.)
#define TRB_ROWS 128
#define TRB_MAXC 64
__global__ void __launch_bounds__(256) k_transpose_batch(const NmfTransposeJob* __restrict__ jobs) {
  __shared__ float tile[TRB_ROWS * (TRB_MAXC + 1)];
  const NmfTransposeJob jb = jobs[blockIdx.y];
  const int C = jb.c, ld = C + 1;
  const size_t n = jb.n;
  if (C == 1) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) jb.dst[i] = jb.src[i];
    return;
  }
  for (size_t r0 = (size_t)blockIdx.x * TRB_ROWS; r0 < n; r0 += (size_t)gridDim.x * TRB_ROWS) {
    const int rows = (int)min((size_t)TRB_ROWS, n - r0);
    const float* sp = jb.src + r0 * C;
    for (int i = threadIdx.x; i < rows * C; i += 256) tile[(i / C) * ld + (i % C)] = sp[i];
    __syncthreads();
    for (int i = threadIdx.x; i < rows * C; i += 256) {
      const int c = i / rows, r = i - c * rows;
      jb.dst[(size_t)c * n + r0 + r] = tile[r * ld + c];
    }
    __syncthreads();
  }
}
extern "C" int nmf_transpose_batch(const NmfTransposeJob* jobs_dev, int n_jobs, int blocks_per_job, void* stream) {
  if (!jobs_dev || n_jobs <= 0 || blocks_per_job <= 0) return NMF_E_ARG;
  k_transpose_batch<<<dim3((unsigned)blocks_per_job, (unsigned)n_jobs), 256, 0, (cudaStream_t)stream>>>(jobs_dev);
  CKL();
  return NMF_OK;
}

// ---- material heads + BRDF MLP operands (modules/render_modules.py:519-574, modules/brdf.py:73-120) in ONE launch ----
// per layer i: wt[k][o] = W[o][k] (fp32, for the SIMT paths), bias copy, and the tensor-core tiles: W (rows out, K in) ->
// [K/8][rows][8] with K zero-padded to 80, the bias in column 66 (it multiplies the constant-1 input), rows padded to
// 64 / 64 / 16 -- once in fp16 (forward, csrc/nmf_mlp_tc.cuh) and once in bf16 (reverse pass, csrc/nmf_mlp_tc_bwd.cuh).
// (The torch version of this was ~40 small launches and 0.3 ms of host time per training iteration.)
__global__ void __launch_bounds__(256) k_pack_shading(const NmfShadingPack a) {
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < 3; ++i) {
    const int O = a.n_out[i], K = a.n_in[i], rows = i == 2 ? 16 : 64;
    if (!a.w[i]) continue;
    for (int e = t0; e < O * K; e += stride) {
      const int o = e / K, k = e - o * K;
      a.wt[i][(size_t)k * O + o] = a.w[i][e];
    }
    for (int e = t0; e < O; e += stride) a.bo[i][e] = a.b[i][e];
    for (int e = t0; e < rows * 80; e += stride) {
      const int r = e / 80, k = e - r * 80;
      float v = 0.f;
      if (r < O) v = k < K ? a.w[i][(size_t)r * K + k] : (k == 66 ? a.b[i][r] : 0.f);
      const size_t q = (size_t)(k >> 3) * rows * 8 + (size_t)r * 8 + (k & 7);
      if (a.w16[i]) ((__half*)a.w16[i])[q] = __float2half_rn(v);
      if (a.wbf[i]) ((__nv_bfloat16*)a.wbf[i])[q] = __float2bfloat16_rn(v);
    }
  }
  int row0 = 0;
  for (int h = 0; h < 4; ++h) {
    const int R = a.head_rows[h];
    if (a.head_w[h]) {
      for (int e = t0; e < R * 24; e += stride) a.head_w_out[row0 * 24 + e] = a.head_w[h][e];
      for (int e = t0; e < R; e += stride) a.head_b_out[row0 + e] = a.head_b[h][e];
    }
    row0 += R;
  }
}
extern "C" int nmf_pack_shading(const NmfShadingPack* p, void* stream) {
  if (!p) return NMF_E_ARG;
  for (int i = 0; i < 3; ++i)
    if (p->w[i] && (!p->b[i] || !p->wt[i] || !p->bo[i] || p->n_in[i] > 66 || p->n_out[i] > (i == 2 ? 16 : 64))) return NMF_E_ARG;
  for (int h = 0; h < 4; ++h)
    if (p->head_w[h] && (!p->head_b[h] || !p->head_w_out || !p->head_b_out)) return NMF_E_ARG;
  k_pack_shading<<<16, 256, 0, (cudaStream_t)stream>>>(*p);
  CKL();
  return NMF_OK;
}
