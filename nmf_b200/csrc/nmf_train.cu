// nmf_b200 -- training slice, first model: model=tensorf.
//
// nmf_sample_rays_train: AlphaGridSampler.sample(is_train=True) (samplers/alphagrid.py:131-207, 278-370).
// nmf_train_plain: one fused forward + backward of TensorNeRF.forward(is_train=True) for model=tensorf and the
// photometric loss of train.py:576-611, as a fixed sequence of launches (no autograd, no host synchronisation):
//   k_train_sample     warp per ray: keyed jitter, exact fp64 warp scan of the step lengths, AABB + occupancy test
//   k_train_prefix     one CTA: per-ray sample offsets and the dynamic batch truncation (whole_valid)
//   k_train_density    warp per kept ray: ballot compaction of the valid steps, VM density, warp product scan of the
//                      transmittance -> per-sample (ray, z, dist, f, alpha, T, w), per-ray acc
//   k_train_mlp<0>     128 samples per CTA: appearance gather, basis, encoding, 135-128-128-3 MLP; rgb per sample,
//                      w * rgb into the ray
//   k_train_loss       per ray: tonemap, background, loss, dL/d(linear colour), dL/d(acc)
//   k_train_mlp<1>     reloads the tile's forward activations ([feature][sample], stride 136; kept by k_train_mlp<0>) and walks
//                      back: the four tile GEMMs (dW1, dH1, dW0, dX) on the tensor cores as 3xTF32 mma.sync,
//                      weight gradients are per-tile contractions over the 128 samples flushed with REDs; activations'
//                      gradients in place;
//                      encoding, basis_mat and the appearance factors (16-byte REDs into channel-last buffers)
//   k_train_composite_bwd  warp per kept ray: reverse warp scan of dw * w, d sigma, softplus', density-factor REDs
#include <cuda_runtime.h>

#include "nmf_train.cuh"

#define FULL 0xffffffffu
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
#define CKL() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)
#define TS 136              // shared-memory stride of one feature row (136 = 8 mod 32): columns [k*TS + t] and the mma B
                            // fragments [(k0 + lane%4)*TS + n0 + lane/4] are bank-conflict-free; 16-byte aligned rows
#define TILE 128
#define MLP_T 256           // threads of the MLP tile kernel (8 warps: two per scheduler)

static int t_sms = 0;
static int t_sm_count() {
  if (!t_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&t_sms, cudaDevAttrMultiProcessorCount, dev);
    if (t_sms <= 0) t_sms = 148;
  }
  return t_sms;
}
static int t_check_scene(const NmfScene* s) {
  if (!s) return NMF_E_ARG;
  if (s->n_steps <= 0 || s->n_steps > NMF_MAX_STEPS) return NMF_E_UNSUPPORTED;
  for (int p = 0; p < 3; ++p)
    if (!s->dval[p] || !s->lval[p] || s->plane_w[p] < 2 || s->plane_h[p] < 2 || s->line_n[p] < 2) return NMF_E_ARG;
  if (s->plane_w[1] != s->plane_w[0] || s->line_n[2] != s->plane_w[0] || s->plane_w[2] != s->plane_h[0] ||
      s->line_n[1] != s->plane_h[0] || s->plane_h[2] != s->plane_h[1] || s->line_n[0] != s->plane_h[1])
    return NMF_E_UNSUPPORTED;
  if (s->has_occ && (!s->occ_vox || !s->occ_cell || (s->opitch & 31))) return NMF_E_ARG;
  return NMF_OK;
}

// ================================================================================================
// sampling (train)
// ================================================================================================
struct SampleArgs {
  const float* rays; int n; float near_override; uint64_t seed, ray_id0; const uint64_t* ray_ids;
  uint8_t* valid; float* z; int* n_valid;
};
__global__ void __launch_bounds__(256) k_train_sample(const NmfScene s, const SampleArgs a) {
  const int lane = threadIdx.x & 31;
  const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= a.n) return;
  float o[3], d[3];
  for (int i = 0; i < 3; ++i) { o[i] = a.rays[(size_t)ray * 6 + i]; d[i] = a.rays[(size_t)ray * 6 + 3 + i]; }
  const float tmin = nmf_ray_tmin(o, d, s.aabb0, s.aabb1, a.near_override >= 0.f ? a.near_override : s.near, s.far);
  const uint64_t key = nmf_primary_key(a.seed, a.ray_ids ? a.ray_ids[ray] : a.ray_id0 + (uint64_t)ray);
  double carry = 0.0;
  int nv = 0;
  for (int base = 0; base < s.n_steps; base += 32) {
    const int k = base + lane;
    // every partial sum of <= 2048 fp32 step lengths in [stepsize/2, 3 stepsize/2] is exact in fp64 (37 significant
    // bits), so the scan order does not matter: prefix == ATen's sequential fp64 accumulation, rounded to fp32
    double v = k < s.n_steps ? (double)nmf_jitter_step(key, k, s.stepsize) : 0.0;
    for (int off = 1; off < 32; off <<= 1) {
      const double u = __shfl_up_sync(FULL, v, off);
      if (lane >= off) v += u;
    }
    v += carry;
    carry = __shfl_sync(FULL, v, 31);
    if (k < s.n_steps) {
      const float z = NMF_ADD(tmin, (float)v);
      float p[3];
      nmf_step_pos(o, d, z, p);
      bool ok = nmf_inside(p, s.aabb0, s.aabb1);
      if (ok && s.has_occ) {
        float xn[3];
        nmf_normalize_xyz(s, p, xn);
        ok = nmf_occupied(s.occ_vox, s.occ_cell, s.ow, s.oh, s.od, s.opitch, xn[0], xn[1], xn[2]);
      }
      if (a.valid) a.valid[(size_t)ray * s.n_steps + k] = ok;
      a.z[(size_t)ray * s.n_steps + k] = z;
      nv += ok;
    }
  }
  for (int off = 16; off > 0; off >>= 1) nv += __shfl_xor_sync(FULL, nv, off);
  if (lane == 0) a.n_valid[ray] = nv;
}

extern "C" int nmf_sample_rays_train(const NmfScene* scene, const float* rays, int n_rays, float near_override, uint64_t seed,
                                     uint64_t ray_id0, const uint64_t* ray_ids, int max_samples, uint8_t* ray_valid,
                                     float* z_vals, int* n_valid, uint8_t* whole_valid, int* n_kept, void* stream) {
  int st = t_check_scene(scene);
  if (st) return st;
  if (!rays || n_rays <= 0 || !z_vals || !n_valid) return NMF_E_ARG;     // ray_valid is optional
  cudaStream_t cs = (cudaStream_t)stream;
  SampleArgs a{rays, n_rays, near_override, seed, ray_id0, ray_ids, ray_valid, z_vals, n_valid};
  k_train_sample<<<(n_rays + 7) / 8, 256, 0, cs>>>(*scene, a);
  CKL();
  if (whole_valid && n_kept) {
    CK(cudaMemsetAsync(n_kept, 0, 2 * sizeof(int), cs));
    k_train_prefix<<<1, 1024, 0, cs>>>(n_valid, n_rays, max_samples, (int*)nullptr, whole_valid, n_kept);
    CKL();
  }
  return NMF_OK;
}

// ================================================================================================
// fused training step, model=tensorf
// ================================================================================================
struct TWS {
  uint8_t* valid; float* z; int* n_valid; int* offs;
  float* lin;     // [n][3] sum w * rgb  (zeroed)
  float* acc;     // [n]
  float* g_lin;   // [n][3]
  float* g_acc;   // [n]
  int* s_ray; float* s_z; float* s_dist; float* s_f; float* s_alpha; float* s_T; float* s_w; float* s_rgb; float* s_dw;
  // activations of the forward tiles, kept for the backward pass: per 128-sample tile [row][128] floats -- the encoded
  // input X (135 rows) and the two hidden layers after ReLU (128 rows each); 1564 B per sample, written once, read once
  float* act_x; float* act_h1; float* act_h2;
  size_t zero_off, zero_bytes, total;
};
static size_t t_align(size_t x) { return (x + 255) & ~(size_t)255; }
static void t_carve(TWS& w, const NmfScene* s, int n, int cap, char* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += t_align(bytes); return p; };
  w.zero_off = off;
  w.lin = (float*)take((size_t)n * 3 * 4);
  w.acc = (float*)take((size_t)n * 4);
  w.zero_bytes = off - w.zero_off;
  w.valid = (uint8_t*)take((size_t)n * s->n_steps);
  w.z = (float*)take((size_t)n * s->n_steps * 4);
  w.n_valid = (int*)take((size_t)n * 4);
  w.offs = (int*)take((size_t)n * 4);
  w.g_lin = (float*)take((size_t)n * 3 * 4);
  w.g_acc = (float*)take((size_t)n * 4);
  w.s_ray = (int*)take((size_t)cap * 4);
  w.s_z = (float*)take((size_t)cap * 4);
  w.s_dist = (float*)take((size_t)cap * 4);
  w.s_f = (float*)take((size_t)cap * 4);
  w.s_alpha = (float*)take((size_t)cap * 4);
  w.s_T = (float*)take((size_t)cap * 4);
  w.s_w = (float*)take((size_t)cap * 4);
  w.s_rgb = (float*)take((size_t)cap * 3 * 4);
  w.s_dw = (float*)take((size_t)cap * 4);
  const size_t tiles = ((size_t)cap + 127) / 128;
  w.act_x = (float*)take(tiles * 135 * 128 * 4);
  w.act_h1 = (float*)take(tiles * 128 * 128 * 4);
  w.act_h2 = (float*)take(tiles * 128 * 128 * 4);
  w.total = off;
}
extern "C" size_t nmf_train_workspace_bytes(const NmfScene* scene, int n_rays, int cap_samples) {
  if (!scene || n_rays <= 0 || cap_samples <= 0) return 0;
  TWS w;
  t_carve(w, scene, n_rays, cap_samples, nullptr);
  return w.total;
}

// warp per kept ray: compaction + density + transmittance (tensor_nerf.py:19-35, 264-289, 366-369)
struct DensityArgs { const float* rays; const int* n_kept; int cap; unsigned* error; };
__global__ void __launch_bounds__(256) k_train_density(const NmfScene s, const DensityArgs a, const TWS w) {
  const int lane = threadIdx.x & 31;
  const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= a.n_kept[0]) return;
  if (a.n_kept[1] > a.cap) { if (ray == 0 && lane == 0) atomicOr(a.error, NMF_DEV_E_SURVIVORS); return; }
  float o[3], d[3];
  for (int i = 0; i < 3; ++i) { o[i] = a.rays[(size_t)ray * 6 + i]; d[i] = a.rays[(size_t)ray * 6 + 3 + i]; }
  const uint8_t* valid = w.valid + (size_t)ray * s.n_steps;
  const float* zs = w.z + (size_t)ray * s.n_steps;
  int cnt = w.offs[ray];
  float Tc = 1.0f, accsum = 0.f;
  for (int base = 0; base < s.n_steps; base += 32) {
    const int k = base + lane;
    const bool ok = k < s.n_steps && valid[k];
    const unsigned m = __ballot_sync(FULL, ok);
    if (m == 0) continue;
    float fac = 1.0f, alpha = 0.f, f = 0.f, z = 0.f, dist = 0.f;
    if (ok) {
      z = zs[k];
      dist = k + 1 < s.n_steps ? NMF_SUB(zs[k + 1], z) : 0.f;           // alphagrid.py:343-345 (last = 0)
      float p[3], xn[3];
      nmf_step_pos(o, d, z, p);
      nmf_normalize_xyz(s, p, xn);
      const NmfTaps t = nmf_vm_taps(s, xn);
      for (int g = 0; g < 4; ++g) f += nmf_density_group(s, t, g);
      const float sigma = nmf_feature2density(f, s.density_shift);
      dist = dist * s.distance_scale;
      alpha = 1.0f - expf(-sigma * dist);
      fac = 1.0f - alpha + 1e-10f;
    }
    float incl = fac;                        // inclusive product scan in step order (invalid lanes contribute 1)
    for (int off = 1; off < 32; off <<= 1) {
      const float u = __shfl_up_sync(FULL, incl, off);
      if (lane >= off) incl *= u;
    }
    float excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float T = Tc * excl;
    Tc *= __shfl_sync(FULL, incl, 31);
    if (ok) {
      const int i = cnt + __popc(m & ((1u << lane) - 1u));
      const float wt = alpha * T;
      w.s_ray[i] = ray; w.s_z[i] = z; w.s_dist[i] = dist; w.s_f[i] = f; w.s_alpha[i] = alpha; w.s_T[i] = T; w.s_w[i] = wt;
      accsum += wt;
    }
    cnt += __popc(m);
  }
  for (int off = 16; off > 0; off >>= 1) accsum += __shfl_xor_sync(FULL, accsum, off);
  if (lane == 0) w.acc[ray] = accsum;
}

// per ray: loss head
struct LossArgs { const float* gt; const int* n_kept; int n; float lambda_pred; int white_bg; float* rgb_map; float* acc_map; double* loss; };
__global__ void __launch_bounds__(256) k_train_loss(const LossArgs a, const TWS w) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  float lp = 0.f, la = 0.f;
  if (ray < a.n) {
    float map[3] = {0.f, 0.f, 0.f}, acc = 0.f;
    if (ray < a.n_kept[0]) {
      const float bg[3] = {a.white_bg ? 1.f : 0.f, a.white_bg ? 1.f : 0.f, a.white_bg ? 1.f : 0.f};
      acc = w.acc[ray];
      float gl[3], ga;
      lp = nmf_train_loss_ray(w.lin + 3 * (size_t)ray, acc, bg, a.gt + 3 * (size_t)ray, a.lambda_pred, map, gl, &ga);
      la = acc;
      for (int c = 0; c < 3; ++c) w.g_lin[3 * (size_t)ray + c] = gl[c];
      w.g_acc[ray] = ga;
    }
    if (a.rgb_map) for (int c = 0; c < 3; ++c) a.rgb_map[3 * (size_t)ray + c] = map[c];
    if (a.acc_map) a.acc_map[ray] = acc;
  }
  for (int off = 16; off > 0; off >>= 1) { lp += __shfl_xor_sync(FULL, lp, off); la += __shfl_xor_sync(FULL, la, off); }
  if ((threadIdx.x & 31) == 0 && (lp != 0.f || la != 0.f)) {
    atomicAdd(a.loss, (double)lp);
    atomicAdd(a.loss + 1, (double)la);
  }
}

// ------------------------------------------------------------------------------------------------
// the MLP tile kernel.  Shared memory: X [136][TS] (row 135 = 0: K padding), H1 [128][TS], H2 [128][TS], D2 [3][128].
// The five tile GEMMs (two forward layers, dW1, dH1, dW0, dX) run on the tensor cores as 3xTF32 mma.sync.m16n8k8
// (hi*hi + lo*hi + hi*lo: fp32-level accuracy, so the images and gradients stay within the parity tolerances); they
// are true dense contractions over the 128 samples of the tile.  Everything per sample (gathers, encodings, the
// 3-wide output layer, scatters) stays on the SIMT pipes, thread = sample = column.
// ------------------------------------------------------------------------------------------------
#define SM_X 0
#define SM_H1 (136 * TS)
#define SM_H2 (SM_H1 + 128 * TS)
#define SM_D2 (SM_H2 + 128 * TS)
#define SM_FLOATS (SM_D2 + 3 * 128)

// x = hi + lo with hi = x truncated to TF32's 10 explicit mantissa bits (one LOP) and lo = x - hi, exact in fp32; the
// tensor core reads only the TF32 bits of lo, so x is represented to ~2^-21 relative (cvt.rna costs ~6 ALU instructions
// per element on this target and made the GEMM loops issue-bound)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// One warp: acc[mt][nt] (16 x 8 tiles) += sum_k A(m, k) B(k, n), m in [0, 16 MT), n in [0, 8 NT), k in [0, 8 ksteps).
// fa(m, k) / fb(k, n) fetch one element (global weights through the read-only path, or shared memory).
// Fragment layout of m16n8k8: g = lane / 4, t = lane % 4; A: (g, t), (g+8, t), (g, t+4), (g+8, t+4); B: (t, g), (t+4, g);
// C: (g, 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1).
template <int MT, int NT, class FA, class FB>
__device__ __forceinline__ void warp_gemm(float (&acc)[MT][NT][4], int ksteps, FA fa, FB fb, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.f; acc[mt][nt][1] = 0.f; acc[mt][nt][2] = 0.f; acc[mt][nt][3] = 0.f; }
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k0 = 8 * ks;
    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      split_tf32(fa(16 * mt + g, k0 + t), ah[mt][0], al[mt][0]);
      split_tf32(fa(16 * mt + g + 8, k0 + t), ah[mt][1], al[mt][1]);
      split_tf32(fa(16 * mt + g, k0 + t + 4), ah[mt][2], al[mt][2]);
      split_tf32(fa(16 * mt + g + 8, k0 + t + 4), ah[mt][3], al[mt][3]);
    }
    // B fragments four column tiles at a time; the three passes (lo*hi, hi*lo, hi*hi) each sweep all 4 x MT accumulator
    // tiles, so that dependent MMAs on one accumulator are 4 * MT instructions apart (4 warps per SM: the ILP has to
    // come from inside the warp)
    constexpr int NG = NT < 4 ? NT : 4;
#pragma unroll
    for (int n4 = 0; n4 < NT; n4 += NG) {
      uint32_t bh[NG][2], bl[NG][2];
#pragma unroll
      for (int q = 0; q < NG; ++q) {
        split_tf32(fb(k0 + t, 8 * (n4 + q) + g), bh[q][0], bl[q][0]);
        split_tf32(fb(k0 + t + 4, 8 * (n4 + q) + g), bh[q][1], bl[q][1]);
      }
#pragma unroll
      for (int q = 0; q < NG; ++q)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][n4 + q], al[mt], bh[q][0], bh[q][1]);
#pragma unroll
      for (int q = 0; q < NG; ++q)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][n4 + q], ah[mt], bl[q][0], bl[q][1]);
#pragma unroll
      for (int q = 0; q < NG; ++q)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][n4 + q], ah[mt], bh[q][0], bh[q][1]);
    }
  }
}
// visits every accumulator element of the warp: f(row, col, value)
template <int MT, int NT, class F>
__device__ __forceinline__ void warp_epilogue(float (&acc)[MT][NT][4], int lane, F f) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      f(16 * mt + g, 8 * nt + 2 * t, acc[mt][nt][0]);
      f(16 * mt + g, 8 * nt + 2 * t + 1, acc[mt][nt][1]);
      f(16 * mt + g + 8, 8 * nt + 2 * t, acc[mt][nt][2]);
      f(16 * mt + g + 8, 8 * nt + 2 * t + 1, acc[mt][nt][3]);
    }
}
// sum over the 128 samples of row `r` (read by one thread)
__device__ __forceinline__ float tile_rowsum(const float* r) {
  float b = 0.f;
  for (int q = 0; q < TILE / 4; ++q) { const float4 v = ((const float4*)r)[q]; b += v.x + v.y + v.z + v.w; }
  return b;
}
__device__ __forceinline__ float tile_rowdot(const float* r, const float* c) {
  float b = 0.f;
  for (int q = 0; q < TILE / 4; ++q) {
    const float4 v = ((const float4*)r)[q], u = ((const float4*)c)[q];
    b += v.x * u.x + v.y * u.y + v.z * u.z + v.w * u.w;
  }
  return b;
}

// shared-memory tile ([row][TS]) <-> global tile ([row][128]), 16-byte accesses by all threads of the CTA
__device__ __forceinline__ void tile_store(float* g, const float* sm_rows, int rows, int t) {
  for (int i = t; i < rows * 32; i += MLP_T) {
    const int r = i >> 5, c4 = (i & 31) * 4;
    __stcs((float4*)(g + (size_t)r * 128 + c4), *(const float4*)(sm_rows + r * TS + c4));      // streamed: read once, later
  }
}
__device__ __forceinline__ void tile_load(float* sm_rows, const float* g, int rows, int t) {
  for (int i = t; i < rows * 32; i += MLP_T) {
    const int r = i >> 5, c4 = (i & 31) * 4;
    *(float4*)(sm_rows + r * TS + c4) = __ldcs((const float4*)(g + (size_t)r * 128 + c4));
  }
}

struct MlpArgs { const float* rays; const int* n_kept; int cap; NmfPlainGrads g; };
template <int BWD>
__global__ void __launch_bounds__(MLP_T, 1) k_train_mlp(const NmfScene s, const MlpArgs a, const TWS w) {
  extern __shared__ __align__(16) float sm[];
  float* X = sm + SM_X;
  float* H1 = sm + SM_H1;
  float* H2 = sm + SM_H2;
  float* D2 = sm + SM_D2;
  // 8 warps: warp w owns the 32 x 64 output block (rows m0 .. m0+31, columns n0 .. n0+63) of every tile GEMM; the
  // per-sample phases run on the first 128 threads (thread = sample = column)
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, m0 = 32 * (wid & 3), n0 = 64 * (wid >> 2);
  const bool lead = t < TILE;
  const int M = min(a.n_kept[1], a.cap);
  if (a.n_kept[1] > a.cap) return;
  if (lead) X[135 * TS + t] = 0.f;                                  // K padding row of the first layer
  for (int tile = blockIdx.x * TILE; tile < M; tile += gridDim.x * TILE) {
    const int si = tile + t;
    const bool active = lead && si < M;
    int ray = 0;
    float dv[3] = {0.f, 0.f, 0.f};
    NmfTaps tp;
    const size_t tile_id = (size_t)(tile / TILE);
    float rgb[3] = {0.f, 0.f, 0.f};
    if (lead) {
      float o[3] = {0.f, 0.f, 0.f}, p[3], xn[3];
      float z = 0.f;
      if (active) {
        ray = w.s_ray[si];
        z = w.s_z[si];
        for (int i = 0; i < 3; ++i) { o[i] = a.rays[(size_t)ray * 6 + i]; dv[i] = a.rays[(size_t)ray * 6 + 3 + i]; }
      }
      nmf_step_pos(o, dv, z, p);
      nmf_normalize_xyz(s, p, xn);
      tp = nmf_vm_taps(s, xn);
      if (!BWD) {
        float coef[72], feat[24];
        nmf_app_coef(s, tp, coef);
        for (int oo = 0; oo < 24; ++oo) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 72; ++j) acc += __ldg(s.basis_t + j * 24 + oo) * coef[j];
          feat[oo] = active ? acc : 0.f;
        }
        nmf_plain_encode(feat, dv, X + t, TS);
      }
    }
    float acc[2][8][4];
    if (BWD) {
      // the forward pass kept this tile's activations (1564 B per sample through HBM instead of recomputing the
      // gather, the encoding and both 128-wide layers: a third of the backward pass's MMA work)
      tile_load(X, w.act_x + tile_id * 135 * 128, 135, t);
      tile_load(H1, w.act_h1 + tile_id * 128 * 128, 128, t);
      tile_load(H2, w.act_h2 + tile_id * 128 * 128, 128, t);
      if (active) for (int c = 0; c < 3; ++c) rgb[c] = w.s_rgb[3 * (size_t)si + c];
      __syncthreads();
    } else {
      __syncthreads();
      tile_store(w.act_x + tile_id * 135 * 128, X, 135, t);
      // layer 1: H1[j][n] = relu(b0[j] + sum_k W0[j][k] X[k][n])      (W0 as stored: (out, in), row stride 135)
      warp_gemm<2, 8>(acc, 17,
                       [&](int m, int k) { return k < 135 ? __ldg(s.plain_w0 + (m0 + m) * 135 + k) : 0.f; },
                       [&](int k, int n) { return X[k * TS + n0 + n]; }, lane);
      warp_epilogue<2, 8>(acc, lane, [&](int r, int c, float v) { H1[(m0 + r) * TS + n0 + c] = fmaxf(v + __ldg(s.plain_b0 + m0 + r), 0.f); });
      __syncthreads();
      tile_store(w.act_h1 + tile_id * 128 * 128, H1, 128, t);
      // layer 2
      warp_gemm<2, 8>(acc, 16, [&](int m, int k) { return __ldg(s.plain_w1 + (m0 + m) * 128 + k); },
                       [&](int k, int n) { return H1[k * TS + n0 + n]; }, lane);
      warp_epilogue<2, 8>(acc, lane, [&](int r, int c, float v) { H2[(m0 + r) * TS + n0 + c] = fmaxf(v + __ldg(s.plain_b1 + m0 + r), 0.f); });
      __syncthreads();
      tile_store(w.act_h2 + tile_id * 128 * 128, H2, 128, t);
      // output layer (3 wide) per sample
      float o3[3] = {__ldg(s.plain_b2), __ldg(s.plain_b2 + 1), __ldg(s.plain_b2 + 2)};
      if (lead) for (int k = 0; k < 128; ++k) {
        const float hv = H2[k * TS + t];
        const float* w2 = s.plain_w2t + k * 3;
        o3[0] += hv * __ldg(w2); o3[1] += hv * __ldg(w2 + 1); o3[2] += hv * __ldg(w2 + 2);
      }
      for (int c = 0; c < 3; ++c) rgb[c] = nmf_sigmoid(o3[c]);
    }
    if (!BWD) {
      if (active) {
        const float wt = w.s_w[si];
        for (int c = 0; c < 3; ++c) {
          w.s_rgb[3 * (size_t)si + c] = rgb[c];
          atomicAdd(w.lin + 3 * (size_t)ray + c, wt * rgb[c]);
        }
      }
      continue;      // the next tile's writes to X / H1 / H2 are ordered behind this tile's reads by its own barriers
    }
    // ---- backward ----
    float dpre[3] = {0.f, 0.f, 0.f};
    if (active) {
      const float wt = w.s_w[si];
      float dw = w.g_acc[ray];
      for (int c = 0; c < 3; ++c) {
        const float gl = w.g_lin[3 * (size_t)ray + c];
        dw += gl * rgb[c];
        dpre[c] = wt * gl * rgb[c] * (1.0f - rgb[c]);
      }
      w.s_dw[si] = dw;
    }
    if (lead) for (int c = 0; c < 3; ++c) D2[c * 128 + t] = dpre[c];
    __syncthreads();
    if (lead) {  // (a) dW2t[k = t][c], db2
      const float* hrow = H2 + t * TS;
      atomicAdd(a.g.w2t + t * 3, tile_rowdot(hrow, D2));
      atomicAdd(a.g.w2t + t * 3 + 1, tile_rowdot(hrow, D2 + 128));
      atomicAdd(a.g.w2t + t * 3 + 2, tile_rowdot(hrow, D2 + 256));
      if (t < 3) atomicAdd(a.g.b2 + t, tile_rowsum(D2 + t * 128));
    }
    __syncthreads();
    // (b) dh2 in place (own column)
    if (lead) for (int k = 0; k < 128; ++k) {
      const float hv = H2[k * TS + t];
      const float* w2 = s.plain_w2t + k * 3;
      H2[k * TS + t] = hv > 0.f ? __ldg(w2) * dpre[0] + __ldg(w2 + 1) * dpre[1] + __ldg(w2 + 2) * dpre[2] : 0.f;
    }
    __syncthreads();
    // (c) dW1t[k][j] += sum_n H1[k][n] dH2[j][n]  (M = k, N = j, contraction over the tile's samples);  db1
    warp_gemm<2, 8>(acc, 16, [&](int m, int k) { return H1[(m0 + m) * TS + k]; }, [&](int k, int n) { return H2[(n0 + n) * TS + k]; }, lane);
    warp_epilogue<2, 8>(acc, lane, [&](int r, int c, float v) { atomicAdd(a.g.w1t + (m0 + r) * 128 + n0 + c, v); });
    if (lead) atomicAdd(a.g.b1 + t, tile_rowsum(H2 + t * TS));
    __syncthreads();
    // (d) dH1[k][n] = [H1[k][n] > 0] sum_j W1[j][k] dH2[j][n], in place (every element has one owner)
    warp_gemm<2, 8>(acc, 16, [&](int m, int k) { return __ldg(s.plain_w1t + (m0 + m) * 128 + k); },
                     [&](int k, int n) { return H2[k * TS + n0 + n]; }, lane);
    warp_epilogue<2, 8>(acc, lane, [&](int r, int c, float v) { float* q = H1 + (m0 + r) * TS + n0 + c; *q = *q > 0.f ? v : 0.f; });
    __syncthreads();
    // (e) dW0t[i][j] += sum_n X[i][n] dH1[j][n]: rows 0..127 as above, rows 128..134 as one 16-row tile split over the warps' columns;  db0
    warp_gemm<2, 8>(acc, 16, [&](int m, int k) { return X[(m0 + m) * TS + k]; }, [&](int k, int n) { return H1[(n0 + n) * TS + k]; }, lane);
    warp_epilogue<2, 8>(acc, lane, [&](int r, int c, float v) { atomicAdd(a.g.w0t + (m0 + r) * 128 + n0 + c, v); });
    {
      float acc1[1][2][4];
      warp_gemm<1, 2>(acc1, 16, [&](int m, int k) { return X[(128 + m) * TS + k]; }, [&](int k, int n) { return H1[(16 * wid + n) * TS + k]; }, lane);
      warp_epilogue<1, 2>(acc1, lane, [&](int r, int c, float v) { if (r < 7) atomicAdd(a.g.w0t + (128 + r) * 128 + 16 * wid + c, v); });
    }
    if (lead) atomicAdd(a.g.b0 + t, tile_rowsum(H1 + t * TS));
    // (f) dX[i][n] = sum_j W0[j][i] dH1[j][n] for input rows 0..127 (features and their encodings are rows 0..122) -> H2
    warp_gemm<2, 8>(acc, 16, [&](int m, int k) { return __ldg(s.plain_w0t + (m0 + m) * 128 + k); },
                     [&](int k, int n) { return H1[k * TS + n0 + n]; }, lane);
    warp_epilogue<2, 8>(acc, lane, [&](int r, int c, float v) { H2[(m0 + r) * TS + n0 + c] = v; });
    __syncthreads();
    float dfeat[24];
    if (lead) {
#pragma unroll
      for (int oo = 0; oo < 24; ++oo)
        dfeat[oo] = nmf_plain_encode_bwd(X + t, TS, oo, H2[oo * TS + t], H2[(27 + 2 * oo) * TS + t], H2[(28 + 2 * oo) * TS + t],
                                         H2[(75 + 2 * oo) * TS + t], H2[(76 + 2 * oo) * TS + t]);
    }
    __syncthreads();                              // every thread is done with X before it is overwritten
    if (lead) {
      for (int oo = 0; oo < 24; ++oo) X[oo * TS + t] = active ? dfeat[oo] : 0.f;
      float coef[72];                             // gathered again (cheap next to the MLP) rather than kept live
      nmf_app_coef(s, tp, coef);
      for (int j = 0; j < 72; ++j) X[(24 + j) * TS + t] = active ? coef[j] : 0.f;
    }
    __syncthreads();
    // (g) d basis_t[j][o] += sum_n coef_j[n] dfeat_o[n]
    for (int idx = t; idx < 72 * 24; idx += MLP_T)
      atomicAdd(a.g.basis_t + idx, tile_rowdot(X + (24 + idx / 24) * TS, X + (idx % 24) * TS));
    // (h) appearance factors (own sample)
    if (active) {
      float dcoef[72];
      for (int j = 0; j < 72; ++j) {
        float accj = 0.f;
#pragma unroll
        for (int oo = 0; oo < 24; ++oo) accj += __ldg(s.basis_t + j * 24 + oo) * dfeat[oo];
        dcoef[j] = accj;
      }
      nmf_app_bwd(s, tp, dcoef, a.g.a_plane, a.g.a_line);
    }
    __syncthreads();                              // the next tile's forward overwrites X / H1 / H2
  }
}

// warp per kept ray, samples walked from the last to the first
struct CompArgs { const float* rays; const int* n_kept; int cap; NmfPlainGrads g; };
__global__ void __launch_bounds__(256) k_train_composite_bwd(const NmfScene s, const CompArgs a, const TWS w) {
  const int lane = threadIdx.x & 31;
  const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= a.n_kept[0] || a.n_kept[1] > a.cap) return;
  const int b = w.offs[ray], n = w.n_valid[ray];
  float o[3], d[3];
  for (int i = 0; i < 3; ++i) { o[i] = a.rays[(size_t)ray * 6 + i]; d[i] = a.rays[(size_t)ray * 6 + 3 + i]; }
  float carry = 0.f;                              // sum of dw_j w_j over the samples after the current group
  for (int hi = n; hi > 0; hi -= 32) {
    const int i = hi - 1 - lane;                  // lane 0 = last sample of the group
    const bool ok = i >= 0;
    const float dw = ok ? w.s_dw[b + i] : 0.f, wt = ok ? w.s_w[b + i] : 0.f;
    float incl = dw * wt;                         // inclusive scan over lanes = over later samples
    for (int off = 1; off < 32; off <<= 1) {
      const float u = __shfl_up_sync(FULL, incl, off);
      if (lane >= off) incl += u;
    }
    const float suffix = carry + incl - dw * wt;
    carry += __shfl_sync(FULL, incl, 31);
    if (ok) {
      const float f = w.s_f[b + i];
      const float dsigma = nmf_composite_bwd(dw, w.s_T[b + i], w.s_alpha[b + i], w.s_dist[b + i], suffix);
      const float df = dsigma * nmf_feature2density_grad(f, s.density_shift);
      if (df != 0.f) {
        float p[3], xn[3];
        nmf_step_pos(o, d, w.s_z[b + i], p);
        nmf_normalize_xyz(s, p, xn);
        const NmfTaps t = nmf_vm_taps(s, xn);
        nmf_density_bwd(s, t, df, a.g.d_plane, a.g.d_line);
      }
    }
  }
}

extern "C" int nmf_train_plain(const NmfScene* scene, const NmfTrain* tp, const float* rays, const float* gt,
                               const NmfPlainGrads* grads, const NmfTrainOut* out, void* workspace, size_t workspace_bytes,
                               void* stream) {
  int st = t_check_scene(scene);
  if (st) return st;
  if (!tp || !rays || !gt || !grads || !out || !workspace || tp->n_rays <= 0 || tp->cap_samples <= 0) return NMF_E_ARG;
  if (scene->model != 1 || !scene->plain_w0t || !scene->plain_w0 || !scene->plain_w1 || !scene->aval[0] || !scene->basis_t)
    return NMF_E_UNSUPPORTED;
  if (!out->loss || !out->n_kept || !out->error || !out->whole_valid) return NMF_E_ARG;
  const int n = tp->n_rays, cap = tp->cap_samples;
  TWS w;
  t_carve(w, scene, n, cap, (char*)workspace);
  if (workspace_bytes < w.total) return NMF_E_WORKSPACE;
  cudaStream_t cs = (cudaStream_t)stream;
  // zero: per-ray accumulators, outputs, gradients
  CK(cudaMemsetAsync((char*)workspace + w.zero_off, 0, w.zero_bytes, cs));
  CK(cudaMemsetAsync(out->loss, 0, 2 * sizeof(double), cs));
  CK(cudaMemsetAsync(out->n_kept, 0, 2 * sizeof(int), cs));
  CK(cudaMemsetAsync(out->error, 0, sizeof(unsigned), cs));
  for (int p = 0; p < 3; ++p) {
    const size_t hw = (size_t)scene->plane_h[p] * scene->plane_w[p];
    if (!grads->d_plane[p] || !grads->d_line[p] || !grads->a_plane[p] || !grads->a_line[p]) return NMF_E_ARG;
    CK(cudaMemsetAsync(grads->d_plane[p], 0, hw * 16 * 4, cs));
    CK(cudaMemsetAsync(grads->a_plane[p], 0, hw * 24 * 4, cs));
    CK(cudaMemsetAsync(grads->d_line[p], 0, (size_t)scene->line_n[p] * 16 * 4, cs));
    CK(cudaMemsetAsync(grads->a_line[p], 0, (size_t)scene->line_n[p] * 24 * 4, cs));
  }
  if (!grads->basis_t || !grads->w0t || !grads->b0 || !grads->w1t || !grads->b1 || !grads->w2t || !grads->b2) return NMF_E_ARG;
  CK(cudaMemsetAsync(grads->basis_t, 0, 72 * 24 * 4, cs));
  CK(cudaMemsetAsync(grads->w0t, 0, 135 * 128 * 4, cs));
  CK(cudaMemsetAsync(grads->b0, 0, 128 * 4, cs));
  CK(cudaMemsetAsync(grads->w1t, 0, 128 * 128 * 4, cs));
  CK(cudaMemsetAsync(grads->b1, 0, 128 * 4, cs));
  CK(cudaMemsetAsync(grads->w2t, 0, 128 * 3 * 4, cs));
  CK(cudaMemsetAsync(grads->b2, 0, 3 * 4, cs));

  SampleArgs sa{rays, n, -1.0f, tp->seed, tp->ray_id0, tp->ray_ids, w.valid, w.z, w.n_valid};
  const int warp_blocks = (n + 7) / 8;
  k_train_sample<<<warp_blocks, 256, 0, cs>>>(*scene, sa);
  CKL();
  k_train_prefix<<<1, 1024, 0, cs>>>(w.n_valid, n, tp->max_samples, w.offs, out->whole_valid, out->n_kept);
  CKL();
  DensityArgs da{rays, out->n_kept, cap, out->error};
  k_train_density<<<warp_blocks, 256, 0, cs>>>(*scene, da, w);
  CKL();
  const size_t smem = (size_t)SM_FLOATS * sizeof(float);
  CK(cudaFuncSetAttribute(k_train_mlp<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_train_mlp<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int tiles = (cap + TILE - 1) / TILE;
  const int grid = tiles < t_sm_count() ? tiles : t_sm_count();      // one resident CTA per SM (202 KB of shared memory)
  MlpArgs ma{rays, out->n_kept, cap, *grads};
  k_train_mlp<0><<<grid, MLP_T, smem, cs>>>(*scene, ma, w);
  CKL();
  LossArgs la{gt, out->n_kept, n, tp->lambda_pred, tp->white_bg, out->rgb_map, out->acc_map, out->loss};
  k_train_loss<<<(n + 255) / 256, 256, 0, cs>>>(la, w);
  CKL();
  k_train_mlp<1><<<grid, MLP_T, smem, cs>>>(*scene, ma, w);
  CKL();
  CompArgs ca{rays, out->n_kept, cap, *grads};
  k_train_composite_bwd<<<warp_blocks, 256, 0, cs>>>(*scene, ca, w);
  CKL();
  return NMF_OK;
}

// ================================================================================================
// resolution schedule: TensoRF.upsample (fields/tensoRF.py:208-227) = F.interpolate(bilinear, align_corners=True)
// of every plane (1,C,H,W) and line (1,C,N,1), in the reference's own parameter layout
// ================================================================================================
__global__ void __launch_bounds__(256) k_upsample(const float* src, int C, int H, int W, float* dst, int H2, int W2) {
  const size_t n = (size_t)C * H2 * W2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W2), y = (int)((i / W2) % H2), c = (int)(i / ((size_t)W2 * H2));
    int x0, x1, y0, y1;
    float wx0, wx1, hy0, hy1;
    nmf_resize_tap(x, W, W2, &x0, &x1, &wx0, &wx1);
    nmf_resize_tap(y, H, H2, &y0, &y1, &hy0, &hy1);
    dst[i] = nmf_resize_pixel(src + (size_t)c * H * W, W, y0, y1, hy0, hy1, x0, x1, wx0, wx1);
  }
}
extern "C" int nmf_upsample_bilinear(const float* src, int C, int H, int W, float* dst, int H2, int W2, void* stream) {
  if (!src || !dst || C <= 0 || H <= 0 || W <= 0 || H2 <= 0 || W2 <= 0) return NMF_E_ARG;
  const size_t n = (size_t)C * H2 * W2;
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)t_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  k_upsample<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, C, H, W, dst, H2, W2);
  CKL();
  return NMF_OK;
}


// ================================================================================================
// optimiser step on the device (train.py:443-467 Adam + LambdaLR, :675-678 density L1, :752-754 clip + step)
// ================================================================================================
// Every kernel streams its tensors once (grid-stride, 16-byte accesses when the segment is aligned): HBM-bound.
__global__ void __launch_bounds__(256) k_l1_reg(const float* __restrict__ p, size_t n, float coef, float* __restrict__ g,
                                                double* sum_abs) {
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = p[i];
    acc += (double)fabsf(v);
    if (g) g[i] += nmf_l1_grad(v, coef);
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(FULL, acc, off);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && sum_abs) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(sum_abs, t);
  }
}

__global__ void __launch_bounds__(256) k_sq_norm(const float* __restrict__ g, size_t n, double* out) {
  double acc = 0.0;
  const size_t n4 = ((uintptr_t)g & 15) == 0 ? n / 4 : 0;
  const float4* g4 = (const float4*)g;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += (double)g[i] * g[i];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(FULL, acc, off);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, size_t n, const NmfAdamScalars h, const double* sq_norm,
                                              const float* __restrict__ control) {
  if (control && control[1] != 0.f) return;                       // the step overflowed a list: no update (NmfAdam.control)
  const float gs = control ? control[0] : h.grad_scale;
  const float gmul = gs * (sq_norm ? nmf_clip_coef(*sq_norm, gs, h.max_norm) : 1.0f);
  const bool al = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  const size_t n4 = al ? n / 4 : 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 P = ((float4*)p)[i], M = ((float4*)m)[i], V = ((float4*)v)[i];
    const float4 G = ((const float4*)g)[i];
    nmf_adam_elem(&P.x, G.x, &M.x, &V.x, h, gmul);
    nmf_adam_elem(&P.y, G.y, &M.y, &V.y, h, gmul);
    nmf_adam_elem(&P.z, G.z, &M.z, &V.z, h, gmul);
    nmf_adam_elem(&P.w, G.w, &M.w, &V.w, h, gmul);
    ((float4*)p)[i] = P; ((float4*)m)[i] = M; ((float4*)v)[i] = V;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    nmf_adam_elem(p + i, g[i], m + i, v + i, h, gmul);
}

static int stream_grid(size_t n, int per_thread) {
  const size_t blocks = (n + (size_t)256 * per_thread - 1) / ((size_t)256 * per_thread);
  const size_t cap = (size_t)t_sm_count() * 8;          // one resident wave of 256-thread CTAs
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

extern "C" int nmf_l1_reg(const float* param, size_t n, float coef, float* grad, double* sum_abs, void* stream) {
  if (!param || n == 0) return NMF_E_ARG;
  k_l1_reg<<<stream_grid(n, 4), 256, 0, (cudaStream_t)stream>>>(param, n, coef, grad, sum_abs);
  CKL();
  return NMF_OK;
}

extern "C" int nmf_grad_sq_norm(const float* grad, size_t n, double* sq_norm, void* stream) {
  if (!grad || !sq_norm || n == 0) return NMF_E_ARG;
  k_sq_norm<<<stream_grid(n, 8), 256, 0, (cudaStream_t)stream>>>(grad, n, sq_norm);
  CKL();
  return NMF_OK;
}

extern "C" int nmf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, const NmfAdam* a,
                             const double* sq_norm, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !a || n == 0 || a->step < 1) return NMF_E_ARG;
  NmfAdamScalars h;
  const double b1 = a->beta1, b2 = a->beta2;
  h.one_minus_b1 = (float)(1.0 - b1);
  h.b2 = (float)b2;
  h.one_minus_b2 = (float)(1.0 - b2);
  h.eps = a->eps;
  h.weight_decay = a->weight_decay;
  h.step_size = (float)((double)a->lr / (1.0 - pow(b1, (double)a->step)));
  h.bc2_sqrt = (float)sqrt(1.0 - pow(b2, (double)a->step));
  h.grad_scale = a->grad_scale;
  h.max_norm = a->max_norm;
  k_adam<<<stream_grid(n, 8), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, h, sq_norm, a->control);
  CKL();
  return NMF_OK;
}
