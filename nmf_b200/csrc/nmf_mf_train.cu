// nmf_b200 -- reverse pass of the MICROFACET model on the device.
//
// nmf_train_microfacet = the training forward (nmf_render_rays_train's launch sequence, csrc/nmf_kernels.cu, which in train
// mode keeps every valid sample and the survivor -> sample maps) + the photometric / prediction / orientation loss of
// train.py:586-650 + the hand-written backward of every stage the reference differentiates with autograd:
//   models/microfacet.py:271-673 (Fresnel mix, per-sample means, bounce rays), brdf_samplers/ggx.py:61-226 (d L / d roughness,
//   d L / d N once detach_N is off), modules/brdf.py:177-261 (BRDF MLP), modules/integral_equirect.py:409-504 (map, mip bias,
//   direction), modules/render_modules.py:519-574 (material heads), fields/tensoRF.py:181-205,392-405 (VM factors, basis_mat),
//   fields/tensor_base.py:107-129 (normals, create_graph=True), modules/tensor_nerf.py:19-35,291-317 (compositing, re-traced rays).
// The per-element math is csrc/nmf_microfacet_bwd.cuh (host-checked against the oracle's autograd); this file is the kernels:
//   k_mf_loss          per primary ray: tonemap, loss, d loss / d (linear rgb, acc)
//   k_mf_tangent1      per level-1 bounce ray: forward-mode d comb / d (direction of its re-traced ray), three tangents, segmented
//   k_mf_sec_tangent   sums into the 3x3 Jacobian d rgb1 / d direction of every re-traced ray (+ its background lookup)
//   k_mf_bounce_bwd<L> per bounce ray: GGX sample with its roughness / normal tangents, incoming radiance (environment with
//                      direction derivative, or the re-traced ray's radiance and Jacobian), Fresnel-mix backward; segmented sums
//                      of d R0, d diffuse, d roughness, d N and d weight back to the sample; environment-map scatter, d mipbias,
//                      upstream of the BRDF MLP and of the re-traced ray
//   k_mf_sec_bwd       per re-traced ray: background gradient, upstream of its samples
//   k_mf_mlp_bwd       128 bounce rays per tile: BRDF MLP recomputed in fp32, walked back; the weight gradients are tile
//                      contractions over the rays kept in registers across the CTA's tiles (one flush per CTA), d feature
//                      segment-summed to the sample
//   k_mf_sample_bwd<L> per bounce sample: material heads, basis_mat (tile contractions), appearance factors, the normal path
//                      (orientation loss; bounce direction once detach_N is off) into the derivative-plane gradient images
//   k_mf_ori_bwd       survivors without a bounce sample but a back-facing normal: orientation loss only
//   k_mf_composite_bwd<L> warp per ray over ALL its valid samples: reverse scan, softplus', density factors
// Gradients ACCUMULATE into NmfMicrofacetGrads (the caller zeroes them once per optimiser step and finishes the environment and
// normal images with nmf_env_lookup_bwd_finish / nmf_vm_normals_bwd_finish).
#include <cuda_runtime.h>
#include <stdlib.h>

#include "nmf_render_ws.cuh"
#include "nmf_microfacet_bwd.cuh"
#include "nmf_mlp_tc_bwd.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
#define CKL() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

int nmf_render_impl_train(const NmfScene* scene, const NmfRender* rp, const NmfRenderTrain* tr, const float* rays,
                          const NmfImages* out, const NmfCounters* counters, void* workspace, size_t workspace_bytes,
                          void* stream_, WS* ws_out);

static int m_sms = 0;
static int m_sm_count() {
  if (!m_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&m_sms, cudaDevAttrMultiProcessorCount, dev);
    if (m_sms <= 0) m_sms = 148;
  }
  return m_sms;
}

// persistent grids are sized to exactly one resident wave (SMs x CTAs that fit): the tile lists of a training batch are short
// (1-3 tiles per CTA), so a grid larger than the resident set costs a whole extra round of tiles on the SMs that get a late CTA
template <class K>
static int m_resident(K kernel, int threads, size_t smem, int* cache) {
  if (!*cache) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess || nb < 1) nb = 1;
    *cache = nb * m_sm_count();
  }
  return *cache;
}

template <int N>
__device__ __forceinline__ void seg_sumN(float (&v)[N], const Seg& g, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float a = __shfl_down_sync(FULL, v[i], off);
      if (lane + off <= g.last) v[i] += a;
    }
  }
}

// (chunk, ray index within the chunk's region, rays in the region) of thread `tid` of tile `tile`
__device__ __forceinline__ void mf_locate(const int* tile_start, int n_chunks, const int* ray_count, int cap_rays, int tile, int tid,
                                          int& chunk, int& n, int& r) {
  int lo = 0, hi = n_chunks;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(tile_start + mid) <= tile) lo = mid; else hi = mid;
  }
  chunk = lo;
  n = min(ray_count[chunk], cap_rays);
  r = (tile - __ldg(tile_start + chunk)) * MLP_THREADS + tid;
}

// ------------------------------------------------------------------------------------------------
// loss head (train.py:586-650 through modules/tonemap.py:38-49): per kept primary ray
// ------------------------------------------------------------------------------------------------
struct MfLossArgs {
  const float* accum0; const float* acc0; const uint8_t* whole; const float* gt; int n; float lambda_pred;
  float* g_lin0; double* loss; const float* stat4;
};
__global__ void __launch_bounds__(256) k_mf_loss(const MfLossArgs a) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  float lp = 0.f, la = 0.f;
  if (ray < a.n) {
    float gl[3] = {0.f, 0.f, 0.f}, ga = 0.f;
    if (a.whole[ray]) {
      const float bg[3] = {1.f, 1.f, 1.f};
      float map[3];
      const float acc = a.acc0[ray];
      lp = nmf_train_loss_ray(a.accum0 + (size_t)ray * A_N + A_RGB, acc, bg, a.gt + 3 * (size_t)ray, a.lambda_pred, map, gl, &ga);
      la = acc;
    }
    *(float4*)(a.g_lin0 + 4 * (size_t)ray) = make_float4(gl[0], gl[1], gl[2], ga);
  }
  for (int off = 16; off > 0; off >>= 1) { lp += __shfl_xor_sync(FULL, lp, off); la += __shfl_xor_sync(FULL, la, off); }
  if ((threadIdx.x & 31) == 0 && (lp != 0.f || la != 0.f)) {
    atomicAdd(a.loss, (double)lp);
    atomicAdd(a.loss + 1, (double)la);
  }
  if (ray == 0) a.loss[2] = (double)a.stat4[0];          // ori_loss = sum w min(v.n, 0)^2 (tensor_nerf.py:573-583), one chunk
}

__global__ void k_mf_zero_bgrad(float* bgrad, const int* n_bs, int cap_bs) {
  const size_t n = (size_t)min(*n_bs, cap_bs) * NMF_BGRAD;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) bgrad[i] = 0.f;
}

// what every bounce-ray kernel reads of its sample record
struct MfSample {
  nmf_v3 V, N; float rough, w; int count; float f0[3], diffuse[3]; uint32_t ray, roff, vidx; float offu, offv;
};
__device__ __forceinline__ MfSample mf_read_sample(const BSample* b) {
  MfSample m;
  const float4 q0 = *(const float4*)b->pos, q1 = *(const float4*)b->V, q2 = *(const float4*)b->N;
  const float4 q3 = *(const float4*)b->f0, q4 = *(const float4*)b->diffuse, f5 = ((const float4*)b->frame)[5];
  m.V = nmf_mk3(q1.x, q1.y, q1.z); m.N = nmf_mk3(q2.x, q2.y, q2.z);
  m.rough = q1.w; m.w = q0.w; m.count = max(__float_as_int(q2.w), 1);
  m.f0[0] = q3.x; m.f0[1] = q3.y; m.f0[2] = q3.z; m.ray = __float_as_uint(q3.w);
  m.diffuse[0] = q4.x; m.diffuse[1] = q4.y; m.diffuse[2] = q4.z; m.roff = __float_as_uint(q4.w);
  m.vidx = b->pad; m.offu = f5.y; m.offv = f5.z;
  return m;
}

// ------------------------------------------------------------------------------------------------
// level-1 tangents: d rgb1 / d direction of every re-traced ray (modules/tensor_nerf.py:291-317: the recursion sees the
// parent's bounce direction as its ray direction; positions are detached in the field, so only the level-1 view vector and
// the background lookup move with it).  jac1[ray][3 c + k] = d rgb1_k / d direction_c.
// ------------------------------------------------------------------------------------------------
struct MfTangArgs {
  const BSample* bs; const BRay* brays; const uint32_t* owner; const int* ray_count; int cap_rays; const int* tile_start; int n_chunks;
  float* jac1;
};
__global__ void __launch_bounds__(MLP_THREADS) k_mf_tangent1(const NmfScene s, const MfTangArgs a) {
  const int n_tiles = a.tile_start[a.n_chunks];
  const int lane = threadIdx.x & 31;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int chunk, n, r;
    mf_locate(a.tile_start, a.n_chunks, a.ray_count, a.cap_rays, tile, threadIdx.x, chunk, n, r);
    const uint32_t key = r < n ? a.owner[(size_t)chunk * a.cap_rays + r] : NMF_NO_OWNER;
    const bool active = key != NMF_NO_OWNER;
    float j0[3] = {0.f, 0.f, 0.f}, j1[3] = {0.f, 0.f, 0.f}, j2[3] = {0.f, 0.f, 0.f};
    uint32_t ray1 = 0;
    if (active) {
      const MfSample m = mf_read_sample(a.bs + key);
      ray1 = m.ray;
      const int j = r - (int)m.roff;
      const float u1 = nmf_wrap01(__ldg(s.sobol + 2 * j) + m.offu), u2 = nmf_wrap01(__ldg(s.sobol + 2 * j + 1) + m.offv);
      const BRay* o = a.brays + (size_t)chunk * a.cap_rays + r;
      const float4 qa = *(const float4*)o->L, qb = *(const float4*)o->bw;
      const float bw[3] = {qb.x, qb.y, qb.z};
      const float sw = m.w / (float)m.count;
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        NmfDual3 Vd = nmf_d3k(m.V);                       // V1 = -direction: d V1 / d direction_c = -e_c
        if (c == 0) Vd.x.d = -1.0f; else if (c == 1) Vd.y.d = -1.0f; else Vd.z.d = -1.0f;
        float comb[3], dcomb[3];
        nmf_bounce_ray_tangent(s, Vd, m.N, m.f0, m.diffuse, m.rough, u1, u2, qa.w, bw, comb, dcomb);
        float* dst = c == 0 ? j0 : (c == 1 ? j1 : j2);
        dst[0] = sw * dcomb[0]; dst[1] = sw * dcomb[1]; dst[2] = sw * dcomb[2];
      }
    }
    const Seg seg = seg_setup(key, lane);
    seg_sum3(j0, seg, lane); seg_sum3(j1, seg, lane); seg_sum3(j2, seg, lane);
    if (active && seg.head) {
      float* J = a.jac1 + (size_t)ray1 * 12;
#pragma unroll
      for (int k = 0; k < 3; ++k) { atomicAdd(J + k, j0[k]); atomicAdd(J + 3 + k, j1[k]); atomicAdd(J + 6 + k, j2[k]); }
    }
  }
}
__global__ void k_mf_sec_tangent(const NmfScene s, const float* rays1, const float* mip1, const float* acc1, const int* n_sec,
                                 int max_retrace, int n, float* jac1) {
  const NmfEnvDyn ed = nmf_env_dyn_load(s);     // mipbias and pole means: by value or from NmfScene.env_dyn
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int chunk = i / max_retrace;
  if (i - chunk * max_retrace >= n_sec[chunk]) return;
  const nmf_v3 d = nmf_mk3(rays1[6 * (size_t)i + 3], rays1[6 * (size_t)i + 4], rays1[6 * (size_t)i + 5]);
  const float t = 1.0f - acc1[i];
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    NmfDual3 Dd = nmf_d3k(d);
    if (c == 0) Dd.x.d = 1.0f; else if (c == 1) Dd.y.d = 1.0f; else Dd.z.d = 1.0f;
    float bg[3], dbg[3];
    nmf_env_lookup1_d(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot, Dd, mip1[i], bg, dbg);
    for (int k = 0; k < 3; ++k) jac1[(size_t)i * 12 + 3 * c + k] += t * dbg[k];
  }
}

// ------------------------------------------------------------------------------------------------
// bounce rays, reverse (models/microfacet.py:352-613 under autograd)
// ------------------------------------------------------------------------------------------------
struct MfBounceBwdArgs {
  const BSample* bs; const BRay* brays; const uint32_t* owner; const int* ray_count; int cap_rays; const int* tile_start; int n_chunks;
  const float* g_lin;                 // [rays of this level][4]: d loss / d (linear rgb, acc)
  const float* rgb1; const float* jac1; float* g_lin1; int max_retrace;     // level 0: the re-traced rays
  float* bgrad; float* vdw; float4* dout;
  float* gsat; float* d_mipbias; int detach_N;
};
template <int LEVEL>
#ifndef NMF_BB_MINBLOCKS
#define NMF_BB_MINBLOCKS 3      // 168 registers: 3 CTAs per SM (uncapped the kernel takes 212 and runs 15 % slower); 4 or 5 change nothing
#endif
__global__ void __launch_bounds__(MLP_THREADS, NMF_BB_MINBLOCKS) k_mf_bounce_bwd(const NmfScene s, const MfBounceBwdArgs a) {
  const NmfEnvDyn ed = nmf_env_dyn_load(s);     // mipbias and pole means: by value or from NmfScene.env_dyn
  const int n_tiles = a.tile_start[a.n_chunks];
  const int lane = threadIdx.x & 31;
  float* g_top = a.gsat + (size_t)s.env_h * s.env_w * 4;
  float* g_bot = g_top + 4;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int chunk, n, r;
    mf_locate(a.tile_start, a.n_chunks, a.ray_count, a.cap_rays, tile, threadIdx.x, chunk, n, r);
    const uint32_t key = r < n ? a.owner[(size_t)chunk * a.cap_rays + r] : NMF_NO_OWNER;
    const bool active = key != NMF_NO_OWNER;
    // segment sums back to the sample: d R0 (3) | d diffuse (3) | d roughness | d N (3) | d weight
    float red[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) red[i] = 0.f;
    float mb = 0.f;
    uint32_t vidx = 0;
    if (r < n) a.dout[(size_t)chunk * a.cap_rays + r] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      const MfSample m = mf_read_sample(a.bs + key);
      vidx = m.vidx;
      const int j = r - (int)m.roff;
      const float u1 = nmf_wrap01(__ldg(s.sobol + 2 * j) + m.offu), u2 = nmf_wrap01(__ldg(s.sobol + 2 * j + 1) + m.offv);
      const BRay* o = a.brays + (size_t)chunk * a.cap_rays + r;
      const float4 qa = *(const float4*)o->L, qb = *(const float4*)o->bw;
      const float mip = qa.w;
      const float bw[3] = {qb.x, qb.y, qb.z};
      const int rslot = LEVEL == 0 ? __float_as_int(qb.w) : -1;
      const float4 gl4 = *(const float4*)(a.g_lin + 4 * (size_t)m.ray);
      const float gl[3] = {gl4.x, gl4.y, gl4.z};
      const float inv_m = 1.0f / (float)m.count;
      const float gm[3] = {m.w * gl[0] * inv_m, m.w * gl[1] * inv_m, m.w * gl[2] * inv_m};
      // the sample with its roughness tangent (brdf_samplers/ggx.py:61-226; `a` detached :116, pdf under no_grad :218)
      const NmfGGXdr dg = nmf_ggx_sample_dr(u1, u2, m.V, m.N, m.rough);
      float inc[3], tang[3];
      const float* J = nullptr;
      if (rslot >= 0) {
        const size_t gi = (size_t)chunk * (size_t)a.max_retrace + (size_t)rslot;
        const float* src = a.rgb1 + gi * 4;
        inc[0] = src[0]; inc[1] = src[1]; inc[2] = src[2];
        J = a.jac1 + gi * 12;
#pragma unroll
        for (int k = 0; k < 3; ++k) tang[k] = J[k] * dg.dL.x + J[3 + k] * dg.dL.y + J[6 + k] * dg.dL.z;
      } else {
        const NmfDual3 Ld = nmf_d3(nmf_dmk(dg.L.x, dg.dL.x), nmf_dmk(dg.L.y, dg.dL.y), nmf_dmk(dg.L.z, dg.dL.z));
        nmf_env_lookup1_d(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot, Ld, mip, inc, tang);
      }
      const float vh = nmf_dot(m.V, dg.H);
      const float cost = fabsf(vh), svh = vh > 0.f ? 1.0f : (vh < 0.f ? -1.0f : 0.f);
      float a_R0[3], a_inc[3], a_bw[3], a_diff[3];
      const float dcost = nmf_fresnel_mix_bwd(m.f0, cost, inc, bw, m.diffuse, gm, a_R0, a_inc, a_bw, a_diff);
      float dr = dcost * svh * nmf_dot(m.V, dg.dH);
#pragma unroll
      for (int k = 0; k < 3; ++k) { red[k] = a_R0[k]; red[3 + k] = a_diff[k]; dr += a_inc[k] * tang[k]; }
      red[6] = dr;
      if (!a.detach_N) {         // microfacet.py:352-353: the same two paths per column of N
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          const NmfGGXdr dn = nmf_ggx_sample_dN(u1, u2, m.V, m.N, m.rough, c);
          float tn[3];
          if (J) {
#pragma unroll
            for (int k = 0; k < 3; ++k) tn[k] = J[k] * dn.dL.x + J[3 + k] * dn.dL.y + J[6 + k] * dn.dL.z;
          } else {
            const NmfDual3 Ln = nmf_d3(nmf_dmk(dn.L.x, dn.dL.x), nmf_dmk(dn.L.y, dn.dL.y), nmf_dmk(dn.L.z, dn.dL.z));
            float inc_n[3];
            nmf_env_lookup1_d(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot, Ln, mip, inc_n, tn);
          }
          red[7 + c] = dcost * svh * nmf_dot(m.V, dn.dH) + a_inc[0] * tn[0] + a_inc[1] * tn[1] + a_inc[2] * tn[2];
        }
      }
      // d weight of the sample: gl . comb / count (comb = the ray's share of the sample's reflected radiance)
      float dwc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float F = nmf_fresnel(m.f0[k], cost);
        dwc += gl[k] * (F * inc[k] * bw[k] + (1.0f - F) * m.diffuse[k]);
      }
      red[10] = dwc * inv_m;
      // upstream of the BRDF MLP: bw = sigmoid(out + bias)   (modules/brdf.py:237-239)
      a.dout[(size_t)chunk * a.cap_rays + r] = make_float4(a_bw[0] * bw[0] * (1.0f - bw[0]), a_bw[1] * bw[1] * (1.0f - bw[1]),
                                                            a_bw[2] * bw[2] * (1.0f - bw[2]), 0.f);
      if (rslot >= 0) {          // upstream of the re-traced ray (one bounce ray per slot)
        float* g1 = a.g_lin1 + ((size_t)chunk * (size_t)a.max_retrace + (size_t)rslot) * 4;
        g1[0] = a_inc[0]; g1[1] = a_inc[1]; g1[2] = a_inc[2];
      } else if (a_inc[0] != 0.f || a_inc[1] != 0.f || a_inc[2] != 0.f) {
        nmf_env_lookup1_bwd_map(a.gsat, s.env_h, s.env_w, ed.mipbias, dg.L, mip, a_inc, g_top, g_bot);
        float rgb[3], dmb[3];
        nmf_env_lookup1_dmipbias(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot, dg.L, mip, rgb, dmb);
        mb = a_inc[0] * dmb[0] + a_inc[1] * dmb[1] + a_inc[2] * dmb[2];
      }
    }
    const Seg seg = seg_setup(key, lane);
    seg_sumN<11>(red, seg, lane);
    if (active && seg.head) {
      float* G = a.bgrad + (size_t)key * NMF_BGRAD;
#pragma unroll
      for (int i = 0; i < 10; ++i) if (red[i] != 0.f) atomicAdd(G + i, red[i]);
      atomicAdd(a.vdw + vidx, red[10]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mb += __shfl_xor_sync(FULL, mb, off);
    if (lane == 0 && mb != 0.f) atomicAdd(a.d_mipbias, mb);
  }
}

// re-traced rays: rgb1 = sum w refl + (1 - acc) env(d, mip)   (tensor_nerf.py:460-468 with tonemap=False): background
// gradient into the map, and the upstream of the ray's samples: d weight base = -g . bg
__global__ void k_mf_sec_bwd(const NmfScene s, const float* rays1, const float* mip1, const float* acc1, const int* n_sec,
                             int max_retrace, int n, float* g_lin1, float* gsat, float* d_mipbias) {
  const NmfEnvDyn ed = nmf_env_dyn_load(s);     // mipbias and pole means: by value or from NmfScene.env_dyn
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float mb = 0.f;
  if (i < n) {
    const int chunk = i / max_retrace;
    if (i - chunk * max_retrace < n_sec[chunk]) {
      float* g = g_lin1 + 4 * (size_t)i;
      const float a_inc[3] = {g[0], g[1], g[2]};
      const nmf_v3 d = nmf_mk3(rays1[6 * (size_t)i + 3], rays1[6 * (size_t)i + 4], rays1[6 * (size_t)i + 5]);
      float bg[3], dmb[3];
      nmf_env_lookup1_dmipbias(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot, d, mip1[i], bg, dmb);
      g[3] = -(a_inc[0] * bg[0] + a_inc[1] * bg[1] + a_inc[2] * bg[2]);
      const float t = 1.0f - acc1[i];
      const float gb[3] = {t * a_inc[0], t * a_inc[1], t * a_inc[2]};
      if (gb[0] != 0.f || gb[1] != 0.f || gb[2] != 0.f) {
        float* g_top = gsat + (size_t)s.env_h * s.env_w * 4;
        nmf_env_lookup1_bwd_map(gsat, s.env_h, s.env_w, ed.mipbias, d, mip1[i], gb, g_top, g_top + 4);
        mb = gb[0] * dmb[0] + gb[1] * dmb[1] + gb[2] * dmb[2];
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mb += __shfl_xor_sync(FULL, mb, off);
  if ((threadIdx.x & 31) == 0 && mb != 0.f) atomicAdd(d_mipbias, mb);
}

// ------------------------------------------------------------------------------------------------
// BRDF MLP 66 -> 64 -> 64 -> 4 backward (modules/brdf.py:177-261 under autograd), fp32.
// Shared memory (floats): weights (layout of mlp_load_weights) | X [128][73] | H1 [128][65] | H2 [128][65] | DH [128][68] | DO [128][4]
//   thread = row phases read / write their own row (odd strides: conflict-free); the weight-gradient phases read rows as
//   broadcasts (X, H1) and 16-byte pieces (DH, stride 68) and keep their outputs in registers across all tiles of the CTA.
// ------------------------------------------------------------------------------------------------
#define MB_W 8712
#define MB_XS 73
#define MB_HS 65
#define MB_DS 68
#define MB_OFF_X MB_W
#define MB_OFF_H1 (MB_OFF_X + 128 * MB_XS)
#define MB_OFF_H2 (MB_OFF_H1 + 128 * MB_HS)
#define MB_OFF_DH (MB_OFF_H2 + 128 * MB_HS)
#define MB_OFF_DO (MB_OFF_DH + 128 * MB_DS)
#define MB_FLOATS (MB_OFF_DO + 128 * 4)

struct MfMlpArgs {
  const BSample* bs; const uint32_t* owner; const int* ray_count; int cap_rays; const int* tile_start; int n_chunks;
  const float4* dout; float* bgrad;
  float* w0t; float* b0; float* w1t; float* b1; float* w2t; float* b2;
};
__global__ void __launch_bounds__(MLP_THREADS, 1) k_mf_mlp_bwd(const NmfScene s, const MfMlpArgs a) {
  extern __shared__ __align__(16) float msm[];
  float* W0 = msm; float* B0 = W0 + 66 * 64; float* W1 = B0 + 64; float* B1 = W1 + 64 * 64; float* W2 = B1 + 64; float* B2 = W2 + 256;
  float* X = msm + MB_OFF_X; float* H1 = msm + MB_OFF_H1; float* H2 = msm + MB_OFF_H2; float* DH = msm + MB_OFF_DH;
  float* DO = msm + MB_OFF_DO;
  const int t = threadIdx.x, lane = t & 31;
  for (int i = t; i < 66 * 64; i += MLP_THREADS) W0[i] = s.brdf_w0t[i];
  for (int i = t; i < 64 * 64; i += MLP_THREADS) W1[i] = s.brdf_w1t[i];
  for (int i = t; i < 256; i += MLP_THREADS) W2[i] = s.brdf_w2t[i];
  if (t < 64) { B0[t] = s.brdf_b0[t]; B1[t] = s.brdf_b1[t]; }
  if (t < 4) B2[t] = s.brdf_b2[t];
  for (int k = 66; k < MB_XS; ++k) X[t * MB_XS + k] = 0.f;      // K padding of the dW0 blocks (rows 66..71)
  __syncthreads();
  // weight-gradient blocks of this thread (kept in registers over all tiles): dW0t rows 9 kg .. 9 kg + 8, dW1t rows 8 kg ..
  // 8 kg + 7, columns 4 jg .. 4 jg + 3; threads < 64 also own dW2t row t, db0[t], db1[t]; threads 64..67 own db2
  const int kg = t >> 4, jg = t & 15;
  float g0[9][4], g1[8][4], g2[4] = {0.f, 0.f, 0.f, 0.f}, gb0 = 0.f, gb1 = 0.f, gb2 = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) g0[i][0] = g0[i][1] = g0[i][2] = g0[i][3] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) g1[i][0] = g1[i][1] = g1[i][2] = g1[i][3] = 0.f;
  const int n_tiles = a.tile_start[a.n_chunks];
  float* xr = X + t * MB_XS; float* h1r = H1 + t * MB_HS; float* h2r = H2 + t * MB_HS; float* dhr = DH + t * MB_DS;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int chunk, n, r;
    mf_locate(a.tile_start, a.n_chunks, a.ray_count, a.cap_rays, tile, t, chunk, n, r);
    const uint32_t key = r < n ? a.owner[(size_t)chunk * a.cap_rays + r] : NMF_NO_OWNER;
    const bool active = key != NMF_NO_OWNER;
    float4 dout = make_float4(0.f, 0.f, 0.f, 0.f);
    // ---- phase A: this row's input, forward, d h2 ----
    if (active) {
      const BSample* b = a.bs + key;
      const float4 q1 = *(const float4*)b->V, q2 = *(const float4*)b->N, q4 = *(const float4*)b->diffuse;
      const nmf_v3 V = nmf_mk3(q1.x, q1.y, q1.z), N = nmf_mk3(q2.x, q2.y, q2.z);
      const float rough = q1.w;
      const int j = r - (int)__float_as_uint(q4.w);
      const float4* fq = (const float4*)b->frame;
      const float4 f0 = fq[0], f1 = fq[1], f2 = fq[2], f3 = fq[3], f4 = fq[4], f5 = fq[5];
      NmfGGXFrame fr;
      fr.t = nmf_mk3(f0.x, f0.y, f0.z); fr.b = nmf_mk3(f0.w, f1.x, f1.y); fr.V_l = nmf_mk3(f1.z, f1.w, f2.x);
      fr.Vs = nmf_mk3(f2.y, f2.z, f2.w); fr.T1 = nmf_mk3(f3.x, f3.y, f3.z); fr.T2 = nmf_mk3(f3.w, f4.x, f4.y);
      fr.a = f4.z;
      const float u1 = nmf_wrap01(__ldg(s.sobol + 2 * j) + f5.y), u2 = nmf_wrap01(__ldg(s.sobol + 2 * j + 1) + f5.z);
      const NmfGGX g = nmf_ggx_sample_f(fr, u1, u2, V, N, rough);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float4 f = *(const float4*)(b->feat + 4 * i);
        xr[4 * i] = f.x; xr[4 * i + 1] = f.y; xr[4 * i + 2] = f.z; xr[4 * i + 3] = f.w;
      }
      float e[18];
      nmf_ish18_s(g.half_l, f4.w, f5.x, e);
#pragma unroll
      for (int i = 0; i < 18; ++i) xr[24 + i] = e[i];
      xr[42] = g.half_l.x; xr[43] = g.half_l.y; xr[44] = g.half_l.z;
      nmf_ish18_s(g.diff_l, f4.w, f5.x, e);
#pragma unroll
      for (int i = 0; i < 18; ++i) xr[45 + i] = e[i];
      xr[63] = g.diff_l.x; xr[64] = g.diff_l.y; xr[65] = g.diff_l.z;
      dout = a.dout[(size_t)chunk * a.cap_rays + r];
    } else {
      for (int i = 0; i < 66; ++i) xr[i] = 0.f;
    }
    *(float4*)(DO + 4 * t) = dout;
    float h[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) h[i] = B0[i];
#pragma unroll 2
    for (int k = 0; k < 66; ++k) {
      const float xv = xr[k];
      const float4* wr = (const float4*)(W0 + k * 64);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 wv = wr[q];
        h[4 * q] += xv * wv.x; h[4 * q + 1] += xv * wv.y; h[4 * q + 2] += xv * wv.z; h[4 * q + 3] += xv * wv.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) { h1r[i] = fmaxf(h[i], 0.f); h[i] = B1[i]; }
#pragma unroll 2
    for (int k = 0; k < 64; ++k) {
      const float xv = h1r[k];
      const float4* wr = (const float4*)(W1 + k * 64);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 wv = wr[q];
        h[4 * q] += xv * wv.x; h[4 * q + 1] += xv * wv.y; h[4 * q + 2] += xv * wv.z; h[4 * q + 3] += xv * wv.w;
      }
    }
    // d h2 = [h2 > 0] W2 dout, parked as 16-byte pieces
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = 4 * q + i;
        const float hv = fmaxf(h[k], 0.f);
        h2r[k] = hv;
        const float4 wv = *(const float4*)(W2 + 4 * k);
        v[i] = hv > 0.f ? wv.x * dout.x + wv.y * dout.y + wv.z * dout.z : 0.f;
        h[k] = v[i];
      }
      *(float4*)(dhr + 4 * q) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    // ---- phase B: dW1t, dW2t, db1, db2 over the tile's rows ----
#pragma unroll 2
    for (int row = 0; row < 128; ++row) {
      const float4 d4 = *(const float4*)(DH + row * MB_DS + 4 * jg);
      const float* hp = H1 + row * MB_HS + 8 * kg;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float hv = hp[i];
        g1[i][0] += hv * d4.x; g1[i][1] += hv * d4.y; g1[i][2] += hv * d4.z; g1[i][3] += hv * d4.w;
      }
    }
    if (t < 64) {
#pragma unroll 4
      for (int row = 0; row < 128; ++row) {
        const float hv = H2[row * MB_HS + t];
        const float4 d4 = *(const float4*)(DO + 4 * row);
        g2[0] += hv * d4.x; g2[1] += hv * d4.y; g2[2] += hv * d4.z;
        gb1 += DH[row * MB_DS + t];
      }
    } else if (t < 67) {
      for (int row = 0; row < 128; ++row) gb2 += DO[4 * row + (t - 64)];
    }
    __syncthreads();
    // ---- phase C: d h1 = [h1 > 0] W1 d h2, in place (every row has one owner; d h2 of this row is in registers) ----
#pragma unroll 1
    for (int q = 0; q < 16; ++q) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = 4 * q + i;
        const float4* wr = (const float4*)(W1 + k * 64);
        float acc = 0.f;
#pragma unroll
        for (int jq = 0; jq < 16; ++jq) {
          const float4 wv = wr[jq];
          acc += wv.x * h[4 * jq] + wv.y * h[4 * jq + 1] + wv.z * h[4 * jq + 2] + wv.w * h[4 * jq + 3];
        }
        v[i] = h1r[k] > 0.f ? acc : 0.f;
      }
      *(float4*)(dhr + 4 * q) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    // ---- phase D: dW0t, db0 ----
#pragma unroll 2
    for (int row = 0; row < 128; ++row) {
      const float4 d4 = *(const float4*)(DH + row * MB_DS + 4 * jg);
      const float* xp = X + row * MB_XS + 9 * kg;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const float xv = xp[i];
        g0[i][0] += xv * d4.x; g0[i][1] += xv * d4.y; g0[i][2] += xv * d4.z; g0[i][3] += xv * d4.w;
      }
    }
    if (t < 64) {
#pragma unroll 4
      for (int row = 0; row < 128; ++row) gb0 += DH[row * MB_DS + t];
    }
    // ---- phase E: d feature (rows 0..23 of the input; the encodings reach the MLP detached, microfacet.py:461-472) ----
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float4 v = *(const float4*)(dhr + 4 * q);
      h[4 * q] = v.x; h[4 * q + 1] = v.y; h[4 * q + 2] = v.z; h[4 * q + 3] = v.w;
    }
    const Seg seg = seg_setup(key, lane);
#pragma unroll 1
    for (int k3 = 0; k3 < 8; ++k3) {
      float df[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float4* wr = (const float4*)(W0 + (3 * k3 + i) * 64);
        float acc = 0.f;
#pragma unroll
        for (int jq = 0; jq < 16; ++jq) {
          const float4 wv = wr[jq];
          acc += wv.x * h[4 * jq] + wv.y * h[4 * jq + 1] + wv.z * h[4 * jq + 2] + wv.w * h[4 * jq + 3];
        }
        df[i] = active ? acc : 0.f;
      }
      seg_sum3(df, seg, lane);
      if (active && seg.head) {
        float* G = a.bgrad + (size_t)key * NMF_BGRAD + 10 + 3 * k3;
        atomicAdd(G, df[0]); atomicAdd(G + 1, df[1]); atomicAdd(G + 2, df[2]);
      }
    }
    __syncthreads();              // the next tile overwrites X / H1 / H2 / DH / DO
  }
  // ---- flush: one atomic per weight per CTA ----
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int k = 9 * kg + i;
    if (k < 66) {
#pragma unroll
      for (int c = 0; c < 4; ++c) if (g0[i][c] != 0.f) atomicAdd(a.w0t + k * 64 + 4 * jg + c, g0[i][c]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) if (g1[i][c] != 0.f) atomicAdd(a.w1t + (8 * kg + i) * 64 + 4 * jg + c, g1[i][c]);
  if (t < 64) {
    for (int c = 0; c < 3; ++c) if (g2[c] != 0.f) atomicAdd(a.w2t + 4 * t + c, g2[c]);
    if (gb0 != 0.f) atomicAdd(a.b0 + t, gb0);
    if (gb1 != 0.f) atomicAdd(a.b1 + t, gb1);
  } else if (t < 67) {
    if (gb2 != 0.f) atomicAdd(a.b2 + (t - 64), gb2);
  }
}


// ------------------------------------------------------------------------------------------------
// BRDF MLP backward on tcgen05 (csrc/nmf_mlp_tc_bwd.cuh): BF16 operands, fp32 accumulators in TMEM, weight gradients
// accumulated in TMEM over all tiles of the CTA.  scene.mlp_mode == 0 selects it; the fp32 kernel above is mlp_mode == 1.
// ------------------------------------------------------------------------------------------------
struct MfMlpTcArgs {
  const BSample* bs; const uint32_t* owner; const int* ray_count; int cap_rays; const int* tile_start; int n_chunks;
  const float4* dout; float* bgrad;
  float* w0t; float* b0; float* w1t; float* b1; float* w2t; float* b2;
  const void* w0b; const void* w1b; const void* w2b;
};
template <int MIXED>
__global__ void __launch_bounds__(MLP_THREADS, 1) k_mf_mlp_bwd_tc(const NmfScene s, const MfMlpTcArgs a) {
  extern __shared__ __align__(128) char tsm[];
  TbMlp tc;
  tb_init(tc, tsm, a.w0b, a.w1b, a.w2b);
  const int t = threadIdx.x, lane = t & 31;
  const uint32_t taddr = tc.tmem + ((uint32_t)(t & ~31) << 16);
  uint4* xrow = (uint4*)(tsm + TB_OFF_X) + t;          // + chunk * 128
  uint4* h1row = (uint4*)(tsm + TB_OFF_H1) + t;
  uint4* h2row = (uint4*)(tsm + TB_OFF_H2) + t;
  uint4* g2row = (uint4*)(tsm + TB_OFF_G2) + t;
  uint4* g1row = (uint4*)(tsm + TB_OFF_G1) + t;
  uint4* gorow = (uint4*)(tsm + TB_OFF_GO) + t;
  const int n_tiles = a.tile_start[a.n_chunks];
  bool first = true;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int chunk, n, r;
    mf_locate(a.tile_start, a.n_chunks, a.ray_count, a.cap_rays, tile, t, chunk, n, r);
    const uint32_t key = r < n ? a.owner[(size_t)chunk * a.cap_rays + r] : NMF_NO_OWNER;
    const bool active = key != NMF_NO_OWNER;
    // ---- this ray's input row (BF16) and output gradient ----
    float x[TC_K0];
#pragma unroll
    for (int i = 0; i < TC_K0; ++i) x[i] = 0.f;
    float4 dout = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      const BSample* b = a.bs + key;
      const float4 q1 = *(const float4*)b->V, q2 = *(const float4*)b->N, q4 = *(const float4*)b->diffuse;
      const nmf_v3 V = nmf_mk3(q1.x, q1.y, q1.z), N = nmf_mk3(q2.x, q2.y, q2.z);
      const int j = r - (int)__float_as_uint(q4.w);
      const float4* fq = (const float4*)b->frame;
      const float4 f0 = fq[0], f1 = fq[1], f2 = fq[2], f3 = fq[3], f4 = fq[4], f5 = fq[5];
      NmfGGXFrame fr;
      fr.t = nmf_mk3(f0.x, f0.y, f0.z); fr.b = nmf_mk3(f0.w, f1.x, f1.y); fr.V_l = nmf_mk3(f1.z, f1.w, f2.x);
      fr.Vs = nmf_mk3(f2.y, f2.z, f2.w); fr.T1 = nmf_mk3(f3.x, f3.y, f3.z); fr.T2 = nmf_mk3(f3.w, f4.x, f4.y);
      fr.a = f4.z;
      const float u1 = nmf_wrap01(__ldg(s.sobol + 2 * j) + f5.y), u2 = nmf_wrap01(__ldg(s.sobol + 2 * j + 1) + f5.z);
      const NmfGGX g = nmf_ggx_sample_f(fr, u1, u2, V, N, q1.w);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float4 f = *(const float4*)(b->feat + 4 * i);
        x[4 * i] = f.x; x[4 * i + 1] = f.y; x[4 * i + 2] = f.z; x[4 * i + 3] = f.w;
      }
      nmf_ish18_s(g.half_l, f4.w, f5.x, &x[24]);
      x[42] = g.half_l.x; x[43] = g.half_l.y; x[44] = g.half_l.z;
      nmf_ish18_s(g.diff_l, f4.w, f5.x, &x[45]);
      x[63] = g.diff_l.x; x[64] = g.diff_l.y; x[65] = g.diff_l.z;
      x[TC_ONE] = 1.f;
      dout = a.dout[(size_t)chunk * a.cap_rays + r];
    }
#pragma unroll
    for (int kc = 0; kc < TC_KC; ++kc) {
      uint4 v;
      if (MIXED) {
        v.x = tc_pack(x[8 * kc], x[8 * kc + 1]); v.y = tc_pack(x[8 * kc + 2], x[8 * kc + 3]);
        v.z = tc_pack(x[8 * kc + 4], x[8 * kc + 5]); v.w = tc_pack(x[8 * kc + 6], x[8 * kc + 7]);
      } else {
        v.x = tb_pack(x[8 * kc], x[8 * kc + 1]); v.y = tb_pack(x[8 * kc + 2], x[8 * kc + 3]);
        v.z = tb_pack(x[8 * kc + 4], x[8 * kc + 5]); v.w = tb_pack(x[8 * kc + 6], x[8 * kc + 7]);
      }
      xrow[kc * TC_ROWS] = v;
      if (kc >= 8) { h1row[kc * TC_ROWS] = v; h2row[kc * TC_ROWS] = v; }      // x64, x65, the constant 1 (biases), zeros
    }
    gorow[0] = make_uint4(tb_pack(dout.x, dout.y), tb_pack(dout.z, 0.f), 0u, 0u);
    gorow[TC_ROWS] = make_uint4(0u, 0u, 0u, 0u);
    // ---- forward, layer 1 ----
    tb_publish();
    if (t == 0) { tc_fence_after(); tb_gemm_kk<MIXED>(tc, TB_COL_D, TB_OFF_X, TB_OFF_W0, 64, TC_KC / 2); tc_commit(tc_smem_u32(tsm + TB_OFF_BAR)); }
    tb_wait(tc);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float h[32];
      tc_ld32(taddr + TB_COL_D + 32 * half, h);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        h1row[(4 * half + q) * TC_ROWS] = MIXED
            ? make_uint4(tc_pack_relu(h[8 * q], h[8 * q + 1]), tc_pack_relu(h[8 * q + 2], h[8 * q + 3]),
                         tc_pack_relu(h[8 * q + 4], h[8 * q + 5]), tc_pack_relu(h[8 * q + 6], h[8 * q + 7]))
            : make_uint4(tb_pack_relu(h[8 * q], h[8 * q + 1]), tb_pack_relu(h[8 * q + 2], h[8 * q + 3]),
                         tb_pack_relu(h[8 * q + 4], h[8 * q + 5]), tb_pack_relu(h[8 * q + 6], h[8 * q + 7]));
    }
    // ---- forward, layer 2 ----
    tb_publish();
    if (t == 0) { tc_fence_after(); tb_gemm_kk<MIXED>(tc, TB_COL_D, TB_OFF_H1, TB_OFF_W1, 64, TC_KC / 2); tc_commit(tc_smem_u32(tsm + TB_OFF_BAR)); }
    tb_wait(tc);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float h[32];
      tc_ld32(taddr + TB_COL_D + 32 * half, h);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        h2row[(4 * half + q) * TC_ROWS] = MIXED
            ? make_uint4(tc_pack_relu(h[8 * q], h[8 * q + 1]), tc_pack_relu(h[8 * q + 2], h[8 * q + 3]),
                         tc_pack_relu(h[8 * q + 4], h[8 * q + 5]), tc_pack_relu(h[8 * q + 6], h[8 * q + 7]))
            : make_uint4(tb_pack_relu(h[8 * q], h[8 * q + 1]), tb_pack_relu(h[8 * q + 2], h[8 * q + 3]),
                         tb_pack_relu(h[8 * q + 4], h[8 * q + 5]), tb_pack_relu(h[8 * q + 6], h[8 * q + 7]));
    }
    // ---- d H2 = dOut W2 (masked below);  d W2^T += H2^T dOut ----
    tb_publish();
    if (t == 0) {
      tc_fence_after();
      tb_gemm_data<MIXED>(tc, TB_COL_D, TB_OFF_GO, TB_OFF_W2, 16, 64, 1);
      tb_gemm_wgrad<MIXED>(tc, TB_COL_W2, TB_OFF_H2, TB_OFF_GO, 16, first);
      tc_commit(tc_smem_u32(tsm + TB_OFF_BAR));
    }
    tb_wait(tc);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float h[32];
      tc_ld32(taddr + TB_COL_D + 32 * half, h);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 m = h2row[(4 * half + q) * TC_ROWS];          // relu mask: the forward activation of this ray
        const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          o[e] = tb_pack((mm[e] & 0x7fffu) != 0u ? h[8 * q + 2 * e] : 0.f, (mm[e] & 0x7fff0000u) != 0u ? h[8 * q + 2 * e + 1] : 0.f);
        g2row[(4 * half + q) * TC_ROWS] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    // ---- d H1 = dH2 W1 (masked below);  d W1^T += H1^T dH2 ----
    tb_publish();
    if (t == 0) {
      tc_fence_after();
      tb_gemm_data<MIXED>(tc, TB_COL_D, TB_OFF_G2, TB_OFF_W1, 64, 64, 4);
      tb_gemm_wgrad<MIXED>(tc, TB_COL_W1, TB_OFF_H1, TB_OFF_G2, 64, first);
      tc_commit(tc_smem_u32(tsm + TB_OFF_BAR));
    }
    tb_wait(tc);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float h[32];
      tc_ld32(taddr + TB_COL_D + 32 * half, h);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 m = h1row[(4 * half + q) * TC_ROWS];
        const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          o[e] = tb_pack((mm[e] & 0x7fffu) != 0u ? h[8 * q + 2 * e] : 0.f, (mm[e] & 0x7fff0000u) != 0u ? h[8 * q + 2 * e + 1] : 0.f);
        g1row[(4 * half + q) * TC_ROWS] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    // ---- d X[:, :32] = dH1 W0;  d W0^T += X^T dH1 ----
    tb_publish();
    if (t == 0) {
      tc_fence_after();
      tb_gemm_data<MIXED>(tc, TB_COL_D, TB_OFF_G1, TB_OFF_W0, 64, 32, 4);
      tb_gemm_wgrad<MIXED>(tc, TB_COL_W0, TB_OFF_X, TB_OFF_G1, 64, first);
      tc_commit(tc_smem_u32(tsm + TB_OFF_BAR));
    }
    tb_wait(tc);
    first = false;
    {
      float df[32];
      tc_ld32(taddr + TB_COL_D, df);
      const Seg seg = seg_setup(key, lane);
#pragma unroll
      for (int k3 = 0; k3 < 8; ++k3) {
        float v[3] = {active ? df[3 * k3] : 0.f, active ? df[3 * k3 + 1] : 0.f, active ? df[3 * k3 + 2] : 0.f};
        seg_sum3(v, seg, lane);
        if (active && seg.head) {
          float* G = a.bgrad + (size_t)key * NMF_BGRAD + 10 + 3 * k3;
          atomicAdd(G, v[0]); atomicAdd(G + 1, v[1]); atomicAdd(G + 2, v[2]);
        }
      }
    }
    // the next tile's stores to the operand tiles are ordered behind this tile's MMAs by the commit / wait above, and behind
    // every thread's TMEM reads by the tb_publish() barrier that precedes the next MMA
  }
  // ---- weight gradients: lane = input index (row 66 = the bias), one atomic per weight per CTA ----
  if (!first) {
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if ((t & ~31) < TC_ONE + 1) {          // warp-uniform: tcgen05.ld is warp-collective (.sync.aligned)
      float* dst0 = t < 66 ? a.w0t + (size_t)t * 64 : (t == TC_ONE ? a.b0 : nullptr);
      float* dst1 = t < 64 ? a.w1t + (size_t)t * 64 : (t == TC_ONE ? a.b1 : nullptr);
      float* dst2 = t < 64 ? a.w2t + (size_t)t * 4 : (t == TC_ONE ? a.b2 : nullptr);
      float g[32];
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        tc_ld32(taddr + TB_COL_W0 + 32 * half, g);
        if (dst0) for (int i = 0; i < 32; ++i) if (g[i] != 0.f) atomicAdd(dst0 + 32 * half + i, g[i]);
        tc_ld32(taddr + TB_COL_W1 + 32 * half, g);
        if (dst1) for (int i = 0; i < 32; ++i) if (g[i] != 0.f) atomicAdd(dst1 + 32 * half + i, g[i]);
      }
      float g4[4];
      tc_ld4(taddr + TB_COL_W2, g4);
      if (dst2) for (int i = 0; i < 3; ++i) if (g4[i] != 0.f) atomicAdd(dst2 + i, g4[i]);
    }
  }
  tb_free(tc);
}

// ------------------------------------------------------------------------------------------------
// per surviving sample: material heads, basis_mat, appearance factors, normals
// (modules/render_modules.py:519-574, fields/tensoRF.py:402-405, fields/tensor_base.py:107-129, tensor_nerf.py:573-583)
// ------------------------------------------------------------------------------------------------
struct MfGradPtrs {
  float* d_plane[3]; float* d_line[3]; float* a_plane[3]; float* a_line[3]; float* basis_t; float* head_w; float* head_b;
  float* gpack[3]; float* glpack[3];
};
struct MfSampleBwdArgs {
  const BSample* bs; const int* n_bs; int cap_bs;             // the bounce samples (dense: every thread has work)
  const float* rays; const float* zvals; int n_steps; const Surv* surv; const int* n_surv; int cap_surv;   // k_mf_ori_bwd
  const uint32_t* survv; const int* survslot; const float* bgrad; float* vdw;
  float lambda_ori; int detach_N; float min_rough; MfGradPtrs g; int cap_vs;
};
#define SB_T 128
#ifndef NMF_SB_UNROLL
#define NMF_SB_UNROLL 1        // gather groups of the normal in flight per thread (experiments: NMF_NVCC_EXTRA=-DNMF_SB_UNROLL=2)
#endif
constexpr int kSbUnroll = NMF_SB_UNROLL;
#define SB_FS 25
#define SB_CS 73
#define SB_FLOATS (11 * 24 + 12 + SB_T * (SB_FS + 12 + SB_CS + SB_FS) + 72 * 24)
template <int LEVEL>
__global__ void __launch_bounds__(SB_T) k_mf_sample_bwd(const NmfScene s, const MfSampleBwdArgs a) {
  extern __shared__ __align__(16) float ssm[];
  float* sW = ssm; float* sB = sW + 11 * 24;
  float* F = sB + 12;                    // [128][25] un-noised feature
  float* DL = F + SB_T * SB_FS;          // [128][12] d loss / d head pre-activations
  float* CO = DL + SB_T * 12;            // [128][73] appearance coefficients
  float* DF = CO + SB_T * SB_CS;         // [128][25] d loss / d feature
  float* BT = DF + SB_T * SB_FS;         // [72][24] basis_t: the two 24 x 72 matrix-vector products read it as broadcast float4 rows
  const int t = threadIdx.x;
  for (int q = t; q < 11 * 24; q += SB_T) sW[q] = s.head_w[q];
  if (t < 11) sB[t] = s.head_b[t];
  for (int q = t; q < 72 * 24; q += SB_T) BT[q] = s.basis_t[q];
  __syncthreads();
  const int n = min(*a.n_bs, a.cap_bs);
  for (int base = blockIdx.x * SB_T; base < n; base += gridDim.x * SB_T) {
    const int si = base + t;
    float* fr = F + t * SB_FS; float* dl = DL + t * 12; float* co = CO + t * SB_CS; float* dfr = DF + t * SB_FS;
#pragma unroll
    for (int i = 0; i < 24; ++i) { fr[i] = 0.f; dfr[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 11; ++i) dl[i] = 0.f;
    const int slot = si < n ? si : -1;
    // (a slot whose ray range did not fit was never written: the call reports the overflow; its stale record is skipped)
    if (slot >= 0 && __float_as_int(((const float4*)a.bs[slot].N)->w) > 0 && a.bs[slot].pad < (uint32_t)a.cap_vs) {
      const BSample* b = a.bs + slot;
      const float4 q0 = *(const float4*)b->pos, q1 = *(const float4*)b->V;
      const float p[3] = {q0.x, q0.y, q0.z};
      const float w_s = q0.w;
      float xn[3];
      nmf_normalize_xyz(s, p, xn);
      const NmfTaps tp = nmf_vm_taps(s, xn);
      float grad[3] = {0.f, 0.f, 0.f};
#pragma unroll kSbUnroll
      for (int l = 0; l < 8; ++l) nmf_normal_lane(s, tp, l, grad);
      const nmf_v3 nrm = nmf_normal_from_grad(s, grad);
      const nmf_v3 V = nmf_mk3(q1.x, q1.y, q1.z);
      const float vn = nmf_dot(V, nrm);
      const float sgn = vn > 0.f ? 1.f : (vn < 0.f ? -1.f : 0.f);
      float dn[3] = {0.f, 0.f, 0.f};
      if (LEVEL == 0 && a.lambda_ori != 0.f && vn < 0.f) {
        // ori_lambda * sum w min(v.n, 0)^2 (tensor_nerf.py:573-583): to the weight and, through the normal, to the density factors
        atomicAdd(a.vdw + b->pad, a.lambda_ori * vn * vn);
        const float k2 = a.lambda_ori * w_s * 2.0f * vn;
        dn[0] = k2 * V.x; dn[1] = k2 * V.y; dn[2] = k2 * V.z;
      }
      if (slot >= 0) {
        const float* G = a.bgrad + (size_t)slot * NMF_BGRAD;
        float coef[72];
        nmf_app_coef(s, tp, coef);
#pragma unroll 1
        for (int j = 0; j < 72; ++j) co[j] = coef[j];
        // feat = basis_t^T coef (tensoRF.py:405): register-blocked -- one broadcast float4 row piece of basis_t per 4 FMAs, the
        // 72 coefficients and 24 outputs statically indexed (a rolled output loop kept both operands in local / global memory:
        // two L1 loads per FMA, 13 stall cycles per issued instruction)
        float feat[24];
#pragma unroll
        for (int oo = 0; oo < 24; ++oo) feat[oo] = 0.f;
#pragma unroll
        for (int j = 0; j < 72; ++j) {
          const float cj = coef[j];
#pragma unroll
          for (int q4 = 0; q4 < 6; ++q4) {
            const float4 b4 = *(const float4*)(BT + j * 24 + 4 * q4);
            feat[4 * q4] += b4.x * cj; feat[4 * q4 + 1] += b4.y * cj; feat[4 * q4 + 2] += b4.z * cj; feat[4 * q4 + 3] += b4.w * cj;
          }
        }
#pragma unroll
        for (int oo = 0; oo < 24; ++oo) fr[oo] = feat[oo];
        float sh[9];
        nmf_sh9(nrm, sh);
        float g_alb[3], dR0[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float e = 0.f;
#pragma unroll
          for (int i = 0; i < 9; ++i) e += __ldg(s.sh_conv + i * 3 + c) * sh[i];
          g_alb[c] = G[3 + c] * e;               // diffuse = albedo * E(n), E under no_grad (microfacet.py:304-316)
          dR0[c] = G[c];
        }
        float drough = G[6];
        {                                         // r1.clip(min=min_rough) (microfacet.py:361-363): closed-interval gate
          float lin = sB[9];
#pragma unroll
          for (int i = 0; i < 24; ++i) lin += sW[9 * 24 + i] * feat[i];
          const float rough = nmf_clampf(nmf_sigmoid(lin + s.roughness_bias) / 2.0f, 1e-2f, 1.0f);
          if (rough < a.min_rough) drough = 0.f;
        }
        float dlin[11];
        nmf_heads_dlin(feat, sW, sB, s.diffuse_mul, s.diffuse_bias, s.f0_bias, s.roughness_bias, g_alb, dR0, drough, dlin);
        float df[24];
#pragma unroll
        for (int i = 0; i < 11; ++i) dl[i] = dlin[i];
#pragma unroll 1
        for (int kk = 0; kk < 24; ++kk) {
          float acc = G[10 + kk];
#pragma unroll
          for (int h = 0; h < 11; ++h) acc += dlin[h] * sW[h * 24 + kk];
          df[kk] = acc;
          dfr[kk] = acc;
        }
        float dcoef[72];
#pragma unroll
        for (int j = 0; j < 72; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int q4 = 0; q4 < 6; ++q4) {
            const float4 b4 = *(const float4*)(BT + j * 24 + 4 * q4);
            acc += b4.x * df[4 * q4] + b4.y * df[4 * q4 + 1] + b4.z * df[4 * q4 + 2] + b4.w * df[4 * q4 + 3];
          }
          dcoef[j] = acc;
        }
        float* ap[3] = {a.g.a_plane[0], a.g.a_plane[1], a.g.a_plane[2]};
        float* al[3] = {a.g.a_line[0], a.g.a_line[1], a.g.a_line[2]};
        nmf_app_bwd(s, tp, dcoef, ap, al);
        if (!a.detach_N) { dn[0] += sgn * G[7]; dn[1] += sgn * G[8]; dn[2] += sgn * G[9]; }
      } else {
#pragma unroll 1
        for (int j = 0; j < 72; ++j) co[j] = 0.f;
      }
      if (dn[0] != 0.f || dn[1] != 0.f || dn[2] != 0.f) {
        float dgrad[3];
        nmf_normal_vec_bwd(s, grad, dn, dgrad);
        if (dgrad[0] != 0.f || dgrad[1] != 0.f || dgrad[2] != 0.f) {
          float* gp[3] = {a.g.gpack[0], a.g.gpack[1], a.g.gpack[2]};
          float* gl[3] = {a.g.glpack[0], a.g.glpack[1], a.g.glpack[2]};
          nmf_normal_bwd4(s, tp, dgrad, gp, gl);
        }
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < 72; ++j) co[j] = 0.f;
    }
    __syncthreads();
    // tile contractions: d head_w[h][k] = sum_j dlin_j[h] feat_j[k];  d basis_t[jj][oo] = sum_j coef_j[jj] dfeat_j[oo].
    // Threads 0..95 own a 3 x 6 register block of d basis_t each (9 shared loads per 18 FMAs; one output per thread and trip
    // cost 2 loads per FMA), threads 96..127 the 275 head outputs.
    if (t < 96) {
      const int jb = (t >> 2) * 3, ob = (t & 3) * 6;
      float acc[3][6];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 6; ++q) acc[r][q] = 0.f;
#pragma unroll 4
      for (int j = 0; j < SB_T; ++j) {
        const float c0 = CO[j * SB_CS + jb], c1 = CO[j * SB_CS + jb + 1], c2 = CO[j * SB_CS + jb + 2];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const float d = DF[j * SB_FS + ob + q];
          acc[0][q] += c0 * d; acc[1][q] += c1 * d; acc[2][q] += c2 * d;
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 6; ++q)
          if (acc[r][q] != 0.f) atomicAdd(a.g.basis_t + (jb + r) * 24 + ob + q, acc[r][q]);
    } else {
      for (int idx = t - 96; idx < 11 * 24 + 11; idx += 32) {
        float acc = 0.f;
        if (idx < 11 * 24) {
          const int h = idx / 24, kk = idx - 24 * h;
#pragma unroll 8
          for (int j = 0; j < SB_T; ++j) acc += DL[j * 12 + h] * F[j * SB_FS + kk];
          if (acc != 0.f) atomicAdd(a.g.head_w + idx, acc);
        } else {
          const int h = idx - 11 * 24;
#pragma unroll 8
          for (int j = 0; j < SB_T; ++j) acc += DL[j * 12 + h];
          if (acc != 0.f) atomicAdd(a.g.head_b + h, acc);
        }
      }
    }
    __syncthreads();
  }
}

// survivors WITHOUT a bounce sample whose normal faces away from the viewer (marked -2 by k_shade): only the orientation loss
// reaches them (tensor_nerf.py:573-583).  Rare once a scene has formed; thread per survivor with an early exit.
__global__ void __launch_bounds__(128) k_mf_ori_bwd(const NmfScene s, const MfSampleBwdArgs a) {
  const int n = min(*a.n_surv, a.cap_surv);
  for (int si = blockIdx.x * blockDim.x + threadIdx.x; si < n; si += gridDim.x * blockDim.x) {
    if (a.survslot[si] != -2) continue;
    const Surv sv = a.surv[si];
    const int ray = (int)sv.ray, k = (int)sv.step;
    float o[3], d[3], p[3], xn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { o[i] = __ldg(a.rays + (size_t)ray * 6 + i); d[i] = __ldg(a.rays + (size_t)ray * 6 + 3 + i); }
    nmf_step_pos(o, d, a.zvals[(size_t)ray * a.n_steps + k], p);
    nmf_normalize_xyz(s, p, xn);
    const NmfTaps tp = nmf_vm_taps(s, xn);
    float grad[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int l = 0; l < 8; ++l) nmf_normal_lane(s, tp, l, grad);
    const nmf_v3 nrm = nmf_normal_from_grad(s, grad);
    const nmf_v3 V = nmf_mk3(-d[0], -d[1], -d[2]);
    const float vn = nmf_dot(V, nrm);
    if (!(vn < 0.f)) continue;
    atomicAdd(a.vdw + a.survv[si], a.lambda_ori * vn * vn);
    const float k2 = a.lambda_ori * sv.w * 2.0f * vn;
    const float dn[3] = {k2 * V.x, k2 * V.y, k2 * V.z};
    float dgrad[3];
    nmf_normal_vec_bwd(s, grad, dn, dgrad);
    if (dgrad[0] != 0.f || dgrad[1] != 0.f || dgrad[2] != 0.f) {
      float* gp[3] = {a.g.gpack[0], a.g.gpack[1], a.g.gpack[2]};
      float* gl[3] = {a.g.glpack[0], a.g.glpack[1], a.g.glpack[2]};
      nmf_normal_bwd4(s, tp, dgrad, gp, gl);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// compositing backward over ALL valid samples of a ray (modules/tensor_nerf.py:19-35; zero-weight samples carry a density
// gradient too), softplus', density factors (fields/tensoRF.py:392-400)
// ------------------------------------------------------------------------------------------------
struct MfCompArgs {
  const float* rays; const float* zvals; int n_steps; int n; int group; const int* n_active; const uint8_t* whole;
  const int* nvalid; const int* vbase; const VSmp* vs; const float* vdw; int cap_vs; const float* g_lin;
  float* d_plane[3]; float* d_line[3];
};
template <int LEVEL>
__global__ void __launch_bounds__(256) k_mf_composite_bwd(const NmfScene s, const MfCompArgs a) {
  const int lane = threadIdx.x & 31;
  const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= a.n) return;
  if (LEVEL == 0) { if (!a.whole[ray]) return; }
  else { const int chunk = ray / a.group; if (ray - chunk * a.group >= a.n_active[chunk]) return; }
  const int n = a.nvalid[ray], b = a.vbase[ray];
  if (n <= 0 || b < 0 || b + n > a.cap_vs) return;
  float o[3], d[3];
  for (int i = 0; i < 3; ++i) { o[i] = a.rays[(size_t)ray * 6 + i]; d[i] = a.rays[(size_t)ray * 6 + 3 + i]; }
  const float base_dw = a.g_lin[4 * (size_t)ray + 3];
  const float* zrow = a.zvals + (size_t)ray * a.n_steps;
  float* gp[3] = {a.d_plane[0], a.d_plane[1], a.d_plane[2]};
  float* gl[3] = {a.d_line[0], a.d_line[1], a.d_line[2]};
  float carry = 0.f;                              // sum of dw_j w_j over the samples after the current group
  for (int hi = n; hi > 0; hi -= 32) {
    const int i = hi - 1 - lane;                  // lane 0 = last sample of the group
    const bool ok = i >= 0;
    VSmp v; v.k = 0; v.f = 0.f; v.alpha = 0.f; v.T = 0.f;
    float dw = 0.f;
    if (ok) { v = a.vs[b + i]; dw = base_dw + a.vdw[b + i]; }
    const float wt = v.alpha * v.T;
    float incl = dw * wt;
    for (int off = 1; off < 32; off <<= 1) {
      const float u = __shfl_up_sync(FULL, incl, off);
      if (lane >= off) incl += u;
    }
    const float suffix = carry + incl - dw * wt;
    carry += __shfl_sync(FULL, incl, 31);
    if (ok) {
      const int k = (int)v.k;
      const float z = zrow[k];
      const float dist = (k + 1 < a.n_steps ? NMF_SUB(zrow[k + 1], z) : 0.f) * s.distance_scale;
      const float dsigma = nmf_composite_bwd(dw, v.T, v.alpha, dist, suffix);
      const float df = dsigma * nmf_feature2density_grad(v.f, s.density_shift);
      if (df != 0.f) {
        float p[3], xn[3];
        nmf_step_pos(o, d, z, p);
        nmf_normalize_xyz(s, p, xn);
        const NmfTaps t = nmf_vm_taps(s, xn);
        nmf_density_bwd(s, t, df, gp, gl);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int nmf_train_microfacet(const NmfScene* scene, const NmfRender* rp_in, const NmfRenderTrain* tr,
                                    const NmfMicrofacetTrain* tp, const float* rays, const float* gt,
                                    const NmfMicrofacetGrads* grads, const NmfImages* out, const NmfCounters* counters,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
  if (!scene || !rp_in || !tr || !tp || !rays || !gt || !grads || !out || !workspace || !tp->loss) return NMF_E_ARG;
  if (scene->model != 0) return NMF_E_UNSUPPORTED;
  for (int p = 0; p < 3; ++p)
    if (!grads->d_plane[p] || !grads->d_line[p] || !grads->a_plane[p] || !grads->a_line[p] || !grads->normals.gpack[p] ||
        !grads->normals.glpack[p])
      return NMF_E_ARG;
  if (!grads->basis_t || !grads->head_w || !grads->head_b || !grads->w0t || !grads->b0 || !grads->w1t || !grads->b1 ||
      !grads->w2t || !grads->b2 || !grads->gsat || !grads->d_mipbias)
    return NMF_E_ARG;
  NmfRender rp = *rp_in;
  rp.skip_eps = 0.f;       // the reference shades every sample with a positive weight; the weight cut is an eval-time option
  rp.t_cut = 0.f;
  WS w;
  int st = nmf_render_impl_train(scene, &rp, tr, rays, out, counters, workspace, workspace_bytes, stream_, &w);
  if (st) return st;
  const NmfScene& s = *scene;
  cudaStream_t cs = (cudaStream_t)stream_;
  const int n = rp.n_rays, nc = w.n_chunks;
  const bool retrace = s.max_retrace > 0 && w.n_rays1 > 0;
  CK(cudaMemsetAsync(tp->loss, 0, 3 * sizeof(double), cs));
  k_mf_zero_bgrad<<<m_sm_count() * 4, 256, 0, cs>>>(w.bgrad0, w.n_bs, w.cap_bs0);
  CKL();
  if (retrace) {
    k_mf_zero_bgrad<<<m_sm_count() * 4, 256, 0, cs>>>(w.bgrad1, w.n_bs + 1, w.cap_bs1);
    CKL();
    CK(cudaMemsetAsync(w.jac1, 0, (size_t)w.n_rays1 * 12 * 4, cs));
    CK(cudaMemsetAsync(w.g_lin1, 0, (size_t)w.n_rays1 * 16, cs));
  }
  MfLossArgs la = {w.accum0, w.acc0, tr->whole_valid, gt, n, tp->lambda_pred, w.g_lin0, tp->loss, w.stat4};
  k_mf_loss<<<(n + 255) / 256, 256, 0, cs>>>(la);
  CKL();
  static int g_tang = 0, g_bb0 = 0, g_bb1 = 0, g_sb0 = 0, g_sb1 = 0, g_ori = 0;
  if (retrace) {
    MfTangArgs ta = {w.bs1, w.brays1, w.owner1, w.ray_count1, w.cap_rays1, w.tile_start1, nc, w.jac1};
    k_mf_tangent1<<<m_resident(k_mf_tangent1, MLP_THREADS, 0, &g_tang), MLP_THREADS, 0, cs>>>(s, ta);
    CKL();
    k_mf_sec_tangent<<<(w.n_rays1 + 127) / 128, 128, 0, cs>>>(s, w.rays1, w.mip1, w.acc1, w.n_sec, s.max_retrace, w.n_rays1, w.jac1);
    CKL();
  }
  MfBounceBwdArgs b0 = {};
  b0.bs = w.bs0; b0.brays = w.brays0; b0.owner = w.owner0; b0.ray_count = w.ray_count0; b0.cap_rays = w.cap_rays0;
  b0.tile_start = w.tile_start0; b0.n_chunks = nc; b0.g_lin = w.g_lin0; b0.rgb1 = w.rgb1; b0.jac1 = w.jac1; b0.g_lin1 = w.g_lin1;
  b0.max_retrace = s.max_retrace; b0.bgrad = w.bgrad0; b0.vdw = w.vdw0; b0.dout = w.dout0; b0.gsat = grads->gsat;
  b0.d_mipbias = grads->d_mipbias; b0.detach_N = tp->detach_N;
  k_mf_bounce_bwd<0><<<m_resident(k_mf_bounce_bwd<0>, MLP_THREADS, 0, &g_bb0), MLP_THREADS, 0, cs>>>(s, b0);
  CKL();
  if (retrace) {
    k_mf_sec_bwd<<<(w.n_rays1 + 127) / 128, 128, 0, cs>>>(s, w.rays1, w.mip1, w.acc1, w.n_sec, s.max_retrace, w.n_rays1, w.g_lin1,
                                                         grads->gsat, grads->d_mipbias);
    CKL();
    MfBounceBwdArgs b1 = {};
    b1.bs = w.bs1; b1.brays = w.brays1; b1.owner = w.owner1; b1.ray_count = w.ray_count1; b1.cap_rays = w.cap_rays1;
    b1.tile_start = w.tile_start1; b1.n_chunks = nc; b1.g_lin = w.g_lin1; b1.bgrad = w.bgrad1; b1.vdw = w.vdw1; b1.dout = w.dout1;
    b1.gsat = grads->gsat; b1.d_mipbias = grads->d_mipbias; b1.detach_N = tp->detach_N;
    k_mf_bounce_bwd<1><<<m_resident(k_mf_bounce_bwd<1>, MLP_THREADS, 0, &g_bb1), MLP_THREADS, 0, cs>>>(s, b1);
    CKL();
  }
  static bool attr_done = false;
  if (!attr_done) {
    CK(cudaFuncSetAttribute(k_mf_mlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MB_FLOATS * sizeof(float))));
    CK(cudaFuncSetAttribute(k_mf_mlp_bwd_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_mf_mlp_bwd_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_mf_sample_bwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SB_FLOATS * sizeof(float))));
    CK(cudaFuncSetAttribute(k_mf_sample_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SB_FLOATS * sizeof(float))));
    attr_done = true;
  }
  const bool tc_bwd = s.mlp_mode == 0 && s.brdf_w0b && s.brdf_w1b && s.brdf_w2b;
  // operand formats of the tcgen05 reverse kernel: all BF16.  (FP16 activations / weights with BF16 gradients in one MMA --
  // a_format != b_format -- was tried on a B200: the launch faults, kind::f16 wants both operands in one format.  The
  // instantiation is kept behind NMF_TC_BWD_MIXED=1 only as the record of that experiment.)
  static int tc_mixed = -1;
  if (tc_mixed < 0) { const char* e = getenv("NMF_TC_BWD_MIXED"); tc_mixed = (e && e[0] == '1') ? 1 : 0; }
  for (int lvl = 0; lvl < (retrace ? 2 : 1); ++lvl) {
    const BSample* bs = lvl ? w.bs1 : w.bs0;
    const uint32_t* owner = lvl ? w.owner1 : w.owner0;
    const int* rc = lvl ? w.ray_count1 : w.ray_count0;
    const int cap = lvl ? w.cap_rays1 : w.cap_rays0;
    const int* ts = lvl ? w.tile_start1 : w.tile_start0;
    const float4* dout = lvl ? w.dout1 : w.dout0;
    float* bgrad = lvl ? w.bgrad1 : w.bgrad0;
    if (tc_bwd) {
      MfMlpTcArgs ma = {bs, owner, rc, cap, ts, nc, dout, bgrad, grads->w0t, grads->b0, grads->w1t, grads->b1, grads->w2t, grads->b2,
                        tc_mixed ? s.brdf_w0u : s.brdf_w0b, tc_mixed ? s.brdf_w1u : s.brdf_w1b, tc_mixed ? s.brdf_w2u : s.brdf_w2b};
      if (tc_mixed) k_mf_mlp_bwd_tc<1><<<m_sm_count(), MLP_THREADS, TB_SMEM_BYTES, cs>>>(s, ma);
      else k_mf_mlp_bwd_tc<0><<<m_sm_count(), MLP_THREADS, TB_SMEM_BYTES, cs>>>(s, ma);
    } else {
      MfMlpArgs ma = {bs, owner, rc, cap, ts, nc, dout, bgrad, grads->w0t, grads->b0, grads->w1t, grads->b1, grads->w2t, grads->b2};
      k_mf_mlp_bwd<<<m_sm_count(), MLP_THREADS, MB_FLOATS * sizeof(float), cs>>>(s, ma);
    }
    CKL();
  }
  MfGradPtrs gp;
  for (int p = 0; p < 3; ++p) {
    gp.d_plane[p] = grads->d_plane[p]; gp.d_line[p] = grads->d_line[p]; gp.a_plane[p] = grads->a_plane[p];
    gp.a_line[p] = grads->a_line[p]; gp.gpack[p] = grads->normals.gpack[p]; gp.glpack[p] = grads->normals.glpack[p];
  }
  gp.basis_t = grads->basis_t; gp.head_w = grads->head_w; gp.head_b = grads->head_b;
  MfSampleBwdArgs s0 = {w.bs0, w.n_bs, w.cap_bs0, rays, w.zvals0, s.n_steps, w.surv0, w.n_surv, w.cap_surv0, w.survv0, w.survslot0,
                        w.bgrad0, w.vdw0, tp->lambda_ori, tp->detach_N, tr->min_rough, gp, w.cap_vs0};
  k_mf_sample_bwd<0><<<m_resident(k_mf_sample_bwd<0>, SB_T, SB_FLOATS * sizeof(float), &g_sb0), SB_T, SB_FLOATS * sizeof(float), cs>>>(s, s0);
  CKL();
  if (tp->lambda_ori != 0.f) {
    k_mf_ori_bwd<<<m_resident(k_mf_ori_bwd, 128, 0, &g_ori), 128, 0, cs>>>(s, s0);
    CKL();
  }
  if (retrace) {
    MfSampleBwdArgs s1 = {w.bs1, w.n_bs + 1, w.cap_bs1, w.rays1, w.zvals1, s.n_steps, w.surv1, w.n_surv + 1, w.cap_surv1, w.survv1,
                          w.survslot1, w.bgrad1, w.vdw1, 0.f, tp->detach_N, tr->min_rough, gp, w.cap_vs1};
    k_mf_sample_bwd<1><<<m_resident(k_mf_sample_bwd<1>, SB_T, SB_FLOATS * sizeof(float), &g_sb1), SB_T, SB_FLOATS * sizeof(float), cs>>>(s, s1);
    CKL();
  }
  MfCompArgs c0 = {};
  c0.rays = rays; c0.zvals = w.zvals0; c0.n_steps = s.n_steps; c0.n = n; c0.group = rp.chunk; c0.whole = tr->whole_valid;
  c0.nvalid = w.nvalid0; c0.vbase = w.vbase0; c0.vs = w.vs0; c0.vdw = w.vdw0; c0.cap_vs = w.cap_vs0; c0.g_lin = w.g_lin0;
  for (int p = 0; p < 3; ++p) { c0.d_plane[p] = grads->d_plane[p]; c0.d_line[p] = grads->d_line[p]; }
  k_mf_composite_bwd<0><<<(n + 7) / 8, 256, 0, cs>>>(s, c0);
  CKL();
  if (retrace) {
    MfCompArgs c1 = c0;
    c1.rays = w.rays1; c1.zvals = w.zvals1; c1.n = w.n_rays1; c1.group = s.max_retrace; c1.n_active = w.n_sec; c1.whole = nullptr;
    c1.nvalid = w.nvalid1; c1.vbase = w.vbase1; c1.vs = w.vs1; c1.vdw = w.vdw1; c1.cap_vs = w.cap_vs1; c1.g_lin = w.g_lin1;
    k_mf_composite_bwd<1><<<(w.n_rays1 + 7) / 8, 256, 0, cs>>>(s, c1);
    CKL();
  }
  return NMF_OK;
}
