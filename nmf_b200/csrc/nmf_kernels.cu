// nmf_b200 -- sm_100a kernels of the NMF per-ray render path and the C ABI over them (include/nmf_b200.h).
//
// One call of nmf_render_rays renders every chunk of a ray batch with a fixed sequence of launches (no host
// synchronisation, CUDA-graph capturable).  Phases (DESIGN.md "Kernels"):
//   k_march<0>     warp per ray: slab test, dense step enumeration (four 32-step groups per occupancy round trip, exact
//                  early exit past the box), AABB + occupancy-bit test (bit-exact), VM density gather (4 lanes per
//                  sample), transmittance scan, warp-compacted survivor list
//   k_shade<0>     per warp, 32 survivors: gather phase (8 lanes per sample: appearance taps, 3xTF32 mma.sync basis
//                  contraction, smoothed-gradient normal) + shade phase (1 lane per sample: material heads, SH
//                  irradiance, bounce count, debug maps, A19 sums, bounce-sample record incl. the per-sample GGX frame)
//   k_tile_prefix  flat list of 128-ray tiles over the per-chunk bounce-ray regions + per-tile descriptors (TileWalk)
//   k_bounce<0>    persistent CTAs, thread per bounce ray: Sobol + GGX VNDF sample, ISH encodings, BRDF MLP on
//                  tcgen05 (fp16 operands, TMEM accumulators, TMA-staged weights), retrace score
//   k_select       per chunk (a CTA, or a thread-block cluster with DSMEM histograms when there are few chunks):
//                  radix-select of the top max_retrace scores -> secondary rays
//   k_march<1>, k_shade<1>, k_bounce<1>, k_incoming<1>, k_finish1: the retraced rays (recur = 1)
//   k_incoming<0>  per primary bounce ray: retraced radiance or environment lookup, Fresnel mix, per-sample sums
//   k_reduce0      per bounce sample: mean over its rays, composite into the pixel
//   k_finish0      per ray: tonemap, background, auxiliary maps, per-chunk A19 statistics (in three stages when the
//                  caller wants host buffers: maps leave for the host as soon as they are final)
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include "nmf_field.cuh"
#include "nmf_mlp_tc.cuh"
#include "nmf_train.cuh"   // train mode: jittered distances, dynamic batch truncation

#include "nmf_render_ws.cuh"

// ================================================================================================
// k_march: samplers/alphagrid.py:131-207,278-370 + fields/tensoRF.py:392-400 + tensor_nerf.py:19-35
// ================================================================================================
struct MarchArgs {
  const float* rays;       // level 0: (n,6); level 1: ws.rays1
  int n;                   // rays (level 1: slots)
  int group;               // rays per chunk (level 0: chunk; level 1: max_retrace)
  const int* n_active;     // level 1: n_sec[chunk]
  const uint64_t* keys;    // level 1
  uint64_t seed, ray_id0;
  float skip_eps, t_cut;
  float* tmin; float* acc; float* depth; int* termk; int* nvalid;
  int* n_samples; int* n_cand; double* wsum;
  Surv* surv; int* n_surv; int cap_surv; unsigned* error;
  float* zvals;            // train mode: (n, n_steps) jittered distances (level 0: read, level 1: written here)
  const uint8_t* whole;    // train mode, level 0: rays kept by the dynamic batch truncation
  // train mode: every valid sample is kept for the reverse pass (csrc/nmf_mf_train.cu)
  VSmp* vs; float* vdw; int* vbase; int* n_vs; int cap_vs; uint32_t* survv;
};

// The 8-corner occupancy test of a step that lies exactly on a lattice plane.  Rare (a handful per thousand rays), so it
// is kept out of line: inlined four times it doubled the size of the march loop past the instruction cache.
__device__ __noinline__ bool march_occupied_on_lattice(const uint32_t* vox, const uint32_t* cell, int ow, int oh, int od, int opitch,
                                                       float xn0, float xn1, float xn2) {
  return nmf_occupied(vox, cell, ow, oh, od, opitch, xn0, xn1, xn2);
}

// TRAIN = 1 (TensorNeRF.forward(is_train=True), alphagrid.py:167-173): the distance of dense step k is read from the
// ray's row of jittered cumulative distances instead of tmin + stepsize * k.  Level 0 rows come from k_train_sample (the
// dynamic batch truncation sits between the sampling and the march); level 1 rows are written here (the recursion runs
// with dynamic_batch_size=False, tensor_nerf.py:299).  The distances are still increasing in k, so the exit test holds.
template <int LEVEL, int TRAIN>
__global__ void __launch_bounds__(256, 4) k_march(const NmfScene s, const MarchArgs a) {
  __shared__ uint16_t s_list[8][NMF_MAX_STEPS];
  __shared__ uint32_t s_coarse[NMF_MAX_COARSE_WORDS];
  const bool use_coarse = s.has_occ && s.occ_coarse != nullptr;
  if (use_coarse) {
    const int nw = (s.ocw * s.och * s.ocd + 31) >> 5;
    for (int i = threadIdx.x; i < nw; i += 256) s_coarse[i] = s.occ_coarse[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = s.n_steps;
  const unsigned lt = (1u << lane) - 1u;
  uint16_t* list = s_list[warp];
  for (int ray = blockIdx.x * 8 + warp; ray < a.n; ray += gridDim.x * 8) {
    const int chunk = ray / a.group;
    if (LEVEL == 1 && (ray - chunk * a.group) >= a.n_active[chunk]) continue;
    if (TRAIN && LEVEL == 0 && !a.whole[ray]) {      // alphagrid.py:359-364: the ray is dropped from the batch
      if (lane == 0) { a.tmin[ray] = 0.f; a.acc[ray] = 0.f; a.nvalid[ray] = 0; a.depth[ray] = 0.f; a.termk[ray] = -1; }
      continue;
    }
    float o[3], d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { o[i] = __ldg(a.rays + (size_t)ray * 6 + i); d[i] = __ldg(a.rays + (size_t)ray * 6 + 3 + i); }
    const float near_ = LEVEL == 0 ? s.near : NMF_MUL(3.0f, s.stepsize);     // tensor_nerf.py:302 override_near
    const float tmin = nmf_ray_tmin(o, d, s.aabb0, s.aabb1, near_, s.far);
    uint64_t key = 0;
    if (LEVEL == 1) key = a.keys[ray];
    float* zrow = TRAIN ? a.zvals + (size_t)ray * S : nullptr;
    if (TRAIN && LEVEL == 1) {
      nmf_warp_jitter_z(key, tmin, s.stepsize, S, zrow, lane);
      __syncwarp();
    }
#define MARCH_Z(k_) (TRAIN ? zrow[k_] : nmf_step_z(tmin, s.stepsize, (k_)))
    // ---- pass A: enumerate the dense steps, keep those inside the box and in an occupied cell ----
    // Each coordinate of p_k = o + d * (tmin + step * k) is monotone in k also in fp32 (mul and add are monotone), so
    // the in-box steps are one contiguous range: after the first out-of-box step that follows an in-box one, every
    // later step is out of the box as well and the enumeration can stop (bit-exact, 40 % of the dense steps).
    // The steps are tested four 32-step groups at a time: all four occupancy words are requested before the first is
    // used (the march is bound by the latency of that dependent load, not by instruction issue).
    int nv = 0, cand = 0;
    float usum = 0.f;
    bool seen_inside = false, done = false;
    int k0 = 0;
    for (; k0 < S && !done; k0 += 128) {
      uint32_t word[4];
      int shift[4], state[4];                // 0 = outside the box, 1 = one cell bit decides, 2 = on a lattice plane,
                                             // 3 = no occupancy grid, 4 = in the box but in an empty coarse cell
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + 32 * j + lane;
        state[j] = 0; word[j] = 0; shift[j] = 0;
        if (k < S) {
          float p[3];
          nmf_step_pos(o, d, MARCH_Z(k), p);
          if (nmf_inside(p, s.aabb0, s.aabb1)) {
            ++cand;
            state[j] = 3;
            if (use_coarse && !nmf_occ_coarse(s, s_coarse, p)) {
              state[j] = 4;                  // most of the box: no voxel near this cell, the exact test cannot succeed
            } else if (s.has_occ) {
              float xn[3];
              long long wi;
              nmf_normalize_xyz(s, p, xn);
              if (nmf_occ_fast(s.ow, s.oh, s.od, s.opitch, xn[0], xn[1], xn[2], &wi, &shift[j])) {
                state[j] = 1;
                if (wi >= 0) word[j] = __ldg(s.occ_cell + wi);
              } else {
                state[j] = 2;
              }
            }
          }
          if (LEVEL == 1) usum += nmf_uniform(nmf_mix64(key, (uint64_t)k), NMF_STREAM_BOUNCE);   // pt_selectors.py:25
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + 32 * j + lane;
        bool ok = state[j] == 3 || (state[j] == 1 && ((word[j] >> shift[j]) & 1u));
        if (state[j] == 2) {                 // rare: a sample exactly on a lattice plane takes the 8-corner path
          float p[3], xn[3];
          nmf_step_pos(o, d, MARCH_Z(k), p);
          nmf_normalize_xyz(s, p, xn);
          ok = march_occupied_on_lattice(s.occ_vox, s.occ_cell, s.ow, s.oh, s.od, s.opitch, xn[0], xn[1], xn[2]);
        }
        const unsigned m = __ballot_sync(FULL, ok);
        if (ok) list[nv + __popc(m & lt)] = (uint16_t)k;
        nv += __popc(m);
        const unsigned mb = __ballot_sync(FULL, state[j] != 0);
        if (mb) seen_inside = true;
        if (seen_inside && !(mb >> 31)) done = true;     // the last step of this group is past the exit point
      }
    }
    if (LEVEL == 1) {            // pt_selectors.py:25 adds a keyed uniform for EVERY dense step, also the skipped ones
      for (int k = k0 + lane; k < S; k += 32) usum += nmf_uniform(nmf_mix64(key, (uint64_t)k), NMF_STREAM_BOUNCE);
    }
    __syncwarp();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cand += __shfl_xor_sync(FULL, cand, off);
    int vb = 0;
    if (TRAIN && a.vs) {       // the ray's range in the valid-sample list kept for the reverse pass
      if (lane == 0) { vb = nv ? atomicAdd(a.n_vs, nv) : 0; a.vbase[ray] = vb; }
      vb = __shfl_sync(FULL, vb, 0);
      if (nv && vb + nv > a.cap_vs) { if (lane == 0) atomicOr(a.error, NMF_DEV_E_VSAMPLES); vb = -1; }
    }
    // ---- pass B: density of the valid samples (4 lanes per sample), transmittance, weights ----
    const int sub = lane & 3, sj = lane >> 2;
    const float wcut = a.skip_eps > 0.f ? a.skip_eps / (float)max(nv, 1) : 0.f;
    float T = 1.f, acc = 0.f, depth = 0.f, best_w = 0.f;
    int best_k = 0;
    for (int j0 = 0; j0 < nv; j0 += 8) {
      const int j = j0 + sj;
      const bool active = j < nv;
      const int k = active ? (int)list[j] : 0;
      const float z = MARCH_Z(k);
      float p[3], xn[3];
      nmf_step_pos(o, d, z, p);
      nmf_normalize_xyz(s, p, xn);
      const NmfTaps t = nmf_vm_taps(s, xn);
      float f = nmf_density_group(s, t, sub);
      f += __shfl_xor_sync(FULL, f, 1);
      f += __shfl_xor_sync(FULL, f, 2);
      const float sigma = active ? nmf_feature2density(f, s.density_shift) : 0.f;
      const float z1 = (k + 1 < S) ? MARCH_Z(k + 1) : z;                            // alphagrid.py:348-350
      const float dist = NMF_SUB(z1, z) * s.distance_scale;
      const float alpha = 1.0f - expf(-sigma * dist);                                // tensor_nerf.py:25
      float incl = (1.0f - alpha) + 1e-10f;                                          // :28
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        const float v = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl *= v;
      }
      float excl = __shfl_up_sync(FULL, incl, 4);
      if (lane < 4) excl = 1.f;
      const float Tex = T * excl;
      const float w = alpha * Tex;
      T *= __shfl_sync(FULL, incl, 31);
      const bool mine = active && sub == 0;
      if (TRAIN && a.vs && mine && vb >= 0) {
        VSmp v; v.k = (uint32_t)k; v.f = f; v.alpha = alpha; v.T = Tex;
        a.vs[vb + j] = v;
        a.vdw[vb + j] = 0.f;
      }
      if (mine) {
        acc += w;
        depth += w * z;
        if (w > best_w) { best_w = w; best_k = k; }
      }
      const bool keep = mine && w > 0.f && w >= wcut;
      const unsigned km = __ballot_sync(FULL, keep);
      if (km) {
        int base = 0;
        if (lane == 0) base = atomicAdd(a.n_surv, __popc(km));
        base = __shfl_sync(FULL, base, 0);
        if (keep) {
          const int idx = base + __popc(km & lt);
          if (idx < a.cap_surv) {
            Surv sv; sv.ray = (uint32_t)ray; sv.step = (uint32_t)k; sv.w = w;
            a.surv[idx] = sv;
            if (TRAIN && a.vs) a.survv[idx] = vb >= 0 ? (uint32_t)(vb + j) : 0u;
          } else {
            atomicOr(a.error, NMF_DEV_E_SURVIVORS);
          }
        }
      }
      if (a.t_cut > 0.f && T < a.t_cut) break;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      acc += __shfl_xor_sync(FULL, acc, off);
      depth += __shfl_xor_sync(FULL, depth, off);
      const float ow = __shfl_xor_sync(FULL, best_w, off);
      const int ok_ = __shfl_xor_sync(FULL, best_k, off);
      if (ow > best_w || (ow == best_w && ok_ < best_k)) { best_w = ow; best_k = ok_; }
    }
    if (LEVEL == 1) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) usum += __shfl_xor_sync(FULL, usum, off);
    }
    if (lane == 0) {
      // tensor_nerf.py:505-511: termination point = sample of maximal weight (first index on ties)
      int tk = best_k;
      if (!(best_w > 0.f)) tk = (nv > 0 && list[0] == 0) ? 0 : -1;
      a.tmin[ray] = tmin;
      a.acc[ray] = acc;
      a.nvalid[ray] = nv;
      if (LEVEL == 0) { a.depth[ray] = depth; a.termk[ray] = tk; }
      if (nv) atomicAdd(a.n_samples + chunk, nv);
      if (cand) atomicAdd(a.n_cand + chunk, cand);
      if (LEVEL == 1) atomicAdd(a.wsum + chunk, (double)acc + 1e-3 * (double)usum);
    }
    __syncwarp();
  }
#undef MARCH_Z
}

// ================================================================================================
// k_shade: fields/tensoRF.py:402-405, tensor_base.py:107-129, render_modules.py:553-560,
//          models/microfacet.py:295-349, pt_selectors.py:5-60
// ================================================================================================
struct ShadeArgs {
  const float* rays; const float* tmin;
  const uint64_t* keys;     // level 1
  uint64_t seed, ray_id0;
  int group;                // rays per chunk at this level
  const Surv* surv; const int* n_surv; int cap_surv;
  float* accum;             // level 0: [n][A_N]
  BSample* bs; int* n_bs; int cap_bs;
  int* ray_count; int cap_rays; uint32_t* owner;
  const int* n_samples; const double* wsum;   // level 1 budget
  unsigned* error;
  float4* red;              // level 0: per bounce sample {w, count, ray, flags | rgb sum}
  const float* zvals; int n_steps;   // train mode: jittered distances (n, n_steps)
  float min_rough;          // train mode: Microfacet.min_rough (models/microfacet.py:361-363)
  const uint32_t* survv; int* survslot;   // train mode: survivor -> valid-sample index (in), survivor -> bounce-sample slot (out)
};

// Two phases per warp, 32 surviving samples at a time:
//   gather  8 lanes per sample, 4 samples per round, 8 rounds: the 72 appearance coefficients (6 lanes x 16-byte taps)
//           and the smoothed-gradient normal (three 128-byte pieces per plane row over the 8 lanes).  Every two rounds
//           the basis contraction feat[8 x 24] = coef[8 x 72] * basis_t[72 x 24] (tensoRF.py:405) runs on the tensor
//           cores as 3xTF32 (mma.sync m16n8k8: hi*hi + lo*hi + hi*lo, fp32 accumulate => fp32-level accuracy);
//           features and normal are parked in shared memory
//   shade   1 lane per sample: material heads, SH irradiance, Fresnel, keyed bounce count, debug-map atomics, the
//           bounce-sample record -- nothing here is computed redundantly by the lanes of a group
// The kernel is bound by L1 data-pipe wavefronts (profiles/), which is what this structure minimises: a scalar GEMV
// re-reads both operands from shared memory for every 3 FMAs, the MMA fragments are read once per 8 samples.
#ifndef NMF_SHADE_UNROLL
#define NMF_SHADE_UNROLL 1     // experiments: NMF_NVCC_EXTRA=-DNMF_SHADE_UNROLL=2 (build.py)
#endif
constexpr int kShadeUnroll = NMF_SHADE_UNROLL;
#define SHADE_COEF_LD 76       // floats per staged coefficient row: A-fragment reads (row g, col t) are conflict-free
#define SHADE_FEAT_LD 27       // 24 features + normal per parked sample: lane-per-sample reads are conflict-free
#define SHADE_SMEM_FLOATS (2 * 72 * 24 + 11 * 24 + 16 + 32 + 8 * 8 * SHADE_COEF_LD + 8 * 32 * SHADE_FEAT_LD)
__device__ __forceinline__ uint32_t shade_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void shade_mma(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  // rows 8..15 of the A tile are not used (8 samples per batch): a1 = a3 = 0
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
template <int LEVEL, int TRAIN>
__global__ void __launch_bounds__(256, 3) k_shade(const NmfScene s, const ShadeArgs a) {
  extern __shared__ __align__(16) float shade_sm[];
  uint32_t* s_bhi = (uint32_t*)shade_sm;                 // [72][24] TF32 high part of basis_t
  uint32_t* s_blo = s_bhi + 72 * 24;                     // [72][24] TF32 low part
  float* s_headw = (float*)(s_blo + 72 * 24);            // [11][24]
  float* s_headb = s_headw + 11 * 24;                    // [11] (+5 pad)
  float* s_sh = s_headb + 16;                            // [27] (+5 pad)
  float* s_coef_all = s_sh + 32;                         // [8 warps][8 samples][SHADE_COEF_LD]
  float* s_feat_all = s_coef_all + 8 * 8 * SHADE_COEF_LD;   // [8 warps][32 samples][SHADE_FEAT_LD]
  for (int i = threadIdx.x; i < 72 * 24; i += 256) {
    const float b = s.basis_t[i];
    const uint32_t hi = shade_tf32(b);
    s_bhi[i] = hi;
    s_blo[i] = shade_tf32(b - __uint_as_float(hi));
  }
  for (int i = threadIdx.x; i < 11 * 24; i += 256) s_headw[i] = s.head_w[i];
  if (threadIdx.x < 11) s_headb[threadIdx.x] = s.head_b[threadIdx.x];
  if (threadIdx.x < 27) s_sh[threadIdx.x] = s.sh_conv[threadIdx.x];
  __syncthreads();
  const int n = min(*a.n_surv, a.cap_surv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, l = lane & 7;
  float* s_coef = s_coef_all + warp * 8 * SHADE_COEF_LD;
  float* s_feat = s_feat_all + warp * 32 * SHADE_FEAT_LD;
  for (int wbase = (blockIdx.x * 8 + warp) * 32; wbase < n; wbase += gridDim.x * 256) {
    // ---------------- gather phase ----------------
    // Sample positions first, one lane per sample: the dependent chain survivor record -> ray -> position is paid
    // once per 32 samples (coalesced) instead of once per round; the normalised position waits in the sample's
    // feature row (the row is only overwritten by the contraction that follows the sample's own rounds).
    {
      const int si = wbase + lane;
      Surv sv; sv.ray = 0; sv.step = 0; sv.w = 0.f;
      if (si < n) sv = a.surv[si];
      const int ray = (int)sv.ray;
      float o[3], d[3], p[3], xn[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) { o[i] = __ldg(a.rays + (size_t)ray * 6 + i); d[i] = __ldg(a.rays + (size_t)ray * 6 + 3 + i); }
      nmf_step_pos(o, d, TRAIN ? a.zvals[(size_t)ray * a.n_steps + sv.step] : nmf_step_z(a.tmin[ray], s.stepsize, (int)sv.step), p);
      nmf_normalize_xyz(s, p, xn);
      float* row = s_feat + lane * SHADE_FEAT_LD;
      row[0] = xn[0]; row[1] = xn[1]; row[2] = xn[2];
    }
    __syncwarp();
#pragma unroll kShadeUnroll
    for (int round = 0; round < 8; ++round) {
      float xn[3];
      {
        const float* row = s_feat + (round * 4 + grp) * SHADE_FEAT_LD;
        xn[0] = row[0]; xn[1] = row[1]; xn[2] = row[2];
      }
      const NmfTaps t = nmf_vm_taps(s, xn);
      // appearance coefficients: lanes 0..5 own 4 channels of each plane
      const int row = (round & 1) * 4 + grp;
      if (l < 6) {
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
          const nmf_f4 c = nmf_app_group(s, t, pl, l);
          *(float4*)&s_coef[row * SHADE_COEF_LD + pl * 24 + 4 * l] = c;
        }
      }
      // smoothed-gradient normal, this lane's share
      float grad[3] = {0.f, 0.f, 0.f};
      nmf_normal_lane(s, t, l, grad);
#pragma unroll
      for (int off = 1; off < 8; off <<= 1) {
        grad[0] += __shfl_xor_sync(FULL, grad[0], off);
        grad[1] += __shfl_xor_sync(FULL, grad[1], off);
        grad[2] += __shfl_xor_sync(FULL, grad[2], off);
      }
      if (l == 0) {
        const nmf_v3 nrm = nmf_normal_from_grad(s, grad);
        float* no = s_feat + (round * 4 + grp) * SHADE_FEAT_LD + 24;
        no[0] = nrm.x; no[1] = nrm.y; no[2] = nrm.z;
      }
      __syncwarp();
      if (round & 1) {
        // basis contraction of the 8 staged samples: lane (g, tq) holds A(row g, cols tq, tq+4), B(rows tq, tq+4, col g)
        const int g = lane >> 2, tq = lane & 3;
        float c[3][4];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
#pragma unroll 3
        for (int ks = 0; ks < 9; ++ks) {
          const float a0 = s_coef[g * SHADE_COEF_LD + 8 * ks + tq], a2 = s_coef[g * SHADE_COEF_LD + 8 * ks + tq + 4];
          const uint32_t a0h = shade_tf32(a0), a2h = shade_tf32(a2);
          const uint32_t a0l = shade_tf32(a0 - __uint_as_float(a0h)), a2l = shade_tf32(a2 - __uint_as_float(a2h));
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) {
            const int b0i = (8 * ks + tq) * 24 + 8 * nt + g, b1i = b0i + 4 * 24;
            const uint32_t bh0 = s_bhi[b0i], bh1 = s_bhi[b1i], bl0 = s_blo[b0i], bl1 = s_blo[b1i];
            shade_mma(c[nt], a0l, a2l, bh0, bh1);
            shade_mma(c[nt], a0h, a2h, bl0, bl1);
            shade_mma(c[nt], a0h, a2h, bh0, bh1);
          }
        }
        float* fo = s_feat + ((round >> 1) * 8 + g) * SHADE_FEAT_LD + 2 * tq;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) { fo[8 * nt] = c[nt][0]; fo[8 * nt + 1] = c[nt][1]; }
        __syncwarp();
      }
    }
    // ---------------- shade phase: one lane per sample ----------------
    // Written compactly (rolled channel / noise loops, results stored where they are consumed) because this kernel
    // is instruction-cache bound: the code of both phases has to stream through a 32 KB cache (profiles/).
    const int si = wbase + lane;
    const bool active = si < n;
    Surv sv; sv.ray = 0; sv.step = 0; sv.w = 0.f;
    if (active) sv = a.surv[si];
    const int ray = (int)sv.ray, k = (int)sv.step;
    const float w = sv.w;
    float o[3], d[3], p[3], xn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { o[i] = __ldg(a.rays + (size_t)ray * 6 + i); d[i] = __ldg(a.rays + (size_t)ray * 6 + 3 + i); }
    nmf_step_pos(o, d, TRAIN ? a.zvals[(size_t)ray * a.n_steps + k] : nmf_step_z(a.tmin[ray], s.stepsize, k), p);
    nmf_normalize_xyz(s, p, xn);
    float* fs = s_feat + lane * SHADE_FEAT_LD;
    float* hs = s_coef + lane * 9;         // per-channel head results, parked in the (now idle) coefficient staging rows
    const nmf_v3 nrm = nmf_mk3(fs[24], fs[25], fs[26]);
    const nmf_v3 V = nmf_mk3(-d[0], -d[1], -d[2]);
    const float vn = nmf_dot(V, nrm);
    // bounce count (pt_selectors.py:20-40)
    const int chunk = ray / a.group;
    const uint64_t rkey = LEVEL == 0 ? nmf_mix64(a.seed, a.ray_id0 + (uint64_t)ray) : (active ? a.keys[ray] : 0ull);
    const uint64_t skey = nmf_mix64(rkey, (uint64_t)k);
    const float U = nmf_uniform(skey, NMF_STREAM_BOUNCE);
    float kf;
    if (LEVEL == 0) {
      kf = floorf(w * (float)s.rays_per_ray + U - 0.5f);
    } else {
      const int N = s.max_brdf_rays1 - a.n_samples[chunk];
      const float wsum = fmaxf((float)a.wsum[chunk], 1e-3f);
      const float wj = w + 1e-3f * U;
      kf = N > 0 ? floorf(wj / wsum * (float)N + 1.0f) : floorf(wj / wsum * (float)s.max_brdf_rays1 + 0.5f);
    }
    const int count = active ? (int)nmf_clampf(kf, 0.f, (float)NMF_MAX_BOUNCE) : 0;
    // allocate the bounce-sample slot (one atomic per warp) and the sample's range in its chunk's bounce-ray region
    // (one atomic per warp when all bouncing samples of the warp belong to one chunk, the common case)
    const unsigned bm = __ballot_sync(FULL, count > 0);
    int slot = -1, roff = 0;
    if (bm) {
      const int leader = __ffs(bm) - 1;
      const unsigned lt = (1u << lane) - 1u;
      int sbase = 0;
      if (lane == leader) sbase = atomicAdd(a.n_bs, __popc(bm));
      sbase = __shfl_sync(FULL, sbase, leader);
      const int chunk0 = __shfl_sync(FULL, chunk, leader);
      const bool same = __all_sync(FULL, count == 0 || chunk == chunk0);
      if (same) {
        int incl = count;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int v = __shfl_up_sync(FULL, incl, off);
          if (lane >= off) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        int rbase = 0;
        if (lane == leader) rbase = atomicAdd(a.ray_count + chunk0, total);
        rbase = __shfl_sync(FULL, rbase, leader);
        roff = rbase + incl - count;
      } else if (count > 0) {
        roff = atomicAdd(a.ray_count + chunk, count);
      }
      if (count > 0) {
        slot = sbase + __popc(bm & lt);
        if (slot >= a.cap_bs) {
          atomicOr(a.error, NMF_DEV_E_BSAMPLES);
          slot = -1;
        } else if (roff + count > a.cap_rays) {
          // the slot index is taken but its rays do not fit: leave an empty reduction record behind (every index
          // below n_bs must be readable by k_reduce0) and drop the sample; the call reports the overflow
          atomicOr(a.error, NMF_DEV_E_BRAYS);
          if (LEVEL == 0) { a.red[2 * slot] = make_float4(0.f, 0.f, 0.f, 0.f); a.red[2 * slot + 1] = make_float4(0.f, 0.f, 0.f, 0.f); }
          slot = -1;
        }
      }
    }
    BSample* b = a.bs + (slot >= 0 ? slot : 0);
    // survivor -> bounce-sample slot for the reverse pass; -2 = no bounce sample but a back-facing normal (the orientation
    // loss still has a gradient there), -1 = nothing to do
    if (TRAIN && a.survslot && active) a.survslot[si] = slot >= 0 ? slot : (vn < 0.f ? -2 : -1);
    float* acc = LEVEL == 0 ? a.accum + (size_t)ray * A_N : nullptr;
    // roughness head (render_modules.py:553-560; r2 = r1, microfacet.py:360)
    float lin = s_headb[9];
#pragma unroll
    for (int i = 0; i < 24; ++i) lin += s_headw[9 * 24 + i] * fs[i];
    const float rough = nmf_clampf(nmf_sigmoid(lin + s.roughness_bias) / 2.0f, 1e-2f, 1.0f);
    float sh[9];
    nmf_sh9(nrm, sh);
    const float cost = fabsf(vn);
    // per colour channel: albedo head (row c), f0 head (row 6 + c), SH irradiance, Fresnel; the tint head (rows 3..5)
    // feeds nothing on this path (fresnel mode)
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      float la = s_headb[c], lf = s_headb[6 + c];
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        const float fv = fs[i];
        la += s_headw[c * 24 + i] * fv;
        lf += s_headw[(6 + c) * 24 + i] * fv;
      }
      const float albedo = nmf_clampf(nmf_sigmoid(s.diffuse_mul * la + s.diffuse_bias), 0.f, 1.f);
      const float f0v = nmf_sigmoid(lf + s.f0_bias);
      float e = 0.f;
#pragma unroll
      for (int i = 0; i < 9; ++i) e += s_sh[i * 3 + c] * sh[i];
      const float diffuse = albedo * e;                                  // microfacet.py:316
      const float fresn = nmf_fresnel(f0v, cost);
      if (LEVEL == 0 && active) {
        // debug / auxiliary maps (tensor_nerf.py:495-566, microfacet.py:639-647)
        atomicAdd(acc + A_WN + c, w * (c == 0 ? nrm.x : c == 1 ? nrm.y : nrm.z));
        atomicAdd(acc + A_DIFF + c, w * (1.0f - fresn) * diffuse);
        atomicAdd(acc + A_ALB + c, w * albedo);
      }
      hs[c] = f0v; hs[3 + c] = diffuse; hs[6 + c] = fresn;             // parked: the record is written with 16-byte stores
    }
    if (LEVEL == 0 && active) {
      atomicAdd(acc + A_ROUGH, w * rough);
      if (vn < 0.f) atomicAdd(acc + A_ORI, w * (vn * vn));               // tensor_nerf.py:573-583 ori_loss
    }
    if (slot >= 0) {
      const float sgn = vn > 0.f ? 1.f : (vn < 0.f ? -1.f : 0.f);        // microfacet.py:354-356
      const nmf_v3 Nf = nmf_mk3(nrm.x * sgn, nrm.y * sgn, nrm.z * sgn);
      float* q = (float*)b;                                                // ten 256-bit stores (see BSample)
      const float rough_b = TRAIN ? fmaxf(rough, a.min_rough) : rough;     // microfacet.py:361-363 (bounce rays only)
      nmf_st8(q, make_float4(p[0], p[1], p[2], w), make_float4(V.x, V.y, V.z, rough_b));
      nmf_st8(q + 8, make_float4(Nf.x, Nf.y, Nf.z, __int_as_float(count)), make_float4(hs[0], hs[1], hs[2], __uint_as_float((uint32_t)ray)));
      nmf_st8(q + 16, make_float4(hs[3], hs[4], hs[5], __uint_as_float((uint32_t)roff)),
              make_float4(hs[6], hs[7], hs[8], __uint_as_float(xn[2] < 0.f ? 1u : 0u)));
      if (LEVEL == 0)
        nmf_st8((float*)(a.red + 2 * slot),
                make_float4(w, __int_as_float(count), __uint_as_float((uint32_t)ray), __uint_as_float(xn[2] < 0.f ? 1u : 0u)),
                make_float4(0.f, 0.f, 0.f, 0.f));
      // per-sample part of the GGX sampler and of the ISH encodings, shared by all bounce rays of the sample
      const NmfGGXFrame fr = nmf_ggx_frame(V, Nf, rough_b);
      float s1, s2;
      nmf_ish_scales(rough_b, &s1, &s2);
      // appearance feature + noise (microfacet.py:297, keyed Box-Muller), one feature per trip of a rolled loop
      const uint64_t nseed = nmf_noise_seed(skey);
#pragma unroll 1
      for (int i = 0; i < 12; ++i) {
        float n0, n1;
        nmf_noise_pair(nseed, (uint32_t)i, &n0, &n1);
        fs[2 * i] += s.anoise * n0;
        fs[2 * i + 1] += s.anoise * n1;
      }
      const uint32_t vsi = (TRAIN && a.survv) ? a.survv[si] : 0u;
      nmf_st8(q + 24, make_float4(__uint_as_float((uint32_t)(skey & 0xffffffffull)), __uint_as_float((uint32_t)(skey >> 32)),
                                  __uint_as_float((uint32_t)chunk), __uint_as_float(vsi)),
              make_float4(fs[0], fs[1], fs[2], fs[3]));
      nmf_st8(q + 32, make_float4(fs[4], fs[5], fs[6], fs[7]), make_float4(fs[8], fs[9], fs[10], fs[11]));
      nmf_st8(q + 40, make_float4(fs[12], fs[13], fs[14], fs[15]), make_float4(fs[16], fs[17], fs[18], fs[19]));
      nmf_st8(q + 48, make_float4(fs[20], fs[21], fs[22], fs[23]), make_float4(fr.t.x, fr.t.y, fr.t.z, fr.b.x));
      nmf_st8(q + 56, make_float4(fr.b.y, fr.b.z, fr.V_l.x, fr.V_l.y), make_float4(fr.V_l.z, fr.Vs.x, fr.Vs.y, fr.Vs.z));
      nmf_st8(q + 64, make_float4(fr.T1.x, fr.T1.y, fr.T1.z, fr.T2.x), make_float4(fr.T2.y, fr.T2.z, fr.a, s1));
      nmf_st8(q + 72, make_float4(s2, 0.25f * nmf_uniform(skey, NMF_STREAM_OFF_U), 0.25f * nmf_uniform(skey, NMF_STREAM_OFF_V), 0.f),
              make_float4(0.f, 0.f, 0.f, 0.f));
    }
    // ray -> bounce-sample map of the allocated ranges, written by the whole warp.  A sample whose allocation failed
    // (a list overflowed: the call reports an error) marks its rays with NMF_NO_OWNER so that no consumer follows a
    // stale index.
    unsigned todo = __ballot_sync(FULL, count > 0);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int c_cnt = __shfl_sync(FULL, count, src), c_slot = __shfl_sync(FULL, slot, src);
      const int c_roff = __shfl_sync(FULL, roff, src), c_chunk = __shfl_sync(FULL, chunk, src);
      uint32_t* ow = a.owner + (size_t)c_chunk * a.cap_rays + c_roff;
      const int lim = min(c_cnt, a.cap_rays - c_roff);
      for (int j = lane; j < lim; j += 32) ow[j] = c_slot >= 0 ? (uint32_t)c_slot : NMF_NO_OWNER;
    }
    __syncwarp();
  }
}

// ================================================================================================
// BRDF MLP 66 -> 64 -> 64 -> 4 (modules/brdf.py:73-120,237-239), one row per thread; weights in shared
// memory (transposed, so that the 32 lanes broadcast-read 16 bytes of one weight row at a time)
// ================================================================================================
#define MLP_SMEM_FLOATS (66 * 64 + 64 + 64 * 64 + 64 + 64 * 4 + 4 + 66 * MLP_THREADS)

__device__ __forceinline__ void mlp_load_weights(const NmfScene& s, float* sm) {
  float* w0 = sm; float* b0 = w0 + 66 * 64; float* w1 = b0 + 64; float* b1 = w1 + 64 * 64; float* w2 = b1 + 64; float* b2 = w2 + 256;
  for (int i = threadIdx.x; i < 66 * 64; i += blockDim.x) w0[i] = s.brdf_w0t[i];
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) w1[i] = s.brdf_w1t[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) w2[i] = s.brdf_w2t[i];
  if (threadIdx.x < 64) { b0[threadIdx.x] = s.brdf_b0[threadIdx.x]; b1[threadIdx.x] = s.brdf_b1[threadIdx.x]; }
  if (threadIdx.x < 4) b2[threadIdx.x] = s.brdf_b2[threadIdx.x];
}
// x: this thread's column of the [66][MLP_THREADS] staging buffer (x[k * MLP_THREADS]); overwritten
__device__ __forceinline__ void mlp_forward(const float* sm, float* x, float brdf_bias, float* out3) {
  const float* w0 = sm; const float* b0 = w0 + 66 * 64; const float* w1 = b0 + 64; const float* b1 = w1 + 64 * 64;
  const float* w2 = b1 + 64; const float* b2 = w2 + 256;
  float h[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) h[i] = b0[i];
#pragma unroll 2
  for (int k = 0; k < 66; ++k) {
    const float xv = x[k * MLP_THREADS];
    const float4* wr = (const float4*)(w0 + k * 64);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float4 wv = wr[q];
      h[4 * q] += xv * wv.x; h[4 * q + 1] += xv * wv.y; h[4 * q + 2] += xv * wv.z; h[4 * q + 3] += xv * wv.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 64; ++i) { x[i * MLP_THREADS] = fmaxf(h[i], 0.f); h[i] = b1[i]; }
#pragma unroll 2
  for (int k = 0; k < 64; ++k) {
    const float xv = x[k * MLP_THREADS];
    const float4* wr = (const float4*)(w1 + k * 64);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float4 wv = wr[q];
      h[4 * q] += xv * wv.x; h[4 * q + 1] += xv * wv.y; h[4 * q + 2] += xv * wv.z; h[4 * q + 3] += xv * wv.w;
    }
  }
  float o0 = b2[0], o1 = b2[1], o2 = b2[2];
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    const float hv = fmaxf(h[k], 0.f);
    const float4 wv = *(const float4*)(w2 + 4 * k);
    o0 += hv * wv.x; o1 += hv * wv.y; o2 += hv * wv.z;
  }
  out3[0] = nmf_sigmoid(o0 + brdf_bias);
  out3[1] = nmf_sigmoid(o1 + brdf_bias);
  out3[2] = nmf_sigmoid(o2 + brdf_bias);
}
// modules/brdf.py:216-225: x = [feat | ISH(half) | half | ISH(diff) | diff | 0-pad]; feat is already in x[0..23]
__device__ __forceinline__ void mlp_encode_s(float (&x)[TC_K0], nmf_v3 half_l, nmf_v3 diff_l, float s1, float s2) {
  nmf_ish18_s(half_l, s1, s2, &x[24]);
  x[42] = half_l.x; x[43] = half_l.y; x[44] = half_l.z;
  nmf_ish18_s(diff_l, s1, s2, &x[45]);
  x[63] = diff_l.x; x[64] = diff_l.y; x[65] = diff_l.z;
  x[TC_ONE] = 1.f;                      // carries the biases through the tensor-core GEMMs (nmf_mlp_tc.cuh)
#pragma unroll
  for (int i = TC_ONE + 1; i < TC_K0; ++i) x[i] = 0.f;
}
__device__ __forceinline__ void mlp_encode(float (&x)[TC_K0], nmf_v3 half_l, nmf_v3 diff_l, float rough) {
  float s1, s2;
  nmf_ish_scales(rough, &s1, &s2);
  mlp_encode_s(x, half_l, diff_l, s1, s2);
}
// fp32 SIMT variant (scene.mlp_mode == 1): stage the row in this thread's shared-memory column, then mlp_forward
__device__ __forceinline__ void mlp_simt(const float* sm, float* xcol, const float (&x)[TC_K0], float brdf_bias, float* out3) {
#pragma unroll
  for (int i = 0; i < 66; ++i) xcol[i * MLP_THREADS] = x[i];
  mlp_forward(sm, xcol, brdf_bias, out3);
}

// ================================================================================================
// k_bounce: brdf_samplers/base.py:11-20, ggx.py:61-268, models/microfacet.py:367-472 (+ :561-613 at level 1)
// ================================================================================================
struct BounceArgs {
  const BSample* bs; BRay* brays; const uint32_t* owner; const int* ray_count; int cap_rays;
  unsigned long long* score_sum; float2* scu;     // level 0: retrace scores
  const int* tile_start;    // [n_chunks + 1] exclusive prefix of the chunks' 128-ray tile counts (k_tile_prefix)
  int n_chunks;
  const int2* tile_desc;    // [n_tiles] (chunk, first ray) per tile
};

// tile_start[c] = number of 128-ray tiles in the bounce-ray regions of chunks < c, so that a persistent grid can
// walk one flat, evenly sized work list instead of a (tiles x chunks) grid with ragged rows; tile_desc[t] = (chunk, first
// ray of the tile within the chunk's region): the consumers read it two tiles ahead instead of searching tile_start (a
// chain of dependent loads that nothing can overlap: capture K, profiles/r02_b_ncu_summary.md)
__global__ void __launch_bounds__(1024) k_tile_prefix(const int* ray_count, int cap_rays, int n_chunks, int* tile_start, int2* tile_desc) {
  __shared__ int s_part[1024];
  const int per = (n_chunks + 1023) / 1024;
  const int c0 = threadIdx.x * per;
  int sum = 0;
  for (int c = c0; c < min(c0 + per, n_chunks); ++c) sum += (min(ray_count[c], cap_rays) + MLP_THREADS - 1) / MLP_THREADS;
  s_part[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = s_part[threadIdx.x] - sum;
  for (int c = c0; c < min(c0 + per, n_chunks); ++c) {
    tile_start[c] = run;
    run += (min(ray_count[c], cap_rays) + MLP_THREADS - 1) / MLP_THREADS;
  }
  if (threadIdx.x == 1023) tile_start[n_chunks] = s_part[1023];
  __syncthreads();                 // tile_start (global, written by this block) is visible to the block
  // descriptors: a warp per chunk, lanes across its tiles
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < n_chunks; c += 32) {
    const int t0 = tile_start[c], nt = (min(ray_count[c], cap_rays) + MLP_THREADS - 1) / MLP_THREADS;
    for (int k = lane; k < nt; k += 32) tile_desc[t0 + k] = make_int2(c, k * MLP_THREADS);
  }
}

// The persistent bounce-ray kernels walk tiles blockIdx.x, + gridDim.x, ...  TileWalk keeps the chain descriptor -> owner
// off the critical path: the descriptor is loaded two tiles ahead, the owner (whose address needs the descriptor) one tile
// ahead, so that neither load is consumed in the iteration that issues it.
struct TileWalk {
  int2 d_next;            // descriptor of tile + gridDim.x
  int chunk, n, r;        // current tile
  uint32_t slot;          // owner of this thread's ray in the current tile
  int chunk_n, n_n, r_n;  // next tile
  uint32_t slot_n;
};
__device__ __forceinline__ void tw_fetch(const int2 d, const int* ray_count, int cap_rays, const uint32_t* owner, int& chunk, int& n, int& r,
                                         uint32_t& slot) {
  chunk = d.x;
  n = min(__ldg(ray_count + chunk), cap_rays);
  r = d.y + threadIdx.x;
  slot = r < n ? __ldg(owner + (size_t)chunk * cap_rays + r) : NMF_NO_OWNER;
}
__device__ __forceinline__ void tw_begin(TileWalk& t, const int2* desc, int n_tiles, const int* ray_count, int cap_rays, const uint32_t* owner) {
  const int tile = blockIdx.x, g = gridDim.x;
  t.chunk = t.n = t.r = 0; t.slot = NMF_NO_OWNER;
  t.d_next = make_int2(0, 0);
  if (tile < n_tiles) tw_fetch(__ldg(desc + tile), ray_count, cap_rays, owner, t.chunk, t.n, t.r, t.slot);
  if (tile + g < n_tiles) t.d_next = __ldg(desc + tile + g);
}
// at the top of iteration `tile`: issue the next tile's owner load and the descriptor load of the tile after it
__device__ __forceinline__ void tw_issue(TileWalk& t, int tile, const int2* desc, int n_tiles, const int* ray_count, int cap_rays,
                                         const uint32_t* owner) {
  const int g = gridDim.x;
  t.chunk_n = t.n_n = t.r_n = 0; t.slot_n = NMF_NO_OWNER;
  if (tile + g < n_tiles) tw_fetch(t.d_next, ray_count, cap_rays, owner, t.chunk_n, t.n_n, t.r_n, t.slot_n);
  t.d_next = tile + 2 * g < n_tiles ? __ldg(desc + tile + 2 * g) : make_int2(0, 0);
}
__device__ __forceinline__ void tw_advance(TileWalk& t) { t.chunk = t.chunk_n; t.n = t.n_n; t.r = t.r_n; t.slot = t.slot_n; }

template <int LEVEL, int TC>
__global__ void __launch_bounds__(MLP_THREADS, TC ? 5 : 1) k_bounce(const NmfScene s, const BounceArgs a) {
  extern __shared__ __align__(128) float sm[];
  TcMlp tc;
  float* xcol = nullptr;
  if (TC) {
    tc_mlp_init(tc, sm, s.brdf_w0u, s.brdf_w1u, s.brdf_w2u);
  } else {
    mlp_load_weights(s, sm);
    __syncthreads();
    xcol = sm + (MLP_SMEM_FLOATS - 66 * MLP_THREADS) + threadIdx.x;
  }
  const int n_tiles = a.tile_start[a.n_chunks];
  TileWalk tw;
  tw_begin(tw, a.tile_desc, n_tiles, a.ray_count, a.cap_rays, a.owner);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    tw_issue(tw, tile, a.tile_desc, n_tiles, a.ray_count, a.cap_rays, a.owner);
    const int chunk = tw.chunk, n = tw.n, r = tw.r;
    uint32_t slot = tw.slot;
    BRay* region = a.brays + (size_t)chunk * a.cap_rays;
    const bool active = slot != NMF_NO_OWNER;
    if (!active) slot = 0u;
    if (LEVEL == 0 && r < n && !active) a.scu[(size_t)chunk * a.cap_rays + r] = make_float2(0.f, 0.f);
    const BSample* b = a.bs + slot;
    const int j = active ? r - (int)b->roff : 0;
    // the record in nine 256-bit loads (BSample: ten 32-byte pairs; diffuse|fresn is not needed here)
    const float* rb = (const float*)b;
    float4 q0, q1, q2, qx, kb, ft[6], f0, f1, f2, f3, f4, f5;
    nmf_ld8(rb, q0, q1);                 // pos, w | V, rough
    nmf_ld8(rb + 8, q2, qx);             // N, count | (f0)
    nmf_ld8(rb + 24, kb, ft[0]);         // key, chunk, pad | feat 0..3
    nmf_ld8(rb + 32, ft[1], ft[2]);
    nmf_ld8(rb + 40, ft[3], ft[4]);
    nmf_ld8(rb + 48, ft[5], f0);         // feat 20..23 | frame 0..3
    nmf_ld8(rb + 56, f1, f2);
    nmf_ld8(rb + 64, f3, f4);
    nmf_ld8(rb + 72, f5, qx);
    const nmf_v3 V = nmf_mk3(q1.x, q1.y, q1.z), N = nmf_mk3(q2.x, q2.y, q2.z);
    const float rough = q1.w, w = q0.w;
    const int count = max(__float_as_int(q2.w), 1);
    const uint64_t skey = (uint64_t)__float_as_uint(kb.x) | ((uint64_t)__float_as_uint(kb.y) << 32);
    NmfGGXFrame fr;
    fr.t = nmf_mk3(f0.x, f0.y, f0.z); fr.b = nmf_mk3(f0.w, f1.x, f1.y); fr.V_l = nmf_mk3(f1.z, f1.w, f2.x);
    fr.Vs = nmf_mk3(f2.y, f2.z, f2.w); fr.T1 = nmf_mk3(f3.x, f3.y, f3.z); fr.T2 = nmf_mk3(f3.w, f4.x, f4.y);
    fr.a = f4.z;
    const float ish1 = f4.w, ish2 = f5.x;
    const float u1 = nmf_wrap01(__ldg(s.sobol + 2 * j) + f5.y);
    const float u2 = nmf_wrap01(__ldg(s.sobol + 2 * j + 1) + f5.z);
    const NmfGGX g = nmf_ggx_sample_f(fr, u1, u2, V, N, rough);
    float x[TC_K0];
#pragma unroll
    for (int i = 0; i < 6; ++i) { x[4 * i] = ft[i].x; x[4 * i + 1] = ft[i].y; x[4 * i + 2] = ft[i].z; x[4 * i + 3] = ft[i].w; }
    mlp_encode_s(x, g.half_l, g.diff_l, ish1, ish2);
    float bw[3];
#ifdef NMF_BOUNCE_NO_MLP      // experiment only (NMF_NVCC_EXTRA): the kernel without its MLP, to see what the SIMT part alone costs
    { float acc = 0.f;
#pragma unroll
      for (int i = 0; i < TC_ONE; ++i) acc += x[i];
      bw[0] = bw[1] = bw[2] = 1.0f / (1.0f + expf(-acc)); }
#else
    if (TC) tc_mlp_forward(tc, x, s.brdf_bias, bw);      // all 128 threads: the tile is one tensor-core GEMM
    else if (active) mlp_simt(sm, xcol, x, s.brdf_bias, bw);
#endif
    const float mip = -logf((float)count) - g.logpdf;                    // microfacet.py:445-448
    if (tw.slot_n != NMF_NO_OWNER) {            // the next tile's sample record: its owner arrived under the MLP above
      const char* rec = (const char*)(a.bs + tw.slot_n);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rec));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + 128));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + 256));
    }
    if (LEVEL == 0) {
      float sc = 0.f;
      if (active) {
        // contribution estimate for the retrace selection (microfacet.py:480-504) and its tie-break uniform (:506)
        const float pdf = expf(g.logpdf);
        const float per_ray = fmaxf(bw[0], fmaxf(bw[1], bw[2])) * (nmf_dot(V, N) > 0.f ? 1.f : 0.f) * pdf;
        sc = per_ray * (w / ((float)count + 1e-8f));
        BRay* o = region + r;
        nmf_st8((float*)o, make_float4(g.L.x, g.L.y, g.L.z, mip), make_float4(bw[0], bw[1], bw[2], __int_as_float(-1)));
        const uint64_t rkey = nmf_mix64(skey, (uint64_t)j + NMF_STREAM_RAY0);
        a.scu[(size_t)chunk * a.cap_rays + r] = make_float2(sc, nmf_uniform(rkey, NMF_STREAM_TIE));
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sc += __shfl_xor_sync(FULL, sc, off);
      // integer atomics: the chunk total (and with it the retrace selection) does not depend on arrival order
      if ((threadIdx.x & 31) == 0 && sc > 0.f) atomicAdd(a.score_sum + chunk, (unsigned long long)((double)sc * 4294967296.0));
    } else if (active) {
      // no further retrace at this depth (microfacet.py:561): k_incoming<1> looks every ray up in the environment.
      // (Doing the lookup here saved the 32-byte record but ran 15 % slower: this kernel sits at its register and
      // shared-memory occupancy limit, the lookup kernel does not.)
      BRay* o = region + r;
      nmf_st8((float*)o, make_float4(g.L.x, g.L.y, g.L.z, mip), make_float4(bw[0], bw[1], bw[2], __int_as_float(-1)));
    }
    tw_advance(tw);
  }
  if (TC) tc_mlp_free(tc);
}

// ================================================================================================
// k_select: models/microfacet.py:475-559 -- per chunk, the max_retrace bounce rays of largest
// (normalised contribution + U) become secondary rays.  Three-pass radix select on the float bits.
// ================================================================================================
struct SelectArgs {
  const BSample* bs; BRay* brays; const uint32_t* owner; float2* scu; const int* ray_count; int cap_rays;
  const unsigned long long* score_sum; int max_retrace; int* n_sec;
  float* rays1; float* mip1; uint64_t* key1;
};

// A thread-block CLUSTER per chunk (1, 2, 4 or 8 CTAs, chosen so that few chunks still fill the GPU: a training batch is
// ONE chunk of ~100 k bounce rays): every CTA histograms its share of the rays, the histograms are summed through
// distributed shared memory, every CTA scans the sum redundantly (same threshold everywhere); the tie list, the tie
// counter and the output-slot counter live in CTA 0's shared memory and are reached with DSMEM atomics.
// USE_CL = 0: one CTA per chunk, launched without a cluster (many chunks: a cluster of one only adds barrier cost).
template <int USE_CL>
__global__ void __launch_bounds__(1024) k_select(const SelectArgs a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned CL = USE_CL ? cluster.num_blocks() : 1u, rank = USE_CL ? cluster.block_rank() : 0u;
  auto csync = [&]() { if (USE_CL) cluster.sync(); else __syncthreads(); };
  __shared__ unsigned hist[2048];      // this CTA's histogram (read by the whole cluster)
  __shared__ unsigned hsum[2048];      // the cluster's histogram (private copy)
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned sh_prefix, sh_need, sh_slot, sh_eq;
  __shared__ unsigned long long tie_key[NMF_TIE_CAP];
  __shared__ unsigned long long tie_cut;
  const int chunk = blockIdx.x / CL;
  const int n = min(a.ray_count[chunk], a.cap_rays);
  const int n_re = min(n, a.max_retrace);
  if (threadIdx.x == 0 && rank == 0) a.n_sec[chunk] = n_re;
  if (n_re == 0) return;               // the whole cluster leaves together
  BRay* region = a.brays + (size_t)chunk * a.cap_rays;
  float2* scu = a.scu + (size_t)chunk * a.cap_rays;
  const float total = (float)((double)a.score_sum[chunk] * (1.0 / 4294967296.0));
  const int lane = threadIdx.x & 31;
  const int stride = (int)CL * 1024, first = (int)rank * 1024 + threadIdx.x;
  unsigned* sh_slot0 = USE_CL ? cluster.map_shared_rank(&sh_slot, 0) : &sh_slot;
  unsigned* sh_eq0 = USE_CL ? cluster.map_shared_rank(&sh_eq, 0) : &sh_eq;
  unsigned long long* tie_key0 = USE_CL ? cluster.map_shared_rank(tie_key, 0) : tie_key;
  unsigned long long* tie_cut0 = USE_CL ? cluster.map_shared_rank(&tie_cut, 0) : &tie_cut;
  const unsigned* hs = USE_CL ? hsum : hist;      // the histogram the threshold scan reads
  // pass 0: final score = cc / sum * n_re + U(ray key)   (microfacet.py:504-506); histogram of its top 11 bits.
  // The scores crowd into a handful of exponent bins, so equal bins of a warp are merged into one shared atomic.
  for (int i = threadIdx.x; i < 2048; i += 1024) hist[i] = 0;
  if (threadIdx.x == 0) { sh_slot = 0; sh_eq = 0; tie_cut = ~0ull; }
  __syncthreads();
  for (int r0 = (int)rank * 1024; r0 < n; r0 += stride) {
    const int r = r0 + threadIdx.x;
    unsigned bin = 0xFFFFFFFFu;
    if (r < n) {
      const float2 v = scu[r];
      const float sc = (total > 0.f ? v.x / total * (float)n_re : 0.f) + v.y;
      scu[r].x = sc;
      bin = __float_as_uint(sc) >> 21;
    }
    const unsigned m = __match_any_sync(FULL, bin);
    if (r < n && lane == __ffs(m) - 1) atomicAdd(&hist[bin], (unsigned)__popc(m));
  }
  unsigned prefix = 0, need = (unsigned)n_re;
  // digit 0: bits 31..21, digit 1: bits 20..10, digit 2: bits 9..0
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    const int bins = pass == 2 ? 1024 : 2048;
    if (pass > 0) {
      for (int i = threadIdx.x; i < 2048; i += 1024) hist[i] = 0;
      __syncthreads();
      const unsigned hi_shift = pass == 1 ? 21 : 10;
      for (int r = first; r < n; r += stride) {
        const unsigned bits = __float_as_uint(scu[r].x);
        if ((bits >> hi_shift) == prefix) atomicAdd(&hist[(bits >> shift) & (bins - 1)], 1u);
      }
    }
    csync();                           // every CTA's histogram is complete (also orders the block's own shared writes)
    if (USE_CL) {
      for (int i = threadIdx.x; i < bins; i += 1024) {
        unsigned v = 0;
        for (unsigned c = 0; c < CL; ++c) v += cluster.map_shared_rank(hist, c)[i];
        hsum[i] = v;
      }
      cluster.sync();                  // all remote reads of `hist` are done before anyone clears it again
    }
    // largest bin b whose suffix count reaches `need` (b = 0 if none does): block-wide scan over the bins in descending
    // order, two bins per thread (a serial walk by one thread cost ~30 us per pass: 2048 dependent shared loads)
    {
      const int i0 = 2 * threadIdx.x;                                // descending position: bin = bins - 1 - i
      const unsigned h0 = i0 < bins ? hs[bins - 1 - i0] : 0u, h1 = i0 + 1 < bins ? hs[bins - 2 - i0] : 0u;
      unsigned incl = h0 + h1;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned v = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl += v;
      }
      if (lane == 31) warp_tot[threadIdx.x >> 5] = incl;
      __syncthreads();
      if (threadIdx.x < 32) {
        unsigned t = warp_tot[threadIdx.x];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const unsigned v = __shfl_up_sync(FULL, t, off);
          if (lane >= off) t += v;
        }
        warp_tot[threadIdx.x] = t;                                    // inclusive totals of the warps
        if (threadIdx.x == 31) {                                      // default: nothing reaches `need` -> bin 0
          sh_prefix = prefix << (pass == 0 ? 0 : (pass == 1 ? 11 : 10));
          sh_need = need - (t - hs[0]);
        }
      }
      __syncthreads();
      const unsigned excl = incl - (h0 + h1) + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0u);
      int hit = -1;
      unsigned cum = 0;
      if (i0 < bins - 1 && excl < need && need <= excl + h0) { hit = bins - 1 - i0; cum = excl; }
      else if (i0 + 1 < bins - 1 && excl + h0 < need && need <= excl + h0 + h1) { hit = bins - 2 - i0; cum = excl + h0; }
      if (hit > 0) {
        sh_prefix = (prefix << (pass == 0 ? 0 : (pass == 1 ? 11 : 10))) | (unsigned)hit;
        sh_need = need - cum;
      }
    }
    __syncthreads();
    prefix = sh_prefix;
    need = sh_need;
    __syncthreads();
  }
  // prefix now holds the full 32-bit pattern of the threshold; `need` = how many of the rays equal to it are taken.
  // Scores are sums with a 24-bit uniform, so exact ties at the threshold do occur; they are broken by the rays' keys
  // (a property of the ray, not of its position in the list), which keeps the selection reproducible run to run.
  const unsigned thr = prefix;
  auto ray_key = [&](int r) -> unsigned long long {
    const uint32_t own = a.owner[(size_t)chunk * a.cap_rays + r];
    if (own == NMF_NO_OWNER) return ~0ull;
    const BSample* b = a.bs + own;
    return (unsigned long long)nmf_mix64(b->key, (uint64_t)(r - (int)b->roff) + NMF_STREAM_RAY0);
  };
  for (int r = first; r < n; r += stride) {
    if (__float_as_uint(scu[r].x) == thr) {
      const unsigned e = atomicAdd(sh_eq0, 1u);
      if (e < NMF_TIE_CAP) tie_key0[e] = ray_key(r);
    }
  }
  csync();
  const unsigned n_tie = *sh_eq0;
  csync();             // every thread of the cluster holds the tie count before CTA 0 reuses the counter below (racecheck)
  const bool by_key = n_tie > need && n_tie <= NMF_TIE_CAP;      // otherwise all ties are taken, or (absurdly many) first come
  if (rank == 0 && threadIdx.x == 0) {
    if (by_key) {
      // the `need` smallest keys win: selection sort over a handful of entries
      for (unsigned i = 0; i < need; ++i) {
        unsigned m = i;
        for (unsigned j = i + 1; j < n_tie; ++j) if (tie_key[j] < tie_key[m]) m = j;
        const unsigned long long t = tie_key[i]; tie_key[i] = tie_key[m]; tie_key[m] = t;
      }
      tie_cut = tie_key[need - 1];
    }
    sh_eq = 0;
  }
  csync();
  const unsigned long long cut = *tie_cut0;
  for (int r = first; r < n; r += stride) {
    const unsigned bits = __float_as_uint(scu[r].x);
    bool take = bits > thr;
    if (bits == thr) take = by_key ? ray_key(r) <= cut : atomicAdd(sh_eq0, 1u) < need;
    if (take) {
      const unsigned sl = atomicAdd(sh_slot0, 1u);
      if (sl < (unsigned)n_re) {
        BRay* o = region + r;
        const uint32_t own = a.owner[(size_t)chunk * a.cap_rays + r];
        if (own == NMF_NO_OWNER) continue;                              // overflow case only
        const BSample* b = a.bs + own;
        const float4 q = *(const float4*)o->L;
        const size_t gi = (size_t)chunk * a.max_retrace + sl;
        float* ry = a.rays1 + gi * 6;
        ry[0] = b->pos[0] + q.x * 5e-3f;                                // microfacet.py:449-452
        ry[1] = b->pos[1] + q.y * 5e-3f;
        ry[2] = b->pos[2] + q.z * 5e-3f;
        ry[3] = q.x; ry[4] = q.y; ry[5] = q.z;
        a.mip1[gi] = q.w;
        a.key1[gi] = nmf_mix64(b->key, (uint64_t)(r - (int)b->roff) + NMF_STREAM_RAY0);
        o->slot = (int)sl;
      }
    }
  }
  csync();             // CTA 0's shared memory must outlive every remote access
}

// ================================================================================================
// k_incoming: models/microfacet.py:549-613 -- incoming radiance of every bounce ray (level 0: re-traced radiance or
// environment lookup; level 1: environment), Fresnel mix, and the per-sample sums over the rays as segmented warp
// reductions.  Level 0: spec / tint debug maps go straight to the pixel accumulators, the combined radiance to the
// sample's reduction record.  Level 1: the mean radiance of the sample is weighted straight into its retraced ray.
// ================================================================================================
struct IncomingArgs {
  const BSample* bs; const BRay* brays; const uint32_t* owner; const int* ray_count; int cap_rays; const float* rgb1; int max_retrace;
  float* accum; const int* tile_start; int n_chunks; float4* red; const int2* tile_desc;
};
template <int LEVEL>
__global__ void __launch_bounds__(MLP_THREADS) k_incoming(const NmfScene s, const IncomingArgs a) {
  const NmfEnvDyn ed = nmf_env_dyn_load(s);     // mipbias and pole means: by value or from NmfScene.env_dyn
  const int n_tiles = a.tile_start[a.n_chunks];
  const int lane = threadIdx.x & 31;
  TileWalk tw;
  tw_begin(tw, a.tile_desc, n_tiles, a.ray_count, a.cap_rays, a.owner);
  // the ray record's address needs only the tile descriptor: it is loaded one tile ahead (these records stream from DRAM,
  // written by k_bounce ~3 GB earlier) and the next tile's sample record is pulled into the L2 at the end of the iteration
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 q0 = zero4, q1 = zero4;
  if (tw.r < tw.n) {
    const BRay* o = a.brays + (size_t)tw.chunk * a.cap_rays + tw.r;
    nmf_ld8((const float*)o, q0, q1);
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tw_advance(tw)) {
    tw_issue(tw, tile, a.tile_desc, n_tiles, a.ray_count, a.cap_rays, a.owner);
    float4 q0n = zero4, q1n = zero4;
    if (tw.r_n < tw.n_n) {
      const BRay* o = a.brays + (size_t)tw.chunk_n * a.cap_rays + tw.r_n;
      q0n = __ldcs((const float4*)o->L); q1n = __ldcs((const float4*)o->bw);      // (one 256-bit load measured slower here)
    }
    const int chunk = tw.chunk;
    const uint32_t key = tw.slot;
    const bool active = key != NMF_NO_OWNER;
    float comb[3] = {0.f, 0.f, 0.f}, inc[3] = {0.f, 0.f, 0.f}, bw[3] = {0.f, 0.f, 0.f};
    const BSample* b = a.bs;
    if (active) {
      const int slot = LEVEL == 0 ? __float_as_int(q1.w) : -1;
      b = a.bs + key;
      const nmf_v3 L = nmf_mk3(q0.x, q0.y, q0.z);
      if (slot >= 0) {
        const float* src = a.rgb1 + ((size_t)chunk * (size_t)a.max_retrace + (size_t)slot) * 4;
        inc[0] = src[0]; inc[1] = src[1]; inc[2] = src[2];
      } else {
        if (s.env_sat2) nmf_env_lookup1_pair(s.env_sat2, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot, L, q0.w, inc);
        else nmf_env_lookup1(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot, L, q0.w, inc);
      }
      const float4 qv = *(const float4*)b->V, q3 = *(const float4*)b->f0, q4 = *(const float4*)b->diffuse;
      const nmf_v3 H = nmf_unit(nmf_mk3((qv.x + L.x) / 2.0f, (qv.y + L.y) / 2.0f, (qv.z + L.z) / 2.0f));
      const float ch = fabsf(qv.x * H.x + qv.y * H.y + qv.z * H.z);
      const float F0 = nmf_fresnel(q3.x, ch), F1 = nmf_fresnel(q3.y, ch), F2 = nmf_fresnel(q3.z, ch);
      comb[0] = F0 * inc[0] * q1.x + (1.f - F0) * q4.x;                  // microfacet.py:585-600
      comb[1] = F1 * inc[1] * q1.y + (1.f - F1) * q4.y;
      comb[2] = F2 * inc[2] * q1.z + (1.f - F2) * q4.z;
      bw[0] = q1.x; bw[1] = q1.y; bw[2] = q1.z;
    }
    const Seg seg = seg_setup(key, lane);        // every lane takes part in the shuffles
    seg_sum3(comb, seg, lane);
    if (LEVEL == 0) {
      seg_sum3(inc, seg, lane);
      seg_sum3(bw, seg, lane);
    }
    if (active && seg.head) {
      if (LEVEL == 0) {
        const float4 hdr = a.red[2 * (size_t)key], q5 = *(const float4*)b->fresn;      // {w, count, ray, flags}
        const int cnt = max(__float_as_int(hdr.y), 1);
        const float sw = hdr.x / (float)cnt;
        float* acc = a.accum + (size_t)__float_as_uint(hdr.z) * A_N;
        float* rs = (float*)(a.red + 2 * (size_t)key + 1);
        atomicAdd(rs, comb[0]); atomicAdd(rs + 1, comb[1]); atomicAdd(rs + 2, comb[2]);
        atomicAdd(acc + A_SPEC, sw * inc[0]); atomicAdd(acc + A_SPEC + 1, sw * inc[1]); atomicAdd(acc + A_SPEC + 2, sw * inc[2]);
        atomicAdd(acc + A_TINT, sw * q5.x * bw[0]); atomicAdd(acc + A_TINT + 1, sw * q5.y * bw[1]);
        atomicAdd(acc + A_TINT + 2, sw * q5.z * bw[2]);
        atomicAdd(acc + A_TINTU, (q5.x * bw[0] + q5.y * bw[1] + q5.z * bw[2]) / (float)cnt);   // brdf_reg
      } else {
        // tensor_nerf.py:448-452 at recur 1: sum_samples w * mean_rays(comb) into the retraced ray (its own index)
        const float4 q0 = *(const float4*)b->pos, q2 = *(const float4*)b->N;
        const float sw = q0.w / (float)max(__float_as_int(q2.w), 1);
        float* acc = a.accum + (size_t)b->ray * 4;
        atomicAdd(acc, sw * comb[0]); atomicAdd(acc + 1, sw * comb[1]); atomicAdd(acc + 2, sw * comb[2]);
      }
    }
    if (tw.slot_n != NMF_NO_OWNER) {
      const char* rec = (const char*)(a.bs + tw.slot_n);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rec));          // V, f0, diffuse, fresn: bytes 16 .. 95
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + 64));
    }
    q0 = q0n; q1 = q1n;
  }
}

// ================================================================================================
// k_reduce0: tensor_nerf.py:448-452,528 -- per bounce sample: mean radiance of its rays, weighted into the pixel
// (and into the cross-section map, which clips the sample's colour first)
// ================================================================================================
struct ReduceArgs { const float4* red; const int* n_bs; int cap_bs; float* accum; };
__global__ void __launch_bounds__(256) k_reduce0(const ReduceArgs a) {
  const int n = min(*a.n_bs, a.cap_bs);
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    float4 hdr, sum;                                                              // one 32-byte sector per sample, one load
    nmf_ld8((const float*)(a.red + 2 * (size_t)i), hdr, sum);
    if (__float_as_int(hdr.y) <= 0) continue;                                       // dropped sample (overflow case only)
    const float w = hdr.x, inv = 1.0f / (float)__float_as_int(hdr.y);
    float* acc = a.accum + (size_t)__float_as_uint(hdr.z) * A_N;
    const bool below = __float_as_uint(hdr.w) & 1u;
    const float rgb[3] = {sum.x * inv, sum.y * inv, sum.z * inv};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      atomicAdd(acc + A_RGB + c, w * rgb[c]);
      if (below) atomicAdd(acc + A_CROSS + c, w * nmf_clampf(rgb[c], 0.f, 1.f));
    }
  }
}

// retraced rays: linear radiance + (1 - acc) * env(d, mip)   (tensor_nerf.py:460-468, 657-659 with tonemap=False)
__global__ void k_finish1(const NmfScene s, const float* rays1, const float* mip1, const float* acc1, const float* accum1,
                          const int* n_sec, int max_retrace, int n, float* rgb1) {
  const NmfEnvDyn ed = nmf_env_dyn_load(s);     // mipbias and pole means: by value or from NmfScene.env_dyn
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int chunk = i / max_retrace;
  if (i - chunk * max_retrace >= n_sec[chunk]) return;
  float bg[3];
  nmf_env_lookup1(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot,
                  nmf_mk3(rays1[6 * i + 3], rays1[6 * i + 4], rays1[6 * i + 5]), mip1[i], bg);
  const float t = 1.0f - acc1[i];
  rgb1[4 * i] = accum1[4 * i] + t * bg[0];
  rgb1[4 * i + 1] = accum1[4 * i + 1] + t * bg[1];
  rgb1[4 * i + 2] = accum1[4 * i + 2] + t * bg[2];
}

// ================================================================================================
// model=tensorf plumbing config: view MLP 135 -> 128 -> 128 -> 3 per surviving sample
// (models/tensorf.py:70-97, modules/render_modules.py:201-235 with viewpe = feape = 2)
// ================================================================================================
struct PlainArgs { const float* rays; const float* tmin; const Surv* surv; const int* n_surv; int cap_surv; float* accum; };
__global__ void __launch_bounds__(128) k_shade_plain(const NmfScene s, const PlainArgs a) {
  extern __shared__ __align__(128) float sm[];
  float* xbuf = sm;                       // [135][128]
  float* hbuf = sm + 135 * 128;           // [128][128]
  const int n = min(*a.n_surv, a.cap_surv);
  float* x = xbuf + threadIdx.x;
  float* hb = hbuf + threadIdx.x;
  for (int si = blockIdx.x * 128 + threadIdx.x; si < n; si += gridDim.x * 128) {
    const Surv sv = a.surv[si];
    const int ray = (int)sv.ray;
    float o[3], d[3], p[3], xn[3];
    for (int i = 0; i < 3; ++i) { o[i] = a.rays[(size_t)ray * 6 + i]; d[i] = a.rays[(size_t)ray * 6 + 3 + i]; }
    nmf_step_pos(o, d, nmf_step_z(a.tmin[ray], s.stepsize, (int)sv.step), p);
    nmf_normalize_xyz(s, p, xn);
    const NmfTaps t = nmf_vm_taps(s, xn);
    float coef[72];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
      for (int g = 0; g < 6; ++g) {
        const nmf_f4 c = nmf_app_group(s, t, pl, g);
        coef[pl * 24 + 4 * g] = c.x; coef[pl * 24 + 4 * g + 1] = c.y; coef[pl * 24 + 4 * g + 2] = c.z; coef[pl * 24 + 4 * g + 3] = c.w;
      }
    // x = [feat(24), view(3), sin(feat*1), sin(feat*2) interleaved per feature..., cos..., sin(view..), cos(view..)]
    // positional_encoding (render_modules.py:38-44): pts = (p[...,None] * [1,2]).reshape(..., 2*C); cat(sin, cos)
    for (int oo = 0; oo < 24; ++oo) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 72; ++j) acc += __ldg(s.basis_t + j * 24 + oo) * coef[j];
      x[oo * 128] = acc;
      x[(27 + 2 * oo) * 128] = sinf(acc);
      x[(27 + 2 * oo + 1) * 128] = sinf(acc * 2.0f);
      x[(27 + 48 + 2 * oo) * 128] = cosf(acc);
      x[(27 + 48 + 2 * oo + 1) * 128] = cosf(acc * 2.0f);
    }
    for (int c = 0; c < 3; ++c) {
      x[(24 + c) * 128] = d[c];
      x[(123 + 2 * c) * 128] = sinf(d[c]);
      x[(123 + 2 * c + 1) * 128] = sinf(d[c] * 2.0f);
      x[(123 + 6 + 2 * c) * 128] = cosf(d[c]);
      x[(123 + 6 + 2 * c + 1) * 128] = cosf(d[c] * 2.0f);
    }
    // layer 1 and 2 in two halves of 64 outputs to bound registers
    for (int half = 0; half < 2; ++half) {
      float h[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) h[i] = __ldg(s.plain_b0 + half * 64 + i);
      for (int k = 0; k < 135; ++k) {
        const float xv = x[k * 128];
        const float4* wr = (const float4*)(s.plain_w0t + k * 128 + half * 64);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4 wv = __ldg(wr + q);
          h[4 * q] += xv * wv.x; h[4 * q + 1] += xv * wv.y; h[4 * q + 2] += xv * wv.z; h[4 * q + 3] += xv * wv.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) hb[(half * 64 + i) * 128] = fmaxf(h[i], 0.f);
    }
    float o3[3] = {__ldg(s.plain_b2), __ldg(s.plain_b2 + 1), __ldg(s.plain_b2 + 2)};
    for (int half = 0; half < 2; ++half) {
      float h[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) h[i] = __ldg(s.plain_b1 + half * 64 + i);
      for (int k = 0; k < 128; ++k) {
        const float xv = hb[k * 128];
        const float4* wr = (const float4*)(s.plain_w1t + k * 128 + half * 64);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4 wv = __ldg(wr + q);
          h[4 * q] += xv * wv.x; h[4 * q + 1] += xv * wv.y; h[4 * q + 2] += xv * wv.z; h[4 * q + 3] += xv * wv.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float hv = fmaxf(h[i], 0.f);
        const float* w2 = s.plain_w2t + (half * 64 + i) * 3;
        o3[0] += hv * __ldg(w2); o3[1] += hv * __ldg(w2 + 1); o3[2] += hv * __ldg(w2 + 2);
      }
    }
    float* acc = a.accum + (size_t)ray * A_N;
    const bool below = xn[2] < 0.f;
    for (int c = 0; c < 3; ++c) {
      const float rgb = nmf_sigmoid(o3[c]);
      atomicAdd(acc + A_RGB + c, sv.w * rgb);
      if (below) atomicAdd(acc + A_CROSS + c, sv.w * nmf_clampf(rgb, 0.f, 1.f));
    }
  }
}

// ================================================================================================
// k_finish0: modules/tensor_nerf.py:448-566, 657-673 (eval, white background) + tonemap.py:38-49
// ================================================================================================
struct FinishArgs {
  const float* rays; const float* tmin; const float* acc; const float* depth; const int* termk; const int* nvalid;
  const float* accum; int n; float focal; int model;
  int chunk; float* stat4; int do_stats;
  const float* zvals; int n_steps;    // train mode: jittered distances
};
__global__ void k_finish0(const NmfScene s, const FinishArgs a, const NmfImages out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (a.do_stats) {
    // A19 (tensor_nerf.py:567-649): per-chunk sums behind ori_loss, diffuse_reg, brdf_reg and prediction_loss
    const bool in = i < a.n;
    const float* A = a.accum + (size_t)(in ? i : 0) * A_N;
    float v[4] = {in ? A[A_ORI] : 0.f, in ? A[A_DIFF] + A[A_DIFF + 1] + A[A_DIFF + 2] : 0.f, in ? A[A_TINTU] : 0.f,
                  in ? a.acc[i] : 0.f};
    const int chunk = (in ? i : a.n - 1) / a.chunk;
    const int chunk0 = __shfl_sync(FULL, chunk, 0);
    if (__all_sync(FULL, chunk == chunk0)) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += __shfl_xor_sync(FULL, v[q], off);
      if ((threadIdx.x & 31) == 0)
        for (int q = 0; q < 4; ++q) atomicAdd(a.stat4 + 4 * chunk0 + q, v[q]);
    } else if (in) {
      for (int q = 0; q < 4; ++q) atomicAdd(a.stat4 + 4 * chunk + q, v[q]);
    }
  }
  if (i >= a.n) return;
  const float acc = a.acc[i];
  const float t = 1.0f - acc;
  const float* A = a.accum + (size_t)i * A_N;
  if (out.acc_map) out.acc_map[i] = acc;
  if (out.depth) out.depth[i] = a.depth[i];
  if (out.surf_width) out.surf_width[i] = (int64_t)a.nvalid[i];
  if (out.termination_xyz) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = a.termk[i];
    if (k >= 0) {
      float o[3], d[3];
      for (int c = 0; c < 3; ++c) { o[c] = a.rays[(size_t)i * 6 + c]; d[c] = a.rays[(size_t)i * 6 + 3 + c]; }
      const float z = a.zvals ? a.zvals[(size_t)i * a.n_steps + k] : nmf_step_z(a.tmin[i], s.stepsize, k);
      nmf_step_pos(o, d, z, v);
      v[3] = z / a.focal;                                               // alphagrid.py:200
    }
    for (int c = 0; c < 4; ++c) out.termination_xyz[(size_t)i * 4 + c] = v[c];
  }
  for (int c = 0; c < 3; ++c) {
    const size_t j = (size_t)i * 3 + c;
    if (out.rgb_map) out.rgb_map[j] = nmf_clampf(nmf_srgb(A[A_RGB + c]), 0.f, 1.f) + t;   // bg = white
    if (out.world_normal) out.world_normal[j] = acc * A[A_WN + c] + t;                    // tensor_nerf.py:497-499
    if (out.normal) out.normal[j] = t;                                                    // no predicted normals
    if (out.cross_section) out.cross_section[j] = A[A_CROSS + c];
    if (a.model == 0) {
      if (out.diffuse) out.diffuse[j] = A[A_DIFF + c] + t;
      if (out.tint) out.tint[j] = A[A_TINT + c] + t;
      if (out.roughness) out.roughness[j] = A[A_ROUGH] + t;
      if (out.spec) out.spec[j] = A[A_SPEC + c] + t;
      if (out.albedo) out.albedo[j] = A[A_ALB + c] + t;
    }
  }
}

__global__ void k_export_counters(const WS w, const NmfCounters c, int model) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w.n_chunks) {
    if (c.n_samples0) c.n_samples0[i] = w.n_samples0[i];
    if (c.n_cand) c.n_cand[i] = w.n_cand[i];
    if (c.n_samples1) c.n_samples1[i] = w.n_samples1[i];
    if (c.n_bounce_rays0) c.n_bounce_rays0[i] = w.ray_count0[i];
    if (c.n_bounce_rays1) c.n_bounce_rays1[i] = w.ray_count1[i];
    if (c.n_retrace) c.n_retrace[i] = w.n_sec[i];
    if (c.stat4)
      for (int q = 0; q < 4; ++q) c.stat4[4 * i + q] = w.stat4[4 * i + q];
  }
  if (i == 0) {
    if (c.n_shaded) { c.n_shaded[0] = w.n_surv[0]; c.n_shaded[1] = w.n_surv[1]; }
    if (c.error) *c.error = *w.error;
  }
}

// ================================================================================================
// host side of the C ABI
// ================================================================================================
static int g_sms = 0;
static int sm_count() {
  if (!g_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}
static int blocks_for(long long work_items, int per_block, int max_per_sm) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = (long long)sm_count() * max_per_sm;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}
// grid of a persistent (grid-stride) kernel: exactly the number of CTAs that are resident at once, so that the work is
// dealt in ONE wave (a grid of 8 CTAs/SM for a kernel that fits 7 runs a second, nearly empty wave: +43 % time)
template <class K>
static int resident_grid(K kernel, int threads, size_t smem) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return sm_count() * per_sm;
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
#define CKL() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

static int check_scene(const NmfScene* s) {
  if (!s) return NMF_E_ARG;
  if (s->n_steps <= 0 || s->n_steps > NMF_MAX_STEPS) return NMF_E_UNSUPPORTED;
  for (int p = 0; p < 3; ++p)
    if (!s->dval[p] || !s->lval[p] || s->plane_w[p] < 2 || s->plane_h[p] < 2 || s->line_n[p] < 2) return NMF_E_ARG;
  // one grid size per axis (nmf_vm_taps relies on it): plane p is grid[mat1(p)] x grid[mat0(p)], line p is grid[vec(p)]
  if (s->plane_w[1] != s->plane_w[0] || s->line_n[2] != s->plane_w[0] || s->plane_w[2] != s->plane_h[0] ||
      s->line_n[1] != s->plane_h[0] || s->plane_h[2] != s->plane_h[1] || s->line_n[0] != s->plane_h[1])
    return NMF_E_UNSUPPORTED;
  if (s->has_occ && (!s->occ_vox || !s->occ_cell || (s->opitch & 31))) return NMF_E_ARG;
  if (s->has_occ && s->occ_coarse &&
      (s->ocw != (s->ow + 7) / 8 || s->och != (s->oh + 7) / 8 || s->ocd != (s->od + 7) / 8 ||
       ((long long)s->ocw * s->och * s->ocd + 31) / 32 > NMF_MAX_COARSE_WORDS))
    return NMF_E_ARG;
  return NMF_OK;
}

// ---- optional phase timing: CUDA events recorded on the caller's stream between the phases ----
static const char* g_phase_names[NMF_N_PHASES] = {"march0", "shade0", "bounce0", "select", "march1", "shade1", "bounce1",
                                                   "incoming1", "finish1", "incoming0", "reduce0", "finish"};
static cudaEvent_t g_ev[NMF_N_PHASES + 1];
static bool g_ev_made = false, g_prof_on = false, g_ev_rec[NMF_N_PHASES + 1];
static void prof_mark(int i, cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEventRecord(g_ev[i], st);
  g_ev_rec[i] = true;
}
extern "C" int nmf_profile_enable(int on) {
  if (on && !g_ev_made) {
    for (int i = 0; i <= NMF_N_PHASES; ++i) CK(cudaEventCreate(&g_ev[i]));
    g_ev_made = true;
  }
  g_prof_on = on != 0;
  for (int i = 0; i <= NMF_N_PHASES; ++i) g_ev_rec[i] = false;
  return NMF_OK;
}
extern "C" int nmf_profile_read(float* ms, int n) {
  if (!ms || n < NMF_N_PHASES || !g_ev_made) return NMF_E_ARG;
  int last = 0;
  for (int i = 0; i < NMF_N_PHASES; ++i) {
    ms[i] = 0.f;
    if (!g_ev_rec[i + 1] || !g_ev_rec[last]) continue;
    CK(cudaEventElapsedTime(&ms[i], g_ev[last], g_ev[i + 1]));
    last = i + 1;
  }
  return NMF_OK;
}
extern "C" const char* nmf_profile_phase_name(int i) { return (i >= 0 && i < NMF_N_PHASES) ? g_phase_names[i] : ""; }

extern "C" int nmf_abi_version(void) { return NMF_ABI_VERSION; }

extern "C" size_t nmf_workspace_bytes(const NmfScene* scene, int n_rays, int chunk) {
  if (!scene || n_rays <= 0 || chunk <= 0) return 0;
  WS w;
  carve(w, scene, n_rays, chunk, nullptr);
  return w.total;
}

extern "C" size_t nmf_workspace_bytes_scaled(const NmfScene* scene, int n_rays, int chunk, float cap_scale) {
  if (!scene || n_rays <= 0 || chunk <= 0) return 0;
  WS w;
  carve(w, scene, n_rays, chunk, nullptr, cap_scale);
  return w.total;
}

// Host-buffer variant (nmf_render_rays_host): maps are copied to the host as soon as they are final, on a second stream,
// while the rest of the sequence runs -- geometry maps after the march, material maps after the shade, radiance maps
// at the end.  Stream order is kept with events only (no host synchronisation).
struct StagedCopy {
  const NmfImages* host;
  cudaStream_t copy;
  cudaEvent_t ready[2], done;
};
static void copy_map(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (dst && src) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
}
static void copy_stage(const NmfImages* h, const NmfImages* d, int stage, size_t n, cudaStream_t st) {
  if (stage == 0) {
    copy_map(h->acc_map, d->acc_map, n * 4, st); copy_map(h->depth, d->depth, n * 4, st);
    copy_map(h->surf_width, d->surf_width, n * 8, st); copy_map(h->termination_xyz, d->termination_xyz, n * 16, st);
    copy_map(h->normal, d->normal, n * 12, st);
  } else if (stage == 1) {
    copy_map(h->world_normal, d->world_normal, n * 12, st); copy_map(h->albedo, d->albedo, n * 12, st);
    copy_map(h->roughness, d->roughness, n * 12, st); copy_map(h->diffuse, d->diffuse, n * 12, st);
  } else {
    copy_map(h->rgb_map, d->rgb_map, n * 12, st); copy_map(h->spec, d->spec, n * 12, st);
    copy_map(h->tint, d->tint, n * 12, st); copy_map(h->cross_section, d->cross_section, n * 12, st);
  }
}
static NmfImages stage_images(const NmfImages& o, int stage) {
  NmfImages r = {};
  if (stage == 0) { r.acc_map = o.acc_map; r.depth = o.depth; r.surf_width = o.surf_width; r.termination_xyz = o.termination_xyz; r.normal = o.normal; }
  else if (stage == 1) { r.world_normal = o.world_normal; r.albedo = o.albedo; r.roughness = o.roughness; r.diffuse = o.diffuse; }
  else { r.rgb_map = o.rgb_map; r.spec = o.spec; r.tint = o.tint; r.cross_section = o.cross_section; }
  return r;
}

static int render_impl(const NmfScene* scene, const NmfRender* rp, const float* rays, const NmfImages* out,
                       const NmfCounters* counters, void* workspace, size_t workspace_bytes, void* stream_, const StagedCopy* sc,
                       const NmfRenderTrain* tr = nullptr, WS* ws_out = nullptr) {
  int st = check_scene(scene);
  if (st) return st;
  if (!rp || !rays || !out || !workspace || rp->n_rays <= 0 || rp->chunk <= 0) return NMF_E_ARG;
  if (((uintptr_t)workspace & 255) != 0) return NMF_E_ARG;
  const NmfScene& s = *scene;
  if (s.model == 0 && (!s.aval[0] || !s.dpack[0] || !s.basis_t || !s.head_w || !s.brdf_w0t || !s.brdf_w0u || !s.brdf_w1u || !s.brdf_w2u || !s.sobol || !s.sh_conv || !s.env_sat))
    return NMF_E_ARG;
  if (s.model == 1 && (!s.aval[0] || !s.basis_t || !s.plain_w0t)) return NMF_E_ARG;
  if (s.model == 0 && s.max_retrace > 0 && s.max_brdf_rays1 <= 0) return NMF_E_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  WS w;
  if (tr && (s.model != 0 || rp->n_rays > rp->chunk || !tr->whole_valid || !tr->n_kept)) return NMF_E_UNSUPPORTED;
  carve(w, scene, rp->n_rays, rp->chunk, (char*)workspace, rp->cap_scale, tr != nullptr);
  if (w.total > workspace_bytes) return NMF_E_WORKSPACE;
  if (ws_out) *ws_out = w;
  const int n = rp->n_rays, nc = w.n_chunks;
  CK(cudaMemsetAsync(w.counters_base, 0, w.counters_bytes, stream));
  prof_mark(0, stream);

  MarchArgs m0 = {};
  m0.rays = rays; m0.n = n; m0.group = rp->chunk; m0.seed = rp->seed; m0.ray_id0 = rp->ray_id0;
  m0.skip_eps = rp->skip_eps; m0.t_cut = rp->t_cut;
  m0.tmin = w.tmin0; m0.acc = w.acc0; m0.depth = w.depth0; m0.termk = w.termk0; m0.nvalid = w.nvalid0;
  m0.n_samples = w.n_samples0; m0.n_cand = w.n_cand; m0.wsum = nullptr;
  m0.surv = w.surv0; m0.n_surv = w.n_surv; m0.cap_surv = w.cap_surv0; m0.error = w.error;
  m0.zvals = w.zvals0; m0.whole = tr ? tr->whole_valid : nullptr;
  if (tr) { m0.vs = w.vs0; m0.vdw = w.vdw0; m0.vbase = w.vbase0; m0.n_vs = w.n_vs; m0.cap_vs = w.cap_vs0; m0.survv = w.survv0; }
  static int g_march0 = 0, g_march1 = 0, g_inc0 = 0, g_inc1 = 0;
  if (!g_march0) {
    g_march0 = resident_grid(k_march<0, 0>, 256, 0); g_march1 = resident_grid(k_march<1, 0>, 256, 0);
    g_inc0 = resident_grid(k_incoming<0>, MLP_THREADS, 0); g_inc1 = resident_grid(k_incoming<1>, MLP_THREADS, 0);
  }
  // rays differ a lot in cost: several waves of small CTAs (a multiple of the resident count) balance better than one
  if (tr) {
    // alphagrid.py:167-207 + 353-364: jittered distances and per-ray counts of the whole batch, then the truncation
    st = nmf_sample_rays_train(scene, rays, n, -1.0f, rp->seed, rp->ray_id0, nullptr, tr->max_samples, nullptr, w.zvals0,
                               w.nvalid0, tr->whole_valid, tr->n_kept, stream_);
    if (st) return st;
    k_march<0, 1><<<min(4 * g_march0, blocks_for(n, 8, 1 << 20)), 256, 0, stream>>>(s, m0);
  } else {
    k_march<0, 0><<<min(4 * g_march0, blocks_for(n, 8, 1 << 20)), 256, 0, stream>>>(s, m0);
  }
  CKL();
  FinishArgs fa = {rays, w.tmin0, w.acc0, w.depth0, w.termk0, w.nvalid0, w.accum0, n, rp->focal, s.model, rp->chunk, w.stat4, 0,
                   w.zvals0, s.n_steps};
  if (sc && s.model == 0) {
    k_finish0<<<(n + 127) / 128, 128, 0, stream>>>(s, fa, stage_images(*out, 0));
    CKL();
    CK(cudaEventRecord(sc->ready[0], stream));
    CK(cudaStreamWaitEvent(sc->copy, sc->ready[0], 0));
    copy_stage(sc->host, out, 0, (size_t)n, sc->copy);
  }
  prof_mark(1, stream);

  if (s.model == 0) {
    const bool tcm = s.mlp_mode == 0;
    const size_t mlp_smem = tcm ? (size_t)TC_SMEM_BYTES : MLP_SMEM_FLOATS * sizeof(float);
    // opt-in shared-memory sizes are a per-device function attribute: set them once per device of this process
    static unsigned long long attr_done_mask = 0;
    int dev_id = 0;
    CK(cudaGetDevice(&dev_id));
    const bool attr_done = (attr_done_mask >> (dev_id & 63)) & 1ull;
    if (!attr_done) {
      CK(cudaFuncSetAttribute(k_shade<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SHADE_SMEM_FLOATS * sizeof(float))));
      CK(cudaFuncSetAttribute(k_shade<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SHADE_SMEM_FLOATS * sizeof(float))));
      CK(cudaFuncSetAttribute(k_shade<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SHADE_SMEM_FLOATS * sizeof(float))));
      CK(cudaFuncSetAttribute(k_shade<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SHADE_SMEM_FLOATS * sizeof(float))));
      CK(cudaFuncSetAttribute(k_bounce<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MLP_SMEM_FLOATS * sizeof(float))));
      CK(cudaFuncSetAttribute(k_bounce<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MLP_SMEM_FLOATS * sizeof(float))));
      CK(cudaFuncSetAttribute(k_bounce<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
      CK(cudaFuncSetAttribute(k_bounce<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
      attr_done_mask |= 1ull << (dev_id & 63);
    }
    ShadeArgs h0 = {};
    h0.rays = rays; h0.tmin = w.tmin0; h0.seed = rp->seed; h0.ray_id0 = rp->ray_id0; h0.group = rp->chunk;
    h0.surv = w.surv0; h0.n_surv = w.n_surv; h0.cap_surv = w.cap_surv0; h0.accum = w.accum0;
    h0.bs = w.bs0; h0.n_bs = w.n_bs; h0.cap_bs = w.cap_bs0; h0.ray_count = w.ray_count0; h0.cap_rays = w.cap_rays0;
    h0.owner = w.owner0; h0.error = w.error; h0.red = w.red0;
    h0.zvals = w.zvals0; h0.n_steps = s.n_steps; h0.min_rough = tr ? tr->min_rough : 0.f;
    if (tr) { h0.survv = w.survv0; h0.survslot = w.survslot0; }
    if (tr) k_shade<0, 1><<<sm_count() * 3, 256, SHADE_SMEM_FLOATS * sizeof(float), stream>>>(s, h0);
    else k_shade<0, 0><<<sm_count() * 3, 256, SHADE_SMEM_FLOATS * sizeof(float), stream>>>(s, h0);
    CKL();
    if (sc) {
      k_finish0<<<(n + 127) / 128, 128, 0, stream>>>(s, fa, stage_images(*out, 1));
      CKL();
      CK(cudaEventRecord(sc->ready[1], stream));
      CK(cudaStreamWaitEvent(sc->copy, sc->ready[1], 0));
      copy_stage(sc->host, out, 1, (size_t)n, sc->copy);
    }
    prof_mark(2, stream);

    // persistent grid over the flat tile list: 5 CTAs per SM with the fp16 operand tiles, 3 with the fp32 SIMT staging
    const int gb = sm_count() * (tcm ? 5 : 3);
    k_tile_prefix<<<1, 1024, 0, stream>>>(w.ray_count0, w.cap_rays0, nc, w.tile_start0, w.tile_desc0);
    CKL();
    BounceArgs b0 = {w.bs0, w.brays0, w.owner0, w.ray_count0, w.cap_rays0, w.score_sum, w.scu0, w.tile_start0, nc, w.tile_desc0};
    if (tcm) k_bounce<0, 1><<<gb, MLP_THREADS, mlp_smem, stream>>>(s, b0);
    else k_bounce<0, 0><<<gb, MLP_THREADS, mlp_smem, stream>>>(s, b0);
    CKL();
    prof_mark(3, stream);

    if (s.max_retrace > 0) {
      SelectArgs sa = {w.bs0, w.brays0, w.owner0, w.scu0, w.ray_count0, w.cap_rays0, w.score_sum, s.max_retrace, w.n_sec, w.rays1, w.mip1, w.key1};
      {
        // cluster size: enough CTAs to fill the GPU when there are few chunks (training: one chunk)
        const unsigned cl = nc <= 18 ? 8u : (nc <= 37 ? 4u : (nc <= 74 ? 2u : 1u));
        if (cl == 1u) {
          k_select<0><<<nc, 1024, 0, stream>>>(sa);
        } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)nc * cl); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, k_select<1>, sa));
        }
      }
      CKL();
      prof_mark(4, stream);
      MarchArgs m1 = {};
      m1.rays = w.rays1; m1.n = w.n_rays1; m1.group = s.max_retrace; m1.n_active = w.n_sec; m1.keys = w.key1;
      m1.skip_eps = rp->skip_eps; m1.t_cut = rp->t_cut;
      m1.tmin = w.tmin1; m1.acc = w.acc1; m1.depth = nullptr; m1.termk = nullptr; m1.nvalid = w.nvalid1;
      m1.n_samples = w.n_samples1; m1.n_cand = w.n_cand; m1.wsum = w.wsum1;
      m1.surv = w.surv1; m1.n_surv = w.n_surv + 1; m1.cap_surv = w.cap_surv1; m1.error = w.error;
      m1.zvals = w.zvals1;
      if (tr) { m1.vs = w.vs1; m1.vdw = w.vdw1; m1.vbase = w.vbase1; m1.n_vs = w.n_vs + 1; m1.cap_vs = w.cap_vs1; m1.survv = w.survv1; }
      if (tr) k_march<1, 1><<<min(4 * g_march1, blocks_for(w.n_rays1, 8, 1 << 20)), 256, 0, stream>>>(s, m1);
      else k_march<1, 0><<<min(4 * g_march1, blocks_for(w.n_rays1, 8, 1 << 20)), 256, 0, stream>>>(s, m1);
      CKL();
      prof_mark(5, stream);
      ShadeArgs h1 = {};
      h1.rays = w.rays1; h1.tmin = w.tmin1; h1.keys = w.key1; h1.group = s.max_retrace;
      h1.surv = w.surv1; h1.n_surv = w.n_surv + 1; h1.cap_surv = w.cap_surv1; h1.accum = nullptr;
      h1.bs = w.bs1; h1.n_bs = w.n_bs + 1; h1.cap_bs = w.cap_bs1; h1.ray_count = w.ray_count1; h1.cap_rays = w.cap_rays1;
      h1.owner = w.owner1; h1.n_samples = w.n_samples1; h1.wsum = w.wsum1; h1.error = w.error;
      h1.zvals = w.zvals1; h1.n_steps = s.n_steps; h1.min_rough = tr ? tr->min_rough : 0.f;
      if (tr) { h1.survv = w.survv1; h1.survslot = w.survslot1; }
      if (tr) k_shade<1, 1><<<sm_count() * 3, 256, SHADE_SMEM_FLOATS * sizeof(float), stream>>>(s, h1);
      else k_shade<1, 0><<<sm_count() * 3, 256, SHADE_SMEM_FLOATS * sizeof(float), stream>>>(s, h1);
      CKL();
      prof_mark(6, stream);
      k_tile_prefix<<<1, 1024, 0, stream>>>(w.ray_count1, w.cap_rays1, nc, w.tile_start1, w.tile_desc1);
      CKL();
      BounceArgs b1 = {w.bs1, w.brays1, w.owner1, w.ray_count1, w.cap_rays1, nullptr, nullptr, w.tile_start1, nc, w.tile_desc1};
      if (tcm) k_bounce<1, 1><<<gb, MLP_THREADS, mlp_smem, stream>>>(s, b1);
      else k_bounce<1, 0><<<gb, MLP_THREADS, mlp_smem, stream>>>(s, b1);
      CKL();
      prof_mark(7, stream);
      IncomingArgs i1 = {w.bs1, w.brays1, w.owner1, w.ray_count1, w.cap_rays1, nullptr, 0, w.accum1, w.tile_start1, nc, nullptr, w.tile_desc1};
      k_incoming<1><<<g_inc1, MLP_THREADS, 0, stream>>>(s, i1);
      CKL();
      prof_mark(8, stream);
      k_finish1<<<(w.n_rays1 + 127) / 128, 128, 0, stream>>>(s, w.rays1, w.mip1, w.acc1, w.accum1, w.n_sec, s.max_retrace,
                                                           w.n_rays1, w.rgb1);
      CKL();
      prof_mark(9, stream);
    }
    IncomingArgs ia = {w.bs0, w.brays0, w.owner0, w.ray_count0, w.cap_rays0, w.rgb1, s.max_retrace, w.accum0, w.tile_start0, nc, w.red0, w.tile_desc0};
    k_incoming<0><<<g_inc0, MLP_THREADS, 0, stream>>>(s, ia);
    CKL();
    prof_mark(10, stream);
    ReduceArgs r0 = {w.red0, w.n_bs, w.cap_bs0, w.accum0};
    k_reduce0<<<sm_count() * 4, 256, 0, stream>>>(r0);
    CKL();
    prof_mark(11, stream);
  } else {
    const size_t smem = (135 + 128) * 128 * sizeof(float);
    static bool attr_done2 = false;
    if (!attr_done2) {
      CK(cudaFuncSetAttribute(k_shade_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done2 = true;
    }
    PlainArgs pa = {rays, w.tmin0, w.surv0, w.n_surv, w.cap_surv0, w.accum0};
    k_shade_plain<<<sm_count(), 128, smem, stream>>>(s, pa);
    CKL();
    prof_mark(2, stream);
  }
  fa.do_stats = 1;
  const bool staged = sc && s.model == 0;
  k_finish0<<<(n + 127) / 128, 128, 0, stream>>>(s, fa, staged ? stage_images(*out, 2) : *out);
  CKL();
  if (staged) {
    copy_stage(sc->host, out, 2, (size_t)n, stream);            // the last maps go out on the main stream
    CK(cudaEventRecord(sc->done, sc->copy));
    CK(cudaStreamWaitEvent(stream, sc->done, 0));                // ... which also waits for the early copies
  } else if (sc) {
    for (int g = 0; g < 3; ++g) copy_stage(sc->host, out, g, (size_t)n, stream);
  }
  if (counters) {
    k_export_counters<<<(nc + 127) / 128 > 0 ? (nc + 127) / 128 : 1, 128, 0, stream>>>(w, *counters, s.model);
    CKL();
  }
  prof_mark(12, stream);
  return NMF_OK;
}

extern "C" int nmf_render_rays(const NmfScene* scene, const NmfRender* rp, const float* rays, const NmfImages* out,
                               const NmfCounters* counters, void* workspace, size_t workspace_bytes, void* stream_) {
  return render_impl(scene, rp, rays, out, counters, workspace, workspace_bytes, stream_, nullptr);
}

extern "C" size_t nmf_render_train_workspace_bytes(const NmfScene* scene, int n_rays, float cap_scale) {
  if (!scene || n_rays <= 0) return 0;
  WS w;
  carve(w, scene, n_rays, n_rays, nullptr, cap_scale, true);
  return w.total;
}

extern "C" int nmf_render_rays_train(const NmfScene* scene, const NmfRender* rp, const NmfRenderTrain* tr, const float* rays,
                                     const NmfImages* out, const NmfCounters* counters, void* workspace,
                                     size_t workspace_bytes, void* stream_) {
  if (!tr) return NMF_E_ARG;
  return render_impl(scene, rp, rays, out, counters, workspace, workspace_bytes, stream_, nullptr, tr);
}

// the training forward for csrc/nmf_mf_train.cu: same call, and the carved workspace comes back so that the reverse pass
// can read the records the forward left behind
int nmf_render_impl_train(const NmfScene* scene, const NmfRender* rp, const NmfRenderTrain* tr, const float* rays,
                          const NmfImages* out, const NmfCounters* counters, void* workspace, size_t workspace_bytes,
                          void* stream_, WS* ws_out) {
  if (!tr) return NMF_E_ARG;
  return render_impl(scene, rp, rays, out, counters, workspace, workspace_bytes, stream_, nullptr, tr, ws_out);
}

extern "C" int nmf_render_rays_host(const NmfScene* scene, const NmfRender* rp, const float* rays_host, float* rays_dev,
                                    const NmfImages* out_host, const NmfImages* out_dev, const NmfCounters* counters_host,
                                    const NmfCounters* counters_dev, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!rp || !rays_host || !rays_dev || !out_host || !out_dev) return NMF_E_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t n = (size_t)rp->n_rays;
  CK(cudaMemcpyAsync(rays_dev, rays_host, n * 6 * sizeof(float), cudaMemcpyHostToDevice, stream));
  // one copy stream and three events per process (created on first use; one host thread drives a device)
  static StagedCopy sc = {};
  static bool sc_made = false;
  static int sc_dev = -1;
  int cur_dev = 0;
  CK(cudaGetDevice(&cur_dev));
  if (sc_made && cur_dev != sc_dev) return NMF_E_UNSUPPORTED;     // one device per process (one process per GPU)
  sc_dev = cur_dev;
  if (!sc_made) {
    CK(cudaStreamCreateWithFlags(&sc.copy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&sc.ready[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&sc.ready[1], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&sc.done, cudaEventDisableTiming));
    sc_made = true;
  }
  sc.host = out_host;
  int st = render_impl(scene, rp, rays_dev, out_dev, counters_dev, workspace, workspace_bytes, stream_, &sc);
  if (st) return st;
  if (counters_host && counters_dev) {
    const size_t nc = (n + rp->chunk - 1) / rp->chunk;
#define C2H(field, cnt) \
  if (counters_host->field && counters_dev->field) CK(cudaMemcpyAsync(counters_host->field, counters_dev->field, (cnt) * 4, cudaMemcpyDeviceToHost, stream));
    C2H(n_samples0, nc) C2H(n_samples1, nc) C2H(n_cand, nc) C2H(n_bounce_rays0, nc) C2H(n_bounce_rays1, nc) C2H(n_retrace, nc)
    C2H(n_shaded, 2) C2H(error, 1) C2H(stat4, 4 * nc)
#undef C2H
  }
  return NMF_OK;
}

// ================================================================================================
// plugin-slot operators (unfused)
// ================================================================================================
__global__ void k_sample_rays(const NmfScene s, const float* rays, int n, float near_override, uint8_t* valid, float* zv, int* n_valid) {
  const int lane = threadIdx.x & 31;
  const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n) return;
  float o[3], d[3];
  for (int i = 0; i < 3; ++i) { o[i] = rays[(size_t)ray * 6 + i]; d[i] = rays[(size_t)ray * 6 + 3 + i]; }
  const float tmin = nmf_ray_tmin(o, d, s.aabb0, s.aabb1, near_override >= 0.f ? near_override : s.near, s.far);
  int nv = 0;
  for (int k = lane; k < s.n_steps; k += 32) {
    const float z = nmf_step_z(tmin, s.stepsize, k);
    float p[3];
    nmf_step_pos(o, d, z, p);
    bool ok = nmf_inside(p, s.aabb0, s.aabb1);
    if (ok && s.has_occ) {
      float xn[3];
      nmf_normalize_xyz(s, p, xn);
      ok = nmf_occupied(s.occ_vox, s.occ_cell, s.ow, s.oh, s.od, s.opitch, xn[0], xn[1], xn[2]);
    }
    if (valid) valid[(size_t)ray * s.n_steps + k] = ok;
    if (zv) zv[(size_t)ray * s.n_steps + k] = z;
    nv += ok;
  }
  for (int off = 16; off > 0; off >>= 1) nv += __shfl_xor_sync(FULL, nv, off);
  if (lane == 0 && n_valid) n_valid[ray] = nv;
}
extern "C" int nmf_sample_rays(const NmfScene* scene, const float* rays, int n_rays, float near_override, uint8_t* ray_valid,
                               float* z_vals, int* n_valid, void* stream) {
  int st = check_scene(scene);
  if (st) return st;
  if (!rays || n_rays <= 0) return NMF_E_ARG;
  k_sample_rays<<<(n_rays + 7) / 8, 256, 0, (cudaStream_t)stream>>>(*scene, rays, n_rays, near_override, ray_valid, z_vals, n_valid);
  CKL();
  return NMF_OK;
}

// 4 lanes per point (density), 8 lanes per point (appearance / normals): same lane maps as the fused kernels
__global__ void k_vm_density(const NmfScene s, const float* xyz, int n, int stride, int activate, float* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 2, sub = t & 3;
  const bool active = i < n;
  float xn[3];
  nmf_normalize_xyz(s, xyz + (size_t)(active ? i : 0) * stride, xn);
  const NmfTaps tp = nmf_vm_taps(s, xn);
  float f = nmf_density_group(s, tp, sub);
  f += __shfl_xor_sync(FULL, f, 1);
  f += __shfl_xor_sync(FULL, f, 2);
  if (active && sub == 0) out[i] = activate ? nmf_feature2density(f, s.density_shift) : f;
}
extern "C" int nmf_vm_density(const NmfScene* scene, const float* xyz, int n, int stride, int activate, float* sigma, void* stream) {
  int st = check_scene(scene);
  if (st) return st;
  if (!xyz || !sigma || n <= 0 || stride < 3) return NMF_E_ARG;
  k_vm_density<<<(n * 4 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*scene, xyz, n, stride, activate, sigma);
  CKL();
  return NMF_OK;
}
__global__ void k_vm_app(const NmfScene s, const float* xyz, int n, int stride, float* out) {
  __shared__ float s_coef[32][73];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 3, l = t & 7, sidx = threadIdx.x >> 3;
  const bool active = i < n;
  float xn[3];
  nmf_normalize_xyz(s, xyz + (size_t)(active ? i : 0) * stride, xn);
  const NmfTaps tp = nmf_vm_taps(s, xn);
  if (l < 6) {
    for (int pl = 0; pl < 3; ++pl) {
      const nmf_f4 c = nmf_app_group(s, tp, pl, l);
      float* q = &s_coef[sidx][pl * 24 + 4 * l];
      q[0] = c.x; q[1] = c.y; q[2] = c.z; q[3] = c.w;
    }
  }
  __syncwarp();
  float f0 = 0.f, f1 = 0.f, f2 = 0.f;
  for (int j = 0; j < 72; ++j) {
    const float c = s_coef[sidx][j];
    f0 += __ldg(s.basis_t + j * 24 + l) * c;
    f1 += __ldg(s.basis_t + j * 24 + l + 8) * c;
    f2 += __ldg(s.basis_t + j * 24 + l + 16) * c;
  }
  if (active) { out[(size_t)i * 24 + l] = f0; out[(size_t)i * 24 + l + 8] = f1; out[(size_t)i * 24 + l + 16] = f2; }
}
extern "C" int nmf_vm_appfeature(const NmfScene* scene, const float* xyz, int n, int stride, float* feat, void* stream) {
  int st = check_scene(scene);
  if (st) return st;
  if (!xyz || !feat || n <= 0 || stride < 3 || !scene->aval[0] || !scene->basis_t) return NMF_E_ARG;
  k_vm_app<<<(n * 8 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*scene, xyz, n, stride, feat);
  CKL();
  return NMF_OK;
}
__global__ void k_vm_normals(const NmfScene s, const float* xyz, int n, int stride, float* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 3, l = t & 7;
  const bool active = i < n;
  float xn[3];
  nmf_normalize_xyz(s, xyz + (size_t)(active ? i : 0) * stride, xn);
  const NmfTaps tp = nmf_vm_taps(s, xn);
  float grad[3] = {0.f, 0.f, 0.f};
  nmf_normal_lane(s, tp, l, grad);
  for (int off = 1; off < 8; off <<= 1) {
    grad[0] += __shfl_xor_sync(FULL, grad[0], off);
    grad[1] += __shfl_xor_sync(FULL, grad[1], off);
    grad[2] += __shfl_xor_sync(FULL, grad[2], off);
  }
  const nmf_v3 nn = nmf_normal_from_grad(s, grad);
  if (active && l == 0) { out[(size_t)i * 3] = nn.x; out[(size_t)i * 3 + 1] = nn.y; out[(size_t)i * 3 + 2] = nn.z; }
}
extern "C" int nmf_vm_normals(const NmfScene* scene, const float* xyz, int n, int stride, float* normals, void* stream) {
  int st = check_scene(scene);
  if (st) return st;
  if (!xyz || !normals || n <= 0 || stride < 3 || !scene->dpack[0]) return NMF_E_ARG;
  k_vm_normals<<<(n * 8 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*scene, xyz, n, stride, normals);
  CKL();
  return NMF_OK;
}
__global__ void k_env_lookup(const NmfScene s, const float* dirs, const float* mip, int n, float* out) {
  const NmfEnvDyn ed = nmf_env_dyn_load(s);     // mipbias and pole means: by value or from NmfScene.env_dyn
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float rgb[3];
  nmf_env_lookup1(s.env_sat, s.env_h, s.env_w, ed.mipbias, ed.top, ed.bot,
                  nmf_mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), mip[i], rgb);
  out[3 * i] = rgb[0]; out[3 * i + 1] = rgb[1]; out[3 * i + 2] = rgb[2];
}
extern "C" int nmf_env_lookup(const NmfScene* scene, const float* dirs, const float* mip, int n, float* rgb, void* stream) {
  if (!scene || !scene->env_sat || !dirs || !mip || !rgb || n <= 0) return NMF_E_ARG;
  k_env_lookup<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*scene, dirs, mip, n, rgb);
  CKL();
  return NMF_OK;
}
__global__ void k_ggx(const float* u, const float* V, const float* N, const float* r, int n, float* L, float* logpdf,
                      float* half_l, float* diff_l) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const NmfGGX g = nmf_ggx_sample(u[2 * i], u[2 * i + 1], nmf_mk3(V[3 * i], V[3 * i + 1], V[3 * i + 2]),
                                  nmf_mk3(N[3 * i], N[3 * i + 1], N[3 * i + 2]), r[i]);
  L[3 * i] = g.L.x; L[3 * i + 1] = g.L.y; L[3 * i + 2] = g.L.z;
  logpdf[i] = g.logpdf;
  if (half_l) { half_l[3 * i] = g.half_l.x; half_l[3 * i + 1] = g.half_l.y; half_l[3 * i + 2] = g.half_l.z; }
  if (diff_l) { diff_l[3 * i] = g.diff_l.x; diff_l[3 * i + 1] = g.diff_l.y; diff_l[3 * i + 2] = g.diff_l.z; }
}
extern "C" int nmf_ggx_sample(const float* u, const float* V, const float* N, const float* r, int n, float* L, float* logpdf,
                              float* half_local, float* diff_local, void* stream) {
  if (!u || !V || !N || !r || !L || !logpdf || n <= 0) return NMF_E_ARG;
  k_ggx<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(u, V, N, r, n, L, logpdf, half_local, diff_local);
  CKL();
  return NMF_OK;
}
template <int TC>
__global__ void __launch_bounds__(MLP_THREADS) k_brdf_mlp(const NmfScene s, const float* feat, const float* half_l,
                                                          const float* diff_l, const float* rough, int n, float* out) {
  extern __shared__ __align__(128) float sm[];
  TcMlp tc;
  float* xcol = nullptr;
  if (TC) {
    tc_mlp_init(tc, sm, s.brdf_w0u, s.brdf_w1u, s.brdf_w2u);
  } else {
    mlp_load_weights(s, sm);
    __syncthreads();
    xcol = sm + (MLP_SMEM_FLOATS - 66 * MLP_THREADS) + threadIdx.x;
  }
  for (int i0 = blockIdx.x * MLP_THREADS; i0 < n; i0 += gridDim.x * MLP_THREADS) {
    const int i = i0 + threadIdx.x;
    const bool active = i < n;
    const int ii = active ? i : 0;
    float x[TC_K0];
#pragma unroll
    for (int k = 0; k < 24; ++k) x[k] = feat[(size_t)ii * 24 + k];
    mlp_encode(x, nmf_mk3(half_l[3 * ii], half_l[3 * ii + 1], half_l[3 * ii + 2]),
               nmf_mk3(diff_l[3 * ii], diff_l[3 * ii + 1], diff_l[3 * ii + 2]), rough[ii]);
    float bw[3];
    if (TC) tc_mlp_forward(tc, x, s.brdf_bias, bw);
    else if (active) mlp_simt(sm, xcol, x, s.brdf_bias, bw);
    if (active) { out[3 * i] = bw[0]; out[3 * i + 1] = bw[1]; out[3 * i + 2] = bw[2]; }
  }
  if (TC) tc_mlp_free(tc);
}
extern "C" int nmf_brdf_mlp(const NmfScene* scene, const float* feat, const float* half_local, const float* diff_local,
                            const float* rough, int n, float* out, void* stream) {
  if (!scene || !scene->brdf_w0t || !scene->brdf_w0u || !scene->brdf_w1u || !scene->brdf_w2u || !feat || !half_local || !diff_local || !rough || !out || n <= 0)
    return NMF_E_ARG;
  static bool done = false;
  if (!done) {
    CK(cudaFuncSetAttribute(k_brdf_mlp<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MLP_SMEM_FLOATS * sizeof(float))));
    CK(cudaFuncSetAttribute(k_brdf_mlp<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    done = true;
  }
  if (scene->mlp_mode == 0)
    k_brdf_mlp<1><<<blocks_for(n, MLP_THREADS, 3), MLP_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(*scene, feat, half_local, diff_local, rough, n, out);
  else
    k_brdf_mlp<0><<<blocks_for(n, MLP_THREADS, 3), MLP_THREADS, MLP_SMEM_FLOATS * sizeof(float), (cudaStream_t)stream>>>(*scene, feat, half_local, diff_local, rough, n, out);
  CKL();
  return NMF_OK;
}
__global__ void k_heads(const NmfScene s, const float* feat, int n, float* albedo, float* tint, float* f0, float* r1, float* r2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float lin[11];
  for (int h = 0; h < 11; ++h) {
    float v = s.head_b[h];
    for (int k = 0; k < 24; ++k) v += s.head_w[h * 24 + k] * feat[(size_t)i * 24 + k];
    lin[h] = v;
  }
  for (int c = 0; c < 3; ++c) {
    albedo[3 * i + c] = nmf_clampf(nmf_sigmoid(s.diffuse_mul * lin[c] + s.diffuse_bias), 0.f, 1.f);
    tint[3 * i + c] = nmf_sigmoid(lin[3 + c] + s.tint_bias);
    f0[3 * i + c] = nmf_sigmoid(lin[6 + c] + s.f0_bias);
  }
  r1[i] = nmf_clampf(nmf_sigmoid(lin[9] + s.roughness_bias) / 2.0f, 1e-2f, 1.0f);
  if (r2) r2[i] = nmf_clampf(nmf_sigmoid(lin[10] + s.roughness_bias) / 2.0f, 1e-2f, 1.0f);   // render_modules.py:557-560
}
extern "C" int nmf_material_heads(const NmfScene* scene, const float* feat, int n, float* albedo, float* tint, float* f0,
                                  float* r1, float* r2, void* stream) {
  if (!scene || !scene->head_w || !feat || !albedo || !tint || !f0 || !r1 || n <= 0) return NMF_E_ARG;
  k_heads<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*scene, feat, n, albedo, tint, f0, r1, r2);
  CKL();
  return NMF_OK;
}
// samplers/alphagrid.py:209-247: alpha on the (gz,gy,gx) lattice; lattice point = aabb0*(1-s) + aabb1*s, s = linspace(0,1,g)
__global__ void k_dense_alpha(const NmfScene s, int gx, int gy, int gz, const float* __restrict__ lins, float* alpha) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = t >> 2;
  const int sub = (int)(t & 3);
  const long long total = (long long)gx * gy * gz;
  const bool active = i < total;
  const long long ii = active ? i : 0;
  const int x = (int)(ii % gx), y = (int)((ii / gx) % gy), z = (int)(ii / ((long long)gx * gy));
  // torch.linspace(0, 1, g): start + step*i for the first half, end - step*(g-1-i) for the second
  auto lin = [](int i, int g) {
    const float step = 1.0f / (float)(g - 1);
    return i < g / 2 ? NMF_MUL(step, (float)i) : NMF_SUB(1.0f, NMF_MUL(step, (float)(g - 1 - i)));
  };
  // `lins`: the caller's own torch.linspace(0, 1, g) per axis (x | y | z).  ATen's CPU linspace is vectorised, so its bits
  // depend on the host ISA; a sample that lies exactly on a plane of the current mask's lattice sees one voxel or eight
  // depending on the last bit, so the reference's coordinates are taken as they are rather than re-derived
  const float sx = lins ? lins[x] : lin(x, gx), sy = lins ? lins[gx + y] : lin(y, gy), sz = lins ? lins[gx + gy + z] : lin(z, gz);
  float p[3], xn[3];
  p[0] = NMF_ADD(NMF_MUL(s.aabb0[0], NMF_SUB(1.0f, sx)), NMF_MUL(s.aabb1[0], sx));
  p[1] = NMF_ADD(NMF_MUL(s.aabb0[1], NMF_SUB(1.0f, sy)), NMF_MUL(s.aabb1[1], sy));
  p[2] = NMF_ADD(NMF_MUL(s.aabb0[2], NMF_SUB(1.0f, sz)), NMF_MUL(s.aabb1[2], sz));
  nmf_normalize_xyz(s, p, xn);
  const NmfTaps tp = nmf_vm_taps(s, xn);
  float f = nmf_density_group(s, tp, sub);
  f += __shfl_xor_sync(FULL, f, 1);
  f += __shfl_xor_sync(FULL, f, 2);
  if (active && sub == 0) {
    // alphagrid.py:209-224 compute_alpha: where the CURRENT mask samples to 0 the density is not evaluated (sigma = 0), so
    // a rebuild can only shrink the occupied set (up to the 3^3 dilation that follows)
    const bool masked_out = s.has_occ && !nmf_occupied(s.occ_vox, s.occ_cell, s.ow, s.oh, s.od, s.opitch, xn[0], xn[1], xn[2]);
    const float sigma = masked_out ? 0.f : nmf_feature2density(f, s.density_shift);
    alpha[i] = 1.0f - expf(-sigma * s.stepsize);                         // alphagrid.py:222 (no distance_scale)
  }
}
extern "C" int nmf_dense_alpha(const NmfScene* scene, int gx, int gy, int gz, const float* lins, float* alpha, void* stream) {
  int st = check_scene(scene);
  if (st) return st;
  if (!alpha || gx < 2 || gy < 2 || gz < 2) return NMF_E_ARG;
  const long long total = (long long)gx * gy * gz * 4;
  k_dense_alpha<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*scene, gx, gy, gz, lins, alpha);
  CKL();
  return NMF_OK;
}

// ================================================================================================
// Callers either side of the path
// ================================================================================================
// dataLoader/ray_utils.py:23-43 (get_ray_directions), dataLoader/blender.py:97-120,170-173 (normalise, get_rays):
// the rays of one view are generated on the device from (pose, intrinsics) instead of being stored as (N,6) tensors
struct RayGenArgs { float R[9], T[3]; int W; float fx, fy, cx, cy; const int* pixel_ids; int n; float* rays; };
__global__ void k_generate_rays(const RayGenArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int pix = a.pixel_ids ? a.pixel_ids[i] : i;
  const float px = (float)(pix % a.W) + 0.5f, py = (float)(pix / a.W) + 0.5f;            // create_meshgrid + 0.5
  const float dx = NMF_DIV(NMF_SUB(px, a.cx), a.fx), dy = NMF_DIV(NMF_SUB(py, a.cy), a.fy), dz = 1.0f;
  const float nrm = sqrtf(NMF_ADD(NMF_ADD(NMF_MUL(dx, dx), NMF_MUL(dy, dy)), NMF_MUL(dz, dz)));   // blender.py:108-110
  const float ux = NMF_DIV(dx, nrm), uy = NMF_DIV(dy, nrm), uz = NMF_DIV(dz, nrm);
  float* o = a.rays + (size_t)i * 6;
  o[0] = a.T[0]; o[1] = a.T[1]; o[2] = a.T[2];                                                   // ray_utils.py:84
#pragma unroll
  for (int r = 0; r < 3; ++r) o[3 + r] = ux * a.R[3 * r] + uy * a.R[3 * r + 1] + uz * a.R[3 * r + 2];   // directions @ c2w[:3,:3].T
}
extern "C" int nmf_generate_rays(const float* c2w_host, int H, int W, float fx, float fy, float cx, float cy, const int* pixel_ids,
                                 int n, float* rays, void* stream) {
  if (!c2w_host || !rays || H <= 0 || W <= 0 || n <= 0 || !(fx > 0.f) || !(fy > 0.f)) return NMF_E_ARG;
  RayGenArgs a;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) a.R[3 * r + c] = c2w_host[4 * r + c];
    a.T[r] = c2w_host[4 * r + 3];
  }
  a.W = W; a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.pixel_ids = pixel_ids; a.n = n; a.rays = rays;
  k_generate_rays<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
  CKL();
  return NMF_OK;
}

// renderer.py:399-401: squared error of the 8-bit quantised render against the ground truth, summed in fp64 on the
// device (one 8-byte read-back per image instead of the image itself)
__global__ void k_image_sq_error(const float* rgb, const float* gt, const int* pixel_ids, int n, double* out) {
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t g = (size_t)(pixel_ids ? pixel_ids[i] : i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float q = floorf(nmf_clampf(rgb[(size_t)i * 3 + c], 0.f, 1.f) * 255.0f) / 255.0f;
      const float d = q - nmf_clampf(gt[g + c], 0.f, 1.f);
      acc += (double)(d * d);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(FULL, acc, off);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
extern "C" int nmf_image_sq_error(const float* rgb, const float* gt, const int* pixel_ids, int n, double* sum_sq, void* stream) {
  if (!rgb || !gt || !sum_sq || n <= 0) return NMF_E_ARG;
  CK(cudaMemsetAsync(sum_sq, 0, sizeof(double), (cudaStream_t)stream));
  k_image_sq_error<<<sm_count() * 4, 256, 0, (cudaStream_t)stream>>>(rgb, gt, pixel_ids, n, sum_sq);
  CKL();
  return NMF_OK;
}
