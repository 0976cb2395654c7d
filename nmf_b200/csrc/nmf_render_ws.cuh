// Scratch layout of one nmf_render_rays call ('Per-call scratch'): the records the phases hand over
// (survivors, bounce samples, bounce rays), the per-ray accumulators and the carve() that lays them out in the caller's
// workspace.  Shared by csrc/nmf_kernels.cu (forward) and csrc/nmf_mf_train.cu (microfacet reverse pass).
#pragma once
#include <cuda_runtime.h>
#include "nmf_field.cuh"

#define FULL 0xffffffffu
#define NMF_BRAY_CAP_PER_RAY 160   // bounce rays per primary ray a chunk region can hold (typical: 57)
#define NMF_SURV0_PER_RAY 64       // surviving samples per primary ray (typical: 10-20)
#define NMF_BS0_PER_RAY 24         // bounce samples per primary ray (typical: 12)
#define NMF_SURV1_PER_RAY 256      // per retraced ray
#define NMF_BS1_PER_RAY 128
#define NMF_TIE_CAP 64            // exact score ties at a chunk's retrace threshold that are ordered by ray key
#define NMF_BGRAD 36              // floats per bounce-sample gradient record (reverse pass)
#define NMF_NO_OWNER 0xFFFFFFFFu  // ray -> sample map entry of a ray whose sample could not be allocated (overflow)

struct Surv { uint32_t ray; uint32_t step; float w; };

// 320 bytes = ten 32-byte pairs, 32-byte aligned: k_shade writes a record with ten 256-bit stores and k_bounce reads it with
// nine 256-bit loads (pairs: pos|V, N|f0, diffuse|fresn, key..|feat0, feat1|feat2, feat3|feat4, feat5|frame0, frame1|frame2,
// frame3|frame4, frame5|tail) -- half the load / store instructions per lane, i.e. half the L1 wavefronts of the record traffic
struct __align__(32) BSample {
  float pos[3]; float w;
  float V[3]; float rough;
  float N[3]; int count;
  float f0[3]; uint32_t ray;
  float diffuse[3]; uint32_t roff;
  float fresn[3]; uint32_t flags;
  uint64_t key; uint32_t chunk; uint32_t pad;    // pad: train mode = index of the sample in the level's valid-sample list (VSmp)
  float feat[24];
  // what GGX sampling and the ISH encodings need per SAMPLE, computed once in k_shade instead of once per bounce ray:
  // frame[0..17] = t, b, V_l, Vs, T1, T2 (nmf_ggx_frame), [18] = a, [19..20] = ISH scales s1, s2, [21..22] = the
  // per-sample Sobol offsets 0.25 * U (brdf_samplers/base.py:16-19)
  float frame[24];
  float tail[4];      // padding to 320 bytes
};
static_assert(sizeof(BSample) == 320, "BSample is ten 32-byte pairs");

// train mode: one record per VALID sample of a kept ray (also the zero-weight ones: the compositing backward needs them all)
struct __align__(16) VSmp { uint32_t k; float f; float alpha; float T; };

struct __align__(32) BRay {     // bounce ray record handed from k_bounce to k_incoming: one 256-bit store, one 256-bit load
  float L[3]; float mip;
  float bw[3]; int slot;        // slot: index of the secondary ray that re-traces it, -1 = environment
};

// per-ray accumulators of level 0 (floats)
// A_ORI / A_TINTU feed the A19 statistics: sum w * min(v.n, 0)^2 and the UNWEIGHTED sum of the per-sample tint
enum { A_RGB = 0, A_WN = 3, A_CROSS = 6, A_DIFF = 9, A_TINT = 12, A_SPEC = 15, A_ALB = 18, A_ROUGH = 21, A_ORI = 22, A_TINTU = 23, A_N = 24 };

struct WS {
  // counters (zeroed every call)
  int* n_surv;        // [2]
  int* n_bs;          // [2]
  int* ray_count0;    // [n_chunks]
  int* ray_count1;    // [n_chunks]
  int* n_samples0;    // [n_chunks]
  int* n_samples1;    // [n_chunks]
  int* n_cand;        // [n_chunks]
  int* n_sec;         // [n_chunks]
  unsigned long long* score_sum;   // [n_chunks] sum of the retrace scores in 2^-32 fixed point (order-independent)
  double* wsum1;      // [n_chunks]
  unsigned* error;    // [1]
  float* stat4;       // [n_chunks][4] A19 statistics: sum of w*min(v.n,0)^2, of the diffuse map, of the sample tints, of acc
  int* tile_start0;   // [n_chunks + 1]
  int* tile_start1;   // [n_chunks + 1]
  int2* tile_desc0;   // [n_chunks * ceil(cap_rays0 / 128)] (chunk, first ray) of every 128-ray tile (k_tile_prefix)
  int2* tile_desc1;
  size_t counters_bytes;
  char* counters_base;
  // level 0
  float* tmin0; float* acc0; float* depth0; int* termk0; int* nvalid0; float* accum0;   // accum0 [n_rays][A_N]
  float4* red0;     // [cap_bs0][2]: {w, count, ray, flags} and the sum of the sample's combined bounce radiance (k_incoming -> k_reduce0)
  Surv* surv0; BSample* bs0; BRay* brays0; uint32_t* owner0; float2* scu0;   // scu0: (retrace score, tie-break U) per ray
  // level 1
  float* rays1; float* mip1; uint64_t* key1; float* tmin1; float* acc1; int* nvalid1; float* accum1; float* rgb1;
  Surv* surv1; BSample* bs1; BRay* brays1; uint32_t* owner1;
  // train mode (nmf_render_rays_train): jittered distances per dense step, dynamic batch truncation
  float* zvals0; float* zvals1; uint8_t* whole0; int* n_kept;
  // train mode, kept for the reverse pass (csrc/nmf_mf_train.cu): every valid sample of both levels in march order
  // (vs: step, density feature, alpha, transmittance; vdw: d loss / d weight, accumulated by the reverse kernels), the
  // ray -> first-sample map, survivor -> valid-sample / bounce-sample maps, and the reverse pass's own scratch
  int* n_vs;          // [2] (zeroed with the counters)
  VSmp* vs0; VSmp* vs1; float* vdw0; float* vdw1; int* vbase0; int* vbase1; int cap_vs0, cap_vs1;
  uint32_t* survv0; uint32_t* survv1; int* survslot0; int* survslot1;
  float* g_lin0;      // [n_rays][4]   d loss / d (linear rgb, acc) of the primary rays
  float* g_lin1;      // [n_rays1][4]  d loss / d (rgb1, acc1) of the re-traced rays
  float* jac1;        // [n_rays1][12] d rgb1 / d direction (9 used)
  float* bgrad0; float* bgrad1;   // [cap_bs][NMF_BGRAD] per bounce sample: dR0 3 | ddiffuse 3 | drough | dNf 3 | dnfeat 24
  float4* dout0; float4* dout1;   // per bounce ray: d loss / d (BRDF MLP pre-sigmoid outputs)
  int n_chunks, n_rays1;
  int cap_surv0, cap_bs0, cap_rays0;   // cap_rays0: per chunk
  int cap_surv1, cap_bs1, cap_rays1;   // cap_rays1: per chunk
  size_t total;
};

#define MLP_THREADS 128            // bounce rays per tile (one tcgen05 M = 128 tile / one thread per ray)
static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static inline void carve(WS& w, const NmfScene* s, int n_rays, int chunk, char* base, float cap_scale = 1.0f, bool train = false) {
  const double cs = cap_scale > 0.f ? (double)cap_scale : 1.0;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  int nc = (n_rays + chunk - 1) / chunk;
  int maxre = s->model == 0 ? s->max_retrace : 0;
  w.n_chunks = nc;
  w.n_rays1 = nc * maxre;
  auto cap = [&](double items) { double v = items * cs; return (int)(v < 2.0e9 ? v : 2.0e9); };
  w.cap_surv0 = cap((double)n_rays * NMF_SURV0_PER_RAY);
  w.cap_bs0 = cap((double)n_rays * NMF_BS0_PER_RAY);
  w.cap_rays0 = cap((double)chunk * NMF_BRAY_CAP_PER_RAY);
  w.cap_surv1 = cap((double)w.n_rays1 * NMF_SURV1_PER_RAY);
  w.cap_bs1 = cap((double)w.n_rays1 * NMF_BS1_PER_RAY);
  w.cap_rays1 = maxre > 0 ? s->max_brdf_rays1 + 1024 : 0;
  w.counters_base = take(0);
  w.n_surv = (int*)take(2 * sizeof(int));
  w.n_bs = (int*)take(2 * sizeof(int));
  w.ray_count0 = (int*)take(nc * sizeof(int));
  w.ray_count1 = (int*)take(nc * sizeof(int));
  w.n_samples0 = (int*)take(nc * sizeof(int));
  w.n_samples1 = (int*)take(nc * sizeof(int));
  w.n_cand = (int*)take(nc * sizeof(int));
  w.n_sec = (int*)take(nc * sizeof(int));
  w.score_sum = (unsigned long long*)take(nc * sizeof(unsigned long long));
  w.wsum1 = (double*)take(nc * sizeof(double));
  w.error = (unsigned*)take(sizeof(unsigned));
  w.stat4 = (float*)take((size_t)nc * 4 * sizeof(float));
  w.tile_start0 = (int*)take((nc + 1) * sizeof(int));
  w.tile_start1 = (int*)take((nc + 1) * sizeof(int));
  w.accum0 = (float*)take((size_t)n_rays * A_N * sizeof(float));
  w.accum1 = (float*)take((size_t)w.n_rays1 * 4 * sizeof(float));
  w.n_vs = (int*)take(2 * sizeof(int));
  w.counters_bytes = off;   // everything up to here is zeroed at the start of a call
  w.tmin0 = (float*)take((size_t)n_rays * 4);
  w.acc0 = (float*)take((size_t)n_rays * 4);
  w.depth0 = (float*)take((size_t)n_rays * 4);
  w.termk0 = (int*)take((size_t)n_rays * 4);
  w.nvalid0 = (int*)take((size_t)n_rays * 4);
  w.surv0 = (Surv*)take((size_t)w.cap_surv0 * sizeof(Surv));
  if (s->model == 0) {
    w.bs0 = (BSample*)take((size_t)w.cap_bs0 * sizeof(BSample));
    w.red0 = (float4*)take((size_t)w.cap_bs0 * 2 * sizeof(float4));
    w.brays0 = (BRay*)take((size_t)nc * w.cap_rays0 * sizeof(BRay));
    w.owner0 = (uint32_t*)take((size_t)nc * w.cap_rays0 * 4);
    w.scu0 = (float2*)take((size_t)nc * w.cap_rays0 * sizeof(float2));
    w.rays1 = (float*)take((size_t)w.n_rays1 * 6 * 4);
    w.mip1 = (float*)take((size_t)w.n_rays1 * 4);
    w.key1 = (uint64_t*)take((size_t)w.n_rays1 * 8);
    w.tmin1 = (float*)take((size_t)w.n_rays1 * 4);
    w.acc1 = (float*)take((size_t)w.n_rays1 * 4);
    w.nvalid1 = (int*)take((size_t)w.n_rays1 * 4);
    w.rgb1 = (float*)take((size_t)w.n_rays1 * 4 * 4);
    w.surv1 = (Surv*)take((size_t)w.cap_surv1 * sizeof(Surv));
    w.bs1 = (BSample*)take((size_t)w.cap_bs1 * sizeof(BSample));
    w.brays1 = (BRay*)take((size_t)nc * w.cap_rays1 * sizeof(BRay));
    w.owner1 = (uint32_t*)take((size_t)nc * w.cap_rays1 * 4);
    w.tile_desc0 = (int2*)take((size_t)nc * ((w.cap_rays0 + MLP_THREADS - 1) / MLP_THREADS) * sizeof(int2));
    w.tile_desc1 = (int2*)take((size_t)nc * ((w.cap_rays1 + MLP_THREADS - 1) / MLP_THREADS) * sizeof(int2));
  }
  w.zvals0 = w.zvals1 = nullptr; w.whole0 = nullptr; w.n_kept = nullptr;
  w.vs0 = w.vs1 = nullptr; w.vdw0 = w.vdw1 = nullptr; w.vbase0 = w.vbase1 = nullptr; w.cap_vs0 = w.cap_vs1 = 0;
  w.survv0 = w.survv1 = nullptr; w.survslot0 = w.survslot1 = nullptr; w.g_lin0 = w.g_lin1 = w.jac1 = nullptr;
  w.bgrad0 = w.bgrad1 = nullptr; w.dout0 = w.dout1 = nullptr;
  if (train) {
    w.zvals0 = (float*)take((size_t)n_rays * s->n_steps * 4);
    w.zvals1 = (float*)take((size_t)w.n_rays1 * s->n_steps * 4);
    w.whole0 = (uint8_t*)take((size_t)n_rays);
    w.n_kept = (int*)take(2 * sizeof(int));
    w.cap_vs0 = cap((double)n_rays * 4 * NMF_SURV0_PER_RAY);
    w.cap_vs1 = cap((double)w.n_rays1 * 2 * NMF_SURV1_PER_RAY);
    if ((double)w.cap_vs0 > (double)n_rays * s->n_steps) w.cap_vs0 = (int)((double)n_rays * s->n_steps);
    if ((double)w.cap_vs1 > (double)w.n_rays1 * s->n_steps) w.cap_vs1 = (int)((double)w.n_rays1 * s->n_steps);
    w.vs0 = (VSmp*)take((size_t)w.cap_vs0 * sizeof(VSmp));
    w.vs1 = (VSmp*)take((size_t)w.cap_vs1 * sizeof(VSmp));
    w.vdw0 = (float*)take((size_t)w.cap_vs0 * 4);
    w.vdw1 = (float*)take((size_t)w.cap_vs1 * 4);
    w.vbase0 = (int*)take((size_t)n_rays * 4);
    w.vbase1 = (int*)take((size_t)w.n_rays1 * 4);
    w.survv0 = (uint32_t*)take((size_t)w.cap_surv0 * 4);
    w.survv1 = (uint32_t*)take((size_t)w.cap_surv1 * 4);
    w.survslot0 = (int*)take((size_t)w.cap_surv0 * 4);
    w.survslot1 = (int*)take((size_t)w.cap_surv1 * 4);
    w.g_lin0 = (float*)take((size_t)n_rays * 16);
    w.g_lin1 = (float*)take((size_t)w.n_rays1 * 16);
    w.jac1 = (float*)take((size_t)w.n_rays1 * 12 * 4);
    if (s->model == 0) {
      w.bgrad0 = (float*)take((size_t)w.cap_bs0 * NMF_BGRAD * 4);
      w.bgrad1 = (float*)take((size_t)w.cap_bs1 * NMF_BGRAD * 4);
      w.dout0 = (float4*)take((size_t)nc * w.cap_rays0 * sizeof(float4));
      w.dout1 = (float4*)take((size_t)nc * w.cap_rays1 * sizeof(float4));
    }
  }
  w.total = off;
}

#ifdef __CUDACC__
// Segments = runs of lanes that share `key` (bounce rays of one sample are consecutive).  seg_setup finds, with one
// ballot, whether this lane starts a run and the last lane of its run; seg_sum3 then leaves the run's total in its
// first lane using value shuffles only.  All 32 lanes must call both.
struct Seg { bool head; int last; };
__device__ __forceinline__ Seg seg_setup(uint32_t key, int lane) {
  Seg g;
  const uint32_t prev = __shfl_up_sync(FULL, key, 1);
  g.head = lane == 0 || prev != key;
  const unsigned heads = __ballot_sync(FULL, g.head);
  const unsigned above = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u));
  g.last = above ? __ffs(above) - 2 : 31;
  return g;
}
__device__ __forceinline__ void seg_sum3(float (&v)[3], const Seg& g, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float a0 = __shfl_down_sync(FULL, v[0], off), a1 = __shfl_down_sync(FULL, v[1], off), a2 = __shfl_down_sync(FULL, v[2], off);
    if (lane + off <= g.last) { v[0] += a0; v[1] += a1; v[2] += a2; }
  }
}

#endif
