// BRDF MLP 66 -> 64 -> 64 -> 4 (modules/brdf.py:73-120, 237-239) REVERSE pass on the 5th-generation tensor cores.
//
// One CTA = 128 threads = 128 bounce rays per tile; thread t owns ray t = TMEM lane t.  Per tile the forward is recomputed
// (layers 1, 2) and the backward runs as tcgen05.mma.kind::f16 tile GEMMs with BF16 operands (gradients span 1e-9 .. 1e3 with
// an HDR environment: they need fp32's exponent range, not fp16's) and fp32 accumulators in TMEM:
//   forward    D = X  W0^T, relu -> H1;   D = H1 W1^T, relu -> H2                                   (K-major operands, M=128)
//   d H2       D = dOut W2      (B = the layer-3 weight tile read MN-major), masked by H2 > 0
//   d H1       D = dH2  W1      (B = the layer-2 weight tile read MN-major), masked by H1 > 0
//   d X        D = dH1  W0[:, :32]   (features are inputs 0..23; the encodings reach the MLP detached)
//   d W2^T += H2^T dOut,  d W1^T += H1^T dH2,  d W0^T += X^T dH1    (A and B both read MN-major: the contraction runs over the
//             128 rays of the tile; M = 128 of which 80 rows are real)
// The three weight-gradient accumulators STAY IN TMEM across all tiles a CTA processes (accumulate flag), so the tile loop has
// no weight-gradient epilogue at all: one tcgen05.ld pass and one atomic per weight per CTA at the very end.  The biases ride
// along as the constant-1 input (column 66 of every activation tile), exactly as in the forward (csrc/nmf_mlp_tc.cuh).
//
// Operand layout: canonical no-swizzle core matrices (8 rows x 16 bytes).  An activation tile T[128 rays][80] is stored
//   element (ray r, column c) at byte (c/8) * 2048 + r * 16 + (c%8) * 2
// which is K-major for "T as A (M = ray, K = column)"  [SBO = 128 between 8-ray groups, LBO = 2048 between 16-byte K chunks]
// and at the same time MN-major for "T^T as A or T as B with the rays as K"  [(mn, k) at (mn/8) SBO' + (k/8) LBO' + (k%8) 16 +
// (mn%8) 2 with SBO' = 2048 between 8-column groups, LBO' = 128 between 8-ray groups].  Weight tiles W[rows][80] (the
// forward's B operands, [K/8][rows][8]) are K-major for the forward and MN-major (mn = input column) for the d X / d H GEMMs.
#pragma once
#include "nmf_mlp_tc.cuh"

#define TB_ACT_BYTES (TC_KC * TC_ROWS * 16)            // 20480: one [128][80] bf16 activation tile
#define TB_G_BYTES (8 * TC_ROWS * 16)                  // 16384: one [128][64] bf16 gradient tile
#define TB_O_BYTES (2 * TC_ROWS * 16)                  // 4096:  the [128][16] output-gradient tile
#define TB_OFF_X 0
#define TB_OFF_H1 (TB_OFF_X + TB_ACT_BYTES)
#define TB_OFF_H2 (TB_OFF_H1 + TB_ACT_BYTES)
#define TB_OFF_G2 (TB_OFF_H2 + TB_ACT_BYTES)           // d H2
#define TB_OFF_G1 (TB_OFF_G2 + TB_G_BYTES)             // d H1
#define TB_OFF_GO (TB_OFF_G1 + TB_G_BYTES)             // d Out
#define TB_OFF_W0 (TB_OFF_GO + TB_O_BYTES)
#define TB_OFF_W1 (TB_OFF_W0 + TC_W_BYTES)
#define TB_OFF_W2 (TB_OFF_W1 + TC_W_BYTES)
#define TB_OFF_BAR (TB_OFF_W2 + TC_W2_BYTES)
#define TB_OFF_TMEM (TB_OFF_BAR + 8)
#define TB_SMEM_BYTES (TB_OFF_TMEM + 8)                // (M = 128 reads of an 80-row MN-major tile run 12 KB past it: into the next tile, rows ignored)
#define TB_TMEM_COLS 256
#define TB_COL_D 0          // scratch accumulator of the per-tile GEMMs (64 columns)
#define TB_COL_W0 64        // d W0^T accumulator (64 columns), rows = input index (66 = bias)
#define TB_COL_W1 128       // d W1^T
#define TB_COL_W2 192       // d W2^T (16 columns)

// kind::f16 instruction descriptor: D = fp32, A / B format 0 = F16, 1 = BF16 (bits 7..9 / 10..12), a_major bit 15, b_major
// bit 16 (1 = MN-major), N >> 3 at bit 17, M >> 4 at bit 24
#define TB_IDESC(N, AMN, BMN, AF, BF) ((1u << 4) | ((uint32_t)(AF) << 7) | ((uint32_t)(BF) << 10) | ((uint32_t)(AMN) << 15) | \
                                       ((uint32_t)(BMN) << 16) | (((uint32_t)(N) >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ uint32_t tb_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t tb_pack_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// bf16 pair -> two floats (a bf16 is the upper half of an fp32)
__device__ __forceinline__ float tb_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float tb_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

struct TbMlp {
  char* sm;
  uint32_t tmem, phase;
};

// all 128 threads, once per CTA.  w0b / w1b / w2b: BF16 weight tiles in the forward's operand layout ([K/8][rows][8], biases
// in column 66, K zero-padded to 80; scene.py)
__device__ __forceinline__ void tb_init(TbMlp& c, void* sm_, const void* w0b, const void* w1b, const void* w2b) {
  char* sm = (char*)sm_;
  c.sm = sm;
  const int tid = threadIdx.x;
  const uint32_t bar = tc_smem_u32(sm + TB_OFF_BAR);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(2 * TC_W_BYTES + TC_W2_BYTES))
                 : "memory");
    tc_bulk_load(tc_smem_u32(sm + TB_OFF_W0), w0b, TC_W_BYTES, bar);
    tc_bulk_load(tc_smem_u32(sm + TB_OFF_W1), w1b, TC_W_BYTES, bar);
    tc_bulk_load(tc_smem_u32(sm + TB_OFF_W2), w2b, TC_W2_BYTES, bar);
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(sm + TB_OFF_TMEM)),
                 "r"((uint32_t)TB_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tc_wait(bar, 0);
  c.phase = 1;
  c.tmem = *(volatile uint32_t*)(sm + TB_OFF_TMEM);
}
__device__ __forceinline__ void tb_free(TbMlp& c) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"((uint32_t)TB_TMEM_COLS) : "memory");
}

// everything the MMAs of one step read was written by all threads before this call
__device__ __forceinline__ void tb_publish() {
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
}
__device__ __forceinline__ void tb_wait(TbMlp& c) {
  tc_wait(tc_smem_u32(c.sm + TB_OFF_BAR), c.phase);
  c.phase ^= 1;
  tc_fence_after();
}

// D[128 x n] (+)= A[128 x 16 kchunks...] * B^T, both K-major: activation tile at a_off (K chunks k0 .. k0 + 2 nk), weight tile at
// w_off with `rows` rows
// MIXED = 1: activations and weights are the forward's FP16 tiles (the recomputed forward is then bit-identical to
// k_bounce's, and the weights keep 11 mantissa bits), gradients BF16 (fp32's exponent range); MIXED = 0: everything BF16.
template <int MIXED>
__device__ __forceinline__ void tb_gemm_kk(const TbMlp& c, uint32_t col, uint32_t a_off, uint32_t w_off, uint32_t rows, int nk2) {
  const uint32_t a = tc_smem_u32(c.sm + a_off), w = tc_smem_u32(c.sm + w_off);
  for (int k = 0; k < nk2; ++k)
    tc_mma(c.tmem + col, tc_desc(a + k * 2 * (TC_ROWS * 16), TC_ROWS * 16, 128), tc_desc(w + k * 2 * (rows * 16), rows * 16, 128),
           TB_IDESC(rows, 0, 0, MIXED ? 0 : 1, MIXED ? 0 : 1), k > 0);
}
// D[128 x n] = G[128 x K] * W  with W = a forward weight tile [rows = K][80] read MN-major (mn = input column 0 .. n-1)
template <int MIXED>
__device__ __forceinline__ void tb_gemm_data(const TbMlp& c, uint32_t col, uint32_t g_off, uint32_t w_off, uint32_t rows, uint32_t n,
                                             int nk2) {
  const uint32_t g = tc_smem_u32(c.sm + g_off), w = tc_smem_u32(c.sm + w_off);
  // weight tile: (mn = column, k = row) at (mn/8) * rows*16 + (k/8) * 128 + (k%8) * 16: SBO = rows * 16, LBO = 128
  for (int k = 0; k < nk2; ++k)
    tc_mma(c.tmem + col, tc_desc(g + k * 2 * (TC_ROWS * 16), TC_ROWS * 16, 128), tc_desc(w + k * 256, 128, rows * 16),
           TB_IDESC(n, 0, 1, 1, MIXED ? 0 : 1), k > 0);
}
// D[128 (80 real) x n] += T^T * G over the 128 rays: T = activation tile at t_off (mn = its column), G = gradient tile at g_off
// (mn = its column 0 .. n-1), both MN-major with SBO = 2048 (8-column groups), LBO = 128 (8-ray groups); 8 MMAs of K = 16 rays
template <int MIXED>
__device__ __forceinline__ void tb_gemm_wgrad(const TbMlp& c, uint32_t col, uint32_t t_off, uint32_t g_off, uint32_t n, bool first) {
  const uint32_t t = tc_smem_u32(c.sm + t_off), g = tc_smem_u32(c.sm + g_off);
  for (int k = 0; k < 8; ++k)
    tc_mma(c.tmem + col, tc_desc(t + k * 256, 128, TC_ROWS * 16), tc_desc(g + k * 256, 128, TC_ROWS * 16), TB_IDESC(n, 1, 1, MIXED ? 0 : 1, 1),
           !(first && k == 0));
}
