// BRDF MLP 66 -> 64 -> 64 -> 4 (modules/brdf.py:73-120, 237-239) on the 5th-generation tensor cores.
//
// One CTA = 128 threads = 128 MLP rows (bounce rays); thread t owns row t = TMEM lane t.  All three layers are
// tcgen05.mma.kind::f16 (fp16 operands = the 10-bit mantissa of TF32, fp32 accumulate in TMEM), K padded to 80:
//   layer 1:  D[128x64] = X[128x80] * W0p^T      5 MMAs (M=128, N=64, K=16); x[66] = 1 carries the bias b0
//   layer 2:  D[128x64] = [relu(D) | x64.. ] * W1p^T   5 MMAs; the K chunks 8,9 of X are kept, W1p[:, 66] = b1
//   layer 3:  D[128x16] = [relu(D) | x64.. ] * W2p^T   5 MMAs (N=16); rows 0..2 of W2p are real, W2p[:, 66] = b2
// so no bias add, no fp32 last layer and no weight reads in the epilogues: each epilogue is tcgen05.ld ->
// cvt.rn.relu.f16x2.f32 -> 8 x st.shared.v4 per row.
// Operands live in shared memory in the canonical no-swizzle K-major layout (8-row x 16-byte core matrices):
//   element (row r, k) of an R-row operand at byte  (k/8) * (R*16) + r*16 + (k%8)*2
//   => stride between 8-row groups SBO = 128 B, between 16-byte K chunks LBO = R*16 B.
// Shared memory per CTA: 20 KB (X) + 10 + 10 + 2.5 KB (weights) = 43.5 KB -> 5 CTAs per SM; TMEM: 64 columns.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_ROWS 128
#define TC_K0 80                       // 66 inputs + the constant 1, padded to a multiple of 16
#define TC_KC 10                       // 16-byte K chunks (8 halves each)
#define TC_ONE 66                      // index of the constant-1 input that carries the biases
#define TC_A_BYTES (TC_KC * TC_ROWS * 16)
#define TC_W_BYTES (TC_KC * 64 * 16)   // layers 1, 2: 64 output rows
#define TC_W2_BYTES (TC_KC * 16 * 16)  // layer 3: 16 output rows (3 real)
#define TC_OFF_A 0
#define TC_OFF_W0 (TC_OFF_A + TC_A_BYTES)
#define TC_OFF_W1 (TC_OFF_W0 + TC_W_BYTES)
#define TC_OFF_W2 (TC_OFF_W1 + TC_W_BYTES)
#define TC_OFF_BAR (TC_OFF_W2 + TC_W2_BYTES)
#define TC_OFF_TMEM (TC_OFF_BAR + 8)
#define TC_SMEM_BYTES (TC_OFF_TMEM + 8)
#define TC_TMEM_COLS 64

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// {lo, hi} -> packed f16x2 (lo in bits 15:0 = the lower K index)
__device__ __forceinline__ uint32_t tc_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t tc_pack_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// K-major, no swizzle: LBO = byte distance between the two 16-byte K chunks of one MMA, SBO = between 8-row groups
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::f16: D = fp32 (bit 4), A = B = F16 (format 0), both K-major, M = 128, N in bits 17..22 (N >> 3)
#define TC_IDESC(N) ((1u << 4) | (((uint32_t)(N) >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tTC_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_WAIT_DONE;\n\tbra TC_WAIT_LOOP;\n\tTC_WAIT_DONE:\n\t}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 accumulator columns of this thread's TMEM lane; `taddr` carries the warp's lane offset and the first column
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcMlp {
  char* sm;           // shared memory base (TC_SMEM_BYTES, 128-byte aligned)
  uint32_t tmem;      // TMEM base address of this CTA's 64 columns
  uint32_t phase;     // mbarrier parity of the next completion
};

// Called by all 128 threads once per CTA.  w0u / w1u / w2u: fp16 weights already in the canonical operand layout
// (scene.py: [K/8][rows][8] halves, biases folded into column TC_ONE, K zero-padded to 80), 16-byte aligned.
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tc_bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mlp_init(TcMlp& c, void* sm_, const void* w0u, const void* w1u, const void* w2u) {
  char* sm = (char*)sm_;
  c.sm = sm;
  const int tid = threadIdx.x;
  const uint32_t bar = tc_smem_u32(sm + TC_OFF_BAR);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the three weight tiles arrive by TMA (no register staging, written through the async proxy the MMAs read with);
    // they complete phase 0 of the barrier that later counts the MMA commits
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(2 * TC_W_BYTES + TC_W2_BYTES))
                 : "memory");
    tc_bulk_load(tc_smem_u32(sm + TC_OFF_W0), w0u, TC_W_BYTES, bar);
    tc_bulk_load(tc_smem_u32(sm + TC_OFF_W1), w1u, TC_W_BYTES, bar);
    tc_bulk_load(tc_smem_u32(sm + TC_OFF_W2), w2u, TC_W2_BYTES, bar);
  }
  if (tid < 32) {   // one warp allocates the accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(sm + TC_OFF_TMEM)),
                 "r"((uint32_t)TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();        // barrier initialised and TMEM address published before anyone waits / reads
  tc_fence_after();
  tc_wait(bar, 0);        // weights have landed
  c.phase = 1;
  c.tmem = *(volatile uint32_t*)(sm + TC_OFF_TMEM);
}
__device__ __forceinline__ void tc_mlp_free(TcMlp& c) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"((uint32_t)TC_TMEM_COLS) : "memory");
}

// one layer: 5 MMAs over the K chunks of X and of the weight tile at `w_off` with `n` output rows, then commit
__device__ __forceinline__ void tc_layer(const TcMlp& c, uint32_t w_off, uint32_t n) {
  const uint32_t a = tc_smem_u32(c.sm + TC_OFF_A), w = tc_smem_u32(c.sm + w_off);
#pragma unroll
  for (int k = 0; k < TC_KC / 2; ++k)
    tc_mma(c.tmem, tc_desc(a + k * 2 * (TC_ROWS * 16), TC_ROWS * 16, 128), tc_desc(w + k * 2 * (n * 16), n * 16, 128),
           TC_IDESC(n), k > 0);
  tc_commit(tc_smem_u32(c.sm + TC_OFF_BAR));
}
// hidden-layer epilogue: relu(D) -> fp16 -> K chunks 0..7 of this thread's row of X
__device__ __forceinline__ void tc_hidden_to_a(const TcMlp& c, uint32_t taddr, uint4* arow) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float h[32];
    tc_ld32(taddr + 32 * half, h);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 v;
      v.x = tc_pack_relu(h[8 * q], h[8 * q + 1]);
      v.y = tc_pack_relu(h[8 * q + 2], h[8 * q + 3]);
      v.z = tc_pack_relu(h[8 * q + 4], h[8 * q + 5]);
      v.w = tc_pack_relu(h[8 * q + 6], h[8 * q + 7]);
      arow[(4 * half + q) * TC_ROWS] = v;
    }
  }
}

// All 128 threads call this together.  x[80]: this row's inputs; x[66] must be 1 and x[67..79] must be 0.
// out3: sigmoid(mlp(x)[:3] + brdf_bias)
__device__ __forceinline__ void tc_mlp_forward(TcMlp& c, const float (&x)[TC_K0], float brdf_bias, float* out3) {
  const int tid = threadIdx.x;
  const uint32_t bar = tc_smem_u32(c.sm + TC_OFF_BAR);
  const uint32_t taddr = c.tmem + ((uint32_t)(tid & ~31) << 16);     // lane offset of this warp in bits 31:16
  uint4* arow = (uint4*)(c.sm + TC_OFF_A) + tid;                      // + kchunk * 128
#pragma unroll
  for (int kc = 0; kc < TC_KC; ++kc) {
    uint4 v;
    v.x = tc_pack(x[8 * kc], x[8 * kc + 1]);
    v.y = tc_pack(x[8 * kc + 2], x[8 * kc + 3]);
    v.z = tc_pack(x[8 * kc + 4], x[8 * kc + 5]);
    v.w = tc_pack(x[8 * kc + 6], x[8 * kc + 7]);
    arow[kc * TC_ROWS] = v;
  }
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    tc_layer(c, TC_OFF_W0, 64);
  }
  tc_wait(bar, c.phase);
  c.phase ^= 1;
  tc_fence_after();
  tc_hidden_to_a(c, taddr, arow);
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    tc_layer(c, TC_OFF_W1, 64);
  }
  tc_wait(bar, c.phase);
  c.phase ^= 1;
  tc_fence_after();
  tc_hidden_to_a(c, taddr, arow);
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    tc_layer(c, TC_OFF_W2, 16);
  }
  tc_wait(bar, c.phase);
  c.phase ^= 1;
  tc_fence_after();
  float o[4];
  tc_ld4(taddr, o);
  out3[0] = 1.0f / (1.0f + expf(-(o[0] + brdf_bias)));
  out3[1] = 1.0f / (1.0f + expf(-(o[1] + brdf_bias)));
  out3[2] = 1.0f / (1.0f + expf(-(o[2] + brdf_bias)));
  // the next call overwrites X and the accumulators: every thread's TMEM reads are complete (wait::ld above) and
  // ordered before the next MMA by the fence + __syncthreads at the top of the next call
}
