// BRDF MLP 66 -> 64 -> 64 -> 4 (modules/brdf.py:73-120, 237-239) on the 5th-generation tensor cores.
//
// One CTA = 128 threads = 128 MLP rows (bounce rays); thread t owns row t = TMEM lane t.
//   layer 1:  D[128x64] = X[128x72] * W0^T[72x64]     9 x tcgen05.mma.kind::tf32 (M=128, N=64, K=8), K padded 66 -> 72
//   layer 2:  D[128x64] = relu(D+b0)[128x64] * W1^T   8 x tcgen05.mma.kind::tf32
//   layer 3:  64 -> 3 outputs per row in registers (192 FMA), sigmoid
// Operands live in shared memory in the canonical no-swizzle K-major layout (8-row x 16-byte core matrices):
//   element (row r, k) of an R-row operand at byte  (k/4) * (R*16) + r*16 + (k%4)*4
//   => stride between 8-row groups SBO = 128 B, between 16-byte K chunks LBO = R*16 B.
// Accumulators live in TMEM (64 fp32 columns per CTA) and come back with tcgen05.ld.32x32b.x64.
// Activations and weights are rounded to TF32 (cvt.rna) before the MMA; accumulation is fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_ROWS 128
#define TC_K0 72                       // 66 inputs padded to a multiple of 8
#define TC_A_BYTES (18 * TC_ROWS * 16) // 18 K-chunks of 4 floats
#define TC_W0_FLOATS (18 * 64 * 4)
#define TC_W1_FLOATS (16 * 64 * 4)
// shared memory carve (floats): A | W0 | W1 | b0 | b1 | w2t (64x4) | b2 (4) | mbarrier (2 words) | tmem ptr (1 word) | pad
#define TC_OFF_A 0
#define TC_OFF_W0 (TC_A_BYTES / 4)
#define TC_OFF_W1 (TC_OFF_W0 + TC_W0_FLOATS)
#define TC_OFF_B0 (TC_OFF_W1 + TC_W1_FLOATS)
#define TC_OFF_B1 (TC_OFF_B0 + 64)
#define TC_OFF_W2 (TC_OFF_B1 + 64)
#define TC_OFF_B2 (TC_OFF_W2 + 256)
#define TC_OFF_BAR (TC_OFF_B2 + 4)
#define TC_OFF_TMEM (TC_OFF_BAR + 2)
#define TC_SMEM_FLOATS (TC_OFF_TMEM + 2)
#define TC_SMEM_BYTES (TC_SMEM_FLOATS * 4)
#define TC_TMEM_COLS 64

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tc_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// K-major, no swizzle: LBO = byte distance between the two 16-byte K chunks of one MMA, SBO = between 8-row groups
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 64
#define TC_IDESC ((1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate)
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

// this thread's 64 accumulator columns (its TMEM lane); `taddr` already carries the warp's lane offset
__device__ __forceinline__ void tc_ld64(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcMlp {
  float* sm;          // shared memory base (TC_SMEM_BYTES, 128-byte aligned)
  uint32_t tmem;      // TMEM base address of this CTA's 64 columns
  uint32_t phase;     // mbarrier parity of the next completion
};

// Called by all 128 threads once per CTA.  w0u / w1u: weights already in the canonical operand layout
// (scene.py: [K/4][64][4], TF32-rounded, K zero-padded); b0, b1, w2t [64][4], b2 [4] in fp32.
__device__ __forceinline__ void tc_mlp_init(TcMlp& c, float* sm, const float* w0u, const float* w1u, const float* b0,
                                            const float* b1, const float* w2t, const float* b2) {
  c.sm = sm;
  c.phase = 0;
  const int tid = threadIdx.x;
  for (int i = tid; i < TC_W0_FLOATS / 4; i += TC_ROWS) ((float4*)(sm + TC_OFF_W0))[i] = __ldg((const float4*)w0u + i);
  for (int i = tid; i < TC_W1_FLOATS / 4; i += TC_ROWS) ((float4*)(sm + TC_OFF_W1))[i] = __ldg((const float4*)w1u + i);
  if (tid < 64) { sm[TC_OFF_B0 + tid] = b0[tid]; sm[TC_OFF_B1 + tid] = b1[tid]; }
  for (int i = tid; i < 256; i += TC_ROWS) sm[TC_OFF_W2 + i] = w2t[i];
  if (tid < 4) sm[TC_OFF_B2 + tid] = b2[tid];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(sm + TC_OFF_BAR)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {   // one warp allocates the accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(sm + TC_OFF_TMEM)),
                 "r"((uint32_t)TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_async_smem();      // the weight tiles were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *(volatile uint32_t*)(sm + TC_OFF_TMEM);
}
__device__ __forceinline__ void tc_mlp_free(TcMlp& c) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"((uint32_t)TC_TMEM_COLS) : "memory");
}

// All 128 threads call this together.  x[72]: this row's inputs (x[66..71] must be 0).  out3: sigmoid(mlp(x)[:3] + bias)
__device__ __forceinline__ void tc_mlp_forward(TcMlp& c, const float (&x)[TC_K0], float brdf_bias, float* out3) {
  float* sm = c.sm;
  const int tid = threadIdx.x;
  const uint32_t a_addr = tc_smem_u32(sm + TC_OFF_A);
  const uint32_t bar = tc_smem_u32(sm + TC_OFF_BAR);
  const uint32_t taddr = c.tmem + ((uint32_t)(tid & ~31) << 16);     // lane offset of this warp in bits 31:16
  float4* arow = (float4*)(sm + TC_OFF_A) + tid;                      // + kchunk * 128
#pragma unroll
  for (int kc = 0; kc < 18; ++kc)
    arow[kc * TC_ROWS] = make_float4(tc_tf32(x[4 * kc]), tc_tf32(x[4 * kc + 1]), tc_tf32(x[4 * kc + 2]), tc_tf32(x[4 * kc + 3]));
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t w0 = tc_smem_u32(sm + TC_OFF_W0);
#pragma unroll
    for (int k = 0; k < 9; ++k)
      tc_mma(c.tmem, tc_desc(a_addr + k * 2 * (TC_ROWS * 16), TC_ROWS * 16, 128), tc_desc(w0 + k * 2 * (64 * 16), 64 * 16, 128), k > 0);
    tc_commit(bar);
  }
  tc_wait(bar, c.phase);
  c.phase ^= 1;
  tc_fence_after();
  float h[64];
  tc_ld64(taddr, h);
#pragma unroll
  for (int kc = 0; kc < 16; ++kc) {
    float4 v;
    v.x = tc_tf32(fmaxf(h[4 * kc] + sm[TC_OFF_B0 + 4 * kc], 0.f));
    v.y = tc_tf32(fmaxf(h[4 * kc + 1] + sm[TC_OFF_B0 + 4 * kc + 1], 0.f));
    v.z = tc_tf32(fmaxf(h[4 * kc + 2] + sm[TC_OFF_B0 + 4 * kc + 2], 0.f));
    v.w = tc_tf32(fmaxf(h[4 * kc + 3] + sm[TC_OFF_B0 + 4 * kc + 3], 0.f));
    arow[kc * TC_ROWS] = v;
  }
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t w1 = tc_smem_u32(sm + TC_OFF_W1);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      tc_mma(c.tmem, tc_desc(a_addr + k * 2 * (TC_ROWS * 16), TC_ROWS * 16, 128), tc_desc(w1 + k * 2 * (64 * 16), 64 * 16, 128), k > 0);
    tc_commit(bar);
  }
  tc_wait(bar, c.phase);
  c.phase ^= 1;
  tc_fence_after();
  tc_ld64(taddr, h);
  float o0 = sm[TC_OFF_B2], o1 = sm[TC_OFF_B2 + 1], o2 = sm[TC_OFF_B2 + 2];
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    const float hv = fmaxf(h[k] + sm[TC_OFF_B1 + k], 0.f);
    const float4 wv = *(const float4*)(sm + TC_OFF_W2 + 4 * k);
    o0 += hv * wv.x; o1 += hv * wv.y; o2 += hv * wv.z;
  }
  out3[0] = 1.0f / (1.0f + expf(-(o0 + brdf_bias)));
  out3[1] = 1.0f / (1.0f + expf(-(o1 + brdf_bias)));
  out3[2] = 1.0f / (1.0f + expf(-(o2 + brdf_bias)));
  // the next call overwrites A and the accumulators: every thread's TMEM reads are complete (wait::ld above) and
  // ordered before the next MMA by the fence + __syncthreads at the top of the next call
}
