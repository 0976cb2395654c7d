// nmf_b200 -- measurement helpers (bench.py only; not on the render path).
//
// nmf_bench_gather: the ceiling the gather kernels (k_march, k_shade) are measured against.  Their factor set (78 MB at
// G = 300) is L2-resident, so the HBM copy rate is not their roofline; what bounds them is how many
// independent 16-byte taps per second the L2 / L1 path serves.  This kernel issues exactly that: every thread draws
// `taps` pseudo-random 16-byte loads over a buffer of `n_elems` float4 (8 independent loads in flight per thread, like the
// taps of one VM query), folds them into a checksum and writes one float4.  bytes = threads * taps * 16.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nmf_b200.h"

__global__ void __launch_bounds__(256) k_bench_gather(const float4* __restrict__ buf, unsigned long long n_elems, int taps, float4* sink) {
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long h = tid * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < taps; i += 8) {
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
      v[j] = __ldg(buf + (h % n_elems));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
  }
  sink[tid] = acc;
}

extern "C" int nmf_bench_gather(const void* buf, size_t n_elems, int taps, int n_threads, void* sink, void* stream) {
  if (!buf || !sink || n_elems == 0 || taps <= 0 || (taps & 7) || n_threads <= 0 || (n_threads & 255)) return NMF_E_ARG;
  k_bench_gather<<<n_threads / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)buf, (unsigned long long)n_elems, taps, (float4*)sink);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? NMF_OK : (int)e;
}
