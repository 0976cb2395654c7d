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

// `group` consecutive lanes (1, 2, 4 or 8) read consecutive 16-byte pieces of ONE random segment of 16 * group bytes --
// the granularity of the kernels' taps: a density texel is 64 bytes (4 lanes), an appearance texel 96, a derivative-packed
// texel pair 384 (8 lanes x three 128-byte pieces).  group = 1 is the worst case (half of every 32-byte sector is wasted).
__global__ void __launch_bounds__(256) k_bench_gather(const float4* __restrict__ buf, unsigned long long n_elems, int taps, int group,
                                                      float4* sink) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned gid = tid / (unsigned)group, sub = tid % (unsigned)group;
  const unsigned n_seg = (unsigned)(n_elems / (unsigned)group);          // < 2^32 segments
  unsigned h = gid * 0x9E3779B9u + 0x85EBCA6Bu;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < taps; i += 8) {
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;      // murmur3 finaliser: a few ALU ops per tap
      const unsigned seg = __umulhi(h, n_seg);                                            // uniform in [0, n_seg) without a division
      v[j] = __ldg(buf + (size_t)seg * (unsigned)group + sub);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
  }
  sink[tid] = acc;
}

extern "C" int nmf_bench_gather(const void* buf, size_t n_elems, int taps, int group, int n_threads, void* sink, void* stream) {
  if (!buf || !sink || n_elems < 8 || taps <= 0 || (taps & 7) || n_threads <= 0 || (n_threads & 255)) return NMF_E_ARG;
  if (group != 1 && group != 2 && group != 4 && group != 8) return NMF_E_ARG;
  k_bench_gather<<<n_threads / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)buf, (unsigned long long)n_elems, taps, group, (float4*)sink);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? NMF_OK : (int)e;
}
