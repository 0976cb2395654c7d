// Reverse pass of the analytic normals (DESIGN.md section 9, row "normals"): what autograd does to
// TensorBase.compute_normals (fields/tensor_base.py:107-129) -> GridSampler2D.backward (modules/grid_sample_Cinf.py:109-281)
// w.r.t. the density planes and lines.
//   k_normals_bwd_scatter  thread per sample: nmf_normals_bwd_sample (host-checked, csrc/nmf_microfacet_bwd.cuh) -- recomputes
//                          the taps, d n -> d grad through the normalisation, fp32 atomics into gradient images laid out like
//                          dpack ([h][w][val16 | dx16 | dy16]) and lpack ([n][4][val4 | dy4])
//   k_normals_bwd_planes   thread per (texel, channel): adjoint of the 5x5 smoothed-difference stencil (zero padding 2) over the
//                          dx / dy images + the value image -> d plane, channel-last like NmfPlainGrads.d_plane (accumulated)
//   k_normals_bwd_lines    the same for the lines (only the stencil's centre column meets an (N,1) image)
// The finishing pass streams each 192 B/texel gradient image once (25 taps hit L1/L2): HBM-bound, (192 + 2 x 64) B per texel.
#include <cuda_runtime.h>
#include <stdint.h>
#include "nmf_microfacet_bwd.cuh"

#define CKL() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

struct NormalImgs { float* gpack[3]; float* glpack[3]; };

__global__ void k_normals_bwd_scatter(const NmfScene s, const float* __restrict__ xyz, int n, int stride,
                                      const float* __restrict__ d_normals, const NormalImgs im) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float dn[3] = {d_normals[3 * (size_t)i], d_normals[3 * (size_t)i + 1], d_normals[3 * (size_t)i + 2]};
  float* gp[3] = {im.gpack[0], im.gpack[1], im.gpack[2]};
  float* gl[3] = {im.glpack[0], im.glpack[1], im.glpack[2]};
  nmf_normals_bwd_sample(s, xyz + (size_t)i * stride, dn, gp, gl);
}

__global__ void k_normals_bwd_planes(const float* __restrict__ gpack, int h, int w, const float* __restrict__ kx25,
                                     const float* __restrict__ ky25, float* __restrict__ d_plane) {
  __shared__ float kx[25], ky[25];
  if (threadIdx.x < 25) { kx[threadIdx.x] = kx25[threadIdx.x]; ky[threadIdx.x] = ky25[threadIdx.x]; }
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t texel = idx >> 4;
  const int c = (int)(idx & 15);
  if (texel >= (size_t)h * w) return;
  const int y = (int)(texel / w), x = (int)(texel - (size_t)y * w);
  d_plane[texel * 16 + c] += nmf_plane_grad_finish(gpack, h, w, kx, ky, y, x, c);
}

__global__ void k_normals_bwd_lines(const float* __restrict__ glpack, int n, const float* __restrict__ ky25, float* __restrict__ d_line) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = idx >> 4, c = idx & 15;
  if (i >= n) return;
  d_line[(size_t)i * 16 + c] += nmf_line_grad_finish(glpack, n, ky25, i, c);
}

static bool imgs_ok(const NmfNormalGrads* g) {
  if (!g) return false;
  for (int p = 0; p < 3; ++p) if (!g->gpack[p] || !g->glpack[p]) return false;
  return true;
}

extern "C" int nmf_vm_normals_bwd_scatter(const NmfScene* scene, const float* xyz, int n, int stride, const float* d_normals,
                                          const NmfNormalGrads* imgs, void* stream) {
  if (!scene || !xyz || !d_normals || !imgs_ok(imgs) || n < 0 || stride < 3 || !scene->dpack[0] || !scene->lpack[0]) return NMF_E_ARG;
  if (n == 0) return NMF_OK;
  NormalImgs im;
  for (int p = 0; p < 3; ++p) { im.gpack[p] = imgs->gpack[p]; im.glpack[p] = imgs->glpack[p]; }
  k_normals_bwd_scatter<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*scene, xyz, n, stride, d_normals, im);
  CKL();
  return NMF_OK;
}

extern "C" int nmf_vm_normals_bwd_finish(const NmfScene* scene, const NmfNormalGrads* imgs, const float* kx25, const float* ky25,
                                         float* const* d_plane, float* const* d_line, void* stream) {
  if (!scene || !imgs_ok(imgs) || !kx25 || !ky25 || !d_plane || !d_line) return NMF_E_ARG;
  for (int p = 0; p < 3; ++p) {
    if (!d_plane[p] || !d_line[p]) return NMF_E_ARG;
    const int h = scene->plane_h[p], w = scene->plane_w[p], ln = scene->line_n[p];
    if (h <= 0 || w <= 0 || ln <= 0) return NMF_E_ARG;
    const size_t nt = (size_t)h * w * 16;
    k_normals_bwd_planes<<<(unsigned)((nt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(imgs->gpack[p], h, w, kx25, ky25, d_plane[p]);
    CKL();
    k_normals_bwd_lines<<<(ln * 16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(imgs->glpack[p], ln, ky25, d_line[p]);
    CKL();
  }
  return NMF_OK;
}
