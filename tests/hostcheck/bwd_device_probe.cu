// TEST INFRASTRUCTURE -- compile-only probe: csrc/nmf_microfacet_bwd.cuh must stay device-compilable for sm_100a (the reverse-pass
// kernels of DESIGN.md section 9 will call it).  Compiled by tests/test_host_api.py::test_backward_header_compiles_for_the_device; never linked.
#include "nmf_microfacet_bwd.cuh"
__global__ void k_probe(const NmfScene s, const float* nfeat, const float* V, const float* N, const float* R0, const float* diffuse,
                        const float* rough, const float* u, int n, int m, const float* g, float* dR0, float* ddiff, float* dr,
                        float* dfeat, NmfBrdfGrads bg, float* gsat, float* g_top, float* g_bot, float* dN, float* gpack0, float* glpack0,
                        const float* kx, const float* ky, float* dplane) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  nmf_bounce_sample_bwd(s, nfeat + 24 * i, nmf_mk3(V[3 * i], V[3 * i + 1], V[3 * i + 2]), nmf_mk3(N[3 * i], N[3 * i + 1], N[3 * i + 2]),
                        R0 + 3 * i, diffuse + 3 * i, rough[i], u + (size_t)2 * m * i, m, g + 3 * i, dR0 + 3 * i, ddiff + 3 * i, dr + i,
                        dfeat + 24 * i, bg, gsat, g_top, g_bot, dN + 3 * i);
  float xn[3] = {V[3 * i], V[3 * i + 1], V[3 * i + 2]};
  const NmfTaps t = nmf_vm_taps(s, xn);
  float grad[3] = {0.f, 0.f, 0.f}, dgrad[3];
  for (int l = 0; l < 8; ++l) nmf_normal_lane(s, t, l, grad);
  nmf_normal_vec_bwd(s, grad, dN + 3 * i, dgrad);
  float* gp[3] = {gpack0, gpack0, gpack0};
  float* gl[3] = {glpack0, glpack0, glpack0};
  nmf_normal_bwd(s, t, dgrad, gp, gl);
  dplane[i] = nmf_plane_grad_finish(gpack0, s.plane_h[0], s.plane_w[0], kx, ky, i % s.plane_h[0], i % s.plane_w[0], i & 15);
}
