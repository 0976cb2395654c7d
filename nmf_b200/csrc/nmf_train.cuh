// Per-element math of the TRAINING slice (model=tensorf), shared by the kernels in nmf_train.cu and
// by the host check (tests/hostcheck): train-mode step jitter (samplers/alphagrid.py:167-173), compositing backward
// (modules/tensor_nerf.py:19-35), VM-factor backward (fields/tensoRF.py:181-205, 392-405 through F.grid_sample's
// bilinear backward w.r.t. the input), positional-encoding backward (modules/render_modules.py:38-44) and the
// photometric loss of train.py:597-611 through the sRGB tonemap (modules/tonemap.py:38-49).
#pragma once
#include "nmf_field.cuh"

#define NMF_STREAM_JITTER 36u    // oracle/keyed_rng.py STREAM_JITTER

#ifdef __CUDA_ARCH__
#define NMF_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define NMF_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

NMF_HD uint64_t nmf_primary_key(uint64_t seed, uint64_t ray_id) { return nmf_mix64(seed, ray_id); }

// steps = U * stepsize + stepsize / 2  (alphagrid.py:169-172), U keyed by (ray key, step)
NMF_HD float nmf_jitter_step(uint64_t ray_key, int k, float stepsize) {
  const float u = nmf_uniform(nmf_mix64(ray_key, (uint64_t)k), NMF_STREAM_JITTER);
  return NMF_ADD(NMF_MUL(u, stepsize), NMF_MUL(stepsize, 0.5f));
}

// 16-byte accumulate into a channel-last gradient buffer
NMF_HD void nmf_acc4(float* p, nmf_f4 v, float w) {
  if (w == 0.f) return;                     // out-of-range tap (zeros padding): no gradient
#ifdef __CUDA_ARCH__
  atomicAdd((float4*)p, make_float4(v.x * w, v.y * w, v.z * w, v.w * w));
#else
  p[0] += v.x * w; p[1] += v.y * w; p[2] += v.z * w; p[3] += v.w * w;
#endif
}
NMF_HD nmf_f4 nmf_f4_mul(nmf_f4 a, nmf_f4 b) { nmf_f4 o; o.x = a.x * b.x; o.y = a.y * b.y; o.z = a.z * b.z; o.w = a.w * b.w; return o; }

// Backward of one channel group (4 channels at float offset `off`) of one plane/line pair:
//   coef = bilerp(plane)(u, v) * lerp(line)(w);   d plane taps += dc * line * bilinear weight, d line taps += dc * plane * weight
// vstride: floats per texel of the VALUE buffers (appearance: NMF_APP_STRIDE), stride: of the gradient buffers ([..][C])
NMF_HD void nmf_vm_bwd_group(const float* plane, const float* line, float* gplane, float* gline, int w, int vstride, int stride, int off,
                             const NmfLerp& lx, const NmfLerp& ly, const NmfLerp& ll, nmf_f4 dc) {
  const nmf_f4 pv = nmf_bilerp4(plane, w, vstride, off, lx, ly);
  const nmf_f4 lv = nmf_lerp4(line, vstride, off, ll);
  const nmf_f4 dpv = nmf_f4_mul(dc, lv), dlv = nmf_f4_mul(dc, pv);
  float* r0 = gplane + (size_t)ly.i0 * w * stride + off;
  float* r1 = gplane + (size_t)ly.i1 * w * stride + off;
  nmf_acc4(r0 + (size_t)lx.i0 * stride, dpv, ly.w0 * lx.w0);
  nmf_acc4(r0 + (size_t)lx.i1 * stride, dpv, ly.w0 * lx.w1);
  nmf_acc4(r1 + (size_t)lx.i0 * stride, dpv, ly.w1 * lx.w0);
  nmf_acc4(r1 + (size_t)lx.i1 * stride, dpv, ly.w1 * lx.w1);
  nmf_acc4(gline + (size_t)ll.i0 * stride + off, dlv, ll.w0);
  nmf_acc4(gline + (size_t)ll.i1 * stride + off, dlv, ll.w1);
}
// d(density feature) = df for every one of the 48 products (tensoRF.py:392-400)
NMF_HD void nmf_density_bwd(const NmfScene& s, const NmfTaps& t, float df, float* const* gplane, float* const* gline) {
  nmf_f4 dc; dc.x = dc.y = dc.z = dc.w = df;
  for (int p = 0; p < 3; ++p)
    for (int g = 0; g < 4; ++g)
      nmf_vm_bwd_group(s.dval[p], s.lval[p], gplane[p], gline[p], s.plane_w[p], 16, 16, 4 * g, t.px[p], t.py[p], t.pl[p], dc);
}
// dcoef: gradient of the 72 appearance products (before basis_mat)
NMF_HD void nmf_app_bwd(const NmfScene& s, const NmfTaps& t, const float* dcoef, float* const* gplane, float* const* gline) {
  for (int p = 0; p < 3; ++p)
    for (int g = 0; g < 6; ++g) {
      nmf_f4 dc; dc.x = dcoef[p * 24 + 4 * g]; dc.y = dcoef[p * 24 + 4 * g + 1]; dc.z = dcoef[p * 24 + 4 * g + 2]; dc.w = dcoef[p * 24 + 4 * g + 3];
      nmf_vm_bwd_group(s.aval[p], s.alval[p], gplane[p], gline[p], s.plane_w[p], NMF_APP_STRIDE, 24, 4 * g, t.px[p], t.py[p], t.pl[p], dc);
    }
}
NMF_HD void nmf_app_coef(const NmfScene& s, const NmfTaps& t, float* coef) {
  for (int p = 0; p < 3; ++p)
    for (int g = 0; g < 6; ++g) {
      const nmf_f4 c = nmf_app_group(s, t, p, g);
      float* q = coef + p * 24 + 4 * g;
      q[0] = c.x; q[1] = c.y; q[2] = c.z; q[3] = c.w;
    }
}

// feature -> density backward: d softplus(clamp(f,-15,1e3) + shift) / df   (tensor_base.py:83-85; torch softplus threshold 20)
NMF_HD float nmf_feature2density_grad(float f, float shift) {
  if (f < -15.0f || f > 1000.0f) return 0.f;
  const float x = f + shift;
  return x > 20.0f ? 1.0f : nmf_sigmoid(x);
}

// MLPRender_Fea input (render_modules.py:201-235, viewpe = feape = 2): rows of x, `ts` floats apart
//   [feat 24 | view 3 | sin(feat (x) [1,2]) 48 | cos(..) 48 | sin(view (x) [1,2]) 6 | cos(..) 6]
NMF_HD void nmf_plain_encode(const float* feat, const float* d, float* x, int ts) {
  for (int o = 0; o < 24; ++o) {
    const float f = feat[o];
    x[o * ts] = f;
    x[(27 + 2 * o) * ts] = sinf(f);
    x[(27 + 2 * o + 1) * ts] = sinf(f * 2.0f);
    x[(75 + 2 * o) * ts] = cosf(f);
    x[(75 + 2 * o + 1) * ts] = cosf(f * 2.0f);
  }
  for (int c = 0; c < 3; ++c) {
    x[(24 + c) * ts] = d[c];
    x[(123 + 2 * c) * ts] = sinf(d[c]);
    x[(123 + 2 * c + 1) * ts] = sinf(d[c] * 2.0f);
    x[(129 + 2 * c) * ts] = cosf(d[c]);
    x[(129 + 2 * c + 1) * ts] = cosf(d[c] * 2.0f);
  }
}
// dfeat[o] from dx (gradient of the 135 inputs) and the encoded input itself (sin / cos values are in x)
NMF_HD float nmf_plain_encode_bwd(const float* x, int ts, int o, float dx_f, float dx_s1, float dx_s2, float dx_c1, float dx_c2) {
  const float s1 = x[(27 + 2 * o) * ts], s2 = x[(27 + 2 * o + 1) * ts], c1 = x[(75 + 2 * o) * ts], c2 = x[(75 + 2 * o + 1) * ts];
  return dx_f + c1 * dx_s1 + 2.0f * c2 * dx_s2 - s1 * dx_c1 - 2.0f * s2 * dx_c2;
}

// Per-ray loss head.  rgb_map = srgb(lin).clip(0,1) + (1 - acc) * bg  (tensor_nerf.py:657-673, tonemap.py:38-49);
// loss = sum_c (rgb_map.clip(max=1).clip(0,1) - gt.clip(0,1))^2  (train.py:576,597-601) + lambda_pred * 2 * acc
// (prediction_loss without a normal module, tensor_nerf.py:598-602).  Returns the photometric term; g_lin = dL/d lin,
// *g_acc = dL/d acc.
NMF_HD float nmf_train_loss_ray(const float* lin, float acc, const float* bg, const float* gt, float lambda_pred, float* map,
                                float* g_lin, float* g_acc) {
  float loss = 0.f, ga = 2.0f * lambda_pred;
  for (int c = 0; c < 3; ++c) {
    const float x = lin[c];
    const bool hi = x > 0.0031308f;
    const float t = hi ? 1.055f * powf(fmaxf(x, 0.0031308f), 1.0f / 2.4f) - 0.055f : 12.92f * x;
    const float tc = nmf_clampf(t, 0.f, 1.f);
    const float m = tc + (1.0f - acc) * bg[c];
    map[c] = m;
    const float y = nmf_clampf(fminf(m, 1.0f), 0.f, 1.f);
    const float diff = y - nmf_clampf(gt[c], 0.f, 1.f);
    loss += diff * diff;
    const float gm = (m >= 0.f && m <= 1.0f) ? 2.0f * diff : 0.f;           // torch.clamp passes the gradient on the closed interval
    const float dt = hi ? (1.055f / 2.4f) * powf(x, 1.0f / 2.4f - 1.0f) : 12.92f;
    g_lin[c] = (t >= 0.f && t <= 1.0f) ? gm * dt : 0.f;
    ga -= gm * bg[c];
  }
  *g_acc = ga;
  return loss;
}

// compositing backward (tensor_nerf.py:19-35): w_i = alpha_i T_i, T_i = prod_{j<i} (1 - alpha_j + 1e-10);
// suffix = sum_{j>i} dw_j w_j.  Returns d sigma_i;  alpha = 1 - exp(-sigma * dist)
NMF_HD float nmf_composite_bwd(float dw, float T, float alpha, float dist, float suffix) {
  const float dalpha = dw * T - suffix / (1.0f - alpha + 1e-10f);
  return dalpha * dist * (1.0f - alpha);
}

// F.interpolate(mode="bilinear", align_corners=True) source tap for output index `o` (fields/tensoRF.py:208-227 ->
// ATen UpSample.h: scale = (in - 1) / (out - 1) in fp32, src = scale * o, i0 = min(floor(src), in - 1),
// lambda1 = clamp(src - i0, 0, 1), i1 = i0 + (i0 < in - 1))
NMF_HD void nmf_resize_tap(int o, int in, int out, int* i0, int* i1, float* l0, float* l1) {
  if (in == out) { *i0 = *i1 = o; *l0 = 1.0f; *l1 = 0.0f; return; }
  const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f;
  const float src = NMF_MUL(scale, (float)o);
  int a = (int)floorf(src);
  if (a > in - 1) a = in - 1;
  *i0 = a;
  *i1 = a + (a < in - 1 ? 1 : 0);
  *l1 = fminf(fmaxf(NMF_SUB(src, (float)a), 0.0f), 1.0f);
  *l0 = NMF_SUB(1.0f, *l1);
}
NMF_HD float nmf_resize_pixel(const float* plane, int W, int y0, int y1, float hy0, float hy1, int x0, int x1, float wx0, float wx1) {
  const float top = NMF_ADD(NMF_MUL(wx0, plane[(size_t)y0 * W + x0]), NMF_MUL(wx1, plane[(size_t)y0 * W + x1]));
  const float bot = NMF_ADD(NMF_MUL(wx0, plane[(size_t)y1 * W + x0]), NMF_MUL(wx1, plane[(size_t)y1 * W + x1]));
  return NMF_ADD(NMF_MUL(hy0, top), NMF_MUL(hy1, bot));
}

// ---- optimiser step (train.py:443-467, 675-678, 752-754) ----
// Hyper-parameters of one fused update, prepared on the host in double like torch does for its Python scalars.
struct NmfAdamScalars {
  float one_minus_b1, b2, one_minus_b2, eps, weight_decay;
  float step_size;         // lr / (1 - beta1^t)
  float bc2_sqrt;          // sqrt(1 - beta2^t)
  float grad_scale;        // 1 / lbatch_size (train.py:709: total_loss / lbatch_size)
  float max_norm;          // params.clip_grad (<= 0: no clipping)
};
// torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6)); the norm is of the SCALED gradient
NMF_HD float nmf_clip_coef(double sq_norm, float grad_scale, float max_norm) {
  if (!(max_norm > 0.f)) return 1.0f;
  const float total = grad_scale * (float)sqrt(sq_norm);
  const float c = max_norm / (total + 1e-6f);
  return c < 1.0f ? c : 1.0f;
}
// torch.optim.Adam, one element (torch/optim/adam.py _single_tensor_adam, amsgrad=False): L2 weight decay added to the
// gradient, exp_avg.lerp_(g, 1 - b1), exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2),
// denom = sqrt(exp_avg_sq) / sqrt(bias_correction2) + eps, param.addcdiv_(exp_avg, denom, -step_size)
NMF_HD void nmf_adam_elem(float* p, float g, float* m, float* v, const NmfAdamScalars& h, float gmul) {
  g = NMF_MUL(g, gmul);
  if (h.weight_decay != 0.f) g = NMF_ADD(g, NMF_MUL(h.weight_decay, *p));
  const float m1 = NMF_ADD(*m, NMF_MUL(NMF_SUB(g, *m), h.one_minus_b1));
  const float v1 = NMF_ADD(NMF_MUL(*v, h.b2), NMF_MUL(h.one_minus_b2, NMF_MUL(g, g)));
  const float denom = NMF_ADD(sqrtf(v1) / h.bc2_sqrt, h.eps);
  *p = NMF_SUB(*p, NMF_MUL(h.step_size, m1 / denom));
  *m = m1;
  *v = v1;
}
// d/dp of weight * mean|p| (TensorVMSplit.density_L1, fields/tensoRF.py:332-340): coef = weight / numel
NMF_HD float nmf_l1_grad(float p, float coef) { return p > 0.f ? coef : (p < 0.f ? -coef : 0.f); }

#ifdef __CUDACC__
// One warp writes the jittered distances of a ray: z_k = tmin + fp32(sum_{j<=k} step_j).  Every partial sum of <= 2048
// fp32 step lengths in [stepsize/2, 3 stepsize/2] is exact in fp64 (37 significant bits), so the scan order does not
// matter: the prefix equals ATen's sequential fp64 accumulation (CPU cumsum), rounded to fp32 per element.
__device__ __forceinline__ void nmf_warp_jitter_z(uint64_t key, float tmin, float stepsize, int S, float* zrow, int lane) {
  double carry = 0.0;
  for (int base = 0; base < S; base += 32) {
    const int k = base + lane;
    double v = k < S ? (double)nmf_jitter_step(key, k, stepsize) : 0.0;
    for (int off = 1; off < 32; off <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, v, off);
      if (lane >= off) v += u;
    }
    v += carry;
    carry = __shfl_sync(0xffffffffu, v, 31);
    if (k < S) zrow[k] = NMF_ADD(tmin, (float)v);
  }
}

// one CTA of 1024 threads: inclusive sums of n_valid -> whole_valid, offsets, kept counts (alphagrid.py:353-364).
// n_kept[0..1] must be zero on entry.
static __global__ void __launch_bounds__(1024) k_train_prefix(const int* n_valid, int n, int max_samples, int* offs, uint8_t* whole,
                                                              int* n_kept) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int b = t * per, e = min(n, b + per);
  long long sum = 0;
  for (int i = b; i < e; ++i) sum += n_valid[i];
  part[t] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const long long v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  const long long total = part[1023];
  const bool trunc = max_samples > 0 && total > (long long)max_samples;
  long long run = t ? part[t - 1] : 0;
  int last = -1;
  long long last_incl = 0;
  for (int i = b; i < e; ++i) {
    const long long incl = run + n_valid[i];
    const bool keep = !trunc || incl < (long long)max_samples;   // the inclusive sums are monotone: kept rays are a prefix
    whole[i] = keep;
    if (offs) offs[i] = (int)run;            // exclusive prefix (used for kept rays only)
    if (keep) { last = i; last_incl = incl; }
    run = incl;
  }
  if (last >= 0) {
    atomicMax(n_kept, last + 1);
    atomicMax(n_kept + 1, (int)last_incl);
  }
}
#endif
