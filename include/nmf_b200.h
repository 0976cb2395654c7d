/* nmf_b200 -- C ABI of the B200-native NMF per-ray render path.
 *
 * This is the drop-in boundary of the hot path.  The reference (half-potato/nmf) has no native
 * operator boundary on this path: it is pure PyTorch behind hydra `_target_` plugin slots
 * (SURVEY.md section 8b).  The Python plugin classes in `nmf_b200/` keep those slots and call the
 * entry points below through ctypes; every entry point names the reference code it replaces
 * (file:line relative to the reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is device memory unless the name ends in
 *     `_host`; all buffers are caller-allocated (PyTorch owns them) and must stay alive until
 *     the stream reaches the end of the call;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), launches a fixed
 *     sequence of kernels and never synchronises, so a call can be captured in a CUDA graph;
 *   - return value: 0 = ok, < 0 = NMF_E_* (argument problems, detected on the host before any
 *     launch), > 0 = a cudaError_t from a launch.  Device-side capacity overflows are reported
 *     in NmfCounters.error (checked by the caller after it synchronises);
 *   - fp32 everywhere; integer / occupancy work is bit-exact with the reference's fp32 op order.
 */
#ifndef NMF_B200_H
#define NMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMF_ABI_VERSION 10
#define NMF_APP_STRIDE 24      /* floats per appearance texel; 32 (a 128-byte texel) was measured: no gain, +9 MB of L2 footprint */

#define NMF_OK 0
#define NMF_E_ARG (-1)          /* null pointer / non-positive size                                   */
#define NMF_E_UNSUPPORTED (-2)  /* shape outside what the kernels are compiled for                     */
#define NMF_E_WORKSPACE (-3)    /* workspace buffer smaller than nmf_workspace_bytes() says            */

#define NMF_DENSITY_COMP 16     /* configs/field/tensorf.yaml:6  density_n_comp                        */
#define NMF_APP_COMP 24         /* configs/field/tensorf.yaml:7  appearance_n_comp                     */
#define NMF_APP_DIM 24          /* configs/field/tensorf.yaml:8  app_dim                               */
#define NMF_BRDF_IN 66          /* modules/brdf.py:96-120  24 + 2*(18+3)                               */
#define NMF_BRDF_HID 64         /* configs/model/microfacet_tensorf2.yaml:93 hidden_w                  */
#define NMF_MAX_STEPS 2048      /* dense steps per ray the march kernel keeps a bitmask for            */
#define NMF_MAX_COARSE_WORDS 2048 /* coarse occupancy bit-field held in shared memory by k_march            */
#define NMF_MAX_BOUNCE 400      /* modules/pt_selectors.py:39                                          */
#define NMF_PLAIN_IN 135        /* modules/render_modules.py:201-235 with viewpe=2, feape=2            */
#define NMF_PLAIN_HID 128       /* configs/model/tensorf.yaml featureC                                 */

/* device error bits (NmfCounters.error) */
#define NMF_DEV_E_SURVIVORS 1u  /* surviving-sample list overflowed                                    */
#define NMF_DEV_E_BSAMPLES 2u   /* bounce-sample list overflowed                                       */
#define NMF_DEV_E_BRAYS 4u      /* a chunk's bounce-ray region overflowed                              */
#define NMF_DEV_E_VSAMPLES 8u   /* train mode: the valid-sample list kept for the reverse pass overflowed */

/* ------------------------------------------------------------------------------------------------
 * Scene: every weight the path reads, in the layouts the kernels gather from (DESIGN.md "HBM layout").
 * Built once per weight update by nmf_b200/scene.py from a reference-format state_dict.
 * plane p pairs matMode[p] = (0,1),(0,2),(1,2) with vecMode[p] = 2,1,0   (fields/tensoRF.py:40-41)
 * ------------------------------------------------------------------------------------------------ */
typedef struct NmfScene {
  /* geometry: fields/tensor_base.py:56-62,219-232 ; samplers/alphagrid.py:96-103 */
  float aabb0[3], aabb1[3];
  float inv_aabb2[3];      /* 2 / (aabb1 - aabb0) */
  float stepsize;          /* min(units) * step_ratio */
  float near, far;
  float distance_scale;    /* configs/field/tensorf.yaml:4 */
  float density_shift;     /* tensor_base.py:85 */
  int n_steps;             /* nSamples: dense steps per ray */
  int plane_w[3], plane_h[3], line_n[3];

  /* occupancy (samplers/alphagrid.py:6-60): bit (z*oh + y)*opitch + x of `occ_vox` is alpha_volume[z][y][x] > 0;
   * `occ_cell` holds the OR over the 8 corners of cell (x0,y0,z0) (in-range corners only). opitch % 32 == 0. */
  const uint32_t* occ_vox;
  const uint32_t* occ_cell;
  int ow, oh, od, opitch;
  int has_occ;
  /* conservative coarse occupancy, a pure accelerator that never changes a result (optional: occ_coarse may be NULL):
   * flat bit (cz*och + cy)*ocw + cx is set iff a voxel with index in [8c-1, 8c+9] on every axis is set, i.e. iff the
   * exact test can succeed for a sample whose cell index -- computed with one multiply, occ_scale = (size-1)/aabbSize,
   * so possibly one cell off -- lands in coarse cell c.  At most NMF_MAX_COARSE_WORDS 32-bit words (it is held in
   * shared memory by the march). */
  const uint32_t* occ_coarse;
  int ocw, och, ocd;
  float occ_scale[3];

  /* density factors.  dval: [h][w][16] plane values; dpack: [h][w][val16 | dx16 | dy16] where dx/dy are the
   * smoothed-difference planes of modules/grid_sample_Cinf.py:218-242; lines: lval [n][16], lpack [n][4][val4,dy4] */
  const float* dval[3];
  const float* dpack[3];
  const float* lval[3];
  const float* lpack[3];
  /* appearance factors: [h][w][NMF_APP_STRIDE], [n][NMF_APP_STRIDE] (24 channels per texel; the stride is a build-time constant
   * so that padded texels can be tried: 128-byte texels never straddle a line, which removes a third of the L1 wavefronts of
   * these taps, but k_shade did not get faster -- profiles/r02_d_*); basis_t: [72][24] = rf.basis_mat.weight transposed
   * (tensoRF.py:293) */
  const float* aval[3];
  const float* alval[3];
  const float* basis_t;

  /* microfacet material heads (modules/render_modules.py:447-574): head_w [11][24] rows = diffuse(3), tint(3),
   * f0(3), roughness(2); head_b [11] */
  const float* head_w;
  const float* head_b;
  float diffuse_mul, diffuse_bias, tint_bias, f0_bias, roughness_bias;
  /* BRDF MLP (modules/brdf.py:73-120): transposed weights w0t [66][64], w1t [64][64], w2t [64][4]; biases */
  const float* brdf_w0t; const float* brdf_b0;
  const float* brdf_w1t; const float* brdf_b1;
  const float* brdf_w2t; const float* brdf_b2;
  float brdf_bias;
  float anoise;            /* models/microfacet.py:297 */
  const float* sobol;      /* [1024][2] brdf_samplers/base.py:6-9 */
  const float* sh_conv;    /* [9][3] clamped-cosine-convolved SH of the env (integral_equirect.py:324-360, sh.py:149-157) */

  /* environment (modules/integral_equirect.py): sat [eh][ew][4] (rgb + pad) summed-area table of exp(bg)/1000 */
  const float* env_sat;
  int env_h, env_w;
  float env_mipbias;
  float env_top[3], env_bot[3];   /* mean of first / last row of exp(bg) (:498-502) */

  /* model=tensorf plumbing config: view MLP 135->128->128->3 (modules/render_modules.py:201-235), transposed */
  const float* plain_w0t; const float* plain_b0;
  const float* plain_w1t; const float* plain_b1;
  const float* plain_w2t; const float* plain_b2;
  /* the same weights as stored by the reference, (out, in) row-major: read by the training backward (nmf_train_plain) */
  const float* plain_w0; const float* plain_w1;

  /* bounce budgets: models/microfacet.py:327-349, modules/pt_selectors.py:5-60 */
  int rays_per_ray;        /* 128 */
  int max_brdf_rays1;      /* max_brdf_rays[1] = 450000 */
  int max_retrace;         /* max_retrace_rays[0] = 1000; 0 disables the secondary level */
  int model;               /* 0 = microfacet, 1 = plain view-MLP (models/tensorf.py) */

  /* BRDF MLP operands for the tensor-core path: fp16 weights of the three layers in the canonical K-major operand
   * layout of tcgen05.mma ([K/8][rows][8 halves]; rows = 64, 64, 16; K zero-padded to 80; the bias of each layer
   * sits in column 66, which multiplies a constant-1 input); see csrc/nmf_mlp_tc.cuh */
  const void* brdf_w0u;
  const void* brdf_w1u;
  const void* brdf_w2u;
  int mlp_mode;            /* 0 = tcgen05 kind::f16 (fp16 operands, fp32 accumulate), 1 = fp32 SIMT */
  /* the same three weight tiles in BF16 for the reverse pass (csrc/nmf_mlp_tc_bwd.cuh: gradients need fp32's exponent
   * range); optional: NULL selects the fp32 reverse kernel */
  const void* brdf_w0b;
  const void* brdf_w1b;
  const void* brdf_w2b;
  /* optional (NULL = not used): the summed-area table with BOTH x-taps of a bilinear lookup in one record,
   * [h][w][8] = { S(y,x).rgb, 0, S(y,min(x+1,w-1)).rgb, 0 }: one aligned 32-byte load (LDG.256) per row of a lookup corner
   * instead of two 16-byte loads -- the environment kernels are bound by L1 data-pipe wavefronts, one per lane-load.
   * Built by nmf_env_pair_sat; costs h*w*32 bytes of L2 footprint, so the host only builds it for maps up to 512 x 1024. */
  const float* env_sat2;
  /* optional (NULL = use env_mipbias / env_top / env_bot above): 7 device floats { mipbias, top rgb, bottom rgb } written by
   * nmf_env_build_sat_dev -- the per-step values of a training run stay on the device (no host round trip per iteration) */
  const float* env_dyn;
} NmfScene;

/* per-call render parameters */
typedef struct NmfRender {
  int n_rays;
  int chunk;               /* renderer.py:60 chunk (4096): bounds the scope of the per-chunk retrace selection */
  float focal;
  uint64_t seed;           /* keyed RNG: every random number is a function of (seed, ray id, step, bounce ray) */
  uint64_t ray_id0;        /* global index of rays[0] (so that a sharded / batched render draws the same numbers) */
  float skip_eps;          /* shade a sample only if w >= skip_eps / n_valid(ray); 0 = shade every valid sample */
  float t_cut;             /* stop evaluating density once transmittance < t_cut; 0 = never */
  int white_bg;            /* 1: primary background is white (tensor_nerf.py:215,478) */
  float cap_scale;         /* scales the capacities of the scratch lists (0 = 1.0).  A scene that overflows one of them
                            * (NmfCounters.error) is re-rendered with a larger scale and nmf_workspace_bytes_scaled() */
} NmfRender;

/* outputs of one TensorNeRF.forward in eval mode (modules/tensor_nerf.py:448-566, 657-673); any pointer may be NULL */
typedef struct NmfImages {
  float* rgb_map;          /* (n,3) */
  float* acc_map;          /* (n)   */
  float* depth;            /* (n)   */
  float* world_normal;     /* (n,3) */
  float* normal;           /* (n,3) */
  float* termination_xyz;  /* (n,4) */
  int64_t* surf_width;     /* (n)   */
  float* cross_section;    /* (n,3) */
  float* diffuse;          /* (n,3) */
  float* tint;             /* (n,3) */
  float* roughness;        /* (n,3) */
  float* spec;             /* (n,3) */
  float* albedo;           /* (n,3) */
} NmfImages;

/* per-chunk statistics, written on the device (int32 [n_chunks] each unless stated) */
typedef struct NmfCounters {
  int* n_samples0;         /* statistics["n_samples"][0]: valid primary samples of the chunk (tensor_nerf.py:266) */
  int* n_samples1;         /* statistics["n_samples"][1]: valid samples of the retraced rays                       */
  int* n_cand;             /* in-AABB candidate steps (occupancy lookups), level 0 + level 1                       */
  int* n_bounce_rays0;     /* bounce rays drawn at recur 0                                                        */
  int* n_bounce_rays1;     /* bounce rays drawn at recur 1 (all go to the environment)                            */
  int* n_retrace;          /* rays actually re-traced (<= max_retrace)                                            */
  int* n_shaded;           /* [2] (not per chunk): samples that survived the weight cut at level 0 / 1            */
  unsigned* error;         /* [1] NMF_DEV_E_* bits                                                                 */
  /* float [n_chunks][4]: the sums behind the statistics TensorNeRF.forward returns without debug maps
   * (modules/tensor_nerf.py:567-649): [0] ori_loss = sum w*min(v.n,0)^2, [1] 3*diffuse_reg = sum of the diffuse map,
   * [2] sum over samples and channels of the tint debug value (brdf_reg = [2] / (3*n_samples0)), [3] sum of acc
   * (prediction_loss = 2*[3] without a normal module) */
  float* stat4;
} NmfCounters;

int nmf_abi_version(void);

/* Optional phase timing for bench.py / profiling: when enabled, nmf_render_rays records a CUDA event on the caller's
 * stream after each phase; nmf_profile_read (after the stream is synchronised) returns the elapsed milliseconds of
 * the NMF_N_PHASES phases of the LAST call: march0, shade0, bounce0, select, march1, shade1, bounce1, incoming1, finish1,
 * incoming0, reduce0, finish. */
#define NMF_N_PHASES 12
int nmf_profile_enable(int on);
int nmf_profile_read(float* ms, int n);
const char* nmf_profile_phase_name(int i);

/* Bytes of scratch nmf_render_rays needs for `n_rays` rays in chunks of `chunk` (NmfRender.cap_scale = 1 / as given). */
size_t nmf_workspace_bytes(const NmfScene* scene, int n_rays, int chunk);
size_t nmf_workspace_bytes_scaled(const NmfScene* scene, int n_rays, int chunk, float cap_scale);

/* The fused path: replaces renderer.chunk_renderer (renderer.py:56-106) + TensorNeRF.forward
 * (modules/tensor_nerf.py:210-674) in eval mode for all chunks of `rays` at once.
 * rays: (n_rays, 6) [origin, direction].  workspace: nmf_workspace_bytes() bytes, 256-byte aligned. */
int nmf_render_rays(const NmfScene* scene, const NmfRender* rp, const float* rays, const NmfImages* out,
                    const NmfCounters* counters, void* workspace, size_t workspace_bytes, void* stream);

/* Same, with HOST buffers: copies `rays_host` in and every non-NULL image / counter out.  Maps are copied as soon as
 * they are final -- geometry maps after the march, material maps after the shade -- on a library-owned copy stream
 * that is ordered against `stream` with events only, so the transfers overlap the rest of the sequence; everything
 * is complete when `stream` reaches the end of the call (the copies are asynchronous only if the host buffers are
 * pinned).  One device per process.  `out_host` / `counters_host` hold HOST
 * pointers; `rays_dev`, `out_dev`, `counters_dev` are the device staging buffers of the same shapes. */
int nmf_render_rays_host(const NmfScene* scene, const NmfRender* rp, const float* rays_host, float* rays_dev,
                         const NmfImages* out_host, const NmfImages* out_dev, const NmfCounters* counters_host,
                         const NmfCounters* counters_dev, void* workspace, size_t workspace_bytes, void* stream);

/* ---- plugin-slot operators (unfused; what the reference's duck-typed plugins expose) ---- */

/* AlphaGridSampler.sample in eval mode (samplers/alphagrid.py:131-207, 278-370): dense validity mask
 * ray_valid (n_rays, n_steps) uint8, z_vals (n_rays, n_steps), and the number of valid samples per ray. */
int nmf_sample_rays(const NmfScene* scene, const float* rays, int n_rays, float near_override, uint8_t* ray_valid,
                    float* z_vals, int* n_valid, void* stream);

/* TensorVMSplit.compute_densityfeature (fields/tensoRF.py:392-400 + tensor_base.py:83-93): xyz (n, stride) */
int nmf_vm_density(const NmfScene* scene, const float* xyz, int n, int stride, int activate, float* sigma, void* stream);
/* TensorVMSplit.compute_appfeature (fields/tensoRF.py:402-405): -> (n, 24) */
int nmf_vm_appfeature(const NmfScene* scene, const float* xyz, int n, int stride, float* feat, void* stream);
/* TensorBase.compute_normals (fields/tensor_base.py:107-129, modules/grid_sample_Cinf.py:109-281): -> (n, 3) */
int nmf_vm_normals(const NmfScene* scene, const float* xyz, int n, int stride, float* normals, void* stream);
/* IntegralEquirect.forward (modules/integral_equirect.py:409-504): dirs (n,3), mip (n) -> (n,3) */
int nmf_env_lookup(const NmfScene* scene, const float* dirs, const float* mip, int n, float* rgb, void* stream);
/* GGXSampler.sample (brdf_samplers/ggx.py:61-226) for one bounce ray per row: u (n,2), V (n,3), N (n,3), r (n)
 * -> L (n,3), logpdf (n), half_local (n,3), diff_local (n,3) */
int nmf_ggx_sample(const float* u, const float* V, const float* N, const float* r, int n, float* L, float* logpdf,
                   float* half_local, float* diff_local, void* stream);
/* MLPBRDF.forward (modules/brdf.py:177-261): feat (n,24), half_local (n,3), diff_local (n,3), rough (n) -> (n,3) */
int nmf_brdf_mlp(const NmfScene* scene, const float* feat, const float* half_local, const float* diff_local,
                 const float* rough, int n, float* out, void* stream);
/* RandHydraMLPDiffuse.forward (modules/render_modules.py:519-574): feat (n,24) -> albedo, tint, f0 (n,3), r1 (n) and,
 * when r2 != NULL, the second roughness channel r2 (n) (the render path sets r2 = r1, models/microfacet.py:360) */
int nmf_material_heads(const NmfScene* scene, const float* feat, int n, float* albedo, float* tint, float* f0,
                       float* r1, float* r2, void* stream);
/* occupancy rebuild, alpha stage of AlphaGridSampler.getDenseAlpha (samplers/alphagrid.py:209-247):
 * alpha[z][y][x] = 1 - exp(-sigma(lattice point) * stepsize) on a (gz,gy,gx) lattice spanning the aabb; sigma = 0 where the
 * scene's CURRENT occupancy samples to 0 (compute_alpha, :209-224).  lins (device, optional): gx + gy + gz floats, the
 * caller's torch.linspace(0, 1, g) of the three axes (x | y | z) -- NULL: an emulation of ATen's scalar formula */
int nmf_dense_alpha(const NmfScene* scene, int gx, int gy, int gz, const float* lins, float* alpha, void* stream);

/* ---- the callers either side of the path ---- */

/* Ray generation of one view on the device: dataLoader/ray_utils.py:23-43 (get_ray_directions: pixel centres + 0.5,
 * ((i - cx) / fx, (j - cy) / fy, 1)), dataLoader/blender.py:108-110 (normalise), ray_utils.py:67-89 (get_rays).
 * c2w_host: HOST pointer to the 3x4 row-major camera-to-world matrix already in the OpenCV convention
 * (blender.py:146: transform_matrix @ diag(1,-1,-1,1)).  pixel_ids (device, optional): ray i is pixel pixel_ids[i]
 * (row-major id = y * W + x), e.g. the shuffle of renderer.py:130-132; NULL = all n = H*W pixels in order.
 * rays: (n, 6) device output [origin, unit direction]. */
int nmf_generate_rays(const float* c2w_host, int H, int W, float fx, float fy, float cx, float cy, const int* pixel_ids,
                      int n, float* rays, void* stream);

/* Evaluation metric of renderer.py:399-401 on the device: sum over rays and channels of
 * (floor(clip(rgb, 0, 1) * 255) / 255 - clip(gt, 0, 1))^2 in fp64 -> sum_sq[0] (device).  gt is indexed by
 * pixel_ids[i] when given (rgb is in render order, gt in image order).  PSNR = -10 log10(sum_sq / (3 n)). */
int nmf_image_sq_error(const float* rgb, const float* gt, const int* pixel_ids, int n, double* sum_sq, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training slice, first model: model=tensorf (models/tensorf.py + MLPRender_Fea).
 * ------------------------------------------------------------------------------------------------ */

/* AlphaGridSampler.sample with is_train=True (samplers/alphagrid.py:167-173, 353-364): jittered cumulative steps
 * z = t_min + cumsum(U * stepsize + stepsize / 2) with U keyed by (seed, ray id, step) -- ray id = ray_ids[i] when
 * given (device, uint64), else ray_id0 + i; the cumulative sum is exact in fp64 and rounded to fp32 per prefix (ATen's
 * CPU cumsum), so z is bit-reproducible.  Outputs cover ALL n_rays rays: ray_valid (n, n_steps) uint8, z_vals
 * (n, n_steps), n_valid (n); whole_valid (n) uint8 = rays kept by the dynamic batch truncation (all 1 unless
 * max_samples > 0 and the batch holds more than max_samples valid samples; then cumsum(n_valid) < max_samples);
 * n_kept (int[2]) = {kept rays (a prefix of the batch), valid samples of the kept rays}.  ray_valid may be NULL. */
int nmf_sample_rays_train(const NmfScene* scene, const float* rays, int n_rays, float near_override, uint64_t seed,
                          uint64_t ray_id0, const uint64_t* ray_ids, int max_samples, uint8_t* ray_valid, float* z_vals,
                          int* n_valid, uint8_t* whole_valid, int* n_kept, void* stream);

/* gradients of one training step, fp32, device, CHANNEL-LAST like the factors of NmfScene (the host permutes them
 * back to the reference's (1,C,H,W) parameters); zeroed by nmf_train_plain before it accumulates */
typedef struct NmfPlainGrads {
  float* d_plane[3];       /* [h][w][16]  rf.density_rf.app_plane.p */
  float* d_line[3];        /* [n][16]     rf.density_rf.app_line.p  */
  float* a_plane[3];       /* [h][w][24]  rf.app_rf.app_plane.p     */
  float* a_line[3];        /* [n][24]     rf.app_rf.app_line.p      */
  float* basis_t;          /* [72][24]    rf.basis_mat.weight^T     */
  float* w0t; float* b0;   /* [135][128], [128]  model.diffuse_module.mlp.0 (transposed weight) */
  float* w1t; float* b1;   /* [128][128], [128]  mlp.2 */
  float* w2t; float* b2;   /* [128][3],   [3]    mlp.4 */
} NmfPlainGrads;

typedef struct NmfTrain {
  int n_rays;
  float focal;
  uint64_t seed, ray_id0;
  const uint64_t* ray_ids; /* optional (device): global ray ids that key the jitter */
  int max_samples;         /* AlphaGridSampler.max_samples (<= 0: no truncation) */
  int cap_samples;         /* capacity of the per-sample scratch; more valid samples than this sets out->error */
  float lambda_pred;       /* weight of statistics["prediction_loss"] = 2 * sum(acc) (tensor_nerf.py:598-602) */
  int white_bg;            /* 1: white background (tensor_nerf.py:215,478) */
} NmfTrain;

typedef struct NmfTrainOut {
  float* rgb_map;          /* (n,3) rows of the kept rays (a prefix), rest zero */
  float* acc_map;          /* (n) */
  uint8_t* whole_valid;    /* (n) statistics["whole_valid"] */
  double* loss;            /* [2]: sum of squared clipped errors (train.py:597-601), sum of acc */
  int* n_kept;             /* [2]: kept rays, valid samples (statistics["n_samples"][0]) */
  unsigned* error;         /* [1]: NMF_DEV_E_SURVIVORS when cap_samples was too small (gradients are then invalid) */
} NmfTrainOut;

size_t nmf_train_workspace_bytes(const NmfScene* scene, int n_rays, int cap_samples);

/* One training forward + backward of model=tensorf on a ray batch, fused on the device: TensorNeRF.forward
 * (is_train=True; modules/tensor_nerf.py:210-674) -> models/tensorf.py:70-97 -> the photometric loss of
 * train.py:576-611 (+ lambda_pred * prediction_loss), then the hand-written backward of every stage (compositing,
 * softplus density, VM factors, basis_mat, positional encoding, the 135-128-128-3 MLP) into `grads`.
 * rays (n,6), gt (n,3) device.  No autograd, no host synchronisation. */
int nmf_train_plain(const NmfScene* scene, const NmfTrain* tp, const float* rays, const float* gt,
                    const NmfPlainGrads* grads, const NmfTrainOut* out, void* workspace, size_t workspace_bytes,
                    void* stream);

/* Training FORWARD of the microfacet model: TensorNeRF.forward(is_train=True, draw_debug=False)
 * (modules/tensor_nerf.py:210-674 with models/microfacet.py:271-673) through the fused render kernels --
 * jittered cumulative distances at recur 0 and in the re-traced rays (samplers/alphagrid.py:167-173; the recursion
 * passes is_train on, tensor_nerf.py:303), the dynamic batch truncation at recur 0 only (alphagrid.py:353-364; the
 * recursion runs with dynamic_batch_size=False, tensor_nerf.py:299), the bounce roughness clipped from below
 * (microfacet.py:361-363).  Loss-side quantities come back as in eval mode: rgb_map / acc_map rows of the kept rays
 * (a prefix of the batch: rows >= n_kept[0] are background), NmfCounters.stat4 = the A19 regulariser sums,
 * n_samples0 / n_samples1 = statistics["n_samples"].  The batch is ONE forward call: rp->n_rays <= rp->chunk.
 * The jitter is keyed by (rp->seed, rp->ray_id0 + i, step).  Forward only: the microfacet backward is not built. */
typedef struct NmfRenderTrain {
  int max_samples;         /* AlphaGridSampler.max_samples (<= 0: no truncation) */
  float min_rough;         /* Microfacet.min_rough */
  uint8_t* whole_valid;    /* (n) out, device: statistics["whole_valid"] */
  int* n_kept;             /* [2] out, device: kept rays, valid primary samples of the kept rays */
} NmfRenderTrain;
size_t nmf_render_train_workspace_bytes(const NmfScene* scene, int n_rays, float cap_scale);
int nmf_render_rays_train(const NmfScene* scene, const NmfRender* rp, const NmfRenderTrain* tr, const float* rays,
                          const NmfImages* out, const NmfCounters* counters, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ---- optimiser step on the device (the update half of train.py:497-813) ----
 * All three stream their tensors once; `sum_abs` / `sq_norm` are device fp64 accumulators the caller zeroes. */

/* TensorVMSplit.density_L1 (fields/tensoRF.py:332-340) for one factor: sum_abs += sum |param|, and when grad != NULL
 * grad += coef * sign(param) with coef = L1_reg_weight / numel (train.py:675-678: the mean over the factor). */
int nmf_l1_reg(const float* param, size_t n, float coef, float* grad, double* sum_abs, void* stream);
/* sq_norm += sum grad^2 in fp64: the total norm torch.nn.utils.clip_grad_norm_ (train.py:752-753) needs */
int nmf_grad_sq_norm(const float* grad, size_t n, double* sq_norm, void* stream);
/* One torch.optim.Adam update (train.py:443-457, 754; amsgrad off, L2 weight_decay) of one parameter segment, fused
 * with the loss normalisation and the gradient clipping: g = grad * grad_scale * min(1, max_norm / (grad_scale *
 * sqrt(*sq_norm) + 1e-6)).  sq_norm (device) may be NULL (no clipping); step >= 1 is the update count of this
 * optimiser instance (bias correction); lr already includes the LambdaLR factor (train.py:458-466). */
typedef struct NmfAdam {
  float lr, beta1, beta2, eps, weight_decay;
  int step;
  float grad_scale;        /* 1 / lbatch_size (train.py:709) */
  float max_norm;          /* params.clip_grad; <= 0: off */
  /* optional device memory { grad_scale, skip }: when not NULL the loss normaliser is read from control[0] instead of
   * grad_scale (the number of kept rays of a training step stays on the device) and control[1] != 0 turns the whole update into
   * a no-op (a device-side list overflowed during the step: the host repeats the iteration with larger buffers) */
  const float* control;
} NmfAdam;
int nmf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, const NmfAdam* adam,
                  const double* sq_norm, void* stream);

/* Reverse pass of IntegralEquirect lookups w.r.t. the map (modules/integral_equirect.py:263-273 activation_fn / calc_sat,
 * 409-504 forward, under autograd) -- the environment stage of the microfacet backward.
 * Step 1, per batch of lookups: dirs (n,3), mip (n), g (n,3) = d loss / d rgb are scattered with the forward's own box
 * walk into gsat = [h][w][4] floats followed by 8 floats (pole-row terms: top rgb, pad, bottom rgb, pad); the caller
 * zeroes gsat (h*w*4 + 8 floats) before the first batch of an optimiser step; batches accumulate. */
int nmf_env_lookup_bwd_scatter(const NmfScene* scene, const float* dirs, const float* mip, const float* g, int n,
                               float* gsat, void* stream);
/* Step 2, once per optimiser step: adjoint of the double cumsum (two reverse prefix sums, gsat is overwritten), the
 * pole-row means and the clipped exp activation: d_bg_mat (3,h,w) += ..., and when not NULL d_brightness[0] += ...,
 * d_mul[0] += ... (device scalars).  bg_mat is the (3,h,w) parameter, brightness / mul its scalar parameters. */
int nmf_env_lookup_bwd_finish(float* gsat, int h, int w, const float* bg_mat, float brightness, float mul,
                              float* d_bg_mat, float* d_brightness, float* d_mul, void* stream);
/* The same with the parameters read from device memory: scalars_dev = { brightness, mul, mipbias } (fp32), so that a training
 * iteration never needs them on the host. */
int nmf_env_lookup_bwd_finish_dev(float* gsat, int h, int w, const float* bg_mat, const float* scalars_dev, float* d_bg_mat,
                                  float* d_brightness, float* d_mul, void* stream);

/* d loss / d IntegralEquirect.mipbias of a batch of lookups (the box size moves with the bias: sa2mip,
 * integral_equirect.py:373-397): d_mipbias[0] (device float) += sum_i g_i . d rgb_i / d mipbias. */
int nmf_env_lookup_bwd_mipbias(const NmfScene* scene, const float* dirs, const float* mip, const float* g, int n,
                               float* d_mipbias, void* stream);

/* Reverse pass of TensorBase.compute_normals (fields/tensor_base.py:107-129 -> modules/grid_sample_Cinf.py:109-281 under
 * autograd) w.r.t. the density planes / lines -- the normal stage of the microfacet backward (detach_N off) and of ori_loss.
 * Gradient images are laid out like the scene's derivative-packed factors and zeroed by the caller before the first batch
 * of an optimiser step: gpack[p] = [h][w][48] (d value | d dx | d dy), glpack[p] = [n][4][8] (val4 | dy4). */
typedef struct NmfNormalGrads {
  float* gpack[3];
  float* glpack[3];
} NmfNormalGrads;
/* Step 1, per batch: xyz (n, stride), d_normals (n,3) = d loss / d compute_normals(xyz); batches accumulate. */
int nmf_vm_normals_bwd_scatter(const NmfScene* scene, const float* xyz, int n, int stride, const float* d_normals,
                               const NmfNormalGrads* imgs, void* stream);
/* Step 2, once per optimiser step: adjoint of the 5x5 smoothed-difference stencils kx25 / ky25 (device, row-major 5x5;
 * grid_sample_Cinf.py:218-242) -> d_plane[p] [h][w][16] += ..., d_line[p] [n][16] += ... (channel-last, the layout of
 * NmfPlainGrads.d_plane / d_line). */
int nmf_vm_normals_bwd_finish(const NmfScene* scene, const NmfNormalGrads* imgs, const float* kx25, const float* ky25,
                              float* const* d_plane, float* const* d_line, void* stream);

/* Reverse pass of RandHydraMLPDiffuse.forward (modules/render_modules.py:519-574 under autograd): feat (n,24) and the
 * upstream gradients g_albedo (n,3) (= d loss / d diffuse), g_f0 (n,3), g_rough (n) (= d loss / d r1)
 * -> d_head_w [11][24] += ..., d_head_b [11] += ... (rows as NmfScene.head_w: diffuse 3, tint 3 (always zero: the tint
 * head feeds nothing on this path), f0 3, roughness 2), d_feat (n,24) written. */
int nmf_material_heads_bwd(const NmfScene* scene, const float* feat, const float* g_albedo, const float* g_f0,
                           const float* g_rough, int n, float* d_head_w, float* d_head_b, float* d_feat, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Training step of the MICROFACET model (configs #3 / #4): forward + loss + reverse pass.
 * Replaces, for model=microfacet_tensorf2, what train.py:540-712 does with autograd: TensorNeRF.forward(is_train=True)
 * (modules/tensor_nerf.py:210-674 -> models/microfacet.py:271-673 -> brdf_samplers/ggx.py:61-226, modules/brdf.py:177-261,
 * modules/render_modules.py:519-574, modules/integral_equirect.py:409-504, fields/tensor_base.py:107-129 with
 * create_graph=True, one re-traced level), the loss  sum (clip(rgb) - clip(gt))^2 + lambda_pred * prediction_loss +
 * lambda_ori * ori_loss  (train.py:586-657 with the lambdas of configs/model/microfacet_tensorf2.yaml:205-218; the other
 * regularisers have weight 0 there), and total_loss.backward().
 * ------------------------------------------------------------------------------------------------ */
typedef struct NmfMicrofacetGrads {   /* device, fp32; every field ACCUMULATES: the caller zeroes once per optimiser step */
  float* d_plane[3];       /* [h][w][16]  rf.density_rf.app_plane.p   (channel-last, like NmfPlainGrads)              */
  float* d_line[3];        /* [n][16]     rf.density_rf.app_line.p                                                     */
  float* a_plane[3];       /* [h][w][24]  rf.app_rf.app_plane.p                                                        */
  float* a_line[3];        /* [n][24]     rf.app_rf.app_line.p                                                         */
  float* basis_t;          /* [72][24]    rf.basis_mat.weight^T                                                        */
  float* head_w;           /* [11][24]    model.diffuse_module.{diffuse,tint,f0,roughness}_mlp.0.weight (rows 3|3|3|2)  */
  float* head_b;           /* [11]                                                                                     */
  float* w0t; float* b0;   /* [66][64], [64]  model.brdf.mlp.0 (transposed weight)                                     */
  float* w1t; float* b1;   /* [64][64], [64]  model.brdf.mlp.2                                                         */
  float* w2t; float* b2;   /* [64][4],  [4]   model.brdf.mlp.4                                                         */
  float* gsat;             /* [env_h][env_w][4] + 8: scatter image of the environment lookups; nmf_env_lookup_bwd_finish
                            * turns it into d bg_mat / d brightness / d mul once per optimiser step                    */
  float* d_mipbias;        /* [1]  bg_module.mipbias                                                                   */
  NmfNormalGrads normals;  /* derivative-plane gradient images of the normal path; nmf_vm_normals_bwd_finish adds their
                            * stencil adjoint to d_plane / d_line once per optimiser step                              */
} NmfMicrofacetGrads;

typedef struct NmfMicrofacetTrain {
  float lambda_pred;       /* params.pred_lambda (3e-4): weight of prediction_loss = 2 sum(acc) (no normal module)     */
  float lambda_ori;        /* params.ori_lambda (0.1): weight of ori_loss = sum w min(v.n, 0)^2                        */
  int detach_N;            /* Microfacet.detach_N (models/microfacet.py:117-118, 352-353)                              */
  double* loss;            /* [3] out, device: photometric sum, sum of acc (prediction_loss / 2), ori_loss             */
} NmfMicrofacetTrain;

/* One forward + backward of a ray batch (rp->n_rays <= rp->chunk, as nmf_render_rays_train; rp->skip_eps / t_cut are
 * ignored: the reference shades every sample of positive weight).  rays (n,6), gt (n,3) device: gt row i belongs to ray i
 * (the kept rays are a prefix).  out / counters / tr as in nmf_render_rays_train (rgb_map and acc_map rows of the kept
 * rays, whole_valid, n_kept, n_samples, the A19 sums).  workspace: nmf_render_train_workspace_bytes().  Asynchronous on
 * `stream`; device-side list overflows are reported in NmfCounters.error (gradients are then incomplete: grow and repeat). */
int nmf_train_microfacet(const NmfScene* scene, const NmfRender* rp, const NmfRenderTrain* tr, const NmfMicrofacetTrain* tp,
                         const float* rays, const float* gt, const NmfMicrofacetGrads* grads, const NmfImages* out,
                         const NmfCounters* counters, void* workspace, size_t workspace_bytes, void* stream);

/* Resolution schedule (fields/tensor_base.py:234-243 -> fields/tensoRF.py:208-227, 408-413): TensoRF.upsample is
 * F.interpolate(mode="bilinear", align_corners=True) of every factor.  src (C,H,W) -> dst (C,H2,W2), both in the
 * reference's own parameter layout (a line (1,C,N,1) is H = N, W = 1). */
int nmf_upsample_bilinear(const float* src, int C, int H, int W, float* dst, int H2, int W2, void* stream);

/* ---- scene re-pack: rebuilt from the parameters after every optimiser step of a training run (csrc/nmf_repack.cu) ---- */

/* One factor in the reference's layout -- a plane (1,C,H,W) or a line (1,C,N,1) passed as H = N, W = 1 -- into the
 * channel-last gather layouts of NmfScene: val [H][W][C] (C = 24: [H][W][NMF_APP_STRIDE], padding untouched) and, for the
 * density factors (C = 16), pack = [H][W][val | dx | dy]
 * (lines: [N][4][val4 | dy4]) with the smoothed-difference planes of modules/grid_sample_Cinf.py:218-242 (5x5
 * cross-correlation with kx25 / ky25, zero padding 2).  val or pack may be NULL. */
int nmf_pack_factor(const float* src, int C, int H, int W, const float* kx25, const float* ky25, float* val, float* pack,
                    void* stream);

/* IntegralEquirect tables (modules/integral_equirect.py:263-273, 431-433, 498-502): act = exp(min(brightness + mul *
 * bg_mat, 20)) (optional out, (3,h,w)), sat4 [h][w][4] = cumsum_x(cumsum_y(act / 1000)) with fp64 accumulation and a
 * rounding to fp32 after each scan (ATen's CPU cumsum), pole_sums[6] (device, fp64) = sums of the first / last row of act
 * per channel.  scratch_c1: 3*h*w floats. */
int nmf_env_pair_sat(const float* sat4, int h, int w, float* sat8, void* stream);    /* NmfScene.env_sat2 from sat4 */
/* nmf_env_build_sat with device-resident parameters: scalars_dev = { brightness, mul, mipbias } (fp32 device memory); also writes
 * env_dyn[7] = { mipbias, pole_sums[0..2] / w, pole_sums[3..5] / w } = what NmfScene.env_dyn points at. */
int nmf_env_build_sat_dev(const float* bg_mat, int h, int w, const float* scalars_dev, float* scratch_c1, float* act, float* sat4,
                          double* pole_sums, float* env_dyn, void* stream);
int nmf_env_build_sat(const float* bg_mat, int h, int w, float brightness, float mul, float* scratch_c1, float* act,
                      float* sat4, double* pole_sums, void* stream);

/* AlphaGridSampler.updateAlphaMask after the dense alpha (samplers/alphagrid.py:256-261): clamp, 3^3 max-pool (padding 1),
 * threshold -> the occupancy bit-fields of NmfScene (vox, cell: gz*gy*pitch/32 words each; coarse: optional,
 * ceil(gx/8)*ceil(gy/8)*ceil(gz/8) bits) and, optionally, the 0/1 float volume (gz,gy,gx) the reference keeps. */
int nmf_occupancy_from_alpha(const float* alpha, int gx, int gy, int gz, float thres, int pitch, uint32_t* vox, uint32_t* cell,
                             uint32_t* coarse, float* volume, void* stream);

/* Material heads and BRDF MLP of NmfScene from the reference's parameters in one launch (after every optimiser step):
 * w[i] / b[i] = model.brdf.mlp.{0,2,4}.weight (n_out, n_in) / bias (modules/brdf.py:73-120; 64x66, 64x64, 4x64); outputs:
 * wt[i] = the transposed fp32 weights (brdf_w{i}t), bo[i] = bias copies (brdf_b{i}), w16[i] / wbf[i] = the tensor-core operand
 * tiles brdf_w{i}u (fp16) / brdf_w{i}b (bf16): [80/8][rows][8], rows = 64, 64, 16, bias in input column 66, zero padding.
 * head_w[h] / head_b[h] = model.diffuse_module.{diffuse,tint,f0,roughness}_mlp.0.weight (rows,24) / bias
 * (modules/render_modules.py:519-574), concatenated into head_w_out [11][24] / head_b_out [11].  NULL w[i] / head_w[h]: skipped. */
typedef struct NmfShadingPack {
  const float* w[3];
  const float* b[3];
  int32_t n_out[3], n_in[3];
  const float* head_w[4];
  const float* head_b[4];
  int32_t head_rows[4];
  float* wt[3];
  float* bo[3];
  void* w16[3];
  void* wbf[3];
  float* head_w_out;
  float* head_b_out;
} NmfShadingPack;
int nmf_pack_shading(const NmfShadingPack* p, void* stream);

/* Gradient hand-over after a training step: the kernels accumulate gradients channel-last ([texel][channel]); the
 * reference's parameters are channel-first ((1,C,H,W) planes, (1,C,N,1) lines, (out,in) linear weights; fields/tensoRF.py:
 * 42-75, modules/brdf.py:73-120).  One launch moves every tensor: job j writes dst[c * n + i] = src[i * c + ch] (c = 1: a copy).
 * jobs_dev: device array.  c <= 64. */
typedef struct NmfTransposeJob {
  const float* src;
  float* dst;
  uint64_t n;      /* rows of src (texels) */
  int32_t c;       /* columns of src (channels) */
  int32_t pad;
} NmfTransposeJob;
int nmf_transpose_batch(const NmfTransposeJob* jobs_dev, int n_jobs, int blocks_per_job, void* stream);

/* Measurement helper (bench.py): `n_threads` threads (a multiple of 256) each issue `taps` (a multiple of 8) independent
 * 16-byte loads over `buf` (n_elems float4) and write one float4 of `sink` (n_threads float4); `group` (1, 2, 4, 8)
 * consecutive lanes read consecutive pieces of one pseudo-random segment of 16 * group bytes (the granularity of the
 * kernels' texel taps).  Timed by the caller; bytes = n_threads * taps * 16.  With an L2-resident buffer this is the
 * gather ceiling k_march / k_shade are reported against (their factor set is L2-resident: the HBM copy rate is not
 * their roofline). */
int nmf_bench_gather(const void* buf, size_t n_elems, int taps, int group, int n_threads, void* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NMF_B200_H */
