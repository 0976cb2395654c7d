"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference (half-potato/nmf) from /root/reference.

This file is not part of the product.  It only works inside the build container (where
`/root/reference` is mounted read-only); it does not travel to the GPU box.  It is used by
`oracle/make_golden.py` to (a) pin `oracle/nmf_oracle.py` (the CPU restatement) against the
reference's own PyTorch code and (b) generate the small golden fixtures under `tests/golden/`.

Nothing under `nmf_b200/` may import this module.

The reference needs a few non-arithmetic third-party modules that are absent here (hydra,
icecream, imageio, ...).  They are stubbed in `sys.modules` before the import;
none of them participates in hot-path arithmetic.
"""
import functools
import importlib
import os
import sys
import types
import warnings

import torch  # must be imported BEFORE the permissive `warp` stub is installed

REFERENCE_ROOT = os.environ.get("NMF_REFERENCE_ROOT", "/root/reference")


class _Any:
    """Permissive object: any attribute / call returns itself; used as decorator it is identity."""

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return self

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return self


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "modules"))


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    warnings.filterwarnings("ignore", category=FutureWarning)
    for name in ["icecream", "hydra", "hydra.utils", "omegaconf", "imageio", "matplotlib",
                 "matplotlib.pyplot", "plotly", "plotly.express", "plotly.graph_objects", "plyfile",
                 "skimage", "skimage.measure", "kornia", "lpips"]:
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["icecream"].ic = lambda *a, **k: (a[0] if len(a) == 1 else a)
    if not hasattr(sys.modules["kornia"], "create_meshgrid"):
        def create_meshgrid(H, W, normalized_coordinates=False):
            # only dataLoader/ray_utils.py uses it: (1,H,W,2) with [...,0]=x, [...,1]=y
            ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32),
                                    torch.arange(W, dtype=torch.float32), indexing="ij")
            return torch.stack([xs, ys], -1)[None]
        sys.modules["kornia"].create_meshgrid = create_meshgrid
    if "warp" not in sys.modules:
        w = types.ModuleType("warp")

        def _wg(n):
            if n.startswith("__"):
                raise AttributeError(n)
            return _Any()

        w.__getattr__ = _wg
        sys.modules["warp"] = w
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def build_reference_model(aabb, near_far, grid_size=(300, 300, 300), bg_resolution=512, seed=0,
                          model_name="microfacet_tensorf2"):
    """Hand-instantiates the reference TensorNeRF exactly as hydra's `_partial_` would
    (kwargs = yaml keys of configs/model/microfacet_tensorf2.yaml:5-157 and configs/field/tensorf.yaml:4-48).
    """
    install_stubs()
    from modules.tensor_nerf import TensorNeRF
    from samplers.alphagrid import AlphaGridSampler
    from fields.tensoRF import TensorVMSplit
    from models.microfacet import Microfacet
    from brdf_samplers.ggx import GGXSampler
    from modules.brdf import MLPBRDF
    from modules.ish import ListISH
    from modules.render_modules import RandHydraMLPDiffuse, MLPRender_Fea
    from modules.integral_equirect import IntegralEquirect
    from modules.tonemap import SRGBTonemap
    from models.tensorf import TensoRF as PlainTensoRF

    P = functools.partial
    torch.manual_seed(seed)
    rf = P(TensorVMSplit, distance_scale=25, density_n_comp=16, appearance_n_comp=24, app_dim=24,
           step_ratio=0.5, density_res_multi=1, contract_space=False, smoothing=1, activation="softplus",
           interp_mode="bilinear", init_mode="rand", d_init_val=0.1, app_init_val=0.1, density_shift=-4,
           numer_grad=True, dbasis=False, grid_size=torch.tensor(list(grid_size)), N_voxel_init=262144,
           N_voxel_final=27000000, upsamp_list=[500, 1000, 2000, 3000, 4000, 5500, 7000], lr=2e-2,
           lr_net=1e-3, triplanar=False, num_pretrain=0, calibrate=False)
    sampler = P(AlphaGridSampler, enable_alpha_mask=True, update_list=[2000, 3000, 4000, 5500, 7000],
                max_samples=200000)
    if model_name == "microfacet_tensorf2":
        model = P(Microfacet, percent_bright=0.0, min_rough_start=0.0, min_rough_decay=0.999,
                  max_brdf_rays=[650000, 450000], conserve_energy=True, target_num_samples=[1000000],
                  russian_roulette=False, max_retrace_rays=[1000], start_std=0.0, std_decay=1.0,
                  cold_start_bg_iters=0, detach_N_iters=0, anoise=0.25, no_emitters=True,
                  diffuse_mixing_mode="fresnel", freeze=False, rays_per_ray=128, test_rays_per_ray=128,
                  brdf_sampler=P(GGXSampler),
                  brdf=P(MLPBRDF, mul_LdotN=False, feape=0, dotpe=-1, h_encoder=ListISH(degs=[0, 1, 2, 4]),
                         d_encoder=ListISH(degs=[0, 1, 2, 4]), hidden_w=64, num_layers=3,
                         initializer="kaiming", bias=0, activation="sigmoid", lr=1e-3),
                  diffuse_module=P(RandHydraMLPDiffuse, pospe=-1, feape=0, roughness_view_encoder=None,
                                   roughness_cfg=dict(hidden_w=64, num_layers=1), hidden_w=64, num_layers=1,
                                   initializer="xavier_sigmoid", lr=1e-3, start_roughness=0.35, tint_bias=0,
                                   diffuse_bias=-0.619, diffuse_mul=1.5, roughness_bias=-1),
                  visibility_module=None)
        eval_batch_size = 4096
    elif model_name == "tensorf":
        model = P(PlainTensoRF, diffuse_module=P(MLPRender_Fea, featureC=128, viewpe=2, feape=2))
        eval_batch_size = 10240
    else:
        raise ValueError(model_name)
    bg_module = IntegralEquirect(bg_resolution=bg_resolution, mipbias=1, activation="exp", lr=0.02,
                                 init_val=-0.6, mul_lr=0, brightness_lr=0, betas=[0.9, 0.99],
                                 mul_betas=[0.9, 0.9], mipbias_lr=1e-4, mipnoise=0.0)
    t = TensorNeRF(rf=rf, model=model, aabb=aabb, near_far=near_far, sampler=sampler, tonemap=SRGBTonemap(),
                   bg_module=bg_module, recur_alpha_thres=1e-3, lr_scale=1, infinity_border=False,
                   eval_batch_size=eval_batch_size, recur_stepmul=0.5, hdr=False, bg_noise=0.0,
                   bg_noise_decay=0.999, use_predicted_normals=False, orient_world_normals=True,
                   align_pred_norms=True, detach_inter=False, geonorm_iters=-1, geonorm_interp_iters=1000,
                   contraction="AABB")
    t.sampler.update(t.rf, init=True)
    return t
