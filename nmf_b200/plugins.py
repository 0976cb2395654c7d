"""Host-side mirror of the reference's hydra plugin slots for the render path.

Every class keeps the reference's constructor keywords, attribute names and state_dict keys, so that a reference
config (`_target_` remapped by nmf_b200/config.py) and a reference checkpoint load unchanged; the arithmetic is done
by the sm_100a kernels behind the C ABI (nmf_b200/ops.py).  Reference classes mirrored (file:line in the reference):

    TensorNeRF            modules/tensor_nerf.py:36-674          (composition root; forward = fused CUDA path)
    AlphaGridSampler      samplers/alphagrid.py:63-370, AlphaGridMask :6-60
    TensorVMSplit         fields/tensoRF.py:255-445 + fields/tensor_base.py:32-252
    Microfacet            models/microfacet.py:18-673
    PlainTensoRF          models/tensorf.py:9-97                  (model=tensorf plumbing config)
    GGXSampler            brdf_samplers/ggx.py:60-268, brdf_samplers/base.py:3-24
    MLPBRDF / ListISH     modules/brdf.py:72-261, modules/ish.py:94-105
    RandHydraMLPDiffuse   modules/render_modules.py:447-574,  MLPRender_Fea :201-235
    IntegralEquirect      modules/integral_equirect.py:176-504
    SRGBTonemap           modules/tonemap.py:38-49

Scope: eval-mode forward (the hot path of BASELINE.json); the training FORWARD (`is_train=True`: jittered steps,
dynamic batch truncation, A19 statistics) for both models; forward + backward (`TensorNeRF.train_step`) for
model=tensorf.  The microfacet backward is not built: `train_step` raises for it.
"""
import math

import torch
import torch.nn as nn

from . import _lib, ops
from .scene import DEFAULT_HP, DeviceScene, step_size_and_count


def _n_to_reso(n_voxels, bbox):
    """utils.N_to_reso (utils.py:55-58)"""
    xyz_min, xyz_max = bbox
    dim = len(xyz_min)
    voxel_size = ((xyz_max - xyz_min).prod() / n_voxels).pow(1 / dim)
    return ((xyz_max - xyz_min) / voxel_size).long().tolist()


class TVLoss(nn.Module):
    """utils.TVLoss (utils.py:139-151): total variation of a plane (1,C,H,W) or a line (1,C,N,1)."""

    def forward(self, x):
        if x.shape[-1] == 1:
            return (x[:, :, 1:, :] - x[:, :, :-1, :]).abs().mean()
        h_tv = x[:, :, 1:, :-1] - x[:, :, :-1, :-1]
        w_tv = x[:, :, :-1, 1:] - x[:, :, :-1, :-1]
        return (w_tv ** 2 + h_tv ** 2 + 1e-5).sqrt().mean()


class SRGBTonemap(nn.Module):
    def forward(self, img, noclip=False):
        limit = 0.0031308
        out = torch.where(img > limit, 1.055 * (img.clip(min=limit) ** (1.0 / 2.4)) - 0.055, 12.92 * img)
        return out if noclip else out.clip(0, 1)

    def inverse(self, img):
        limit = 0.04045
        return torch.where(img > limit, ((img + 0.055) / 1.055) ** 2.4, img / 12.92)


# ------------------------------------------------------------------------------------------------------------
# field
# ------------------------------------------------------------------------------------------------------------
class _Factors(nn.Module):
    """Parameter container with the key names of fields/tensoRF.py:28-75 (app_plane.{i}, app_line.{i})."""

    def __init__(self, grid_size, dim, init_val):
        super().__init__()
        G = int(grid_size)
        mat, vec = [[0, 1], [0, 2], [1, 2]], [2, 1, 0]
        self.app_plane = nn.ParameterList([nn.Parameter(init_val * torch.rand(1, dim, G, G)) for _ in vec])
        self.app_line = nn.ParameterList([nn.Parameter(init_val * torch.rand(1, dim, G, 1)) for _ in vec])
        self._dim = dim

    def dim(self):
        return self._dim * 3


class TensorVMSplit(nn.Module):
    """rf slot.  Constructor keywords of configs/field/tensorf.yaml."""

    def __init__(self, aabb, grid_size=None, density_n_comp=16, appearance_n_comp=24, app_dim=24, step_ratio=0.5,
                 density_res_multi=1, contract_space=False, smoothing=1, activation="softplus", interp_mode="bilinear",
                 init_mode="rand", d_init_val=0.1, app_init_val=0.1, density_shift=-4, numer_grad=True, dbasis=False,
                 N_voxel_init=262144, N_voxel_final=27000000, upsamp_list=(), lr=2e-2, lr_net=1e-3, triplanar=False,
                 num_pretrain=0, calibrate=False, distance_scale=25, **kwargs):
        super().__init__()
        if activation != "softplus" or interp_mode != "bilinear" or dbasis or triplanar or contract_space or smoothing != 1:
            raise _lib.NmfError("TensorVMSplit: the kernels implement activation=softplus, interp_mode=bilinear, "
                                "dbasis=False, triplanar=False, contract_space=False, smoothing=1 (field=tensorf)")
        if density_n_comp != 16 or appearance_n_comp != 24 or app_dim != 24:
            raise _lib.NmfError("TensorVMSplit: kernels are compiled for density_n_comp=16, appearance_n_comp=24, app_dim=24")
        aabb = torch.as_tensor(aabb, dtype=torch.float32)
        self.lr, self.lr_net = lr, lr_net
        self.activation, self.num_pretrain, self.density_shift = activation, num_pretrain, density_shift
        self.contract_space, self.distance_scale, self.calibrate = contract_space, distance_scale, calibrate
        self.density_n_comp, self.app_n_comp, self.app_dim = [density_n_comp] * 3, [appearance_n_comp] * 3, app_dim
        self.step_ratio, self.separate_appgrid, self.smoothing = step_ratio, True, smoothing
        self.upsamp_list = list(upsamp_list)
        self.N_voxel_list = (torch.round(torch.linspace(N_voxel_init ** (1 / 3), N_voxel_final ** (1 / 3),
                                                        len(self.upsamp_list) + 1) ** 3).long()).tolist()[1:]
        self.register_buffer("aabb", aabb)
        self.register_buffer("aabbSize", aabb[1] - aabb[0])
        self.register_buffer("invaabbSize", 2.0 / (aabb[1] - aabb[0]))
        self.register_buffer("aabbDiag", torch.sqrt(torch.sum(torch.square(aabb[1] - aabb[0]))))
        grid_size = torch.tensor(_n_to_reso(N_voxel_init, aabb)) if grid_size is None else torch.as_tensor(grid_size)
        self.update_stepSize(grid_size)
        G = int(self.grid_size[0])
        self.density_rf = _Factors(G, density_n_comp, d_init_val)
        self.app_rf = _Factors(G, appearance_n_comp, app_init_val)
        self.basis_mat = nn.Linear(3 * appearance_n_comp, app_dim, bias=False)
        self.dbasis_mat = nn.Linear(3 * density_n_comp, 1, bias=False)
        self._scene, self._scene_key = None, None

    def update_stepSize(self, grid_size):
        """fields/tensor_base.py:219-232"""
        grid_size = torch.as_tensor(grid_size).long().to(self.aabb.device)
        self.register_buffer("grid_size", grid_size)
        self.register_buffer("units", self.aabbSize / (grid_size - 1))
        self.register_buffer("stepsize", torch.min(self.units) * self.step_ratio)
        self.nSamples = int((self.aabbDiag / self.stepsize).item()) + 1

    def get_device(self):
        return self.aabbSize.device

    def get_optparam_groups(self, lr_scale=1):
        g = []
        for rf in (self.density_rf, self.app_rf):
            g += [{"params": rf.app_plane.parameters(), "lr": self.lr * lr_scale},
                  {"params": rf.app_line.parameters(), "lr": self.lr * lr_scale}]
        g += [{"params": self.basis_mat.parameters(), "lr": self.lr_net * lr_scale}]
        return g

    def check_schedule(self, iter, batch_mul):
        """fields/tensor_base.py:234-243: at the iterations of upsamp_list the factors are resampled to the next
        resolution of N_voxel_list; True tells the caller to rebuild the optimiser (train.py:806-809)."""
        upsamp_list = [i * batch_mul for i in self.upsamp_list]
        if iter in upsamp_list:
            self.upsample_volume_grid(_n_to_reso(self.N_voxel_list[upsamp_list.index(iter)], self.aabb))
            return True
        return False

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        """fields/tensoRF.py:208-227, 408-413: every plane (grid[mat1], grid[mat0]) and line (grid[vec]) through the
        CUDA bilinear resize (nmf_upsample_bilinear = F.interpolate(align_corners=True)), then update_stepSize."""
        res = [int(r) for r in res_target]
        mat, vec = [[0, 1], [0, 2], [1, 2]], [2, 1, 0]
        for rf in (self.app_rf, self.density_rf):
            for i in range(3):
                rf.app_plane[i] = nn.Parameter(ops.upsample_bilinear(rf.app_plane[i].data, (res[mat[i][1]], res[mat[i][0]])))
                rf.app_line[i] = nn.Parameter(ops.upsample_bilinear(rf.app_line[i].data, (res[vec[i]], 1)))
        self.update_stepSize(res)
        self._scene = None

    def normalize_coord(self, xyz_sampled):
        coords = (xyz_sampled[..., :3] - self.aabb[0]) * self.invaabbSize - 1
        return torch.cat((coords, xyz_sampled[..., 3:4]), dim=-1)

    def field_state(self, prefix="rf."):
        return {prefix + k: v for k, v in self.state_dict().items()}

    def _param_key(self):
        return tuple(p._version for p in self.parameters()) + (tuple(self.grid_size.tolist()), str(self.get_device()),
                                                               float(self.density_shift), float(self.distance_scale))

    def scene(self):
        """DeviceScene holding only the factors (field-only plugin calls); rebuilt when a parameter changed."""
        key = self._param_key()
        if self._scene is None or key != self._scene_key:
            self._scene = DeviceScene(self.field_state(), self.aabb, (0.0, 1.0), self.grid_size.tolist(),
                                      device=self.get_device(), model="field", distance_scale=self.distance_scale,
                                      density_shift=self.density_shift, step_ratio=self.step_ratio)
            self._scene_key = key
        return self._scene

    def compute_densityfeature(self, xyz_sampled, activate=True):
        return ops.vm_density(self.scene(), xyz_sampled, activate)

    def compute_appfeature(self, xyz_sampled):
        return ops.vm_appfeature(self.scene(), xyz_sampled)

    def compute_normals(self, xyz_sampled):
        return ops.vm_normals(self.scene(), xyz_sampled)

    @torch.no_grad()
    def accumulate_normals_grad(self, xyz_sampled, grad_normals):
        """Reverse pass of `compute_normals` for one batch (what autograd does through tensor_base.py:107-129 and
        grid_sample_Cinf.py:109-281): grad_normals (n,3) = d loss / d compute_normals(xyz_sampled).  Batches accumulate on
        the device until `finish_normals_grad`."""
        sc = self.scene()
        acc = getattr(self, "_normals_acc", None)
        if acc is None or acc.scene is not sc:
            acc = self._normals_acc = ops.NormalsGrad(sc)
        acc.scatter(xyz_sampled, grad_normals)

    @torch.no_grad()
    def finish_normals_grad(self):
        """Once per optimiser step: adjoint of the smoothed-difference stencil, added to density_rf.app_plane[p].grad /
        app_line[p].grad (created when absent, like autograd's accumulation)."""
        acc = getattr(self, "_normals_acc", None)
        if acc is None:
            return
        d_plane, d_line = acc.finish()
        for params, grads in ((self.density_rf.app_plane, d_plane), (self.density_rf.app_line, d_line)):
            for prm, g in zip(params, grads):
                g = g.to(prm.dtype)
                prm.grad = g if prm.grad is None else prm.grad + g

    def feature2density(self, density_features):
        return torch.nn.functional.softplus(density_features.clamp(-15, 1e3) + self.density_shift)

    def density_L1(self):
        return sum(torch.mean(torch.abs(p)) + torch.mean(torch.abs(l))
                   for p, l in zip(self.density_rf.app_plane, self.density_rf.app_line))

    def TV_loss_density(self, reg):
        """fields/tensoRF.py:342-350 (weight 0 in the shipped configs: plain tensor ops, not a hot path)"""
        return sum(reg(p) * 1e-2 + reg(l) * 1e-3 for p, l in zip(self.density_rf.app_plane, self.density_rf.app_line))

    def TV_loss_app(self, reg, start_ind=0, end_ind=-1):
        """fields/tensoRF.py:352-360"""
        return sum(reg(p) * 1e-2 + reg(l) * 1e-3 for p, l in zip(self.app_rf.app_plane, self.app_rf.app_line))


# ------------------------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------------------------
class AlphaGridMask(nn.Module):
    def __init__(self, aabb, alpha_volume, align_corners=True):
        super().__init__()
        self.register_buffer("aabb", aabb)
        aabbSize = self.aabb[1] - self.aabb[0]
        self.register_buffer("grid_size", torch.LongTensor([alpha_volume.shape[-1], alpha_volume.shape[-2], alpha_volume.shape[-3]]))
        self.register_buffer("invgrid_size", 1.0 / aabbSize * 2)
        self.register_buffer("alpha_volume", alpha_volume.view(1, 1, *alpha_volume.shape[-3:]))


class AlphaGridSampler(nn.Module):
    def __init__(self, aabb, enable_alpha_mask=False, threshold=1e-4, multiplier=1, near_far=(2, 6), nEnvSamples=0,
                 alphaMask_thres=0.001, update_list=(), max_samples=-1):
        super().__init__()
        if int(multiplier) != 1:
            raise _lib.NmfError("AlphaGridSampler: multiplier != 1 is not implemented")
        self.aabb = torch.as_tensor(aabb, dtype=torch.float32)
        self.enable_alpha_mask, self.alphaMask = enable_alpha_mask, None
        self.threshold, self.nEnvSamples, self.multiplier = threshold, nEnvSamples, 1
        self.near_far, self.update_list = list(near_far), list(update_list)
        self.grid_size = [128, 128, 128]
        self.alphaMask_thres, self.max_samples = alphaMask_thres, max_samples

    def check_schedule(self, iteration, batch_mul, rf):
        if iteration in self.update_list:
            self.update(rf)
        return False

    @torch.no_grad()
    def update(self, rf, init=False):
        self.aabb, self.units, self.contract_space = rf.aabb, rf.units, rf.contract_space
        self.nSamples, self.stepsize = rf.nSamples, rf.stepsize
        if not init:
            self.updateAlphaMask(rf, rf.grid_size)
            self.grid_size = rf.grid_size

    def _scene(self, rf):
        sc = rf.scene()
        vol = None if self.alphaMask is None else self.alphaMask.alpha_volume
        if getattr(sc, "_mask_src", None) is not vol:
            sc.set_alpha_volume(vol)
            sc._mask_src = vol
        sc.c.near, sc.c.far = float(self.near_far[0]), float(self.near_far[1])
        return sc

    @torch.no_grad()
    def updateAlphaMask(self, rf, grid_size=(200, 200, 200)):
        """samplers/alphagrid.py:249-276 (dense alpha by the CUDA kernel, 3^3 dilation, threshold)"""
        sc = self._scene(rf)
        sc.hp["alpha_mask_thres"] = self.alphaMask_thres
        vol = sc.update_alpha_mask([int(g) for g in grid_size])
        self.alphaMask = AlphaGridMask(rf.aabb, vol).to(rf.get_device())
        sc._mask_src = self.alphaMask.alpha_volume
        idx = torch.nonzero(vol > 0.5)
        gs = torch.tensor([int(g) for g in grid_size], device=vol.device)
        lo, hi = idx.amin(0).flip(0), idx.amax(0).flip(0)      # (z,y,x) -> (x,y,z)
        new_aabb = torch.stack([rf.aabb[0] + (rf.aabb[1] - rf.aabb[0]) * lo / (gs - 1),
                                rf.aabb[0] + (rf.aabb[1] - rf.aabb[0]) * hi / (gs - 1)])
        return new_aabb

    @torch.no_grad()
    def sample(self, rays_chunk, focal, rf, override_near=None, is_train=False, dynamic_batch_size=True,
               override_alpha_thres=None, stepmul=1, ndc_ray=False, **args):
        """samplers/alphagrid.py:278-370: xyzs (M,4), ray_valid (b,S), S, z_vals (b,S), dists (b,S), whole_valid (B).
        is_train=True: jittered cumulative steps keyed by (seed, ray id, step) -- keywords `seed`, `ray_id0` / `ray_ids` --
        and, with dynamic_batch_size, the truncation to the rays whose cumulative sample count stays below max_samples."""
        if ndc_ray:
            raise NotImplementedError("AlphaGridSampler.sample: the NDC path is not implemented")
        sc = self._scene(rf)
        rays = rays_chunk[:, :6].contiguous()
        whole_valid = torch.ones(rays.shape[0], dtype=torch.bool, device=rays.device)
        if is_train:
            from . import train
            ray_valid, z_vals, _, whole, kept = train.sample_rays_train(
                sc, rays, seed=args.get("seed", 0), ray_id0=args.get("ray_id0", 0), ray_ids=args.get("ray_ids"),
                max_samples=self.max_samples if dynamic_batch_size else -1, override_near=override_near)
            n = int(kept[0])                                   # the kept rays are a prefix of the batch
            if n < rays.shape[0]:
                whole_valid = whole
                rays, ray_valid, z_vals = rays[:n], ray_valid[:n], z_vals[:n]
        else:
            ray_valid, z_vals, _ = ops.sample_rays(sc, rays, override_near)
        pts = rays[:, None, :3] + rays[:, None, 3:6] * z_vals[..., None]
        xyzs = torch.cat([pts, z_vals[..., None] / focal], dim=-1)[ray_valid]
        dists = torch.cat((z_vals[:, 1:] - z_vals[:, :-1], torch.zeros_like(z_vals[:, :1])), dim=-1)
        return xyzs, ray_valid, z_vals.shape[1], z_vals, dists, whole_valid


# ------------------------------------------------------------------------------------------------------------
# shading model and its sub-plugins
# ------------------------------------------------------------------------------------------------------------
class ListISH(nn.Module):
    def __init__(self, degs):
        super().__init__()
        if list(degs) != [0, 1, 2, 4]:
            raise _lib.NmfError("ListISH: kernels implement degs=[0,1,2,4]")
        self.degs = list(degs)

    def dim(self):
        return 18


class GGXSampler(nn.Module):
    def __init__(self, max_samples=1024):
        super().__init__()
        self.max_samples = max_samples
        self.sampler = torch.quasirandom.SobolEngine(dimension=2, scramble=True)
        self.register_buffer("angs", self.sampler.draw(max_samples))

    def draw(self, B, num_samples):
        angs = self.angs.reshape(1, self.max_samples, 2)[:, :num_samples, :].expand(B, num_samples, 2)
        offset = torch.rand(B, 1, 2, device=angs.device) * 0.25
        return (angs + offset) % 1.0

    def sample(self, u1, u2, dir_out, normal, r1, r2, ray_mask, **kwargs):
        """brdf_samplers/ggx.py:61-226: returns (L (R,3), row_world_basis (R,3,3) with columns t,b,n, logpdf (R))"""
        n, m = ray_mask.shape
        ri, _ = torch.where(ray_mask)
        u = torch.stack([u1[ray_mask], u2[ray_mask]], dim=-1)
        V, N, r = dir_out[ri], normal[ri], r1.reshape(n, -1)[:, 0][ri]
        L, logpdf, _, _ = ops.ggx_sample(u, V, N, r)
        up = torch.where(N[:, 2:3].abs() < 0.999, torch.tensor([0.0, 0.0, 1.0], device=N.device),
                         torch.tensor([-1.0, 0.0, 0.0], device=N.device))
        unit = lambda v: v / (v ** 2).sum(-1, keepdim=True).clip(min=torch.finfo(torch.float32).eps).sqrt()
        t = unit(torch.linalg.cross(up, N))
        b = unit(torch.linalg.cross(N, t))
        return L, torch.stack([t, b, N], dim=-1), logpdf


class MLPBRDF(nn.Module):
    def __init__(self, in_channels, h_encoder=None, d_encoder=None, v_encoder=None, n_encoder=None, l_encoder=None,
                 feape=0, dotpe=-1, activation="sigmoid", mul_LdotN=False, bias=0, lr=1e-3, hidden_w=64, num_layers=3,
                 initializer="kaiming", **kwargs):
        super().__init__()
        if not (feape == 0 and dotpe < 0 and activation == "sigmoid" and not mul_LdotN and hidden_w == 64 and num_layers == 3
                and v_encoder is None and n_encoder is None and l_encoder is None):
            raise _lib.NmfError("MLPBRDF: kernels implement the microfacet_tensorf2 configuration (66 -> 64 -> 64 -> 4, sigmoid)")
        self.in_channels, self.bias, self.lr, self.init_val = in_channels, bias, lr, 0.5
        self.activation_name = activation
        self._scene, self._scene_key = None, None
        self.h_encoder, self.d_encoder = h_encoder, d_encoder
        self.in_mlpC = in_channels + 2 * (18 + 3)
        self.mlp = nn.Sequential(nn.Linear(self.in_mlpC, 64), nn.ReLU(inplace=True), nn.Linear(64, 64), nn.ReLU(inplace=True),
                                 nn.Linear(64, 4))
        for m in self.mlp:
            if isinstance(m, nn.Linear):
                nn.init.kaiming_uniform_(m.weight, nonlinearity="relu")
                nn.init.zeros_(m.bias)

    def _own_scene(self):
        """A device scene that only carries this MLP (nmf_brdf_mlp), rebuilt when a weight or the bias changed."""
        dev = self.mlp[0].weight.device
        key = (tuple(p._version for p in self.parameters()), float(self.bias), str(dev))
        if self._scene is None or key != self._scene_key:
            st = {f"model.brdf.{k}": v for k, v in self.state_dict().items()}
            self._scene = DeviceScene.shading_only(st, device=dev, heads=False, brdf=True, brdf_bias=float(self.bias))
            self._scene_key = key
        return self._scene

    @torch.no_grad()
    def forward(self, V, L, N, H, local_v, half_vec, diff_vec, efeatures, eax, eay=None):
        """modules/brdf.py:177-261 in the microfacet_tensorf2 configuration: the MLP sees [feature, ISH(half_vec),
        half_vec, ISH(diff_vec), diff_vec] (V, L, N, H, local_v only feed options that are off) -> (n,3) BRDF weight."""
        return ops.brdf_mlp(self._own_scene(), efeatures, half_vec.reshape(-1, 3), diff_vec.reshape(-1, 3), eax.reshape(-1))

    @torch.no_grad()
    def calibrate(self, efeatures, bg_brightness):
        """modules/brdf.py:141-176: shifts `bias` so that the mean BRDF weight over random directions is
        init_val / bg_brightness (inverse-sigmoid domain)."""
        n, dev = efeatures.shape[0], efeatures.device
        unit = lambda v: v / (v ** 2).sum(-1, keepdim=True).clip(min=torch.finfo(torch.float32).eps).sqrt()
        rand_vecs = lambda: unit(2 * torch.rand((n, 3), device=dev) - 1)
        L, norms = rand_vecs(), rand_vecs()
        norms = (L * norms).sum(dim=-1, keepdim=True) * norms
        weight = self(rand_vecs(), L, norms, rand_vecs(), rand_vecs(), rand_vecs(), rand_vecs(), efeatures,
                      torch.rand(n, device=dev), torch.rand(n, device=dev))
        target = self.init_val / float(bg_brightness)
        self.bias += math.log(target / (1 - target)) - float((weight / (1 - weight)).log().mean())


class RandHydraMLPDiffuse(nn.Module):
    def __init__(self, in_channels, pospe=-1, feape=0, roughness_view_encoder=None, roughness_cfg=None, hidden_w=64,
                 num_layers=1, initializer="xavier_sigmoid", lr=1e-3, start_roughness=0.35, tint_bias=0, diffuse_bias=-0.619,
                 diffuse_mul=1.5, roughness_bias=-1, **kwargs):
        super().__init__()
        if num_layers != 1 or pospe >= 0 or feape > 0 or roughness_view_encoder is not None:
            raise _lib.NmfError("RandHydraMLPDiffuse: kernels implement single-Linear heads on the raw feature (num_layers=1)")
        self.in_channels, self.lr, self.start_roughness = in_channels, lr, start_roughness
        self.tint_bias, self.diffuse_bias, self.diffuse_mul, self.roughness_bias = tint_bias, diffuse_bias, diffuse_mul, roughness_bias
        self.f0_bias = kwargs.get("f0_bias", 0)
        self._scene, self._scene_key = None, None
        for name, od in (("diffuse", 3), ("tint", 3), ("f0", 3), ("roughness", 2)):
            lin = nn.Linear(in_channels, od)
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)
            setattr(self, f"{name}_mlp", nn.Sequential(lin))

    def _own_scene(self):
        dev = self.diffuse_mlp[0].weight.device
        biases = dict(diffuse_bias=float(self.diffuse_bias), diffuse_mul=float(self.diffuse_mul), tint_bias=float(self.tint_bias),
                      roughness_bias=float(self.roughness_bias), f0_bias=float(self.f0_bias))
        key = (tuple(p._version for p in self.parameters()), tuple(biases.values()), str(dev))
        if self._scene is None or key != self._scene_key:
            st = {f"model.diffuse_module.{k}": v for k, v in self.state_dict().items()}
            self._scene = DeviceScene.shading_only(st, device=dev, heads=True, brdf=False, **biases)
            self._scene_key = key
        return self._scene

    @torch.no_grad()
    def forward(self, pts, viewdirs, features, std=0, **kwargs):
        """modules/render_modules.py:519-574 (nmf_material_heads): (diffuse, tint, dict(diffuse, r1, r2, f0, tint))."""
        if std != 0:
            raise NotImplementedError("RandHydraMLPDiffuse: std != 0 (material noise) is not built")
        a, t, f0, r1, r2 = ops.material_heads(self._own_scene(), features, with_r2=True)
        return a, t, dict(diffuse=a, r1=r1.reshape(-1, 1), r2=r2.reshape(-1, 1), f0=f0, tint=t)

    @torch.no_grad()
    def calibrate(self, mean_brightness, conserve_energy, *args, **kwargs):
        """modules/render_modules.py:632-642: moves diffuse_bias / roughness_bias so that the mean albedo is
        (0.5 or 0.25) / mean_brightness and the mean roughness is start_roughness (inverse-sigmoid domain)."""
        inv = lambda x: (x / (1 - x)).log() if torch.is_tensor(x) else math.log(x / (1 - x))
        diffuse, tint, extra = self(*args, **kwargs)
        v = (0.25 if not conserve_energy else 0.5) / float(mean_brightness)
        self.diffuse_bias += inv(v) - float(inv(diffuse).mean())
        rough = (extra["r1"] + extra["r2"]) / 2 / 2
        self.roughness_bias += inv(self.start_roughness) - float(inv(rough).mean())


class MLPRender_Fea(nn.Module):
    def __init__(self, in_channels, viewpe=2, feape=2, featureC=128, **kwargs):
        super().__init__()
        if viewpe != 2 or feape != 2 or featureC != 128:
            raise _lib.NmfError("MLPRender_Fea: kernels implement viewpe=2, feape=2, featureC=128 (model=tensorf)")
        in_mlpC = 2 * viewpe * 3 + 2 * feape * in_channels + 3 + in_channels
        self.mlp = nn.Sequential(nn.Linear(in_mlpC, featureC), nn.ReLU(inplace=True), nn.Linear(featureC, featureC),
                                 nn.ReLU(inplace=True), nn.Linear(featureC, 3))
        nn.init.constant_(self.mlp[-1].bias, 0)

    def calibrate(self, *args):
        return              # modules/render_modules.py:222-223


class Microfacet(nn.Module):
    """model slot (models/microfacet.py).  Owns the sub-plugins; its arithmetic runs inside the fused kernels."""

    def __init__(self, app_dim, brdf, brdf_sampler, diffuse_module, anoise=0.25, rays_per_ray=128, test_rays_per_ray=128,
                 max_brdf_rays=(650000, 450000), max_retrace_rays=(1000,), target_num_samples=(1000000,),
                 percent_bright=0.0, min_rough_start=0.0, min_rough_decay=0.999, conserve_energy=True,
                 russian_roulette=False, start_std=0.0, std_decay=1.0, cold_start_bg_iters=0, detach_N_iters=0,
                 no_emitters=True, diffuse_mixing_mode="fresnel", freeze=False, visibility_module=None,
                 std_decay_interval=10, **kwargs):
        super().__init__()
        if diffuse_mixing_mode != "fresnel" or not no_emitters or visibility_module is not None or russian_roulette or percent_bright:
            raise _lib.NmfError("Microfacet: kernels implement diffuse_mixing_mode=fresnel, no_emitters, no visibility module")
        self.brdf = brdf(in_channels=app_dim)
        self.brdf_sampler = brdf_sampler(max_samples=1024)
        self.diffuse_module = diffuse_module(in_channels=app_dim)
        self.anoise, self.rays_per_ray, self.test_rays_per_ray = anoise, rays_per_ray, test_rays_per_ray
        self.max_brdf_rays, self.max_retrace_rays = list(max_brdf_rays), list(max_retrace_rays)
        self.start_max_retrace_rays = list(max_retrace_rays)
        self.mean_ratios, self.ratio_list, self.conserve_energy = None, None, conserve_energy
        self.target_num_samples = list(target_num_samples)
        # training schedule state (models/microfacet.py:53-58, 70-71)
        self.min_rough, self.min_rough_decay = min_rough_start, min_rough_decay
        self.std, self.std_decay, self.std_decay_interval = start_std, std_decay, std_decay_interval
        self.detach_N_iters, self.detach_N = detach_N_iters, True
        self.outputs = {"diffuse": 3, "roughness": 1, "tint": 3, "spec": 3}
        self.needs_normals = lambda recur: True

    def hyper(self):
        d, b = self.diffuse_module, self.brdf
        return dict(model="microfacet", anoise=self.anoise, rays_per_ray=self.test_rays_per_ray,
                    max_brdf_rays=tuple(self.max_brdf_rays), max_retrace_rays=tuple(self.max_retrace_rays),
                    diffuse_bias=d.diffuse_bias, diffuse_mul=d.diffuse_mul, roughness_bias=d.roughness_bias,
                    tint_bias=d.tint_bias, f0_bias=getattr(d, "f0_bias", 0.0), brdf_bias=b.bias)

    def check_schedule(self, iter, batch_mul, **kw):
        """models/microfacet.py:112-121"""
        if iter % 10 == 0:
            self.min_rough *= self.min_rough_decay
        if iter > batch_mul * self.detach_N_iters:
            self.detach_N = False
        if iter % self.std_decay_interval == 0:
            self.std *= self.std_decay
        return False

    def get_optparam_groups(self, lr_scale=1):
        return [{"params": self.diffuse_module.parameters(), "lr": self.diffuse_module.lr * lr_scale},
                {"params": self.brdf.parameters(), "lr": self.brdf.lr * lr_scale}]

    def reset_counter(self):
        """models/microfacet.py:236-239"""
        self.max_retrace_rays = list(self.start_max_retrace_rays)
        self.mean_ratios = None
        self.ratio_list = None

    def update_n_samples(self, n_samples):
        """models/microfacet.py:241-268, the adaptive retrace controller fed by statistics["n_samples"][1:] (train.py:627):
        max_retrace_rays[i] follows target_num_samples[i] * min(recent rays-per-sample ratios), capped by max_brdf_rays."""
        if len(n_samples) != len(self.max_retrace_rays):
            return
        ratios = [(n_rays / n_sample) if n_sample > 0 else 1e-3 for n_rays, n_sample in zip(self.max_retrace_rays, n_samples)]
        if self.ratio_list is None:
            self.ratio_list = [[r, 1e-3] if r is not None else [] for r in ratios]
        else:
            self.ratio_list = [[r for r in ([ratio] + rlist) if r is not None][:20] for ratio, rlist in zip(ratios, self.ratio_list)]
        self.mean_ratios = [min(rlist) if len(rlist) > 0 else None for rlist in self.ratio_list]
        self.max_retrace_rays = [min(int(target * ratio + 1), maxv) if ratio is not None else prev
                                 for target, ratio, maxv, prev in zip(self.target_num_samples, self.mean_ratios,
                                                                      self.max_brdf_rays[:-1], self.max_retrace_rays)]

    def calibrate(self, args, xyz, feat, bg_brightness, save_config=True):
        """models/microfacet.py:79-96 (train.py:437): bias calibration of the material heads and the BRDF MLP against the
        environment's mean brightness; the new biases are written back into the run config."""
        unit = lambda v: v / (v ** 2).sum(-1, keepdim=True).clip(min=torch.finfo(torch.float32).eps).sqrt()
        self.diffuse_module.calibrate(bg_brightness, self.conserve_energy, xyz, unit(torch.rand_like(xyz[:, :3])), feat)
        self.brdf.calibrate(feat, bg_brightness)
        if save_config and args is not None:
            args.model.arch.model.brdf.bias = self.brdf.bias
            args.model.arch.model.diffuse_module.diffuse_bias = self.diffuse_module.diffuse_bias
            args.model.arch.model.diffuse_module.roughness_bias = self.diffuse_module.roughness_bias
        return args

    def forward(self, *a, **kw):
        raise NotImplementedError("Microfacet.forward runs inside the fused kernels: call TensorNeRF.forward / render_chunks")


class PlainTensoRF(nn.Module):
    """model=tensorf (models/tensorf.py): colour = view MLP of the appearance feature."""

    def __init__(self, app_dim, diffuse_module, **kwargs):
        super().__init__()
        self.diffuse_module = diffuse_module(in_channels=app_dim)
        self.outputs = {}
        self.needs_normals = lambda recur: False
        self.max_retrace_rays = []

    def hyper(self):
        return dict(model="plain")

    def check_schedule(self, iter, batch_mul, **kw):
        return False

    def get_optparam_groups(self, lr_scale=1):
        return [{"params": self.diffuse_module.parameters(), "lr": 1e-3 * lr_scale}]

    def update_n_samples(self, n_samples):
        pass

    def reset_counter(self):
        pass

    def calibrate(self, args, *fargs, **kwargs):
        return args          # models/tensorf.py:22-23


class IntegralEquirect(nn.Module):
    def __init__(self, bg_resolution, init_val, activation="identity", mipbias=0, mipnoise=0, lr=0.15, mipbias_lr=1e-3,
                 brightness_lr=0.01, mul_lr=0.01, mul_betas=(0.9, 0.999), betas=(0.9, 0.99)):
        super().__init__()
        if activation != "exp" or mipnoise != 0:
            raise _lib.NmfError("IntegralEquirect: kernels implement activation=exp, mipnoise=0 (microfacet_tensorf2)")
        self.bg_mat = nn.Parameter(init_val * torch.ones((1, 3, bg_resolution, 2 * bg_resolution)))
        self.register_parameter("mipbias", nn.Parameter(torch.tensor(mipbias, dtype=float)))
        self.register_parameter("brightness", nn.Parameter(torch.tensor(0.0, dtype=float)))
        self.register_parameter("mul", nn.Parameter(torch.tensor(1.0, dtype=float)))
        self.mipnoise, self.lr, self.mul_lr, self.mipbias_lr, self.brightness_lr = mipnoise, lr, mul_lr, mipbias_lr, brightness_lr
        self.betas, self.mul_betas = list(betas), list(mul_betas)
        self._scene, self._key, self._grad_acc = None, None, None

    @property
    def bg_resolution(self):
        return self.bg_mat.shape[2]

    def get_device(self):
        return self.bg_mat.device

    def activation_fn(self, x):
        return torch.exp((self.brightness + self.mul * x).clip(max=20))

    def hw(self):
        return self.bg_mat.shape[-2], self.bg_mat.shape[-1]

    @torch.no_grad()
    def calc_envmap_psnr(self, gt_im, fH=500):
        """modules/integral_equirect.py:290-321 (host-side metric, renderer.py:305-313): the ground-truth panorama is
        flipped and rolled by half its width, both maps are resized (nearest) to (fH, 2 fH), an affine colour map
        pred -> gt is fitted by least squares (sklearn LinearRegression in the reference) and the PSNR of the clipped
        squared error is returned."""
        import numpy as np
        fW = 2 * fH
        gt = np.asarray(gt_im)
        gW = gt.shape[1]
        gt = gt[:, ::-1]
        gt = np.concatenate([gt[:, gW // 2:], gt[:, :gW // 2]], axis=1)
        resize = lambda im: torch.nn.functional.interpolate(im.permute(2, 0, 1).unsqueeze(0), (fH, fW)).squeeze(0).permute(1, 2, 0)
        Y = resize(torch.as_tensor(gt.copy()).float()).reshape(-1, 3).double().numpy()
        X = resize(self.activation_fn(self.bg_mat[0]).permute(1, 2, 0).detach().float().cpu()).reshape(-1, 3).double().numpy()
        A = np.concatenate([X, np.ones((X.shape[0], 1))], axis=1)
        coef, *_ = np.linalg.lstsq(A, Y, rcond=None)
        err = ((A @ coef - Y) ** 2).clip(min=0, max=1)
        return float(-10.0 * np.log(err.mean()) / np.log(10.0))

    def mean_color(self):
        return self.activation_fn(self.bg_mat).reshape(-1, 3).mean(dim=0)     # integral_equirect.py:286-287 (sic)

    def get_optparam_groups(self, lr_scale=1):
        return [{"params": [self.bg_mat], "lr": self.lr * lr_scale, "betas": self.betas},
                {"params": [self.mipbias], "lr": self.mipbias_lr * lr_scale},
                {"params": [self.brightness], "lr": self.brightness_lr * lr_scale},
                {"params": [self.mul], "lr": self.mul_lr * lr_scale, "betas": self.mul_betas}]

    def scene(self):
        key = tuple(p._version for p in self.parameters()) + (str(self.get_device()),)
        if self._scene is None or key != self._key:
            self._scene = DeviceScene.env_only(self.bg_mat, self.mipbias, self.brightness, self.mul, self.get_device())
            self._key = key
        return self._scene

    @torch.no_grad()
    def forward(self, viewdirs, saSample, max_level=None):
        """integral_equirect.py:409-504"""
        return ops.env_lookup(self.scene(), viewdirs, saSample)

    @torch.no_grad()
    def accumulate_grad(self, viewdirs, saSample, grad_rgb):
        """Reverse pass of `forward` w.r.t. the map for one batch of lookups (what autograd does to the reference's module:
        integral_equirect.py:263-273, 409-504).  grad_rgb (n,3) = d loss / d forward(viewdirs, saSample).  Batches
        accumulate on the device until `finish_grad`."""
        sc = self.scene()
        if self._grad_acc is None or self._grad_acc.scene is not sc:
            self._grad_acc = ops.EnvMapGrad(sc)
        self._grad_acc.scatter(viewdirs, saSample, grad_rgb)

    @torch.no_grad()
    def finish_grad(self):
        """Once per optimiser step: adds the accumulated gradient to bg_mat.grad / brightness.grad / mul.grad (created when
        absent, like autograd's accumulation): bg_mat, brightness, mul and mipbias (the box size moves with the bias,
        integral_equirect.py:373-397: nmf_env_lookup_bwd_mipbias)."""
        if self._grad_acc is None:
            return
        d_bg, d_br, d_mul, d_mb = self._grad_acc.finish(self.bg_mat, self.brightness, self.mul)
        for p, g in ((self.bg_mat, d_bg), (self.brightness, d_br), (self.mul, d_mul), (self.mipbias, d_mb)):
            g = g.to(p.dtype).reshape(p.shape)
            p.grad = g if p.grad is None else p.grad + g

    @torch.no_grad()
    def get_spherical_harmonics(self, G, mipval=-5):
        """integral_equirect.py:324-360: (coeffs (9,3), conv_coeffs / pi (9,3)) -- the second value already carries the 1 / pi
        of the Lambertian BRDF, as the reference returns it (:359)."""
        sc = self.scene()
        conv_pi = sc.sh_irradiance(G, mipval)                    # sh_A * coeffs / pi, in the kernels' basis (nmf_sh9)
        # modules/sh.py:67-73 has all-positive degree-2 constants; nmf_sh9 carries a minus on the yz and xz terms (the same
        # sign in projection and evaluation, so E(n) is identical): hand the coefficients out in the reference's convention
        sign = torch.tensor([1, 1, 1, 1, 1, -1, 1, -1, 1], device=conv_pi.device, dtype=conv_pi.dtype).reshape(-1, 1)
        conv_pi = conv_pi * sign
        al2 = torch.tensor([math.pi] + [2 * math.pi / 3] * 3 + [math.pi / 4] * 5, device=conv_pi.device).reshape(-1, 1)
        return conv_pi * math.pi / al2, conv_pi


# ------------------------------------------------------------------------------------------------------------
# composition root
# ------------------------------------------------------------------------------------------------------------
class TensorNeRF(nn.Module):
    def __init__(self, rf, model, aabb, near_far, sampler, tonemap=None, bg_module=None, normal_module=None, alphaMask=None,
                 infinity_border=False, recur_stepmul=1, recur_alpha_thres=1e-3, detach_inter=False, hdr=False, bg_noise=0,
                 bg_noise_decay=0.999, use_predicted_normals=True, orient_world_normals=False, align_pred_norms=True,
                 eval_batch_size=512, geonorm_iters=-1, geonorm_interp_iters=1, lr_scale=1, contraction="AABB", **kwargs):
        super().__init__()
        if normal_module is not None or hdr or infinity_border:
            raise _lib.NmfError("TensorNeRF: normal_module / hdr / infinity_border are not part of the implemented path")
        aabb = torch.as_tensor(aabb, dtype=torch.float32)
        self.rf = rf(aabb=aabb)
        self.normal_module = None
        self.sampler = sampler(near_far=near_far, aabb=aabb)
        self.model = model(self.rf.app_dim)
        self.bg_module = bg_module
        self.tonemap = SRGBTonemap() if tonemap is None else tonemap
        self.lr_scale, self.hdr, self.eval_batch_size = lr_scale, hdr, eval_batch_size
        self.recur_stepmul, self.recur_alpha_thres = recur_stepmul, recur_alpha_thres
        self.near_far = list(near_far)
        self.bg_noise, self.bg_noise_decay = bg_noise, bg_noise_decay
        self.skip_eps, self.t_cut, self.seed, self.mlp = ops.DEFAULT_SKIP_EPS, ops.DEFAULT_T_CUT, 20211200, "f16"
        self._scene, self._scene_key, self._bufs, self._calls = None, None, None, 0
        self._train_bufs = None
        self._render_train_bufs = None

    def get_device(self):
        return self.rf.units.device

    def get_optparam_groups(self):
        g = self.rf.get_optparam_groups(self.lr_scale) + self.model.get_optparam_groups(self.lr_scale)
        if isinstance(self.bg_module, nn.Module):
            g += self.bg_module.get_optparam_groups(self.lr_scale)
        return g

    def check_schedule(self, iter, batch_mul):
        """modules/tensor_nerf.py:177-195: the model's, the sampler's and the field's schedules; a True from any of them
        means the caller re-creates the optimiser (train.py:806-813) and the sampler re-reads stepsize / nSamples."""
        r = bool(self.model.check_schedule(iter, batch_mul))
        r |= bool(self.sampler.check_schedule(iter, batch_mul, self.rf))
        r |= bool(self.rf.check_schedule(iter, batch_mul))
        if r:
            self.sampler.update(self.rf, init=True)
        self.bg_noise *= self.bg_noise_decay
        return r

    def save(self, path, config):
        """modules/tensor_nerf.py:120-134: same wire format as the reference"""
        torch.save({"config": config, "state_dict": self.state_dict()}, path)

    def invalidate(self):
        self._scene = None

    def scene(self):
        """The packed device scene; rebuilt when a parameter, the occupancy volume or the device changed."""
        vol = None if self.sampler.alphaMask is None else self.sampler.alphaMask.alpha_volume
        key = (tuple(p._version for p in self.parameters()), id(vol), str(self.get_device()), self.mlp,
               tuple(self.rf.grid_size.tolist()))
        hp = self.model.hyper()
        hp.update(distance_scale=self.rf.distance_scale, density_shift=self.rf.density_shift, step_ratio=self.rf.step_ratio,
                  mlp=self.mlp)
        norm = lambda v: tuple(v) if isinstance(v, (list, tuple)) else v
        hkey = tuple(sorted((k, norm(v)) for k, v in hp.items()))
        fixed = ("distance_scale", "density_shift", "step_ratio", "mlp", "model")     # baked into packed tensors / step count
        if self._scene is None or key != self._scene_key or any(self._scene.hp.get(k) != hp[k] for k in fixed):
            self._scene = DeviceScene(self.state_dict(), self.rf.aabb, self.sampler.near_far, self.rf.grid_size.tolist(),
                                      alpha_volume=vol, device=self.get_device(), **hp)
            self._scene_key, self._hyper_key, self._bufs, self._render_train_bufs = key, hkey, None, None
        elif hkey != getattr(self, "_hyper_key", None):
            # calibrated biases / controller-moved ray budgets: scalar fields only, scratch sizes may change
            self._scene.update_hyper(**{k: v for k, v in hp.items() if k not in fixed})
            self._hyper_key, self._bufs, self._render_train_bufs = hkey, None, None
        return self._scene

    @torch.no_grad()
    def render_chunks(self, rays, focal, chunk=4096, ray_id0=0, is_train=False, ndc_ray=False, N_samples=-1, **kw):
        """All chunks of `rays` in one asynchronous launch sequence (the B200-first replacement of the per-chunk host
        loop of renderer.py:72-104).  Returns (images, statistics) with the keys of TensorNeRF.forward.  The images are
        BORROWED views of this module's persistent output buffers: the next render call overwrites them (forward and
        renderer.chunk_renderer hand out copies, like the reference's fresh tensors)."""
        if ndc_ray:
            raise NotImplementedError("TensorNeRF: only the non-NDC render path is implemented")
        if is_train:
            raise NotImplementedError("TensorNeRF.render_chunks is the eval driver; the training forward is one batch per "
                                      "call: TensorNeRF.forward(is_train=True)")
        sc = self.scene()
        n = rays.shape[0]
        if self._bufs is not None and (self._bufs.n_rays < n or self._bufs.chunk != chunk):
            self._bufs = None
        if self._bufs is None:
            self._bufs = ops.RenderBuffers(sc, n, chunk, ops.image_keys(sc))
        ims, st = ops.render_rays(sc, rays.to(self.get_device()), focal, chunk=chunk, seed=self.seed, ray_id0=ray_id0,
                                  skip_eps=self.skip_eps, t_cut=self.t_cut, buffers=self._bufs)
        self._bufs = st["buffers"]            # grown if a scratch list overflowed
        stats = dict(recur=0, whole_valid=torch.ones(n, dtype=torch.bool, device=rays.device),
                     n_samples=st["n_samples"], n_retrace=st["n_retrace"])
        # A19 (modules/tensor_nerf.py:567-649): per-chunk regulariser inputs, one list entry per reference forward call
        env_reg = self._envmap_reg()
        for k in ("ori_loss", "diffuse_reg", "brdf_reg", "prediction_loss", "distortion_loss"):
            stats[k] = [c[k] for c in st["statistics"]]
        stats["envmap_reg"] = [env_reg] * len(st["statistics"])
        return ims, stats

    def train_step(self, rays, gt, focal=1.0, ray_ids=None, lambda_pred=0.0, lambda_ori=0.0):
        """Forward + backward of one training iteration (train.py:540-712) in ONE fused device call: the training forward of
        modules/tensor_nerf.py:210-674 (jittered sampling, dynamic batch truncation), the photometric loss of train.py:597-601
        + lambda_pred * prediction_loss + lambda_ori * ori_loss, and the hand-written backward (nmf_train_plain for
        model=tensorf, nmf_train_microfacet for model=microfacet_tensorf2: both shading levels, detach_N as scheduled).
        Accumulates into the `.grad` of every parameter (a SUM over rays; the caller divides by the batch size like
        train.py:709) and returns (loss, images, statistics).  No autograd graph is involved."""
        from . import train
        sc = self.scene()
        dev = self.get_device()
        if sc.hp["model"] == "microfacet":
            return self._train_step_microfacet(sc, rays.to(dev), gt.to(dev), focal, lambda_pred, lambda_ori)
        out = train.train_plain(sc, rays.to(dev), gt.to(dev), focal=focal,
                                seed=self.seed + self._calls, ray_ids=ray_ids, max_samples=self.sampler.max_samples,
                                lambda_pred=lambda_pred, buffers=self._train_bufs)
        self._train_bufs = out["buffers"]
        self._calls += 1
        self._add_grads(out["grads"].reference_layout())
        stats = dict(recur=0, whole_valid=out["whole_valid"], n_samples=[out["n_samples"]],
                     prediction_loss=2.0 * out["sum_acc"], ori_loss=0.0, diffuse_reg=0.0, brdf_reg=0.0, distortion_loss=0.0,
                     envmap_reg=self._envmap_reg())
        images = dict(rgb_map=out["rgb_map"][:out["n_rays"]], acc_map=out["acc_map"][:out["n_rays"]])
        return out["loss_photo"] + lambda_pred * stats["prediction_loss"], images, stats

    def _add_grads(self, grads):
        params = dict(self.named_parameters())
        for k, g in grads.items():
            p = params.get(k)
            if p is None:
                continue
            g = g.reshape(p.shape).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.add_(g)

    def _train_step_microfacet(self, sc, rays, gt, focal, lambda_pred, lambda_ori):
        from . import train
        if self.model.std != 0:
            raise NotImplementedError("Microfacet.std != 0 (material-head noise, render_modules.py:556-558) is not built")
        if self.model.rays_per_ray != self.model.test_rays_per_ray:
            raise NotImplementedError("rays_per_ray != test_rays_per_ray")
        g = getattr(self, "_mf_grads", None)
        if g is None or g.t["gpack0"].shape != sc.keep["dpack0"].shape or g.ew != int(sc.c.env_w) or g.t["gsat"].device != sc.device:
            g = train.MicrofacetGradBuffers(sc)
        g.scene = sc
        self._mf_grads = g
        n = rays.shape[0]
        out = train.train_microfacet(sc, rays, gt, focal=focal, seed=self.seed, ray_id0=self._calls * n,
                                     max_samples=self.sampler.max_samples, min_rough=self.model.min_rough,
                                     lambda_pred=lambda_pred, lambda_ori=lambda_ori, detach_N=self.model.detach_N, grads=g,
                                     zero_grads=True, buffers=self._render_train_bufs)
        self._render_train_bufs = out["buffers"]
        self._calls += 1
        bg = self.bg_module
        g.finish(bg.bg_mat.detach(), bg.brightness.detach(), bg.mul.detach())
        self._add_grads(g.reference_views())
        stats = dict(recur=0, whole_valid=out["whole_valid"], n_samples=list(out["n_samples"]), envmap_reg=self._envmap_reg())
        stats.update(out["statistics"])
        images = dict(rgb_map=out["rgb_map"], acc_map=out["acc_map"])
        loss = out["loss_photo"] + lambda_pred * 2.0 * out["sum_acc"] + lambda_ori * out["ori_loss"]
        return loss, images, stats

    @torch.no_grad()
    def forward_train(self, rays, focal, ray_id0=None):
        """TensorNeRF.forward(is_train=True, draw_debug=False) (modules/tensor_nerf.py:210-674) of the microfacet model
        for one ray batch, through nmf_render_rays_train: jittered steps (also in the re-traced rays), the dynamic batch
        truncation (statistics["whole_valid"], rows of the kept rays only), min_rough, the A19 regulariser inputs.
        Forward only (what an eval of the training statistics needs); `train_step` is the forward + backward call."""
        sc = self.scene()
        if sc.hp["model"] != "microfacet":
            raise NotImplementedError("forward_train: model=tensorf trains through TensorNeRF.train_step")
        if self.model.std != 0:
            raise NotImplementedError("Microfacet.std != 0 (material-head noise, render_modules.py:556-558) is not built")
        if self.model.rays_per_ray != self.model.test_rays_per_ray:
            raise NotImplementedError("rays_per_ray != test_rays_per_ray")
        n = rays.shape[0]
        ims, st = ops.render_rays_train(sc, rays.to(self.get_device()), focal, seed=self.seed,
                                        ray_id0=self._calls * n if ray_id0 is None else ray_id0,
                                        max_samples=self.sampler.max_samples, min_rough=self.model.min_rough,
                                        buffers=self._render_train_bufs)
        self._render_train_bufs = st["buffers"]
        self._calls += 1
        stats = dict(recur=0, whole_valid=st["whole_valid"].to(rays.device), n_samples=list(st["n_samples"]),
                     n_retrace=st["n_retrace"][0], envmap_reg=self._envmap_reg())
        stats.update(st["statistics"])
        return {k: v.clone() for k, v in ims.items()}, stats

    def _envmap_reg(self):
        """modules/tensor_nerf.py:606-610: (bg_module.mean_color().mean() - 0.05).clip(min=0)"""
        if self.bg_module is None or not hasattr(self.bg_module, "mean_color"):
            return 0.0
        return float((self.bg_module.mean_color().detach().mean() - 0.05).clip(min=0))

    @torch.no_grad()
    def forward(self, rays, focal, start_mipval=None, bg_col=None, stepmul=1, recur=0, override_near=None, output_alpha=None,
                dynamic_batch_size=True, gt_normals=None, is_train=False, ndc_ray=False, N_samples=-1, tonemap=True,
                draw_debug=True):
        """modules/tensor_nerf.py:210-674 for one chunk in eval mode (recur=0: the recursion runs on the device)."""
        if recur != 0 or start_mipval is not None or override_near is not None or not tonemap:
            raise NotImplementedError("TensorNeRF.forward: secondary-ray renders are issued by the kernels themselves")
        if is_train:
            if ndc_ray:
                raise NotImplementedError("TensorNeRF: only the non-NDC render path is implemented")
            return self.forward_train(rays, focal)
        ims, stats = self.render_chunks(rays, focal, chunk=rays.shape[0], ray_id0=self._calls * rays.shape[0],
                                        is_train=is_train, ndc_ray=ndc_ray)
        ims = {k: v.clone() for k, v in ims.items()}      # fresh tensors, like the reference (render_chunks lends its buffers)
        self._calls += 1
        for k in ("n_samples", "ori_loss", "diffuse_reg", "brdf_reg", "prediction_loss", "distortion_loss", "envmap_reg"):
            stats[k] = stats[k][0]
        return ims, stats

    @staticmethod
    def load(ckpt, config=None, near_far=None, **kwargs):
        """modules/tensor_nerf.py:136-175: rebuild from a reference-format checkpoint {config, state_dict}.  With an
        external `config` (both reference entry points pass args.model.arch, train.py:80,245) the calibrated biases -- plain
        attributes, not in the state_dict -- are copied over from the checkpoint's own config (tensor_nerf.py:138-146)."""
        from . import config as C
        state = ckpt["state_dict"]
        aabb = state["rf.aabb"]
        own = C.to_plain(ckpt["config"]) if ckpt.get("config") is not None else None
        cfg = own if config is None else C.to_plain(config)
        if config is not None and own is not None:
            def get(d, path):
                for k in path:
                    if not isinstance(d, dict) or k not in d:
                        return None
                    d = d[k]
                return d
            for path in (("model", "brdf", "bias"), ("model", "diffuse_module", "diffuse_bias"),
                         ("model", "diffuse_module", "roughness_bias")):
                v = get(own, path)
                dst = get(cfg, path[:-1])
                if v is not None and isinstance(dst, dict):
                    dst[path[-1]] = v
        cfg["rf"]["grid_size"] = state["rf.grid_size"].tolist()
        t = C.instantiate(cfg)(aabb=aabb, near_far=near_far if near_far is not None else [1, 6])
        if "sampler.alphaMask.alpha_volume" in state:
            vol = state["sampler.alphaMask.alpha_volume"]
            t.sampler.alphaMask = AlphaGridMask(aabb, vol.reshape(vol.shape[-3:]))
        t.load_state_dict(state, strict=False)
        t.sampler.update(t.rf, init=True)
        return t
