"""Synthetic Blender-format scenes and cameras (no dataset exists offline).

A scene is a plain ``dict[str, Tensor]`` that uses the reference's checkpoint key names
(``rf.density_rf.app_plane.0`` ... see /root/reference/modules/tensor_nerf.py:120-175 and
fields/tensoRF.py:49-51,273-294), so the same dict loads into the reference model (the oracle
harness) and into :class:`nmf_b200.tensor_nerf.TensorNeRF`.

The density field is *separable*, so no training is needed: component 0 of plane 0 / line 0 is a
constant -10 (empty space: softplus(-14) ~ 8e-7), components 1..K are ``amp * f_y (x) f_x`` on the
plane and ``f_z`` on the line for K blobs (Gaussians and soft boxes).  All other density factors
are zero.  Appearance factors are ``0.1 * randn`` and the small MLPs use the initialisers the
reference's config selects (xavier-uniform heads, kaiming-uniform BRDF MLP, default Linear init
for the basis matrix).

Rays follow the Blender loader's convention (/root/reference/dataLoader/ray_utils.py:23-43,67-89,
dataLoader/blender.py:46-120): pixel centres +0.5, ``dir = ((i-W/2)/fx, (j-H/2)/fy, 1)``
normalised, rotated by ``c2w @ diag(1,-1,-1)``.
"""
import math

import numpy as np
import torch

DATASET_NEAR_FAR = {  # /root/reference/configs/dataset/*.yaml
    "lego": (2.5, 7.0), "ship": (1.0, 6.0), "materials": (2.0, 6.0),
    "ficus": (1.0, 6.0), "helmet": (3.0, 5.0), "toaster": (2.5, 5.0),
}
DATASET_SEED = {"lego": 0, "ship": 1, "materials": 2, "ficus": 3, "helmet": 4, "toaster": 5}
CAMERA_ANGLE_X = 0.6911112070083618
CAMERA_RADIUS = 4.031128874


def step_size_and_count(aabb, grid_size, step_ratio=0.5):
    """stepsize / nSamples exactly as the reference computes them (fields/tensor_base.py:219-232)."""
    aabb = torch.as_tensor(aabb, dtype=torch.float32)
    gs = torch.as_tensor(list(grid_size), dtype=torch.long)
    aabb_size = aabb[1] - aabb[0]
    units = aabb_size / (gs - 1)
    stepsize = torch.min(units) * step_ratio
    diag = torch.sqrt(torch.sum(torch.square(aabb_size)))
    n_samples = int((diag / stepsize).item()) + 1
    return stepsize, n_samples


def _gauss(x, c, s):
    return torch.exp(-0.5 * ((x - c) / s) ** 2)


def _box(x, lo, hi, slope=60.0):
    return torch.sigmoid(slope * (x - lo)) * torch.sigmoid(slope * (hi - x))


def _blob_layout(seed):
    """K=5 blobs: 2 Gaussians (sigma 0.35, 0.18) and 3 soft boxes, centres depend on the scene seed."""
    g = torch.Generator().manual_seed(1000 + seed)
    c = (torch.rand(5, 3, generator=g) - 0.5) * 0.9
    blobs = [
        ("gauss", c[0] * 0.3, 0.35),
        ("gauss", c[1], 0.18),
        ("box", c[2], 0.22),
        ("box", c[3], 0.15),
        ("box", c[4], 0.12),
    ]
    return blobs


def _linear_default(out_f, in_f, g):
    bound = 1.0 / math.sqrt(in_f)  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound


def _xavier_uniform(out_f, in_f, g, gain=1.0):
    bound = gain * math.sqrt(6.0 / (in_f + out_f))
    return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound


def _kaiming_uniform(out_f, in_f, g):
    bound = math.sqrt(6.0 / in_f)  # gain sqrt(2), a=0
    return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound


def procedural_env_log_radiance(res, seed=0, base=-0.6):
    """(1,3,res,2res) log-radiance map: sky gradient + a few bright lobes (so SAT lookups are non-trivial)."""
    g = torch.Generator().manual_seed(2000 + seed)
    H, W = res, 2 * res
    v = torch.linspace(-1, 1, H).view(H, 1)
    u = torch.linspace(-1, 1, W).view(1, W)
    img = base + 0.6 * (-v) * torch.ones(1, W)  # brighter towards the top rows
    img = img.unsqueeze(0).repeat(3, 1, 1)
    img[2] += 0.25 * (-v)
    for _ in range(4):
        cu, cv = (torch.rand(2, generator=g) * 2 - 1).tolist()
        cv *= 0.6
        s = 0.04 + 0.12 * torch.rand(1, generator=g).item()
        amp = 1.5 + 2.5 * torch.rand(1, generator=g).item()
        col = 0.6 + 0.4 * torch.rand(3, generator=g)
        du = torch.remainder(u - cu + 1, 2) - 1
        lobe = torch.exp(-0.5 * ((du / s) ** 2 + ((v - cv) / s) ** 2))
        img += amp * col.view(3, 1, 1) * lobe.unsqueeze(0)
    return img.unsqueeze(0).contiguous()


def make_scene(name="lego", grid_size=300, bg_resolution=512, n_density=16, n_app=24, app_dim=24,
               amp=40.0, env="procedural", aabb_scale=1.0, full_density=False):
    """Returns (state, meta). ``state`` uses reference checkpoint keys; ``meta`` holds aabb / near_far.
    grid_size: an int (cubic) or (gx, gy, gz); full_density adds small random values to EVERY density factor (the
    blobs only use plane 0 / line 0), so that all three plane/line pairs take part in density and normals."""
    seed = DATASET_SEED.get(name, 0)
    g = torch.Generator().manual_seed(seed)
    gs = [int(grid_size)] * 3 if isinstance(grid_size, int) else [int(v) for v in grid_size]
    aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]]) * aabb_scale
    lins = [torch.linspace(-1, 1, n) for n in gs]
    mat, vec = [[0, 1], [0, 2], [1, 2]], [2, 1, 0]          # fields/tensoRF.py:40-41: plane p is (H = grid[mat1], W = grid[mat0])
    state = {}
    dplanes = [torch.zeros(1, n_density, gs[mat[p][1]], gs[mat[p][0]]) for p in range(3)]
    dlines = [torch.zeros(1, n_density, gs[vec[p]], 1) for p in range(3)]
    dplanes[0][0, 0] = -10.0
    dlines[0][0, 0] = 1.0
    for k, (kind, c, s) in enumerate(_blob_layout(seed)):
        if kind == "gauss":
            fx, fy, fz = _gauss(lins[0], c[0], s), _gauss(lins[1], c[1], s), _gauss(lins[2], c[2], s)
        else:
            fx, fy, fz = (_box(lins[0], c[0] - s, c[0] + s), _box(lins[1], c[1] - s, c[1] + s),
                          _box(lins[2], c[2] - s, c[2] + s))
        # plane 0 = (x -> W, y -> H), line 0 = z   (fields/tensoRF.py:40-41,161-179)
        dplanes[0][0, k + 1] = amp * fy.view(-1, 1) * fx.view(1, -1)
        dlines[0][0, k + 1, :, 0] = fz
    for i in range(3):
        state[f"rf.density_rf.app_plane.{i}"] = dplanes[i]
        state[f"rf.density_rf.app_line.{i}"] = dlines[i]
        state[f"rf.app_rf.app_plane.{i}"] = 0.1 * torch.randn(1, n_app, gs[mat[i][1]], gs[mat[i][0]], generator=g)
        state[f"rf.app_rf.app_line.{i}"] = 0.1 * torch.randn(1, n_app, gs[vec[i]], 1, generator=g)
    if full_density:
        for i in range(3):
            # components 6..15 are unused by the blobs: products of O(0.6) factors bend the density of the blobs'
            # surfaces (and their normals) through all three plane/line pairs without filling empty space
            dplanes[i][0, 6:] = 0.6 * torch.randn(dplanes[i][0, 6:].shape, generator=g)
            dlines[i][0, 6:] = 0.6 * torch.randn(dlines[i][0, 6:].shape, generator=g)
    state["rf.basis_mat.weight"] = _linear_default(app_dim, 3 * n_app, g)
    state["rf.dbasis_mat.weight"] = _linear_default(1, 3 * n_density, g)
    for head, od in (("diffuse", 3), ("tint", 3), ("f0", 3), ("roughness", 2)):
        state[f"model.diffuse_module.{head}_mlp.0.weight"] = _xavier_uniform(od, app_dim, g)
        state[f"model.diffuse_module.{head}_mlp.0.bias"] = torch.zeros(od)
    in_c = app_dim + 2 * (18 + 3)
    for li, (o, i) in zip((0, 2, 4), ((64, in_c), (64, 64), (4, 64))):
        state[f"model.brdf.mlp.{li}.weight"] = _kaiming_uniform(o, i, g)
        state[f"model.brdf.mlp.{li}.bias"] = torch.zeros(o)
    # scrambled Sobol table; the reference draws it at construction (brdf_samplers/base.py:6-9)
    sob = torch.quasirandom.SobolEngine(dimension=2, scramble=True, seed=seed)
    state["model.brdf_sampler.angs"] = sob.draw(1024)
    if env == "procedural":
        state["bg_module.bg_mat"] = procedural_env_log_radiance(bg_resolution, seed)
    else:
        state["bg_module.bg_mat"] = torch.full((1, 3, bg_resolution, 2 * bg_resolution), -0.6)
    state["bg_module.mipbias"] = torch.tensor(1.0, dtype=torch.float64)
    state["bg_module.brightness"] = torch.tensor(0.0, dtype=torch.float64)
    state["bg_module.mul"] = torch.tensor(1.0, dtype=torch.float64)
    meta = dict(name=name, aabb=aabb, near_far=DATASET_NEAR_FAR.get(name, (2.0, 6.0)), grid_size=list(gs),
                bg_resolution=bg_resolution)
    return state, meta


def plain_mlp_state(seed=0):
    """MLPRender_Fea(viewpe=2, feape=2, featureC=128) weights for the model=tensorf plumbing case
    (modules/render_modules.py:201-235: 135 -> 128 -> 128 -> 3, last bias zero)."""
    g = torch.Generator().manual_seed(77 + seed)
    st = {}
    for li, (o, i) in zip((0, 2, 4), ((128, 135), (128, 128), (3, 128))):
        st[f"model.diffuse_module.mlp.{li}.weight"] = _kaiming_uniform(o, i, g)
        st[f"model.diffuse_module.mlp.{li}.bias"] = torch.zeros(o)
    return st


def hemisphere_poses(n=200, seed=1, radius=CAMERA_RADIUS):
    """n camera-to-world matrices (Blender/OpenGL convention) looking at the origin from the upper hemisphere."""
    rs = np.random.RandomState(seed)
    poses = []
    for _ in range(n):
        az = rs.uniform(0, 2 * np.pi)
        el = rs.uniform(np.deg2rad(10), np.deg2rad(70))
        eye = radius * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
        fwd = -eye / np.linalg.norm(eye)            # camera looks along -z (OpenGL)
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
        right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        c2w = np.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up, -fwd, eye
        poses.append(c2w)
    return np.stack(poses).astype(np.float32)


def focal_for(W, camera_angle_x=CAMERA_ANGLE_X):
    return 0.5 * W / math.tan(0.5 * camera_angle_x)


def camera_rays(c2w, H=800, W=800, focal=None, crop=None):
    """(H*W, 6) float32 rays [origin, unit direction] for one pose.

    crop=(y0,y1,x0,x1) returns only that pixel window (row-major)."""
    focal = focal_for(W) if focal is None else focal
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    pose = c2w @ torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))  # blender2opencv
    y0, y1, x0, x1 = (0, H, 0, W) if crop is None else crop
    jj, ii = torch.meshgrid(torch.arange(y0, y1, dtype=torch.float32) + 0.5,
                            torch.arange(x0, x1, dtype=torch.float32) + 0.5, indexing="ij")
    dirs = torch.stack([(ii - W / 2) / focal, (jj - H / 2) / focal, torch.ones_like(ii)], -1)
    dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
    rays_d = (dirs @ pose[:3, :3].T).reshape(-1, 3)
    rays_o = pose[:3, 3].expand(rays_d.shape)
    return torch.cat([rays_o, rays_d], 1).contiguous()
