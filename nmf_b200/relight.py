"""Checkpoint wire format + relight entry: the host-side restatement of train.py:54-188
(`render_test`, `render_only=True`) for the render path.

  * `load_for_render` rebuilds a TensorNeRF mirror from a reference-format checkpoint {config, state_dict}
    (modules/tensor_nerf.py:120-175) and restores the occupancy volume from the checkpoint's own
    `sampler.alphaMask.alpha_volume` (train.py:86-88 re-derives it from that shape).
  * `swap_env` replaces the environment with a fixed `IntegralEquirect` state_dict (train.py:96-131, `fixed_bg`).
    The reference hard-codes bg_resolution=512 there although `backgrounds/forest.th` is 1024x2048, which raises a size
    mismatch as written (SURVEY section 7, quirk list); here the module is sized from the state_dict itself.
  * `relight_sweep` renders (checkpoint x environment) jobs; jobs are independent, so on N GPUs they are dealt
    round-robin to the ranks -- scene-parallel replicas, no collective (BASELINE config #5).
"""
import torch

from . import plugins, renderer


def load_for_render(ckpt, config=None, near_far=None, device="cuda"):
    if isinstance(ckpt, str):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    t = plugins.TensorNeRF.load(ckpt, config, near_far=near_far)
    return t.to(device).eval()


def swap_env(tensorf, bg_state, mipbias=0.0):
    """fixed_bg of train.py:96-131: a new IntegralEquirect holding `bg_state` (a state_dict or a path to one)."""
    if isinstance(bg_state, str):
        bg_state = torch.load(bg_state, map_location="cpu", weights_only=False)
    res = int(bg_state["bg_mat"].shape[-2])
    bg = plugins.IntegralEquirect(bg_resolution=res, mipbias=mipbias, activation="exp", lr=0.0, init_val=-1.897, mul_lr=0.0,
                                  brightness_lr=0.0)
    bg.load_state_dict(bg_state, strict=False)
    tensorf.bg_module = bg.to(tensorf.get_device())
    tensorf.invalidate()
    return tensorf


def load_panorama(path):
    """An equirectangular HDR panorama as a float32 (H, W, 3) RGB tensor.  .exr / .hdr go through OpenCV (the reference
    reads them with imageio, scripts/pano2cube.py:52; neither imageio nor an EXR plugin is a dependency here)."""
    import os
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if im is None:
        raise IOError(f"cannot read {path}")
    im = im[..., :3][..., ::-1].astype("float32")                  # BGR -> RGB
    return torch.from_numpy(im.copy())


def env_from_panorama(pano, resolution=1024, mipbias=0.0, floor=1e-6):
    """An IntegralEquirect state_dict from an equirectangular panorama by DIRECT log-radiance resampling -- the closed
    form of what scripts/pano2cube.py:31-134 fits with 1000 Adam steps (its loss asks IntegralEquirect(direction of
    panorama pixel, tiny solid angle) = pixel colour).  pano2cube's pixel -> direction map (:104-111) is
    theta = row / (H-1) * pi - pi/2, phi = -col / (W-1) * 2 pi - pi, d = (cos phi cos theta, sin phi cos theta, -sin theta);
    the module's own direction -> texel map (modules/integral_equirect.py:409-430) is x = (atan2(d_y, d_x) mod 2 pi - pi) / pi,
    y = -2 atan2(d_z, |d_xy|) / pi.  Composing them: texel row = panorama row (same latitude), texel column ix reads panorama
    column ((1/2 - ix / (w-1)) mod 1) (W-1): a horizontal flip and a roll by half a turn.  Larger panoramas are
    area-averaged down to (resolution, 2 resolution) first; radiance below `floor` is clamped before the log
    (activation = exp, brightness 0, mul 1)."""
    import torch.nn.functional as F
    p = torch.as_tensor(pano, dtype=torch.float32)
    if p.dim() != 3 or p.shape[-1] < 3:
        raise ValueError("env_from_panorama: (H, W, 3) panorama expected")
    img = p[..., :3].permute(2, 0, 1)[None]                          # (1,3,H,W)
    h, w = int(resolution), 2 * int(resolution)
    H, W = img.shape[-2:]
    if H > h and H % h == 0 and W % w == 0:
        img = F.avg_pool2d(img, kernel_size=(H // h, W // w))
    elif (H, W) != (h, w):
        img = F.interpolate(img, size=(h, w), mode="area" if H >= h and W >= w else "bilinear",
                            **({} if H >= h and W >= w else {"align_corners": True}))
    # The module's finest lookup is a half-texel box on the bilinearly interpolated SAT: it returns the mean of texels
    # (r, r+1) x (c, c+1) (modules/integral_equirect.py:18-39 with level 0 of :463-464), i.e. the map shifted by half a
    # texel.  pano2cube's fit absorbs that shift into the map; the closed form samples the panorama half a texel earlier.
    ix = torch.arange(w, dtype=torch.float64) - 0.5
    col = torch.remainder(0.5 - ix / (w - 1), 1.0) * (w - 1)           # panorama column of env texel ix (flip + half-turn roll)
    c0 = col.floor().long() % w
    c1 = (c0 + 1) % w
    f = (col - col.floor()).float()
    env = img[..., c0] * (1 - f) + img[..., c1] * f
    iy = (torch.arange(h, dtype=torch.float64) - 0.5).clamp(0, h - 1)
    r0 = iy.floor().long()
    r1 = (r0 + 1).clamp(max=h - 1)
    fr = (iy - r0).float().reshape(1, 1, h, 1)
    env = env[:, :, r0, :] * (1 - fr) + env[:, :, r1, :] * fr
    return {"bg_mat": torch.log(env.clamp(min=floor)).contiguous(), "mipbias": torch.tensor(float(mipbias), dtype=torch.float64),
            "brightness": torch.tensor(0.0, dtype=torch.float64), "mul": torch.tensor(1.0, dtype=torch.float64)}


def job_slice(n_jobs, rank, world):
    """jobs of `rank`: round-robin over the ranks (independent jobs, no communication)."""
    return list(range(rank, n_jobs, world))


@torch.no_grad()
def relight_sweep(ckpts, envs, poses, H, W, focal, near_far=None, rank=0, world=1, chunk=4096, device="cuda", gt=None):
    """ckpts: {name: checkpoint dict or path}; envs: {name: IntegralEquirect state_dict or path}.  Renders every
    (scene, env) pair of this rank over `poses`; returns {(scene, env): evaluate_views result}."""
    jobs = [(s, e) for s in ckpts for e in envs]
    out = {}
    cache = {}
    for j in job_slice(len(jobs), rank, world):
        s, e = jobs[j]
        if s not in cache:
            cache = {s: load_for_render(ckpts[s], near_far=near_far, device=device)}
        t = swap_env(cache[s], envs[e])
        out[(s, e)] = renderer.evaluate_views(t, poses, H, W, focal, gt_images=None if gt is None else gt.get((s, e)), chunk=chunk)
    return out
