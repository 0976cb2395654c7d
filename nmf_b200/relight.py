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
