"""Render drivers: the host-side mirror of the reference's renderer.py (chunk_renderer :56-106, BundleRender :109-170).

The reference drives one TensorNeRF.forward per 4096-ray chunk from Python and copies every output of every chunk to
the host.  Here all chunks of a ray list go to the GPU in ONE asynchronous call (nmf_render_rays / nmf_render_rays_host):
the chunk size is still honoured -- it bounds the scope of the per-chunk retrace selection exactly as in the reference
(models/microfacet.py:475-510) -- but it no longer costs a launch sequence, a host sync and a D2H copy per chunk.
"""
import ctypes as C

import torch

from . import _lib, ops


class HostRenderer:
    """Pinned host staging + device buffers for rendering ray lists that live in HOST memory (nmf_render_rays_host)."""

    def __init__(self, scene, max_rays, chunk, keys=None):
        self.scene, self.chunk, self.max_rays = scene, chunk, max_rays
        self.keys = ops.image_keys(scene) if keys is None else [k for k in keys if k in ops.IMAGE_SHAPES]
        self.bufs = ops.RenderBuffers(scene, max_rays, chunk, self.keys)
        self.rays_dev = torch.empty(max_rays, 6, device=scene.device)
        self.host_images, self.c_host = {}, _lib.NmfImages()
        for k in self.keys:
            t = torch.empty_like(self.bufs.images[k], device="cpu").pin_memory()
            self.host_images[k] = t
            setattr(self.c_host, k, t.data_ptr())
        self.host_counters, self.c_host_counters = {}, _lib.NmfCounters()
        for k, v in self.bufs.counters.items():
            t = torch.zeros_like(v, device="cpu").pin_memory()
            self.host_counters[k] = t
            setattr(self.c_host_counters, k, t.data_ptr())
        self.h2d_bytes = self.d2h_bytes = 0

    def render(self, rays_host, focal, seed=0, ray_id0=0, skip_eps=ops.DEFAULT_SKIP_EPS, t_cut=ops.DEFAULT_T_CUT):
        """rays_host: (n,6) float32 CPU tensor (pinned for an asynchronous copy).  Returns (images on the host, stats);
        synchronises the stream (the caller reads the result)."""
        if rays_host.is_cuda or rays_host.dtype != torch.float32 or not rays_host.is_contiguous():
            raise _lib.NmfError("HostRenderer.render takes a contiguous float32 CPU tensor")
        n = rays_host.shape[0]
        if n > self.max_rays or rays_host.shape[1] != 6:
            raise _lib.NmfError("ray list larger than the renderer was sized for")
        rp = _lib.NmfRender(n_rays=n, chunk=self.chunk, focal=float(focal), seed=int(seed), ray_id0=int(ray_id0),
                            skip_eps=float(skip_eps), t_cut=float(t_cut), white_bg=1, cap_scale=self.bufs.cap_scale)
        st = _lib.lib().nmf_render_rays_host(self.scene.ref(), C.byref(rp), C.c_void_p(rays_host.data_ptr()),
                                             C.c_void_p(self.rays_dev.data_ptr()), C.byref(self.c_host),
                                             C.byref(self.bufs.c_images), C.byref(self.c_host_counters),
                                             C.byref(self.bufs.c_counters), C.c_void_p(self.bufs.ws_ptr),
                                             self.bufs.ws_bytes, ops._stream())
        _lib.check(st, "nmf_render_rays_host")
        torch.cuda.current_stream().synchronize()
        err = int(self.host_counters["error"][0])
        if err:
            # a scratch list was too small for this scene: grow it like ops.render_rays does and render again
            if self.bufs.cap_scale >= 16:
                raise _lib.NmfOverflow("nmf_render_rays_host: " + "; ".join(m for b, m in _lib.DEV_ERRORS.items() if err & b))
            old = self.bufs
            self.bufs = ops.RenderBuffers(self.scene, self.max_rays, self.chunk, self.keys, cap_scale=old.cap_scale * 2)
            return self.render(rays_host, focal, seed=seed, ray_id0=ray_id0, skip_eps=skip_eps, t_cut=t_cut)
        nc = (n + self.chunk - 1) // self.chunk
        self.h2d_bytes = n * 24
        self.d2h_bytes = sum(v[:n].numel() * v.element_size() for v in self.host_images.values()) + 4 * (6 * nc + 3)
        stats = {k: self.host_counters[k][:nc].tolist() for k in ("n_samples0", "n_samples1", "n_retrace")}
        stats["n_samples"] = [[a, b] for a, b in zip(stats["n_samples0"], stats["n_samples1"])]
        stats["whole_valid"] = torch.ones(n, dtype=torch.bool)
        stats["statistics"] = ops.chunk_statistics(self.host_counters["stat4"][:4 * nc].reshape(nc, 4), stats["n_samples0"])
        return {k: v[:n] for k, v in self.host_images.items()}, stats


class _PinnedRing:
    """Two rotating sets of pinned host buffers for the outputs of chunk_renderer(render2completion=True): the device ->
    host copies are asynchronous (one stream sync per image instead of one blocking pageable copy per map), and a result
    stays valid until the call after next.  (The reference returns fresh pageable tensors; pass fresh=True for that.)"""

    def __init__(self):
        self.sets, self.turn = [{}, {}], 0

    def take(self, key, like, n):
        cur = self.sets[self.turn]
        t = cur.get(key)
        if t is None or t.shape[0] < n or t.shape[1:] != like.shape[1:] or t.dtype != like.dtype:
            t = torch.empty((n,) + tuple(like.shape[1:]), dtype=like.dtype).pin_memory()
            cur[key] = t
        return t[:n]

    def advance(self):
        self.turn ^= 1


def chunk_renderer(rays, tensorf, focal, keys=("rgb_map",), chunk=4096, render2completion=False, fresh=False, **kwargs):
    """renderer.chunk_renderer (renderer.py:56-106): renders `rays` in chunks of `chunk` with `tensorf` and collects
    `keys` (None = everything) from the image and statistics dicts.  Device rays -> device outputs (fresh tensors, as in the
    reference); with render2completion=True outputs are returned on the host like the reference's eval path
    (renderer.py:71,88-97) -- in pinned buffers that stay valid until the call after next, or fresh pageable tensors with
    fresh=True.  In eval mode whole_valid is all-True (no dynamic truncation), so one pass completes every chunk."""
    ims, stats = tensorf.render_chunks(rays, focal, chunk=chunk, **kwargs)
    out_i, out_s = {}, {}
    sel = {k: v for k, v in ims.items() if keys is None or k in keys}
    if render2completion and not fresh:
        ring = getattr(tensorf, "_host_ring", None)
        if ring is None:
            ring = tensorf._host_ring = _PinnedRing()
        for k, v in sel.items():
            h = ring.take(k, v, v.shape[0])
            h.copy_(v, non_blocking=True)
            out_i[k] = h
        ring.advance()
        torch.cuda.current_stream().synchronize()
    else:
        for k, v in sel.items():
            out_i[k] = v.cpu() if render2completion else v.clone()
    for k, v in stats.items():
        if keys is None or k in keys:
            out_s[k] = v
    return out_i, out_s


class BundleRender:
    """renderer.BundleRender (renderer.py:109-170) for bundle_size=1: shuffles the H*W rays of one view so that every
    chunk is a random subset of the image (load balance + the statistical meaning of the per-chunk retrace budget),
    renders, and un-shuffles into (H, W, C) images."""

    def __init__(self, base_renderer, H, W, focal, bundle_size=1, scale_normal=False):
        assert bundle_size == 1, "only bundle_size=1 is implemented (the value every shipped config uses)"
        self.base_renderer, self.H, self.W, self.focal = base_renderer, H, W, focal

    @torch.no_grad()
    def __call__(self, rays, tensorf, **kwargs):
        n = rays.shape[0]
        perm = torch.randperm(n, device=rays.device)
        ims, stats = self.base_renderer(rays[perm], tensorf, keys=None, focal=self.focal, render2completion=True, **kwargs)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(n, device=perm.device)
        inv = inv.cpu()
        out = {}
        for k, v in ims.items():
            v = v[inv]
            out[k] = v.reshape(self.H, self.W, *v.shape[1:])
        return out, stats


@torch.no_grad()
def evaluate_views(tensorf, poses, H, W, focal, gt_images=None, chunk=4096, shuffle=True, seed=0, keys=("rgb_map",)):
    """Eval driver, the device-resident restatement of renderer.evaluate's render + PSNR loop (renderer.py:194-401) and
    of evaluation / evaluation_path (:537-582), without its file output:

      * rays are generated on the device from (pose, intrinsics) (nmf_generate_rays) -- no (N,6) ray tensors in host
        memory, no H2D copy per view;
      * pixels are rendered in a shuffled order like BundleRender (renderer.py:130-132) so that every chunk is a
        random subset of the image, and un-shuffled on the device;
      * PSNR (renderer.py:399-401: 8-bit quantised render vs clipped ground truth) is reduced on the device and read
        back as ONE scalar per view; images stay on the device.

    poses: iterable of 3x4 / 4x4 OpenCV-convention camera-to-world matrices; gt_images: optional (V,H,W,3) tensor.
    Returns dict(images=[{key: (H,W,C) device tensor}], psnr=[float] or None, n_samples=[...])."""
    dev = tensorf.get_device()
    n = H * W
    gen = torch.Generator(device="cpu").manual_seed(seed)
    out_images, sq, n_samples = [], [], []
    rays = torch.empty(n, 6, device=dev)
    for v, pose in enumerate(poses):
        perm = (torch.randperm(n, generator=gen) if shuffle else torch.arange(n)).to(device=dev, dtype=torch.int32)
        ops.generate_rays(pose, H, W, focal, pixel_ids=perm, device=dev, out=rays)
        ims, stats = tensorf.render_chunks(rays, focal, chunk=chunk, ray_id0=v * n)
        n_samples.append(stats["n_samples"])
        if gt_images is not None:
            sq.append(ops.image_sq_error(ims["rgb_map"], gt_images[v].to(dev).reshape(-1, 3), pixel_ids=perm))
        full = {}
        for k in keys:
            t = torch.empty_like(ims[k])
            t[perm.long()] = ims[k]
            full[k] = t.reshape(H, W, *t.shape[1:])
        out_images.append(full)
    psnr = None
    if sq:
        mse = torch.cat(sq).cpu() / (3.0 * n)                      # the only device -> host transfer of the loop
        psnr = (-10.0 * torch.log10(mse)).tolist()
    return dict(images=out_images, psnr=psnr, n_samples=n_samples)
