"""Multi-GPU: ray-batch sharding.  One process per GPU; rays are independent, weights are
replicated (~50 MB), so the render path needs NO data-path collective: each rank renders a contiguous range of
whole chunks and rank 0 gathers the image slices.

Sharding is chunk-aligned and every ray keeps its global id (`ray_id0`), so chunk membership (the scope of the
per-chunk retrace selection, models/microfacet.py:475-510) and every keyed random number are the same as in a
single-GPU render: an N-GPU image equals the 1-GPU image, launch for launch.
"""
import torch


def shard_chunks(n_rays, chunk, rank, world):
    """[start, stop) ray range of `rank`: whole chunks, as even as possible (the first ranks get the extra chunk)."""
    n_chunks = (n_rays + chunk - 1) // chunk
    base, extra = divmod(n_chunks, world)
    c0 = rank * base + min(rank, extra)
    c1 = c0 + base + (1 if rank < extra else 0)
    return min(c0 * chunk, n_rays), min(c1 * chunk, n_rays)


def render_sharded(render_fn, rays, chunk, rank=None, world=None, gather=True, group=None):
    """render_fn(rays_slice, ray_id0) -> dict of per-ray tensors.  Returns the full dict on rank 0 (gather=True) or
    this rank's slice.  The only communication is the final gather of outputs (no collective inside the render)."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = rays.shape[0]
    s0, s1 = shard_chunks(n, chunk, rank, world)
    out = render_fn(rays[s0:s1], s0) if s1 > s0 else {}
    if not gather or world == 1:
        return out
    parts = [None] * world
    dist.gather_object({k: v.cpu() for k, v in out.items()}, parts if rank == 0 else None, dst=0, group=group)
    if rank != 0:
        return None
    keys = next(p for p in parts if p).keys()
    return {k: torch.cat([p[k] for p in parts if p], dim=0) for k in keys}
