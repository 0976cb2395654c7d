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


def gather_ragged(t, counts, rank, world, buffers=None, group=None):
    """Gathers per-rank row blocks of different lengths on rank 0 with ONE collective: every rank pads its (counts[rank], C)
    block to max(counts) rows (dist.gather needs equal shapes).  Returns (the concatenated (sum(counts), C) tensor on rank 0,
    None elsewhere; the staging buffers to pass back in on the next call).  Works with NCCL (device tensors) and gloo."""
    import torch.distributed as dist
    m = max(counts)
    if buffers is None:
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        buffers = (pad, parts)
    pad, parts = buffers
    pad[:t.shape[0]].copy_(t)
    dist.gather(pad, parts, dst=0, group=group)
    if rank != 0:
        return None, buffers
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0), buffers


class FlatGradBucket:
    """The one collective of ray-sharded training (BASELINE config #4): every rank renders its own
    rays, back-propagates locally, then ONE all-reduce(SUM) over a single flat fp32 buffer that aliases every parameter's
    `.grad` (~50 MB at G=300; NVSwitch reduces it in the fabric, so one launch-latency-sized bucket beats many).  The
    loss normaliser becomes the global ray count via `scale`.  (The backward kernels that fill the gradients are the
    next row of SURVEY 8f; this class is the communication half, tested with gloo on CPU and used as-is with NCCL.)"""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 and p.dtype != torch.float64:
                raise ValueError("FlatGradBucket handles floating-point parameters")
            if p.dtype == torch.float32:
                p.grad = self.flat[off:off + p.numel()].view_as(p)      # autograd accumulates in place into the view
            off += p.numel()
        self._offsets = off

    def zero(self):
        self.flat.zero_()

    def allreduce(self, scale=1.0, group=None):
        """Sums the gradients of all ranks in one collective and multiplies by `scale` (e.g. 1 / global ray count)."""
        import torch.distributed as dist
        # float64 scalars of the reference (bg_module.mipbias / brightness / mul) are not views: stage them in and out
        off = 0
        for p in self.params:
            if p.dtype == torch.float64 and p.grad is not None:
                self.flat[off:off + p.numel()] = p.grad.reshape(-1).float()
            off += p.numel()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if scale != 1.0:
            self.flat.mul_(scale)
        off = 0
        for p in self.params:
            if p.dtype == torch.float64 and p.grad is not None:
                p.grad.copy_(self.flat[off:off + p.numel()].view_as(p).double())
            off += p.numel()
        return self.flat
