"""Ray sharding across the GPUs of one box.

Rays are independent, so rendering needs NO data-path collective: rank r of P renders the contiguous slice
`shard_range(N, r, P)` of the flattened ray batch with its own replica of the (2.4 MB) weights.  `gather_image` is the
optional all-gather for callers that need the whole frame on every rank (17 MB per 1600x900 rgb).  One process per GPU,
`torch.distributed` (NCCL on the GPU box, gloo in the CPU tests) for the plumbing.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_rays: int, rank: int, world: int):
    """Contiguous, balanced [start, stop) of rank `rank` (first n % world ranks get one extra ray)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_rays, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rays(ray_batch: torch.Tensor, rank: int | None = None, world: int | None = None) -> torch.Tensor:
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    a, b = shard_range(ray_batch.shape[0], rank, world)
    return ray_batch[a:b]


def gather_image(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather per-rank slices (first dim = rays of `shard_range`) back into the full [n_total, ...] tensor."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    longest = max(b - a for a, b in sizes)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:b - a] for p, (a, b) in zip(parts, sizes)], 0)


def render_sharded(ray_batch: torch.Tensor, render_fn, gather_keys=("rgb_map", "depth_map")):
    """Render this rank's slice with `render_fn(rays) -> dict` and all-gather the requested outputs."""
    n = ray_batch.shape[0]
    out = render_fn(shard_rays(ray_batch))
    return {k: gather_image(out[k], n) for k in gather_keys}


# --------------------------------------------------------------------------------------
# training: every rank draws its own rays, the replicas stay identical by
# summing the (tiny: 2 x 595,844 fp32 = 4.77 MB) parameter gradients with ONE all-reduce per step
# -- the replacement of the reference's DataParallel scatter/replicate/gather (utils/device_utils.py:36-39).
# --------------------------------------------------------------------------------------
class FlatGradients:
    """ONE flat fp32 buffer holding the gradients of every parameter of `nets` (1,191,688 elements =
    4.77 MB for coarse + fine), each `p.grad` a view into it.  The backward kernels accumulate straight into these views
    (no per-parameter AccumulateGrad add, no `torch.cat` / copy around the collective), `all_reduce()` sums the buffer in
    place with a single collective and `zero()` is one memset.  Replaces the gradient handling of the reference's
    DataParallel wrapper (utils/device_utils.py:36-39)."""

    def __init__(self, nets):
        self.nets = [n for n in nets if n is not None]
        slots = [(n, s) for n in self.nets for s in n._slots()]
        dev = slots[0][1][2].device
        self.flat = torch.zeros(sum(p.numel() for _, (_, _, p) in slots), dtype=torch.float32, device=dev)
        off = 0
        for net in self.nets:
            start = off
            for _, _, p in net._slots():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            net._bind_flat_grad(self.flat[start:off])

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, average: bool = True, group=None):
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        if average and self.flat.is_cuda and dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)     # NCCL averages inside the collective: no extra kernel
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.mul_(1.0 / dist.get_world_size(group))

    def release(self):
        for net in self.nets:
            net._bind_flat_grad(None)


def all_reduce_gradients(params, average: bool = True, group=None) -> None:
    """Sum (or average) `.grad` of `params` over the ranks with a single all-reduce on one flat buffer.  The buffer
    covers EVERY parameter on every rank (zeros where a rank has no gradient), so ranks whose batches left different
    heads unused still exchange identically laid-out buffers.  (With `FlatGradients` the gradients already live in one
    buffer: use its `all_reduce()`, which needs no packing.)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    ps = list(params)
    if not ps:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    targets, views = [], []
    for p in ps:
        v = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
        if p.grad is None:
            p.grad = v.clone()
        else:
            targets.append(p.grad)
            views.append(v)
    if targets:
        torch._foreach_copy_(targets, views)     # one multi-tensor kernel instead of one copy per parameter


def broadcast_parameters(params, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters (one flat broadcast).  The copy goes through `p.copy_` under
    no_grad so each parameter's version counter moves and the packed weight images are rebuilt (NeRF.packed)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    ps = list(params)
    flat = torch.cat([p.detach().reshape(-1) for p in ps])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    with torch.no_grad():
        for p in ps:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n
