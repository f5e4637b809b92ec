"""Latent-parallel execution over the GPUs of one box.

After projection the q latent GPs are independent: rank r owns a contiguous block of
latents (Gram, Cholesky, inverse, sweep for those only -- no data-path collective).
The only exchange is ONE all-reduce(sum) per iteration over a flat buffer holding the
loss and every parameter gradient (shared H / M / B gradients are partial sums; the
per-latent kernel gradients are zero outside the owner), issued through
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def latent_block(q: int, rank: int, world: int):
    """Contiguous block [lo, hi) of latents owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(q, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_latents(model, rank: int = None, world: int = None, group=None):
    """Make ``model`` compute only its own block of latents."""
    if rank is None:
        rank = dist.get_rank(group)
    if world is None:
        world = dist.get_world_size(group)
    if world > model.n_latents:
        raise ValueError(f"cannot shard {model.n_latents} latents over {world} ranks")
    model._latent_range = latent_block(model.n_latents, rank, world)
    model._world_size = world
    model._dist_group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
    model._pred_cache = None
    return model


def allreduce_loss_and_grads(loss: torch.Tensor, params, group=None) -> torch.Tensor:
    """One flat all-reduce(sum) of [loss, grad_0, grad_1, ...]; grads are updated in
    place and the global loss value is returned (detached)."""
    params = [p for p in params if p.requires_grad]
    pieces = [loss.detach().reshape(1)]
    for p in params:
        pieces.append((p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1))
    flat = torch.cat(pieces)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 1
    for p in params:
        k = p.numel()
        g = flat[off:off + k].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += k
    return flat[0].clone()
