"""Batch-sharded data parallelism for the label-graph path: one process per GPU, weights and the
label graph replicated, the batch split contiguously across ranks, and ONE flat fp32 gradient all-reduce per
training step (NCCL over NVLink on the GPU box, gloo in the CPU tests).  The forward needs no collective.

Replaces the reference's single-process ``nn.DataParallel`` (main.py:106-108), which re-broadcasts the parameters and
gathers outputs through GPU 0 every step.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment -> (rank, world, local_rank)."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``n`` samples: the first ``n % world`` ranks get one extra sample."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[Optional[torch.Tensor]], rank: int, world: int) -> List[Optional[torch.Tensor]]:
    """Slice every tensor of a batch along dim 0 to this rank's shard (``None`` entries pass through)."""
    n = next(t.shape[0] for t in tensors if t is not None)
    a, b = shard_range(n, rank, world)
    return [None if t is None else t[a:b] for t in tensors]


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world: Optional[int] = None, group=None,
                        local_weight: float = 1.0) -> int:
    """Average gradients over ranks with a single flat all-reduce.

    Parameters whose ``grad`` is ``None`` on this rank (e.g. the reference's dead encoder self-attention,
    lamp/Layers.py:16-18) contribute zeros so every rank reduces the same buffer layout and nothing deadlocks;
    their ``grad`` stays ``None`` unless another rank produced one.  ``local_weight`` lets ranks with uneven shard
    sizes weight their mean-reduced loss (pass ``local_n * world / global_n``).  Returns the number of elements.
    """
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group) if (world is None and dist.is_initialized()) else (world or 1)
    dev = params[0].device
    sizes = [p.numel() for p in params]
    total = sum(sizes)
    have = [p.grad is not None for p in params]
    # one flat buffer: gradients (zeros where this rank has none) + one "some rank has a gradient" flag per parameter
    pieces = [p.grad.reshape(-1).float() if h else torch.zeros(n, dtype=torch.float32, device=dev)
              for p, h, n in zip(params, have, sizes)]
    pieces.append(torch.tensor([1.0 if h else 0.0 for h in have], dtype=torch.float32, device=dev))
    flat = torch.cat(pieces)
    if local_weight != 1.0:
        flat[:total].mul_(local_weight)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat[:total].div_(world)
        any_rank = flat[total:].tolist()  # the only host synchronisation of the step
    else:
        any_rank = [1.0 if h else 0.0 for h in have]
    views = torch.split(flat[:total], sizes)
    dst, src = [], []
    for p, h, flag, v in zip(params, have, any_rank, views):
        if flag <= 0:
            continue
        if h:
            dst.append(p.grad)
            src.append(v.view_as(p).to(p.grad.dtype))
        else:
            p.grad = v.view_as(p).to(p.dtype).clone()
    if dst:
        torch._foreach_copy_(dst, src)
    return total


def data_parallel_step(model, optimizer, loss_fn, batch, world: Optional[int] = None, group=None) -> float:
    """One training step on this rank's shard: forward, backward, flat gradient all-reduce, optimizer step."""
    optimizer.zero_grad(set_to_none=True)
    loss = loss_fn(model, batch)
    loss.backward()
    params = model.get_trainable_parameters() if hasattr(model, 'get_trainable_parameters') else model.parameters()
    allreduce_gradients(list(params), world, group)
    optimizer.step()
    return float(loss.detach())
