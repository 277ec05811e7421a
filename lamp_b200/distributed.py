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


class GradientReducer:
    """Flat-buffer, bucketed, backward-overlapped gradient all-reduce (the replacement of ``nn.DataParallel``'s
    per-step ``reduce_add_coalesced`` to GPU 0, main.py:106-108).

    * ONE pre-allocated fp32 buffer holds every trainable gradient; each ``p.grad`` is a view into it, so autograd
      accumulates straight into the communication buffer -- no ``torch.cat`` before and no copy-back after the
      collective.  Use ``reducer.zero_grad()`` (zeroes the buffer, keeps the views) instead of
      ``optimizer.zero_grad(set_to_none=True)``.
    * the buffer is cut into buckets in REVERSE registration order (about the order in which backward produces
      gradients).  A post-accumulate hook per parameter counts its bucket down; a complete bucket is all-reduced
      asynchronously on the process group's own stream while backward continues (NCCL over NVLink on the GPU box).
      Buckets are always launched in index order, so every rank issues the same sequence of collectives.
    * parameters that never receive a gradient (the reference's dead encoder self-attention, lamp/Layers.py:16-18)
      are found ONCE, by ``calibrate()`` after the first backward (one small MAX all-reduce + one host read, then never
      again): they are laid out behind the live parameters and never communicated; their ``.grad`` stays ``None`` like
      in the reference, so Adam skips them.  ``finish()`` launches whatever the hooks did not (nothing, normally),
      waits, and averages: no per-step host synchronisation.
    * ``local_weight`` (uneven shards): this rank's mean-reduced loss gradient is scaled by ``local_n * world /
      global_n`` before the reduction so that the result equals the single-process gradient of the whole batch."""

    def __init__(self, params: Iterable[torch.nn.Parameter], world: Optional[int] = None, group=None,
                 bucket_mb: float = 16.0):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if (world is None and dist.is_initialized()) else (world or 1)
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.flat: Optional[torch.Tensor] = None
        self.live: List[torch.nn.Parameter] = []
        self.buckets: List[Tuple[int, int, int]] = []  # (start element, end element, number of parameters)
        self._bucket_of = {}
        self._pending: List[int] = []
        self._next = 0
        self._works: list = []
        self._hooks: list = []
        self._fired: set = set()
        self._armed = False
        self.local_weight = 1.0
        self.communicate = True   # False: keep the flat layout but skip the collectives (timing the step without them)
        self.stats = dict(elements=0, buckets=0, launched_from_hooks=0, launched_at_finish=0)
        for p in self.params:  # calibration pass: only record which parameters produce a gradient
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ---- layout, decided once
    def calibrate(self) -> None:
        """Call once, after the first ``backward()`` (before the optimizer step).  Decides the live/dead layout from
        the gradients that exist on ANY rank, moves the existing gradients into the flat buffer and -- for this first
        step only -- reduces them in one blocking call."""
        if self.flat is not None:
            return
        dev = self.params[0].device
        have = torch.tensor([1.0 if (id(p) in self._fired or p.grad is not None) else 0.0 for p in self.params],
                            dtype=torch.float32, device=dev)
        if self.world > 1:
            dist.all_reduce(have, op=dist.ReduceOp.MAX, group=self.group)
        alive = [bool(x) for x in have.tolist()]  # the one host read of the reducer's life
        self.live = [p for p, a in zip(self.params, alive) if a]
        order = list(reversed(self.live))  # gradients of the last layers arrive first
        total = sum(p.numel() for p in order)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off, start, count = 0, 0, 0
        for p in order:
            n = p.numel()
            view = self.flat[off:off + n].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            self._bucket_of[id(p)] = len(self.buckets)
            off += n
            count += 1
            if (off - start) * 4 >= self.bucket_bytes:
                self.buckets.append((start, off, count))
                start, count = off, 0
        if count:
            self.buckets.append((start, off, count))
        self.stats.update(elements=total, buckets=len(self.buckets))
        self._armed = True
        self._reset_step()
        # first step: the hooks were only recording -> reduce everything now
        self.finish()

    def _reset_step(self) -> None:
        self._pending = [b[2] for b in self.buckets]
        self._next = 0
        self._works = []

    # ---- per step
    def zero_grad(self) -> None:
        if self.flat is None:
            for p in self.params:
                p.grad = None
        else:
            self.flat.zero_()
            self._reset_step()

    def _launch(self, i: int) -> None:
        a, b, _ = self.buckets[i]
        chunk = self.flat[a:b]
        if self.local_weight != 1.0:
            chunk.mul_(self.local_weight)
        if self.world > 1 and self.communicate:
            self._works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_grad(self, p) -> None:
        if not self._armed:
            self._fired.add(id(p))
            return
        i = self._bucket_of.get(id(p))
        if i is None:
            return
        self._pending[i] -= 1
        while self._next < len(self.buckets) and self._pending[self._next] <= 0:
            self._launch(self._next)
            self._next += 1
            self.stats['launched_from_hooks'] += 1

    def finish(self) -> int:
        """After ``backward()``: launch the buckets the hooks did not, wait for the collectives (stream wait, not a
        host synchronisation, on NCCL) and average.  -> number of gradient elements reduced."""
        if self.flat is None:
            self.calibrate()
            return self.stats['elements']
        while self._next < len(self.buckets):
            self._launch(self._next)
            self._next += 1
            self.stats['launched_at_finish'] += 1
        for w in self._works:
            w.wait()
        if self.world > 1 and self.communicate:
            self.flat.div_(self.world)
        self._reset_step()
        return self.stats['elements']

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def data_parallel_step(model, optimizer, loss_fn, batch, world: Optional[int] = None, group=None,
                       reducer: Optional[GradientReducer] = None, local_n: Optional[int] = None,
                       global_n: Optional[int] = None) -> float:
    """One training step on this rank's shard: forward, backward, gradient all-reduce, optimizer step.
    ``reducer``: a :class:`GradientReducer` (bucketed, overlapped with backward); without one the step uses the
    single blocking flat all-reduce.  ``local_n`` / ``global_n``: shard and global batch sizes -- when the shards are
    uneven the local mean-loss gradient is weighted by ``local_n * world / global_n`` (see ``allreduce_gradients``).
    Dropout: ranks must draw different masks; seed each rank with ``torch.manual_seed(seed + rank)`` (the attention
    kernels derive their dropout seeds from torch's CPU generator)."""
    w = dist.get_world_size(group) if (world is None and dist.is_initialized()) else (world or 1)
    weight = 1.0 if not (local_n and global_n) else local_n * w / global_n
    if reducer is not None:
        reducer.local_weight = weight
        reducer.zero_grad()
    else:
        optimizer.zero_grad(set_to_none=True)
    loss = loss_fn(model, batch)
    loss.backward()
    if reducer is not None:
        reducer.finish()
    else:
        params = model.get_trainable_parameters() if hasattr(model, 'get_trainable_parameters') else model.parameters()
        allreduce_gradients(list(params), world, group, local_weight=weight)
    optimizer.step()
    return float(loss.detach())
