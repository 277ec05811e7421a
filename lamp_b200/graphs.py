"""CUDA-graph replay of the label-graph forward for serving-style inference.

``LAMP.forward`` (lamp/Models.py:110-137) enqueues ~40 native kernels plus a handful of index computations per
call; at a few milliseconds of GPU work per batch the Python launch path is of the same order and is exposed to host
jitter.  The whole eval forward is shape-static for a fixed ``(batch, seq_len)`` -- the only data-dependent
quantity, the number of non-PAD token rows of the packed batch, lives in a device scalar read by the kernels
(``m_dev``) -- so it can be captured once and replayed with one launch per step.

    runner = lamp_b200.GraphedForward(model, batch=1100, seq_len=300)
    logits, enc_output = runner(src_seq, src_pos)      # host (pinned) or device int64 tensors, [batch, seq_len]

The returned tensors are the graph's static output buffers: they are overwritten by the next call (``.clone()``
to keep them).  The captured forward is the module's own fused path, so results are identical to the eager call.
"""
from __future__ import annotations

import gc
import os
from collections import OrderedDict

import torch

from . import _native as nat

# LAMP.forward serves plain eval calls from a shape-keyed cache of CUDA graphs (EvalGraphCache); LAMP_EVAL_GRAPHS=0: off
EVAL_GRAPHS = os.environ.get('LAMP_EVAL_GRAPHS', '1') != '0'


def _plain_forward(model):
    """The model's forward WITHOUT the eval graph cache (what a capture must record)."""
    return getattr(model, '_forward_impl', model)


class GraphedForward:
    def __init__(self, model, batch: int, seq_len: int, device=None, warmup: int = 2, example=None):
        """``example``: optional ``(src_seq, src_pos)`` used for the warm-up / capture runs (any valid ids; defaults
        to a full-length batch of token id 4)."""
        if model.training:
            raise RuntimeError('GraphedForward captures the eval forward: call model.eval() first')
        p = next(model.parameters())
        nat.require_cuda(p)
        self.model = model
        self.device = torch.device(device) if device is not None else p.device
        self.batch, self.seq_len = batch, seq_len
        self.src_seq = torch.empty((batch, seq_len), dtype=torch.int64, device=self.device)
        self.src_pos = torch.empty((batch, seq_len), dtype=torch.int64, device=self.device)
        if example is not None:
            self.src_seq.copy_(example[0])
            self.src_pos.copy_(example[1])
        else:
            self.src_seq.fill_(4)
            self.src_pos.copy_(torch.arange(1, seq_len + 1, device=self.device).expand(batch, seq_len))
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # weight planes, smem attributes, allocator pools: all outside the graph
                _plain_forward(model)((self.src_seq, self.src_pos), None, None, None)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        from . import ops
        n0 = ops.STATS.launches
        with torch.no_grad(), torch.cuda.graph(self.graph):
            logits, enc_output, *_ = _plain_forward(model)((self.src_seq, self.src_pos), None, None, None)
        self.kernels_per_replay = ops.STATS.launches - n0   # native kernel launches captured in the graph
        self.logits, self.enc_output = logits, enc_output

    def load(self, src_seq: torch.Tensor, src_pos: torch.Tensor) -> None:
        """Copy a batch of token / position ids into the graph's input buffers (H2D when they live on the host)."""
        if tuple(src_seq.shape) != (self.batch, self.seq_len) or tuple(src_pos.shape) != (self.batch, self.seq_len):
            raise RuntimeError(f'GraphedForward was captured for [{self.batch}, {self.seq_len}] inputs, got '
                               f'{tuple(src_seq.shape)} / {tuple(src_pos.shape)}')
        self.src_seq.copy_(src_seq, non_blocking=True)
        self.src_pos.copy_(src_pos, non_blocking=True)

    def replay(self):
        """The graph reads the weight PLANES, not the parameters: they are brought up to date (in place, same buffers)
        first, so a replay after ``optimizer.step()`` / ``load_state_dict`` sees the new weights."""
        from . import ops
        ops.refresh_weight_planes(self.model)
        self.graph.replay()
        ops.STATS.launches += self.kernels_per_replay
        return self.logits, self.enc_output

    def __call__(self, src_seq: torch.Tensor, src_pos: torch.Tensor):
        self.load(src_seq, src_pos)
        return self.replay()


class EvalGraphCache:
    """Shape-keyed cache of CUDA graphs of the fused eval forward, used by ``LAMP.forward`` itself so that the
    reference's own eval loop (test.py:41: ``model(src, adj, None, None)`` batch after batch) gets graph-replay launch
    cost without calling any extra API.

    * key: (batch, padded length, device, precision / path switches).  The reference's loader pads every batch to its
      longest document (utils/data_loader.py:261-279), so the token length varies from batch to batch: it is rounded up
      to a multiple of ``t_bucket`` with PAD tokens (id 0, position 0).  Those are exactly the rows the padding-aware
      encoder drops and the decoder never attends to, so the logits are unchanged; ``enc_output`` is cut back to T.
    * the first call with a key runs eagerly (one-off shapes are never captured), the second captures, later ones replay.
    * weights: planes are refreshed in place before every replay (``ops.refresh_weight_planes``), so an eval epoch
      after a training epoch sees the trained weights; everything else the kernels read (LayerNorm affine, biases,
      embedding tables) is read from the parameters themselves.
    * all graphs share one memory pool (replays are stream-ordered and the outputs handed out are copies), LRU-bounded."""

    def __init__(self, model, max_graphs: int = 16, t_bucket: int = 32):
        self.model = model
        self.max_graphs, self.t_bucket = max_graphs, t_bucket
        self.entries: 'OrderedDict[tuple, dict]' = OrderedDict()
        self.seen: dict = {}
        self.pool = None
        self.replays = self.captures = self.eager_calls = 0

    def __deepcopy__(self, memo):
        return None   # a copied model (copy.deepcopy: train.py:45) builds its own cache on first use; graphs do not copy

    def __reduce__(self):
        return (type(None), ())   # ... and a pickled one carries none

    def _key(self, B: int, Tb: int, dev) -> tuple:
        from . import ops
        return (B, Tb, dev.index, ops.default_precision(), ops.PADDING_AWARE, ops.DEFER_LAYERNORM, ops.FUSE_LAYERNORM)

    def _capture(self, key, B: int, Tb: int, dev, src_seq, src_pos, T: int) -> dict:
        from . import ops
        s_seq = torch.zeros((B, Tb), dtype=torch.int64, device=dev)
        s_pos = torch.zeros((B, Tb), dtype=torch.int64, device=dev)
        s_seq[:, :T].copy_(src_seq)
        s_pos[:, :T].copy_(src_pos)
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        graph = torch.cuda.CUDAGraph()
        n0 = ops.STATS.launches
        torch.cuda.synchronize(dev)
        ops.DEFER_UNPACK = True
        try:
            with torch.no_grad(), torch.cuda.graph(graph, pool=self.pool):
                logits, enc_output, _ = self.model._forward_impl((s_seq, s_pos), None, None, None)
        finally:
            ops.DEFER_UNPACK = False
        packed = getattr(enc_output, '_lamp_packed', None)
        unpack = None if packed is None else packed.get('unpack')
        e = dict(graph=graph, seq=s_seq, pos=s_pos, logits=logits, enc=enc_output, unpack=unpack,
                 kernels=ops.STATS.launches - n0, T=T)
        self.entries[key] = e
        self.captures += 1
        while len(self.entries) > self.max_graphs:
            self.entries.popitem(last=False)
        return e

    def run(self, src_seq: torch.Tensor, src_pos: torch.Tensor):
        from . import ops
        B, T = src_seq.shape
        dev = src_seq.device
        tb = self.t_bucket
        Tb = (T + tb - 1) // tb * tb
        key = self._key(B, Tb, dev)
        e = self.entries.get(key)
        if e is None:
            n = self.seen.get(key, 0)
            self.seen[key] = n + 1
            if n == 0 or src_seq.dtype != torch.int64 or src_pos.dtype != torch.int64:
                self.eager_calls += 1
                logits, enc_output, _ = self.model._forward_impl((src_seq, src_pos), None, None, None)
                return logits, enc_output
            ops.refresh_weight_planes(self.model)
            e = self._capture(key, B, Tb, dev, src_seq, src_pos, T)
        else:
            self.entries.move_to_end(key)
            ops.refresh_weight_planes(self.model)
            if T == Tb:
                e['seq'].copy_(src_seq, non_blocking=True)
                e['pos'].copy_(src_pos, non_blocking=True)
            else:
                e['seq'][:, :T].copy_(src_seq, non_blocking=True)
                e['pos'][:, :T].copy_(src_pos, non_blocking=True)
                if T < e['T']:  # the previous batch of this bucket was longer: blank its tail
                    e['seq'][:, T:].zero_()
                    e['pos'][:, T:].zero_()
            e['T'] = T
        e['graph'].replay()
        self.replays += 1
        ops.STATS.launches += e['kernels']
        if e['unpack'] is not None:
            # padding-aware encoder: un-pack the packed rows the replay left in its static buffer into a FRESH dense
            # [B, T, D] tensor (the same gather the eager forward ends with -- no copy out of a static buffer)
            x32, src_row = e['unpack']
            idx = src_row if T == Tb else src_row.view(B, Tb)[:, :T].reshape(-1)
            enc = ops.gather_rows(x32, idx, x32.shape[1]).view(B, T, x32.shape[1])
        else:
            enc = e['enc'].clone() if T == Tb else e['enc'][:, :T].clone()
        return e['logits'].clone(), enc


class GraphedTrainStep:
    """CUDA-graph replay of one training step of the label-graph model: ``zero_grad -> LAMP.forward -> loss ->
    backward`` (train.py:28-48 without the data loading), optimizer step outside the graph.

    A training step at the reference's batch size (32) is ~350 native kernel launches plus torch glue for ~6 ms of
    GPU work: launched from Python it is host-bound.  The step is shape-static for a fixed ``(batch, seq_len)``, and
    the one thing that must change between replays -- the dropout masks -- is handled on the device: the attention
    kernels add a device-side counter (``ops.TRAIN_SEED_DEV``, advanced inside the graph) to their baked-in seeds,
    torch's own dropouts use the graph-safe Philox offsets of the CUDA generator.

        step = lamp_b200.GraphedTrainStep(model, loss_fn, batch=32, seq_len=300)
        for src_seq, src_pos, target in loader:
            loss = step(src_seq, src_pos, target)      # gradients are in p.grad (static tensors)
            optimizer.step()                           # (or pass a capturable optimizer to put it inside the graph)

    ``loss_fn(logits, target) -> scalar``.  Parameter ``.grad`` tensors are allocated inside the graph's memory pool
    and overwritten by every replay."""

    def __init__(self, model, loss_fn, batch: int, seq_len: int, device=None, warmup: int = 3, example=None,
                 optimizer=None):
        """``optimizer`` (optional): a graph-capturable optimizer -- e.g. ``torch.optim.Adam(..., capturable=True)`` --
        whose ``step()`` is then part of the captured graph, so that a whole training iteration is one launch.  The
        warm-up iterations before the capture DO update the parameters in that case."""
        from . import ops
        if not model.training:
            raise RuntimeError('GraphedTrainStep captures the training step: call model.train() first')
        p = next(model.parameters())
        nat.require_cuda(p)
        self.model, self.loss_fn = model, loss_fn
        self.optimizer_in_graph = optimizer is not None
        self.device = torch.device(device) if device is not None else p.device
        self.batch, self.seq_len = batch, seq_len
        n_labels = model.decoder.n_tgt_vocab
        self.src_seq = torch.empty((batch, seq_len), dtype=torch.int64, device=self.device)
        self.src_pos = torch.empty((batch, seq_len), dtype=torch.int64, device=self.device)
        self.target = torch.zeros((batch, n_labels), dtype=torch.float32, device=self.device)
        if example is not None:
            self.src_seq.copy_(example[0])
            self.src_pos.copy_(example[1])
            if len(example) > 2:
                self.target.copy_(example[2])
        else:
            self.src_seq.fill_(4)
            self.src_pos.copy_(torch.arange(1, seq_len + 1, device=self.device).expand(batch, seq_len))
        self.seed_dev = torch.zeros((1,), dtype=torch.int64, device=self.device)
        ops.TRAIN_SEED_DEV = self.seed_dev

        def body():
            self.seed_dev.add_(1)
            model.zero_grad(set_to_none=True)
            logits, _, _ = model((self.src_seq, self.src_pos), None, None, None)
            loss = loss_fn(logits, self.target)
            loss.backward()
            if optimizer is not None:
                optimizer.step()
            return loss.detach()

        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        # an autograd graph of an earlier iteration that is still referenced (a kept loss, a reference cycle waiting
        # for the collector) keeps its AccumulateGrad nodes and THEIR streams alive; the engine would then make that
        # stream wait on the capturing one, which invalidates the capture.  Collect what can be collected first.
        gc.collect()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.STATS.launches
        with torch.cuda.graph(self.graph):
            self.loss = body()
        self.kernels_per_replay = ops.STATS.launches - n0

    def __call__(self, src_seq: torch.Tensor, src_pos: torch.Tensor, target: torch.Tensor):
        self.src_seq.copy_(src_seq, non_blocking=True)
        self.src_pos.copy_(src_pos, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self.graph.replay()
        from . import ops
        ops.STATS.launches += self.kernels_per_replay
        if self.optimizer_in_graph:
            # the replay changed the parameters without touching their tensor versions: invalidate every cached
            # weight-plane signature so that the next eval forward (eager or captured) re-splits the new weights
            ops.bump_weights_epoch()
        return self.loss
