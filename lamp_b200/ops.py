"""Host-side orchestration of the native kernels (level-1 C ABI) for the label-graph path.

Activations travel between kernels as :class:`Act`: an fp32 matrix ``[rows, D]`` (residual / API tensor) plus
its split-bf16 planes (tensor-core operand form).  Producers (LayerNorm, embedding, GEMM epilogues) emit the planes
directly, so no separate conversion pass runs inside a layer.  Weights are split once and cached per parameter
version.  Everything is enqueued on torch's current CUDA stream; torch only provides memory and streams.

The training path lives here as autograd Functions over the same kernels: ``SDPAFunction``
(attention core with in-kernel dropout / ``lamp_attn_core_bwd``), ``LinearFunction`` (tcgen05 GEMMs for y, dx and dW),
``LayerNormFunction``, ``DiagProjFunction``; see DESIGN.md section 4b.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
import weakref
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _native as nat

_DEFAULT_PRECISION = nat.PREC_FP32


def set_default_precision(name: str) -> None:
    """'fp32' (3-term split-bf16 tensor-core products; meets the 1e-3 fp32 tolerance) or 'bf16'."""
    global _DEFAULT_PRECISION
    _DEFAULT_PRECISION = {'fp32': nat.PREC_FP32, 'bf16': nat.PREC_BF16}[name]


def default_precision() -> int:
    return _DEFAULT_PRECISION


class LaunchStats:
    """Kernel-launch accounting (always on, a counter per entry point) and optional CUDA-event timing of every
    native call (``with ops.STATS.timed(): ...``), used by bench.py for the per-kernel roofline numbers."""

    def __init__(self):
        self.launches = 0
        self.by_kernel: Dict[str, int] = {}
        self._events = None

    def reset(self):
        self.launches = 0
        self.by_kernel = {}

    def timed(self):
        stats = self

        class _Ctx:
            def __enter__(self_inner):
                stats._events = []
                return stats

            def __exit__(self_inner, *exc):
                return False
        return _Ctx()

    def stop_timing(self):
        """-> {kernel: dict(calls, ms, flops, bytes)}; synchronises."""
        ev, self._events = self._events or [], None
        torch.cuda.synchronize()
        out: Dict[str, dict] = {}
        counts: Dict[int, int] = {}  # device row counts, read back once per tensor
        for name, e0, e1, flops, nbytes, rows_dev in ev:
            d = out.setdefault(name, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
            d['calls'] += 1
            d['ms'] += e0.elapsed_time(e1)
            if callable(flops) or callable(nbytes):
                # padding-aware launches: the rows actually processed are only known on the device
                m = None
                if rows_dev is not None:
                    key = id(rows_dev)
                    if key not in counts:
                        counts[key] = int(rows_dev.sum().item())
                    m = counts[key]
                flops = flops(m) if callable(flops) else flops
                nbytes = nbytes(m) if callable(nbytes) else nbytes
            d['flops'] += flops
            d['bytes'] += nbytes
        return out

    def call(self, name: str, n_kernels: int, fn, args, flops=0.0, nbytes=0.0, rows_dev=None):
        """``flops`` / ``nbytes``: algorithmic work of the launch -- numbers, or functions of the row count that the
        kernel reads from the device (``rows_dev``: int32 tensor whose sum is that count; ``None`` is passed to the
        function when the launch is dense), evaluated after the timed region."""
        self.launches += n_kernels
        self.by_kernel[name] = self.by_kernel.get(name, 0) + n_kernels
        if self._events is None:
            nat.check(fn(*args), name)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nat.check(fn(*args), name)
        e1.record()
        self._events.append((name, e0, e1, flops, nbytes, rows_dev))


STATS = LaunchStats()


@dataclass
class DeferredLN:
    """A LayerNorm that has not been applied yet: the planes of the owning :class:`Act` hold the PRE-norm tensor
    ``y`` and ``stats`` the per-row partial sums ``{sum y, sum y^2}`` (float32 ``[rows, nparts, 2]``) written by the
    producing GEMM.  Consumers (the next GEMM's epilogue, the residual add, the label projection) normalise on the
    fly, so the LayerNorm of lamp/SubLayers.py:117,141 costs no pass over HBM of its own."""
    stats: torch.Tensor
    nparts: int
    gamma: torch.Tensor
    beta: torch.Tensor
    eps: float


@dataclass
class Act:
    """fp32 activation ``[rows, cols]`` and/or its planes.  ``bcast``: logical row count when the ``rows``
    physical rows are shared by every sample (label embeddings: rows = L, logical rows = B*L)."""
    f32: Optional[torch.Tensor]
    hi: Optional[torch.Tensor]
    lo: Optional[torch.Tensor]
    rows: int
    cols: int
    bcast_rows: int = 0
    m_dev: Optional[torch.Tensor] = None  # device int32 scalar: rows actually in use (padding-aware packed batch)
    ln: Optional[DeferredLN] = None       # set: hi/lo are the pre-norm tensor of a deferred LayerNorm (f32 is None)

    @property
    def has_planes(self) -> bool:
        return self.hi is not None


def _empty_planes(rows: int, cols: int, prec: int, device) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    hi = torch.empty((rows, cols), dtype=torch.bfloat16, device=device)
    lo = torch.empty((rows, cols), dtype=torch.bfloat16, device=device) if prec == nat.PREC_FP32 else None
    return hi, lo


def split(x: torch.Tensor, prec: int, out=None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """fp32 ``[..., cols]`` (contiguous) -> planes ``[rows, cols]``.  ``out``: existing ``(hi, lo)`` buffers of that
    shape to overwrite (weight planes refreshed in place keep the addresses captured CUDA graphs hold)."""
    nat.require_cuda(x)
    x = x.contiguous()
    cols = x.shape[-1]
    rows = x.numel() // cols
    hi, lo = _empty_planes(rows, cols, prec, x.device) if out is None else out
    STATS.call('split_planes', 1, nat.lib().lamp_split_planes,
               (x.data_ptr(), rows, cols, cols, hi.data_ptr(), nat.ptr(lo), cols, nat.stream()),
               nbytes=rows * cols * (4 + (4 if lo is not None else 2)))
    return hi, lo


def act_from_tensor(x: torch.Tensor, prec: int) -> Act:
    """Wrap an fp32 tensor ``[..., D]``; reuse planes stashed on the tensor by the producing module if still valid."""
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    cols = x.shape[-1]
    rows = x.numel() // cols
    stash = getattr(x, '_lamp_planes', None)
    if stash is not None:
        hi, lo, ver, sprec = stash
        if ver == x._version and sprec == prec and hi.shape == (rows, cols):
            return Act(x.view(rows, cols), hi, lo, rows, cols)
    hi, lo = split(x, prec)
    return Act(x.view(rows, cols), hi, lo, rows, cols)


def stash_planes(x: torch.Tensor, act: Act, prec: int) -> torch.Tensor:
    """Attach the planes to the API tensor so that the next lamp_b200 module can skip the split pass."""
    if act.has_planes:
        x._lamp_planes = (act.hi, act.lo, x._version, prec)
    return x


# Bumped by whatever changes parameters WITHOUT advancing their tensor ``_version`` -- replays of a CUDA graph that
# contains an optimizer step (graphs.GraphedTrainStep with a capturable optimizer).  Part of every weight-plane signature.
WEIGHTS_EPOCH = 0


def bump_weights_epoch() -> None:
    global WEIGHTS_EPOCH
    WEIGHTS_EPOCH += 1


class _PlaneEntry:
    __slots__ = ('sig', 'deps', 'build', 'bufs')

    def __init__(self, sig, deps, build, bufs):
        self.sig, self.deps, self.build, self.bufs = sig, deps, build, bufs


class WeightPlanes:
    """Split-bf16 planes of (a concatenation of) 2-D weights, cached against the parameters' versions.

    * A stale entry is refreshed IN PLACE (same buffers): CUDA graphs that captured a forward keep pointing at valid,
      up-to-date planes after an optimizer step / ``load_state_dict`` (see :func:`refresh_weight_planes`).
    * Entries are keyed by device as well: ``nn.DataParallel`` replicas (the reference's multi-GPU mode, main.py:106-108)
      are shallow copies that share this object, each thread works on its own device's entry under a lock.
    * An entry keeps its parameters alive, so a signature (address, version, shape) can never be recycled."""

    def __init__(self):
        self._cache: Dict[tuple, _PlaneEntry] = {}
        self._lock = threading.Lock()

    # A copy of a module (copy.deepcopy(model) -- the reference does that at train.py:45 -- or pickling) starts with an
    # empty cache of its own: plane buffers belong to the parameters they were split from, and locks do not copy.
    def __deepcopy__(self, memo):
        return WeightPlanes()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()

    @staticmethod
    def _sig(deps, extra) -> tuple:
        return tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in deps) + extra + (WEIGHTS_EPOCH,)

    def _lookup(self, ck: tuple, deps: tuple, extra: tuple, build):
        sig = self._sig(deps, extra)
        hit = self._cache.get(ck)
        if hit is not None and hit.sig == sig:
            return hit.bufs
        with self._lock:
            hit = self._cache.get(ck)
            if hit is not None and hit.sig == sig:
                return hit.bufs
            with torch.no_grad():
                bufs = build(deps, None if hit is None else hit.bufs)
            self._cache[ck] = _PlaneEntry(sig, deps, build, bufs)  # replaced as a whole: readers never see a torn entry
            return bufs

    def refresh(self) -> int:
        """Re-split every entry whose parameters changed (in place).  -> number of entries refreshed."""
        n = 0
        for ck, e in list(self._cache.items()):
            extra = e.sig[len(e.deps):-1]
            if e.sig != self._sig(e.deps, extra):
                self._lookup(ck, e.deps, extra, e.build)
                n += 1
        return n

    @staticmethod
    def _cat(params) -> torch.Tensor:
        mats = [p.detach().reshape(p.shape[0], -1).float() for p in params]  # Conv1d [out,in,1] -> [out,in]
        return mats[0] if len(mats) == 1 else torch.cat(mats, dim=0)

    def get(self, key: str, params, prec: int) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        params = tuple(params)

        def build(deps, old):
            w = self._cat(deps)
            if old is not None and old[0].shape == w.shape and old[0].device == w.device:
                return split(w, prec, out=old)
            return split(w, prec)
        return self._lookup((key, params[0].device.index, prec), params, (prec,), build)

    def get_folded(self, key: str, params, ln: 'DeferredLN', bias, prec: int):
        """Weights of a GEMM whose A operand is a deferred LayerNorm (gamma, beta):
        ``LN(y) W^T + bias = rstd * (y (W*gamma)^T - mean * colsum) + (bias + W beta)``.
        -> (planes of W*diag(gamma), colsum [N] fp32, folded bias [N] fp32), cached per parameter versions."""
        params = tuple(params)
        n = len(params)
        deps = params + (ln.gamma, ln.beta) + ((bias,) if bias is not None else ())

        def build(deps, old):
            w = self._cat(deps[:n])
            gamma, beta = deps[n], deps[n + 1]
            b = deps[n + 2] if len(deps) > n + 2 else None
            wg = (w * gamma.detach().float().unsqueeze(0)).contiguous()
            reuse = old is not None and old[0].shape == wg.shape and old[0].device == wg.device
            hi, lo = split(wg, prec, out=(old[0], old[1]) if reuse else None)
            # colsum from the operand planes themselves, so that `y Wg^T - mean * colsum` cancels exactly
            colsum = (hi.float() if lo is None else hi.float() + lo.float()).sum(dim=1).contiguous()
            biasf = w.double() @ beta.detach().double()
            if b is not None:
                biasf = biasf + b.detach().double()
            biasf = biasf.float().contiguous()
            if reuse:
                old[2].copy_(colsum)
                old[3].copy_(biasf)
                return old
            return hi, lo, colsum, biasf
        return self._lookup((key + '@ln', params[0].device.index, prec), deps, (prec, bias is None), build)


def refresh_weight_planes(model: torch.nn.Module) -> int:
    """Bring every cached weight-plane buffer of ``model`` up to date with its parameters, in place.  Called before a
    captured forward is replayed (graphs.GraphedForward, the eval graph cache of ``LAMP.forward``): the graph reads the
    plane buffers, not the parameters.  A quick (version, address) check makes the common no-change case ~15 us."""
    state = model.__dict__.get('_lamp_planes_state')
    if state is None:
        plist = list(model.parameters())
        holders = [m._wp for m in model.modules() if isinstance(getattr(m, '_wp', None), WeightPlanes)]
        state = model.__dict__['_lamp_planes_state'] = dict(params=plist, holders=holders, stamp=None)
    stamp = (WEIGHTS_EPOCH,) + tuple((p._version, p.data_ptr()) for p in state['params'])
    if stamp == state['stamp']:
        return 0
    n = sum(h.refresh() for h in state['holders'])
    state['stamp'] = stamp
    return n


def gemm(a_hi, a_lo, lda: int, w_hi, w_lo, ldw: int, M: int, N: int, K: int, prec: int, *, bias=None, relu=False,
         residual=None, ldr: int = 0, resid_mod: int = 0, out_f32=None, ldo: int = 0, out_hi=None, out_lo=None,
         ldp: int = 0, m_dev=None) -> None:
    pl = 4 if prec == nat.PREC_FP32 else 2  # bytes per element of a plane pair
    row_bytes = K * pl + (N * 4 if out_f32 is not None else 0) + (N * pl if out_hi is not None else 0) + \
        (N * 4 if residual is not None and not resid_mod else 0)

    def rows(m):
        return M if m is None else min(M, m)
    STATS.call('gemm_planes', 1, nat.lib().lamp_gemm_planes,
               (nat.ptr(a_hi), nat.ptr(a_lo), lda, nat.ptr(w_hi), nat.ptr(w_lo), ldw, M, N, K, prec, nat.ptr(bias),
                int(relu), nat.ptr(residual), ldr, resid_mod, nat.ptr(out_f32), ldo, nat.ptr(out_hi), nat.ptr(out_lo),
                ldp, nat.ptr(m_dev), nat.stream()), flops=lambda m: 2.0 * rows(m) * N * K,
               nbytes=lambda m: rows(m) * row_bytes + N * K * pl, rows_dev=m_dev)


def gemm_drop(a_hi, a_lo, lda: int, w_hi, w_lo, ldw: int, M: int, N: int, K: int, *, bias, p_drop: float, seed: int,
              residual, ldr: int, out_f32, ldo: int) -> None:
    """Training epilogue (3-term products): ``out_f32 = dropout(A W^T + bias) + residual`` in ONE kernel, with the
    counter-hash mask of ``dropout_add`` / ``dropout_split`` (``lamp_gemm_planes_drop``)."""
    STATS.call('gemm_planes', 1, nat.lib().lamp_gemm_planes_drop,
               (nat.ptr(a_hi), nat.ptr(a_lo), lda, nat.ptr(w_hi), nat.ptr(w_lo), ldw, M, N, K, nat.ptr(bias), float(p_drop),
                int(seed), _seed_dev_ptr(), nat.ptr(residual), ldr, 0, nat.ptr(out_f32), ldo, nat.stream()),
               flops=2.0 * M * N * K, nbytes=M * (K * 4 + N * 8) + N * K * 4)


def linear_planes(x: Act, w_hi, w_lo, N: int, prec: int, *, bias=None, relu=False, ld_pad: int = 0) -> Act:
    """planes(x) @ W^T (+bias)(ReLU) -> planes only (operand for the next tensor-core stage).  ``ld_pad``: extra
    (unused) columns in the row pitch of the result, see ``KV_LD_PAD``; the returned Act's ``cols`` is the pitch."""
    ld = N + ld_pad
    hi, lo = _empty_planes(x.rows, ld, prec, x.hi.device)
    gemm(x.hi, x.lo, x.cols, w_hi, w_lo, x.cols, x.rows, N, x.cols, prec, bias=bias, relu=relu, out_hi=hi, out_lo=lo,
         ldp=ld, m_dev=x.m_dev)
    return Act(None, hi, lo, x.rows, ld, x.bcast_rows, x.m_dev)


# Row pitch padding (in bf16 columns) of the label<-input K|V plane matrix.  Its natural pitch, 2*H*d*n_layers columns =
# 4096 B at the bench shape, is a power of two: the attention core's K/V tiles are 64 row segments of 128 B each exactly
# one pitch apart, which aliases onto a few HBM channels.  A pitch that is not a power of two spreads them.
KV_LD_PAD = int(os.environ.get('LAMP_KV_LD_PAD', '0'))


def project(x: Act, wp: WeightPlanes, key: str, params, N: int, prec: int, *, bias=None, relu=False, ld_pad: int = 0) -> Act:
    """``x @ cat(params)^T (+bias)(ReLU)`` -> planes.  ``x`` may be a deferred LayerNorm: the normalisation is then
    folded into the weights and the epilogue of this GEMM (see :class:`DeferredLN`)."""
    if x.ln is None:
        w_hi, w_lo = wp.get(key, params, prec)
        return linear_planes(x, w_hi, w_lo, N, prec, bias=bias, relu=relu, ld_pad=ld_pad)
    ln = x.ln
    wg_hi, wg_lo, colsum, biasf = wp.get_folded(key, params, ln, bias, prec)
    hi, lo = _empty_planes(x.rows, N, prec, x.hi.device)
    M, K = x.rows, x.cols
    pl = 4 if prec == nat.PREC_FP32 else 2
    STATS.call('gemm_planes', 1, nat.lib().lamp_gemm_planes_dln,
               (x.hi.data_ptr(), nat.ptr(x.lo), K, ln.stats.data_ptr(), ln.nparts, float(ln.eps), wg_hi.data_ptr(),
                nat.ptr(wg_lo), K, colsum.data_ptr(), biasf.data_ptr(), M, N, K, prec, int(relu), hi.data_ptr(),
                nat.ptr(lo), N, nat.ptr(x.m_dev), nat.stream()),
               flops=lambda m: 2.0 * (M if m is None else min(M, m)) * N * K,
               nbytes=lambda m: (M if m is None else min(M, m)) * (K * pl + N * pl + 8 * ln.nparts) + N * K * pl,
               rows_dev=x.m_dev)
    return Act(None, hi, lo, x.rows, N, x.bcast_rows, x.m_dev)


def materialize(a: Act, prec: int, *, want_f32: bool = True, want_planes: bool = True, index=None) -> Act:
    """Apply a deferred LayerNorm -> ordinary Act (fp32 and/or planes).  ``index`` (int64 [rows_out]): output row r
    is the LayerNorm of source row index[r] (un-packing gather of the encoder output)."""
    if a.ln is None:
        return a
    ln = a.ln
    rows = a.rows if index is None else index.numel()
    D = a.cols
    dev = a.hi.device
    out = torch.empty((rows, D), dtype=torch.float32, device=dev) if want_f32 else None
    hi, lo = _empty_planes(rows, D, prec, dev) if want_planes else (None, None)
    m_dev = a.m_dev if index is None else None
    pl_in = 2 if a.lo is None else 4
    pl_out = (4 if want_f32 else 0) + (0 if hi is None else (4 if lo is not None else 2))
    STATS.call('ln_apply', 1, nat.lib().lamp_ln_apply,
               (a.hi.data_ptr(), nat.ptr(a.lo), ln.stats.data_ptr(), ln.nparts, ln.gamma.data_ptr(), ln.beta.data_ptr(),
                float(ln.eps), rows, D, nat.ptr(index), nat.ptr(out), nat.ptr(hi), nat.ptr(lo), nat.ptr(m_dev),
                nat.stream()),
               nbytes=lambda m: (rows if m is None else min(rows, m)) * D * (pl_in + pl_out), rows_dev=m_dev)
    return Act(out, hi, lo, rows, D, 0, m_dev)


def act_f32(a: Act) -> torch.Tensor:
    """fp32 view of an activation; reconstructed from the planes when the fp32 copy was never materialised."""
    if a.ln is not None:
        return materialize(a, nat.PREC_FP32, want_f32=True, want_planes=False).f32
    if a.f32 is None:
        a.f32 = a.hi.float() if a.lo is None else a.hi.float() + a.lo.float()
    return a.f32


def linear_residual_f32(x: Act, w_hi, w_lo, N: int, prec: int, residual: Act, *, bias=None) -> torch.Tensor:
    """planes(x) @ W^T (+bias) + residual -> fp32 [rows, N] (pre-LayerNorm).  The residual is read as fp32 when the
    activation has an fp32 copy, otherwise reconstructed from its planes inside the epilogue."""
    y = torch.empty((x.rows, N), dtype=torch.float32, device=x.hi.device)
    mod = residual.rows if residual.bcast_rows else 0
    if residual.f32 is not None:
        gemm(x.hi, x.lo, x.cols, w_hi, w_lo, x.cols, x.rows, N, x.cols, prec, bias=bias, residual=residual.f32,
             ldr=residual.cols, resid_mod=mod, out_f32=y, ldo=N, m_dev=x.m_dev)
        return y
    M, K = x.rows, x.cols
    pl = 4 if prec == nat.PREC_FP32 else 2
    STATS.call('gemm_planes', 1, nat.lib().lamp_gemm_planes_pres,
               (nat.ptr(x.hi), nat.ptr(x.lo), K, nat.ptr(w_hi), nat.ptr(w_lo), K, M, N, K, prec, nat.ptr(bias),
                nat.ptr(residual.hi), nat.ptr(residual.lo), residual.cols, mod, y.data_ptr(), N, None, None, 0,
                nat.ptr(x.m_dev), nat.stream()), flops=lambda m: 2.0 * (M if m is None else min(M, m)) * N * K,
               nbytes=lambda m: (M if m is None else min(M, m)) * (K * pl + N * 4 + N * pl) + N * K * pl,
               rows_dev=x.m_dev)
    return y


def linear_residual_deferred(x: Act, w_hi, w_lo, N: int, prec: int, residual: Act, gamma, beta, eps: float, *,
                             bias=None) -> Act:
    """``y = planes(x) @ W^T (+bias) + residual`` -> planes of y + its row statistics: an Act whose LayerNorm
    (gamma, beta, eps) is deferred to the consumers.  The residual may itself be a deferred LayerNorm."""
    M, K = x.rows, x.cols
    dev = x.hi.device
    hi, lo = _empty_planes(M, N, prec, dev)
    nparts = nat.lib().lamp_gemm_stats_parts(N)
    stats = torch.empty((M, nparts, 2), dtype=torch.float32, device=dev)
    mod = residual.rows if residual.bcast_rows else 0
    rl = residual.ln
    res_f32 = residual.f32 if rl is None else None
    res_hi, res_lo = (None, None) if res_f32 is not None else (residual.hi, residual.lo)
    pl = 4 if prec == nat.PREC_FP32 else 2
    row_bytes = K * pl + N * pl + 8 * nparts + (0 if mod else N * (4 if res_f32 is not None else pl))
    STATS.call('gemm_planes', 1, nat.lib().lamp_gemm_planes_rstats,
               (x.hi.data_ptr(), nat.ptr(x.lo), K, w_hi.data_ptr(), nat.ptr(w_lo), K, M, N, K, prec, nat.ptr(bias),
                nat.ptr(res_f32), nat.ptr(res_hi), nat.ptr(res_lo), residual.cols, mod,
                None if rl is None else rl.stats.data_ptr(), 0 if rl is None else rl.nparts,
                0.0 if rl is None else float(rl.eps), None if rl is None else rl.gamma.data_ptr(),
                None if rl is None else rl.beta.data_ptr(), hi.data_ptr(), nat.ptr(lo), N, stats.data_ptr(),
                nat.ptr(x.m_dev), nat.stream()),
               flops=lambda m: 2.0 * (M if m is None else min(M, m)) * N * K,
               nbytes=lambda m: (M if m is None else min(M, m)) * row_bytes + N * K * pl, rows_dev=x.m_dev)
    return Act(None, hi, lo, M, N, 0, x.m_dev, DeferredLN(stats, nparts, gamma, beta, eps))


NATIVE_ATTENTION_BACKWARD = os.environ.get('LAMP_NATIVE_ATTN_BWD', '1') != '0'  # training: attention core fwd+bwd native
ELIDE_DEAD_ENCODER_ATTENTION = True  # training / composed path: skip the encoder self-attention whose output is discarded
DEFER_UNPACK = False  # set by graphs.EvalGraphCache while it captures: the encoder leaves the un-packing gather to the cache
PADDING_AWARE = True  # GraphEncoder/GraphDecoder compute only non-PAD token rows (results identical, see Encoders.py)
# fc / w_2 GEMMs emit pre-norm planes + row statistics and the LayerNorm is applied by the consumers (no LayerNorm
# kernels inside the stack).  LAMP_DEFER_LN=0/1 overrides the default (benchmarking aid; results agree to fp32 rounding).
DEFER_LAYERNORM = os.environ.get('LAMP_DEFER_LN', '0') != '0'
FUSE_LAYERNORM = os.environ.get('LAMP_FUSE_LN', '0') != '0'  # True: fc / w_2 GEMM epilogue normalises the row on chip when 256 < d_model <= 512 (measured
# slower than GEMM + LayerNorm kernels on B200: the exposed two-pass epilogue costs more than the HBM round trip saves)


def linear_residual_ln(x: Act, w_hi, w_lo, N: int, prec: int, residual: Act, gamma, beta, eps: float, *, bias=None,
                       want_planes: bool = True, want_f32: bool = True) -> Act:
    """LayerNorm(planes(x) @ W^T (+bias) + residual) -> Act (fp32 + planes).  One kernel when the output row fits
    the 512-column TMEM accumulator (the pre-norm tensor never reaches HBM); GEMM + LayerNorm kernels otherwise."""
    if DEFER_LAYERNORM and x.ln is None and N % 8 == 0:
        out = linear_residual_deferred(x, w_hi, w_lo, N, prec, residual, gamma, beta, eps, bias=bias)
        if want_f32:  # the caller needs the real tensor (API boundary, intermediate predictions)
            out = materialize(out, prec, want_f32=True, want_planes=want_planes)
        return out
    if residual.ln is not None:
        residual = materialize(residual, prec, want_f32=True, want_planes=False)
    if not (FUSE_LAYERNORM and 256 < N <= 512 and residual.f32 is not None):
        y = linear_residual_f32(x, w_hi, w_lo, N, prec, residual, bias=bias)
        return layernorm(y, gamma, beta, eps, prec, want_planes=want_planes, want_f32=want_f32, m_dev=x.m_dev)
    dev = x.hi.device
    out = torch.empty((x.rows, N), dtype=torch.float32, device=dev)
    hi, lo = _empty_planes(x.rows, N, prec, dev) if want_planes else (None, None)
    mod = residual.rows if residual.bcast_rows else 0
    M, K = x.rows, x.cols
    pl = 4 if prec == nat.PREC_FP32 else 2
    nbytes = M * K * pl + N * K * pl + M * N * 4 + (M * N * pl if want_planes else 0) + (0 if mod else M * N * 4)
    STATS.call('gemm_ln_planes', 1, nat.lib().lamp_gemm_ln_planes,
               (nat.ptr(x.hi), nat.ptr(x.lo), K, nat.ptr(w_hi), nat.ptr(w_lo), K, M, N, K, prec, nat.ptr(bias),
                nat.ptr(residual.f32), residual.cols, mod, gamma.data_ptr(), beta.data_ptr(), float(eps),
                out.data_ptr(), N, nat.ptr(hi), nat.ptr(lo), N, nat.stream()), flops=2.0 * M * N * K, nbytes=nbytes)
    return Act(out, hi, lo, x.rows, N)


def layernorm(y: torch.Tensor, gamma, beta, eps: float, prec: int, *, add: Optional[Act] = None,
              want_planes: bool = True, want_f32: bool = True, m_dev=None) -> Act:
    rows, D = y.shape
    want_f32 = want_f32 or not want_planes
    out = torch.empty_like(y) if want_f32 else None
    hi, lo = _empty_planes(rows, D, prec, y.device) if want_planes else (None, None)
    add_t, add_mod = (None, 0)
    if add is not None:
        add_t, add_mod = act_f32(add), (add.rows if add.bcast_rows else 0)
    el_bytes = 4 + (4 if want_f32 else 0) + (0 if hi is None else (4 if lo is not None else 2))  # (no tensors in the lambda)
    STATS.call('layernorm', 1, nat.lib().lamp_layernorm,
               (y.data_ptr(), nat.ptr(add_t), add_mod, gamma.data_ptr(), beta.data_ptr(), float(eps), rows, D,
                nat.ptr(out), nat.ptr(hi), nat.ptr(lo), nat.ptr(m_dev), nat.stream()),
               nbytes=lambda m: (rows if m is None else min(rows, m)) * D * el_bytes, rows_dev=m_dev)
    return Act(out, hi, lo, rows, D, 0, m_dev)


def mask_args(mask: Optional[torch.Tensor], B: int, Lq: int, Lk: int):
    """Bool/uint8 mask broadcastable to [B, Lq, Lk] -> (keepalive tensor, ptr, stride_b, stride_q, stride_k).
    Expanded (stride-0) views are passed through untouched: the kernel never needs the tiled copy."""
    if mask is None:
        return None, None, 0, 0, 0
    nat.require_cuda(mask)
    if mask.dim() == 2:
        mask = mask.unsqueeze(0)
    if mask.dtype == torch.bool:
        m8 = mask.view(torch.uint8)
    elif mask.dtype == torch.uint8:
        m8 = mask
    else:
        m8 = (mask != 0).view(torch.uint8)
    m8 = m8.expand(B, Lq, Lk)
    sb, sq, sk = m8.stride()
    return m8, m8.data_ptr(), sb, sq, sk


_MASK_BITS: Dict[tuple, tuple] = {}  # packed label masks, keyed by the byte mask's storage; a handful of entries at most


def mask_bits(m8: torch.Tensor, B: int, Lq: int, Lk: int):
    """Bit-packed form of an expanded uint8 mask view [B, Lq, Lk] (``lamp_pack_mask_bits``) -> (words, mbb, mbq).
    A mask shared by the batch (stride 0: the label-graph mask) is packed once and cached against its storage."""
    sb, sq, sk = m8.stride()
    Bm = 1 if sb == 0 else B
    W = (Lk + 31) // 32
    key = (m8.data_ptr(), m8._version, Lq, Lk, sq, sk) if sb == 0 else None
    hit = _MASK_BITS.get(key) if key is not None else None
    if hit is not None:
        return hit[0], 0, W
    words = torch.empty((Bm, Lq, W), dtype=torch.int32, device=m8.device)
    STATS.call('pack_mask_bits', 1, nat.lib().lamp_pack_mask_bits,
               (m8.data_ptr(), sb, sq, sk, Bm, Lq, Lk, words.data_ptr(), nat.stream()), nbytes=Bm * Lq * Lk)
    if key is not None:
        if len(_MASK_BITS) >= 16:
            _MASK_BITS.clear()
        _MASK_BITS[key] = (words, m8)  # the byte mask is kept alive so that its address cannot be recycled
    return words, (0 if sb == 0 else Lq * W), W


def attention(q: Act, q_col0: int, kv: Act, k_col0: int, v_col0: int, B: int, H: int, Lq: int, Lk: int, d: int,
              prec: int, mask: Optional[torch.Tensor], want_probs: bool, out_f32: bool = False, kv_start=None,
              kv_len=None):
    """-> (O as Act [B*Lq, H*d] (planes, or fp32 when out_f32), probs [H*B, Lq, Lk] or None)."""
    dev = q.hi.device
    hd = H * d
    o_hi = o_lo = o32 = None
    if out_f32:
        o32 = torch.empty((B * Lq, hd), dtype=torch.float32, device=dev)
    else:
        o_hi, o_lo = _empty_planes(B * Lq, hd, prec, dev)
    probs = rmax = rsum = None
    if want_probs:
        probs = torch.empty((H * B, Lq, Lk), dtype=torch.float32, device=dev)
        rmax = torch.empty((H * B * Lq,), dtype=torch.float32, device=dev)
        rsum = torch.empty_like(rmax)
    if kv_len is not None:
        # packed keys: `mask` (optional) is one uint8 per packed key row
        keep = None if mask is None else mask.view(torch.uint8) if mask.dtype == torch.bool else mask
        mptr, sb, sq, sk = (None if keep is None else keep.data_ptr()), 0, 0, 1
    else:
        keep, mptr, sb, sq, sk = mask_args(mask, B, Lq, Lk)
    pl = 4 if prec == nat.PREC_FP32 else 2
    # algorithmic bytes of the attention core: Q + K + V read, O written, 4 B (fp32-equivalent
    # plane pair) or 2 B (bf16) per element; a broadcast Q is read once.
    # With packed keys only sum(kv_len) key rows exist: both figures follow the keys actually attended to.
    qo_bytes = (1 if q.bcast_rows else B) * Lq * hd * pl + B * Lq * hd * (4 if out_f32 else pl)

    def keys(m):
        return B * Lk if m is None else m
    name = 'attn_core_self' if q is kv else 'attn_core_enc'
    if keep is not None and kv_len is None and not want_probs and sq != 0 and (Lk > 128 or sb != 0):
        # a real [.., Lq, Lk] mask that the kernel would have to rebuild for every KV tile / sample: hand it over
        # bit-packed (one word per thread and tile; the shared label-graph mask is packed once and cached)
        words, mbb, mbq = mask_bits(keep, B, Lq, Lk)
        STATS.call(name, 1, nat.lib().lamp_attn_core_planes_mbits,
                   (q.hi.data_ptr(), nat.ptr(q.lo), q.cols, q_col0, 1 if q.bcast_rows else 0,
                    kv.hi.data_ptr(), nat.ptr(kv.lo), kv.cols, k_col0, v_col0, B, H, Lq, Lk, d, float(math.sqrt(d)),
                    prec, words.data_ptr(), mbb, mbq, nat.ptr(o_hi), nat.ptr(o_lo), hd, nat.ptr(o32), hd, nat.stream()),
                   flops=4.0 * H * Lq * d * B * Lk, nbytes=qo_bytes + 2 * B * Lk * hd * pl)
        return Act(o32, o_hi, o_lo, B * Lq, hd), None
    STATS.call(name, 2 if want_probs else 1, nat.lib().lamp_attn_core_planes,
               (q.hi.data_ptr(), nat.ptr(q.lo), q.cols, q_col0, 1 if q.bcast_rows else 0,
                kv.hi.data_ptr(), nat.ptr(kv.lo), kv.cols, k_col0, v_col0, B, H, Lq, Lk, d, float(math.sqrt(d)), prec,
                mptr, sb, sq, sk, nat.ptr(o_hi), nat.ptr(o_lo), hd, nat.ptr(o32), hd, nat.ptr(rmax), nat.ptr(rsum),
                nat.ptr(probs), nat.ptr(kv_start), nat.ptr(kv_len), kv.rows if kv_len is not None else 0,
                nat.stream()), flops=lambda m: 4.0 * H * Lq * d * keys(m),
               nbytes=lambda m: qo_bytes + 2 * keys(m) * hd * pl, rows_dev=kv_len)
    del keep
    return Act(o32, o_hi, o_lo, B * Lq, hd), probs


def sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask, temperature: float, prec: int, want_attn=True):
    """ScaledDotProductAttention on head-major fp32 tensors through the level-2 entry point."""
    nat.require_cuda(q, k, v)
    q, k, v = q.contiguous().float(), k.contiguous().float(), v.contiguous().float()
    N, Lq, d = q.shape
    Lk = k.shape[1]
    L = nat.lib()
    out = torch.empty((N, Lq, d), dtype=torch.float32, device=q.device)
    attn = torch.empty((N, Lq, Lk), dtype=torch.float32, device=q.device) if want_attn else None
    ws = torch.empty((max(L.lamp_sdpa_workspace_bytes(N, Lq, Lk, d), 16),), dtype=torch.uint8, device=q.device)
    keep, mptr, sb, sq, sk = mask_args(mask, N, Lq, Lk)
    STATS.call('sdpa_fwd', 5 if want_attn else 4, L.lamp_sdpa_fwd,
               (q.data_ptr(), k.data_ptr(), v.data_ptr(), mptr, sb, sq, sk, out.data_ptr(), nat.ptr(attn), N, Lq, Lk, d,
                float(temperature), prec, ws.data_ptr(), ws.numel(), nat.stream()), flops=4.0 * N * Lq * Lk * d)
    del keep
    return out, attn


def sdpa_backward(q, k, v, out, probs, attn, grad_out, temperature: float, p_drop: float):
    """Backward of the attention core through ``lamp_attn_core_bwd`` -> (dq, dk, dv).  ``probs``: softmax output
    before dropout [N, Lq, Lk]; ``attn``: after dropout (``None`` / same tensor when dropout is off)."""
    nat.require_cuda(q, k, v, out, probs, grad_out)
    q, k, v, out, probs = (t.contiguous().float() for t in (q, k, v, out, probs))
    grad_out = grad_out.contiguous().float()
    N, Lq, d = q.shape
    Lk = k.shape[1]
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    L = nat.lib()
    ws = torch.empty((max(L.lamp_attn_core_bwd_workspace_bytes(N, Lq, Lk, d), 16),), dtype=torch.uint8, device=q.device)
    a = None if attn is None or attn is probs else attn.contiguous().float()
    STATS.call('attn_core_bwd', 10, L.lamp_attn_core_bwd,
               (q.data_ptr(), k.data_ptr(), v.data_ptr(), grad_out.data_ptr(), out.data_ptr(), probs.data_ptr(),
                nat.ptr(a), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), N, Lq, Lk, d, float(temperature), float(p_drop),
                ws.data_ptr(), ws.numel(), nat.stream()), flops=10.0 * N * Lq * Lk * d)
    return dq, dk, dv


# Device-side dropout counter (int64 [1]) added to every attention-dropout seed; set by graphs.GraphedTrainStep so that
# replays of a captured training step draw fresh masks.  None in eager training (host seeds vary per call).
TRAIN_SEED_DEV: Optional[torch.Tensor] = None


def sdpa_train(q, k, v, mask, temperature: float, prec: int, p_drop: float, seed: int):
    """Training forward of the attention core (dropout inside the kernel) -> (out, attn after dropout, probabilities
    before dropout -- the same tensor as attn when p_drop == 0)."""
    nat.require_cuda(q, k, v)
    q, k, v = q.contiguous().float(), k.contiguous().float(), v.contiguous().float()
    N, Lq, d = q.shape
    Lk = k.shape[1]
    L = nat.lib()
    out = torch.empty((N, Lq, d), dtype=torch.float32, device=q.device)
    attn = torch.empty((N, Lq, Lk), dtype=torch.float32, device=q.device)
    pre = torch.empty_like(attn) if p_drop > 0 else None
    ws = torch.empty((max(L.lamp_sdpa_workspace_bytes(N, Lq, Lk, d), 16),), dtype=torch.uint8, device=q.device)
    keep, mptr, sb, sq, sk = mask_args(mask, N, Lq, Lk)
    STATS.call('sdpa_fwd_train', 5, L.lamp_sdpa_fwd_train,
               (q.data_ptr(), k.data_ptr(), v.data_ptr(), mptr, sb, sq, sk, out.data_ptr(), attn.data_ptr(), nat.ptr(pre),
                N, Lq, Lk, d, float(temperature), prec, float(p_drop), int(seed), nat.ptr(TRAIN_SEED_DEV), ws.data_ptr(),
                ws.numel(), nat.stream()), flops=4.0 * N * Lq * Lk * d)
    del keep
    return out, attn, (attn if pre is None else pre)


class SDPAFunction(torch.autograd.Function):
    """Differentiable attention core on the native kernels: forward = ``lamp_sdpa_fwd_train`` (dropout inside the
    kernel; the probabilities are kept, as the reference keeps ``attn``), backward = ``lamp_attn_core_bwd``.  No
    gradient flows through the returned attention map (the reference's layers never use it for the loss)."""

    @staticmethod
    def forward(ctx, q, k, v, mask, temperature, prec, p_drop=0.0, seed=0):
        out, attn, pre = sdpa_train(q, k, v, mask, temperature, prec, p_drop, seed)
        ctx.save_for_backward(q, k, v, out, pre, attn)
        ctx.temperature, ctx.p_drop = temperature, p_drop
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, grad_out, _grad_attn):
        q, k, v, out, pre, attn = ctx.saved_tensors
        dq, dk, dv = sdpa_backward(q, k, v, out, pre, attn if ctx.p_drop > 0 else None, grad_out, ctx.temperature,
                                   ctx.p_drop)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None, None, None, None, None


# ------------------------------------------------------------------ training path: dense stages
NATIVE_TRAINING = os.environ.get('LAMP_NATIVE_TRAIN', '1') != '0'  # Linear / LayerNorm of the training path on the native kernels


class LinearFunction(torch.autograd.Function):
    """``y = x W^T (+ b)`` for the training path (nn.Linear / Conv1d(k=1) of lamp/SubLayers.py:91-93,110,133).
    forward and ``dx = dy W`` run on the tcgen05 GEMM (3-term split-bf16), ``dW = dy^T x`` and ``db`` on
    ``lamp_gemm_tn_acc``.  x: [..., K] fp32, W: [N, K] (a Conv1d weight [N, K, 1] is viewed as such)."""

    @staticmethod
    def forward(ctx, x, W, b, prec):
        w2 = W.reshape(W.shape[0], -1)
        N, K = w2.shape
        x2 = x.reshape(-1, K).contiguous().float()
        M = x2.shape[0]
        a_hi, a_lo = split(x2, prec)
        w_hi, w_lo = split(w2.detach().contiguous().float(), prec)
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        gemm(a_hi, a_lo, K, w_hi, w_lo, K, M, N, K, prec, bias=None if b is None else b.detach().float().contiguous(),
             out_f32=y, ldo=N)
        # the operand planes of x are what the weight-gradient contraction consumes: keep them instead of x
        ctx.save_for_backward(a_hi, a_lo if a_lo is not None else a_hi.new_empty(0), W,
                              b if b is not None else x2.new_empty(0))
        ctx.prec, ctx.has_bias, ctx.x_shape = prec, b is not None, tuple(x.shape)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        a_hi, a_lo, W, b = ctx.saved_tensors
        a_lo = a_lo if a_lo.numel() else None
        prec = ctx.prec
        w2 = W.reshape(W.shape[0], -1)
        N, K = w2.shape
        dy2 = dy.reshape(-1, N).contiguous().float()
        M = dy2.shape[0]
        dx = dW = db = None
        d_hi, d_lo = split(dy2, prec)
        if ctx.needs_input_grad[0]:
            wt_hi, wt_lo = split(w2.detach().t().contiguous().float(), prec)  # [K, N]: dx = dy (W^T)^T
            dx2 = torch.empty((M, K), dtype=torch.float32, device=dy.device)
            gemm(d_hi, d_lo, N, wt_hi, wt_lo, N, M, K, N, prec, out_f32=dx2, ldo=K)
            dx = dx2.view(ctx.x_shape)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dW2 = torch.zeros((N, K), dtype=torch.float32, device=dy.device)
            db = torch.zeros((N,), dtype=torch.float32, device=dy.device) if ctx.has_bias else None
            STATS.call('gemm_tn', 2 if db is not None else 1, nat.lib().lamp_gemm_tn_acc,
                       (d_hi.data_ptr(), nat.ptr(d_lo), N, a_hi.data_ptr(), nat.ptr(a_lo), K, M, N, K, dW2.data_ptr(),
                        nat.ptr(db), nat.stream()), flops=2.0 * M * N * K)
            dW = dW2.view(W.shape).to(W.dtype)
            if db is not None:
                db = db.to(b.dtype)
        return dx, dW, db, None


class LayerNormFunction(torch.autograd.Function):
    """torch.nn.LayerNorm over the last dimension on the native kernels (forward ``lamp_layernorm``, backward
    ``lamp_layernorm_bwd``) -- lamp/SubLayers.py:117,141 in the training path."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, prec):
        D = x.shape[-1]
        x2 = x.reshape(-1, D).contiguous().float()
        out = layernorm(x2, gamma.detach().float().contiguous(), beta.detach().float().contiguous(), eps, prec,
                        want_planes=False).f32
        ctx.save_for_backward(x2, gamma)
        ctx.eps, ctx.shape = eps, tuple(x.shape)
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, gamma = ctx.saved_tensors
        rows, D = x2.shape
        dy2 = dy.reshape(rows, D).contiguous().float()
        dx = torch.empty_like(x2)
        dg = torch.zeros((D,), dtype=torch.float32, device=dy.device)
        db = torch.zeros_like(dg)
        STATS.call('layernorm_bwd', 1, nat.lib().lamp_layernorm_bwd,
                   (x2.data_ptr(), dy2.data_ptr(), gamma.detach().float().contiguous().data_ptr(), float(ctx.eps), rows, D,
                    dx.data_ptr(), dg.data_ptr(), db.data_ptr(), nat.stream()), nbytes=rows * D * 12)
        return dx.view(ctx.shape), dg.to(gamma.dtype), db.to(gamma.dtype), None, None


class DiagProjFunction(torch.autograd.Function):
    """Differentiable diagonal label projection (lamp/Models.py:124-126: the reference builds the [B, L, L] product
    and keeps its diagonal): forward ``lamp_diag_proj``, backward ``lamp_diag_proj_bwd``."""

    @staticmethod
    def forward(ctx, x, W, bias):
        x = x.contiguous().float()
        ctx.save_for_backward(x, W, bias if bias is not None else x.new_empty(0))
        ctx.has_bias = bias is not None
        return diag_proj(x, W.detach().float().contiguous(), None if bias is None else bias.detach().float().contiguous())

    @staticmethod
    def backward(ctx, g):
        x, W, bias = ctx.saved_tensors
        B, L, D = x.shape
        g = g.contiguous().float()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dW = torch.empty((L, D), dtype=torch.float32, device=x.device)
        db = torch.empty((L,), dtype=torch.float32, device=x.device) if ctx.has_bias else None
        STATS.call('diag_proj_bwd', 2, nat.lib().lamp_diag_proj_bwd,
                   (g.data_ptr(), x.data_ptr(), W.detach().float().contiguous().data_ptr(), B, L, D, nat.ptr(dx),
                    dW.data_ptr(), nat.ptr(db), nat.stream()), nbytes=B * L * D * 8)
        return dx, dW.to(W.dtype), (db.to(bias.dtype) if db is not None else None)


# ------------------------------------------------------------------ fused training sub-layers (N4, second pass)
# One autograd Function per reference sub-layer (PositionwiseFeedForward, MultiHeadAttention).  Inside a Function every
# tensor stays in the layout the tensor-core kernels consume (split-bf16 planes, heads as column slices): no head
# split / merge copies, no fp32 -> planes re-splits of activations that a previous kernel already produced as planes, no
# stored dropout masks (counter-hash masks are recomputed in the backward), ReLU / bias / residual in GEMM epilogues.
FUSED_TRAINING = os.environ.get('LAMP_FUSED_TRAIN', '1') != '0'
# training encoder on the packed non-PAD token rows (one host read of the row count per step; dense inside graph captures)
# dropout + residual inside the fc / w_2 GEMM epilogue (0: separate dropout_add kernel, the round-2 first version)
FUSED_DROPOUT_EPILOGUE = os.environ.get('LAMP_FUSED_DROPOUT', '1') != '0'
PACKED_TRAINING = os.environ.get('LAMP_PACKED_TRAIN', '1') != '0'


def planes_of(x: torch.Tensor, prec: int) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Operand planes of an fp32 activation [..., D]: the ones stashed on the tensor by the lamp_b200 sub-layer that
    produced it (still valid: same version), else a split pass."""
    cols = x.shape[-1]
    rows = x.numel() // cols
    stash = getattr(x, '_lamp_planes', None)
    if stash is not None:
        hi, lo, ver, sprec = stash
        if ver == x._version and sprec == prec and hi.shape == (rows, cols):
            return hi, lo
    return split(x.detach().reshape(rows, cols).float(), prec)


class _TrainWeightCache:
    """Operand planes of the projection weights of the training path, W (forward products) and W^T (input-gradient
    products), row blocks of several parameters side by side where one GEMM consumes them together (Wq | Wk | Wv).
    Weights change every optimizer step, so every entry is rebuilt once per step -- all stale entries of all live
    models by ONE ``lamp_split_planes_multi`` launch at the start of the training forward (``refresh_all``), instead
    of a concatenation, a transposed copy and a split per weight and use.  An entry is validated by the parameters'
    identity (weak references), tensor versions and addresses; a miss or a stale entry met outside ``refresh_all`` is
    (re)built on the spot with its own launch.  Buffers persist (CUDA-graph safe: same addresses every step)."""

    class Entry:
        __slots__ = ('refs', 'transpose', 'hi', 'lo', 'stamp', 'rows', 'cols')

    def __init__(self):
        self.entries = {}
        self.lock = threading.Lock()

    @staticmethod
    def _stamp(ws):
        return tuple((w._version, w.data_ptr()) for w in ws)

    def _jobs(self, e, ws):
        jobs, r0 = [], 0
        for w in ws:
            n, k = w.shape[0], w[0].numel()
            j = nat.LampSplitJob()
            j.src, j.rows, j.cols, j.ld, j.transpose = w.data_ptr(), n, k, k, int(e.transpose)
            if e.transpose:      # dst [K, sum N]: this parameter's block is the column slice r0 .. r0 + n
                j.hi, j.lo, j.ldp = e.hi.data_ptr() + 2 * r0, e.lo.data_ptr() + 2 * r0, e.rows
            else:                # dst [sum N, K]: row block r0 .. r0 + n
                j.hi, j.lo, j.ldp = e.hi.data_ptr() + 2 * r0 * k, e.lo.data_ptr() + 2 * r0 * k, k
            jobs.append(j)
            r0 += n
        return jobs

    @staticmethod
    def _launch(jobs):
        if not jobs:
            return
        arr = (nat.LampSplitJob * len(jobs))(*jobs)
        STATS.call('split_planes', (len(jobs) + 55) // 56, nat.lib().lamp_split_planes_multi,
                   (C.addressof(arr), len(jobs), nat.stream()))

    def get(self, ws, transpose: bool):
        key = tuple(id(w) for w in ws) + (bool(transpose),)
        with self.lock:
            e = self.entries.get(key)
            if e is not None and not all(r() is w for r, w in zip(e.refs, ws)):
                e = None                                  # an id was recycled by another tensor
            if e is None:
                e = self.Entry()
                e.refs, e.transpose = tuple(weakref.ref(w) for w in ws), bool(transpose)
                e.rows, e.cols = sum(w.shape[0] for w in ws), ws[0][0].numel()
                shape = (e.cols, e.rows) if transpose else (e.rows, e.cols)
                e.hi = torch.empty(shape, dtype=torch.bfloat16, device=ws[0].device)
                e.lo = torch.empty(shape, dtype=torch.bfloat16, device=ws[0].device)
                e.stamp = None
                self.entries[key] = e
            stamp = self._stamp(ws)
            if e.stamp != stamp:
                self._launch(self._jobs(e, ws))
                # a build recorded into a CUDA graph proves nothing about later replays: the stamp of an entry that was
                # built under capture only becomes valid through refresh_all, which rebuilds everything under capture
                e.stamp = stamp
            return e.hi, e.lo

    def refresh_all(self, device):
        """Rebuild every stale entry whose parameters live on ``device`` with one launch; forget dead entries."""
        # Under stream capture EVERY entry is rebuilt: the recorded launch is what refreshes the planes on each replay
        # (an optimizer step between replays changes the weights without any host-side check running again).
        capturing = torch.cuda.is_current_stream_capturing()
        with self.lock:
            jobs, fresh = [], []
            for key, e in list(self.entries.items()):
                ws = tuple(r() for r in e.refs)
                if any(w is None for w in ws):
                    del self.entries[key]
                    continue
                if ws[0].device != device:
                    continue
                stamp = self._stamp(ws)
                if capturing or e.stamp != stamp:
                    jobs += self._jobs(e, ws)
                    fresh.append((e, stamp))
            self._launch(jobs)
            for e, stamp in fresh:
                e.stamp = stamp


TRAIN_WEIGHTS = _TrainWeightCache()


def _wplanes(W, prec: int, transpose: bool = False):
    """Operand planes of a projection weight -- or of several weights stacked along their output dimension (``W`` a
    tuple: Wq | Wk | Wv) -- for the training path: [sum N, K], or [K, sum N] with ``transpose``.  Leaf fp32 CUDA
    parameters go through :data:`TRAIN_WEIGHTS`; anything else (DataParallel replicas hold non-leaf views that are
    re-created every forward, other dtypes) is concatenated / transposed / split on the spot."""
    ws = W if isinstance(W, tuple) else (W,)
    if prec == nat.PREC_FP32 and all(w.is_leaf and w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
                                     for w in ws):
        return TRAIN_WEIGHTS.get(ws, transpose)
    w2 = ws[0].detach() if len(ws) == 1 else torch.cat([w.detach() for w in ws], dim=0)
    w2 = w2.reshape(w2.shape[0], -1).float()
    return split(w2.t().contiguous() if transpose else w2.contiguous(), prec)


def _seed_dev_ptr():
    return nat.ptr(TRAIN_SEED_DEV)


def dropout_add(y0: torch.Tensor, x: torch.Tensor, p: float, seed: int) -> torch.Tensor:
    """y = dropout(y0) + x on [rows, D] fp32 (counter-hash mask, see lamp_dropout_add)."""
    rows, D = y0.shape
    y = torch.empty_like(y0)
    STATS.call('dropout_add', 1, nat.lib().lamp_dropout_add,
               (y0.data_ptr(), x.data_ptr(), rows, D, 0, float(p), int(seed), _seed_dev_ptr(), y.data_ptr(), nat.stream()),
               nbytes=rows * D * 12)
    return y


def dropout_split(dy: torch.Tensor, p: float, seed: int):
    """planes(dropout-backward(dy)) for [rows, D] fp32 (the mask of ``dropout_add`` recomputed)."""
    rows, D = dy.shape
    hi, lo = _empty_planes(rows, D, nat.PREC_FP32, dy.device)
    STATS.call('dropout_split', 1, nat.lib().lamp_dropout_split,
               (dy.data_ptr(), rows, D, float(p), int(seed), _seed_dev_ptr(), hi.data_ptr(), lo.data_ptr(), nat.stream()),
               nbytes=rows * D * 8)
    return hi, lo


class _ZeroPool:
    """The zero-initialised gradient accumulators of one backward node (dW / db of the split-K weight-gradient
    kernels, dgamma / dbeta of the LayerNorm backward) carved out of ONE ``torch.zeros`` -- one fill launch per node
    instead of one per tensor.  Slices start on 16-byte boundaries."""

    def __init__(self, device, *shapes):
        self._sizes = [(math.prod(sh) + 3) // 4 * 4 for sh in shapes]
        self._buf = torch.zeros((sum(self._sizes),), dtype=torch.float32, device=device)
        self._shapes, self._next, self._off = list(shapes), 0, 0

    def take(self, shape) -> torch.Tensor:
        want = self._shapes[self._next]
        assert tuple(shape) == tuple(want), (shape, want)
        n = math.prod(want)
        out = self._buf[self._off:self._off + n].view(want)
        self._off += self._sizes[self._next]
        self._next += 1
        return out


def _zeros(zp: Optional[_ZeroPool], shape, device) -> torch.Tensor:
    return zp.take(shape) if zp is not None else torch.zeros(shape, dtype=torch.float32, device=device)


def _layernorm_bwd(y: torch.Tensor, g: torch.Tensor, gamma: torch.Tensor, eps: float, zp: Optional[_ZeroPool] = None,
                   drop=None):
    """LayerNorm backward -> (dy, dgamma, dbeta[, (d_hi, d_lo)]).  ``drop = (p, seed)``: the same kernel also writes
    planes(dropout-backward(dy)) (``lamp_layernorm_bwd_drop``), the operand of the sub-layer's dW / input-gradient
    products -- otherwise that is a separate ``dropout_split`` pass."""
    rows, D = y.shape
    dy = torch.empty_like(y)
    dg = _zeros(zp, (D,), y.device)
    db = _zeros(zp, (D,), y.device)
    gam = gamma.detach().float().contiguous()
    if drop is None or not FUSED_DROPOUT_EPILOGUE:
        STATS.call('layernorm_bwd', 1, nat.lib().lamp_layernorm_bwd,
                   (y.data_ptr(), g.data_ptr(), gam.data_ptr(), float(eps), rows, D, dy.data_ptr(), dg.data_ptr(),
                    db.data_ptr(), nat.stream()), nbytes=rows * D * 12)
        if drop is None:
            return dy, dg, db
        return dy, dg, db, dropout_split(dy, drop[0], drop[1])
    hi, lo = _empty_planes(rows, D, nat.PREC_FP32, y.device)
    STATS.call('layernorm_bwd', 1, nat.lib().lamp_layernorm_bwd_drop,
               (y.data_ptr(), g.data_ptr(), gam.data_ptr(), float(eps), rows, D, dy.data_ptr(), dg.data_ptr(), db.data_ptr(),
                float(drop[0]), int(drop[1]), _seed_dev_ptr(), hi.data_ptr(), lo.data_ptr(), nat.stream()),
               nbytes=rows * D * 16)
    return dy, dg, db, (hi, lo)


def _gemm_tn(d_hi, d_lo, N: int, a_hi, a_lo, K: int, M: int, want_bias: bool, zp: Optional[_ZeroPool] = None):
    """dW [N, K] = dy^T x and db [N] = column sums of dy, both from operand planes."""
    dW = _zeros(zp, (N, K), d_hi.device)
    db = _zeros(zp, (N,), d_hi.device) if want_bias else None
    STATS.call('gemm_tn', 2 if want_bias else 1, nat.lib().lamp_gemm_tn_acc,
               (d_hi.data_ptr(), nat.ptr(d_lo), N, a_hi.data_ptr(), nat.ptr(a_lo), K, M, N, K, dW.data_ptr(), nat.ptr(db),
                nat.stream()), flops=2.0 * M * N * K)
    return dW, db


class FFNTrainFunction(torch.autograd.Function):
    """PositionwiseFeedForward of the training path (lamp/SubLayers.py:125-142) as ONE autograd node:
    forward  GEMM(+b1, ReLU -> planes) -> GEMM(+b2 [+x]) -> dropout + residual -> LayerNorm (fp32 + planes);
    backward LayerNorm bwd -> planes of the dropped gradient -> dW2 | dh (ReLU mask) -> dW1 | dx (+ residual grad)."""

    @staticmethod
    def forward(ctx, x, x_hi, x_lo, W1, b1, W2, b2, gamma, beta, eps, p_drop, seed):
        prec = nat.PREC_FP32
        D = x.shape[-1]
        x2 = x.detach().reshape(-1, D).float().contiguous()
        M, dh = x2.shape[0], W1.shape[0]
        if x_hi is None:
            x_hi, x_lo = split(x2, prec)
        w1h, w1l = _wplanes(W1, prec)
        w2h, w2l = _wplanes(W2, prec)
        h_hi, h_lo = _empty_planes(M, dh, prec, x.device)
        gemm(x_hi, x_lo, D, w1h, w1l, D, M, dh, D, prec, bias=b1.detach().float().contiguous(), relu=True, out_hi=h_hi,
             out_lo=h_lo, ldp=dh)
        y = torch.empty((M, D), dtype=torch.float32, device=x.device)
        b2f = b2.detach().float().contiguous()
        if p_drop > 0 and FUSED_DROPOUT_EPILOGUE:
            gemm_drop(h_hi, h_lo, dh, w2h, w2l, dh, M, D, dh, bias=b2f, p_drop=p_drop, seed=seed, residual=x2, ldr=D,
                      out_f32=y, ldo=D)
        elif p_drop > 0:
            y0 = torch.empty_like(y)
            gemm(h_hi, h_lo, dh, w2h, w2l, dh, M, D, dh, prec, bias=b2f, out_f32=y0, ldo=D)
            y = dropout_add(y0, x2, p_drop, seed)
        else:
            gemm(h_hi, h_lo, dh, w2h, w2l, dh, M, D, dh, prec, bias=b2f, residual=x2, ldr=D, out_f32=y, ldo=D)
        out = layernorm(y, gamma.detach().float().contiguous(), beta.detach().float().contiguous(), eps, prec)
        ctx.save_for_backward(x_hi, x_lo, h_hi, h_lo, y, W1, W2, gamma)
        ctx.eps, ctx.p_drop, ctx.seed, ctx.x_shape = eps, p_drop, seed, tuple(x.shape)
        ctx.mark_non_differentiable(out.hi, out.lo)
        ctx.set_materialize_grads(False)   # no zero tensors for the plane outputs' (never defined) gradients
        return out.f32.view(x.shape), out.hi, out.lo

    @staticmethod
    def backward(ctx, g, _ghi, _glo):
        x_hi, x_lo, h_hi, h_lo, y, W1, W2, gamma = ctx.saved_tensors
        prec = nat.PREC_FP32
        M, D = y.shape
        dh = W1.shape[0]
        if g is None:
            g = torch.zeros_like(y)
        zp = _ZeroPool(y.device, (D,), (D,), (D, dh), (D,), (dh, D), (dh,))
        dy, dgamma, dbeta, (d_hi, d_lo) = _layernorm_bwd(y, g.reshape(M, D).float().contiguous(), gamma, ctx.eps, zp,
                                                         drop=(ctx.p_drop, ctx.seed))
        dW2, db2 = _gemm_tn(d_hi, d_lo, D, h_hi, h_lo, dh, M, True, zp)
        w2t_hi, w2t_lo = _wplanes(W2, prec, transpose=True)          # [dh, D]: dh = d W2
        g_hi, g_lo = _empty_planes(M, dh, prec, y.device)
        gemm(d_hi, d_lo, D, w2t_hi, w2t_lo, D, M, dh, D, prec, out_hi=g_hi, out_lo=g_lo, ldp=dh)
        STATS.call('relu_mask', 1, nat.lib().lamp_relu_mask_planes,
                   (g_hi.data_ptr(), g_lo.data_ptr(), h_hi.data_ptr(), M * dh, nat.stream()), nbytes=M * dh * 10)
        dW1, db1 = _gemm_tn(g_hi, g_lo, dh, x_hi, x_lo, D, M, True, zp)
        w1t_hi, w1t_lo = _wplanes(W1, prec, transpose=True)          # [D, dh]: dx = dh W1 (+ the residual branch)
        dx = torch.empty((M, D), dtype=torch.float32, device=y.device)
        gemm(g_hi, g_lo, dh, w1t_hi, w1t_lo, dh, M, D, dh, prec, residual=dy, ldr=D, out_f32=dx, ldo=D)
        return (dx.view(ctx.x_shape), None, None, dW1.view(W1.shape).to(W1.dtype), db1, dW2.view(W2.shape).to(W2.dtype), db2,
                dgamma, dbeta, None, None, None)


class MHATrainFunction(torch.autograd.Function):
    """MultiHeadAttention of the training path (lamp/SubLayers.py:77-121) as ONE autograd node, self-attention
    (``xkv is None``) or label<-input attention.  Q|K|V live in one [rows, 3*H*d] (or Q and K|V in [rows, H*d] /
    [rows, 2*H*d]) plane matrix, heads are column slices for the attention core in both directions."""

    @staticmethod
    def forward(ctx, xq, xq_hi, xq_lo, xkv, xkv_hi, xkv_lo, Wq, Wk, Wv, Wfc, gamma, beta, eps, mask, H, d, temperature,
                p_attn, p_out, seed_attn, seed_out, want_attn, kv_map):
        """``kv_map`` (label<-input attention over a padding-aware encoder output): ``xkv`` is then the PACKED
        [1, n+1, D] encoder activation (non-PAD token rows + one PAD representative), ``kv_map = (src_row [B*T] int64:
        dense position -> packed row, idx [n] int64: packed row -> dense position, T)``.  K|V are projected on the
        packed rows only and scattered into the dense [B*T, 2*H*d] layout the attention core reads."""
        prec = nat.PREC_FP32
        L = nat.lib()
        B, Lq, D = xq.shape
        self_attn = xkv is None
        Lk = Lq if self_attn else (kv_map[2] if kv_map is not None else xkv.shape[1])
        hd = H * d
        dev = xq.device
        x2 = xq.detach().reshape(-1, D).float().contiguous()
        Mq, Mk = B * Lq, B * Lk
        if xq_hi is None:
            xq_hi, xq_lo = split(x2, prec)
        if self_attn:
            w_hi, w_lo = _wplanes((Wq, Wk, Wv), prec)
            qp = _empty_planes(Mq, 3 * hd, prec, dev)
            gemm(xq_hi, xq_lo, D, w_hi, w_lo, D, Mq, 3 * hd, D, prec, out_hi=qp[0], out_lo=qp[1], ldp=3 * hd)
            kvp, ldq, ldkv, k_col0, v_col0 = qp, 3 * hd, 3 * hd, hd, 2 * hd
        else:
            if xkv_hi is None:
                xkv_hi, xkv_lo = split(xkv.detach().reshape(-1, D).float().contiguous(), prec)
            wq_hi, wq_lo = _wplanes(Wq, prec)
            wkv_hi, wkv_lo = _wplanes((Wk, Wv), prec)
            qp = _empty_planes(Mq, hd, prec, dev)
            gemm(xq_hi, xq_lo, D, wq_hi, wq_lo, D, Mq, hd, D, prec, out_hi=qp[0], out_lo=qp[1], ldp=hd)
            if kv_map is None:
                kvp = _empty_planes(Mk, 2 * hd, prec, dev)
                gemm(xkv_hi, xkv_lo, D, wkv_hi, wkv_lo, D, Mk, 2 * hd, D, prec, out_hi=kvp[0], out_lo=kvp[1], ldp=2 * hd)
            else:
                n_rows = xkv_hi.shape[0]
                kpk = _empty_planes(n_rows, 2 * hd, prec, dev)
                gemm(xkv_hi, xkv_lo, D, wkv_hi, wkv_lo, D, n_rows, 2 * hd, D, prec, out_hi=kpk[0], out_lo=kpk[1], ldp=2 * hd)
                kvp = (kpk[0].index_select(0, kv_map[0]), kpk[1].index_select(0, kv_map[0]))   # dense [B*T, 2hd]
                del kpk
            ldq, ldkv, k_col0, v_col0 = hd, 2 * hd, 0, hd
        o_hi, o_lo = _empty_planes(Mq, hd, prec, dev)
        # want_attn False (nobody reads the attention map: the reference's layers only return it): nothing of size
        # Lq x Lk is written by the forward; the backward rebuilds P from the saved row statistics (recompute form)
        attn = torch.empty((H * B, Lq, Lk), dtype=torch.float32, device=dev) if want_attn else None
        pre = torch.empty_like(attn) if (want_attn and p_attn > 0) else None
        stats = torch.empty((2, H * B * Lq), dtype=torch.float32, device=dev)
        keep, mptr, sb, sq, sk = mask_args(mask, B, Lq, Lk)
        if keep is not None and not want_attn and sq != 0 and (Lk > 128 or sb != 0):
            # per-(query, key) mask of a multi-tile problem: packed bits for the forward core (the label-graph mask is
            # packed once and cached); the backward's recompute kernel reads the byte mask
            words, mbb, mbq = mask_bits(keep, B, Lq, Lk)
            STATS.call('attn_core_train', 1, L.lamp_attn_core_planes_train_mbits,
                       (qp[0].data_ptr(), qp[1].data_ptr(), ldq, 0, 0, kvp[0].data_ptr(), kvp[1].data_ptr(), ldkv, k_col0,
                        v_col0, B, H, Lq, Lk, d, float(temperature), prec, words.data_ptr(), mbb, mbq, o_hi.data_ptr(),
                        o_lo.data_ptr(), hd, stats[0].data_ptr(), stats[1].data_ptr(), float(p_attn), int(seed_attn),
                        _seed_dev_ptr(), nat.stream()), flops=4.0 * H * Lq * d * B * Lk)
        else:
            STATS.call('attn_core_train', 2 if want_attn else 1, L.lamp_attn_core_planes_train,
                       (qp[0].data_ptr(), qp[1].data_ptr(), ldq, 0, 0, kvp[0].data_ptr(), kvp[1].data_ptr(), ldkv, k_col0,
                        v_col0, B, H, Lq, Lk, d, float(temperature), prec, mptr, sb, sq, sk, o_hi.data_ptr(), o_lo.data_ptr(),
                        hd, stats[0].data_ptr(), stats[1].data_ptr(), nat.ptr(attn), nat.ptr(pre), float(p_attn),
                        int(seed_attn), _seed_dev_ptr(), nat.stream()), flops=4.0 * H * Lq * d * B * Lk)
        wfc_hi, wfc_lo = _wplanes(Wfc, prec)
        y = torch.empty((Mq, D), dtype=torch.float32, device=dev)
        if p_out > 0 and FUSED_DROPOUT_EPILOGUE:
            gemm_drop(o_hi, o_lo, hd, wfc_hi, wfc_lo, hd, Mq, D, hd, bias=None, p_drop=p_out, seed=seed_out, residual=x2,
                      ldr=D, out_f32=y, ldo=D)
        elif p_out > 0:
            y0 = torch.empty_like(y)
            gemm(o_hi, o_lo, hd, wfc_hi, wfc_lo, hd, Mq, D, hd, prec, out_f32=y0, ldo=D)
            y = dropout_add(y0, x2, p_out, seed_out)
        else:
            gemm(o_hi, o_lo, hd, wfc_hi, wfc_lo, hd, Mq, D, hd, prec, residual=x2, ldr=D, out_f32=y, ldo=D)
        out = layernorm(y, gamma.detach().float().contiguous(), beta.detach().float().contiguous(), eps, prec)
        empty = xq_hi.new_empty(0)
        emptyf = y.new_empty(0)
        ctx.save_for_backward(xq_hi, xq_lo, empty if self_attn else xkv_hi, empty if self_attn else xkv_lo, qp[0], qp[1],
                              empty if self_attn else kvp[0], empty if self_attn else kvp[1], o_hi, o_lo,
                              attn if want_attn else emptyf, (pre if pre is not None else attn) if want_attn else emptyf, y,
                              Wq, Wk, Wv, Wfc, gamma, stats, keep if keep is not None else empty)
        ctx.cfg = (B, Lq, Lk, D, H, d, float(temperature), float(p_attn), float(p_out), int(seed_out), self_attn, eps,
                   None if self_attn else tuple(xkv.shape), bool(want_attn), int(seed_attn), (sb, sq, sk),
                   TRAIN_SEED_DEV, None if kv_map is None else kv_map[1])
        if want_attn:
            ctx.mark_non_differentiable(out.hi, out.lo, attn)
        else:
            attn = emptyf
            ctx.mark_non_differentiable(out.hi, out.lo, attn)
        ctx.set_materialize_grads(False)
        return out.f32.view(B, Lq, D), out.hi, out.lo, attn

    @staticmethod
    def backward(ctx, g, _ghi, _glo, _gattn):
        (xq_hi, xq_lo, xkv_hi, xkv_lo, q_hi, q_lo, kv_hi, kv_lo, o_hi, o_lo, attn, pre, y, Wq, Wk, Wv, Wfc,
         gamma, stats, keep) = ctx.saved_tensors
        (B, Lq, Lk, D, H, d, temperature, p_attn, p_out, seed_out, self_attn, eps, kv_shape, want_attn, seed_attn,
         (sb, sq, sk), seed_dev, kv_idx) = ctx.cfg
        prec = nat.PREC_FP32
        L = nat.lib()
        hd = H * d
        Mq, Mk = B * Lq, B * Lk
        dev = y.device
        if g is None:
            g = torch.zeros_like(y)
        zp = _ZeroPool(dev, (D,), (D,), (D, hd), *(((3 * hd, D),) if self_attn else ((hd, D), (2 * hd, D))))
        dy, dgamma, dbeta, (d_hi, d_lo) = _layernorm_bwd(y, g.reshape(Mq, D).float().contiguous(), gamma, eps, zp,
                                                         drop=(p_out, seed_out))
        dWfc, _ = _gemm_tn(d_hi, d_lo, D, o_hi, o_lo, hd, Mq, False, zp)
        wfct_hi, wfct_lo = _wplanes(Wfc, prec, transpose=True)       # [hd, D]: dO = d Wfc
        do_hi, do_lo = _empty_planes(Mq, hd, prec, dev)
        gemm(d_hi, d_lo, D, wfct_hi, wfct_lo, D, Mq, hd, D, prec, out_hi=do_hi, out_lo=do_lo, ldp=hd)
        if self_attn:
            dq = _empty_planes(Mq, 3 * hd, prec, dev)
            dkv, kvp, lddq, lddkv, ldq, ldkv = dq, (q_hi, q_lo), 3 * hd, 3 * hd, 3 * hd, 3 * hd
            k_col0, v_col0 = hd, 2 * hd
        else:
            dq = _empty_planes(Mq, hd, prec, dev)
            dkv = _empty_planes(Mk, 2 * hd, prec, dev)
            kvp, lddq, lddkv, ldq, ldkv = (kv_hi, kv_lo), hd, 2 * hd, hd, 2 * hd
            k_col0, v_col0 = 0, hd
        ws = torch.empty((max(L.lamp_attn_bwd_planes_workspace_bytes(B, H, Lq, Lk), 16),), dtype=torch.uint8, device=dev)
        STATS.call('attn_core_bwd', 5, L.lamp_attn_bwd_planes,
                   (q_hi.data_ptr(), q_lo.data_ptr(), ldq, 0, kvp[0].data_ptr(), kvp[1].data_ptr(), ldkv, k_col0, v_col0,
                    do_hi.data_ptr(), do_lo.data_ptr(), o_hi.data_ptr(), o_lo.data_ptr(), hd,
                    pre.data_ptr() if want_attn else None, attn.data_ptr() if (want_attn and p_attn > 0) else None,
                    dq[0].data_ptr(), dq[1].data_ptr(), lddq, 0,
                    dkv[0].data_ptr(), dkv[1].data_ptr(), lddkv, k_col0, v_col0, B, H, Lq, Lk, d, temperature, p_attn,
                    stats[0].data_ptr(), stats[1].data_ptr(), keep.data_ptr() if keep.numel() else None, sb, sq, sk,
                    seed_attn, nat.ptr(seed_dev), ws.data_ptr(), ws.numel(), nat.stream()),
                   flops=(10.0 if want_attn else 12.0) * H * B * Lq * Lk * d)
        dxq = torch.empty((Mq, D), dtype=torch.float32, device=dev)
        dxkv = None
        if self_attn:
            dW, _ = _gemm_tn(dq[0], dq[1], 3 * hd, xq_hi, xq_lo, D, Mq, False, zp)
            dWq, dWk, dWv = dW[:hd], dW[hd:2 * hd], dW[2 * hd:]
            wt_hi, wt_lo = _wplanes((Wq, Wk, Wv), prec, transpose=True)  # [D, 3hd]
            gemm(dq[0], dq[1], 3 * hd, wt_hi, wt_lo, 3 * hd, Mq, D, 3 * hd, prec, residual=dy, ldr=D, out_f32=dxq, ldo=D)
        else:
            dWq, _ = _gemm_tn(dq[0], dq[1], hd, xq_hi, xq_lo, D, Mq, False, zp)
            wqt_hi, wqt_lo = _wplanes(Wq, prec, transpose=True)      # [D, hd]
            gemm(dq[0], dq[1], hd, wqt_hi, wqt_lo, hd, Mq, D, hd, prec, residual=dy, ldr=D, out_f32=dxq, ldo=D)
            if kv_idx is not None:
                # packed encoder rows: gather the key/value gradients of the non-PAD positions; the PAD representative
                # (last packed row) gets exactly zero -- PAD keys are masked, their dK / dV are 0
                n_pk = kv_idx.numel()
                packed = []
                for plane in dkv:   # gather straight into the [n + 1, 2hd] operand (no concatenation pass)
                    buf = torch.empty((n_pk + 1, 2 * hd), dtype=plane.dtype, device=dev)
                    torch.index_select(plane, 0, kv_idx, out=buf[:n_pk])
                    buf[n_pk:].zero_()
                    packed.append(buf)
                dkv = tuple(packed)
                Mk = n_pk + 1
            dWkv, _ = _gemm_tn(dkv[0], dkv[1], 2 * hd, xkv_hi, xkv_lo, D, Mk, False, zp)
            dWk, dWv = dWkv[:hd], dWkv[hd:]
            wkvt_hi, wkvt_lo = _wplanes((Wk, Wv), prec, transpose=True)  # [D, 2hd]
            dxkv = torch.empty((Mk, D), dtype=torch.float32, device=dev)
            gemm(dkv[0], dkv[1], 2 * hd, wkvt_hi, wkvt_lo, 2 * hd, Mk, D, 2 * hd, prec, out_f32=dxkv, ldo=D)
            dxkv = dxkv.view(kv_shape)
        return (dxq.view(B, Lq, D), None, None, dxkv, None, None, dWq.to(Wq.dtype), dWk.to(Wk.dtype), dWv.to(Wv.dtype),
                dWfc.to(Wfc.dtype), dgamma, dbeta, None, None, None, None, None, None, None, None, None, None, None)


class EmbedTrainFunction(torch.autograd.Function):
    """Token (+ position) embedding of the training path (lamp/Encoders.py:66,75) as one autograd node: forward =
    ``lamp_embed`` (fp32 + operand planes for the first FFN), backward = ``lamp_embed_bwd`` -- vector reductions into
    dense gradient tables instead of torch's sort-based embedding backward; the ``padding_idx`` rows get no gradient,
    as with ``nn.Embedding``."""

    @staticmethod
    def forward(ctx, seq, pos, word_w, pos_w, pad_word, pad_pos):
        prec = nat.PREC_FP32
        act = embed(seq, pos, word_w.detach(), None if pos_w is None else pos_w.detach(), prec, want_f32=True)
        ctx.save_for_backward(seq, pos if pos_w is not None else seq)
        ctx.cfg = (tuple(word_w.shape), None if pos_w is None else tuple(pos_w.shape), int(pad_word), int(pad_pos))
        ctx.mark_non_differentiable(act.hi, act.lo)
        ctx.set_materialize_grads(False)
        return act.f32, act.hi, act.lo

    @staticmethod
    def backward(ctx, g, _ghi, _glo):
        seq, pos = ctx.saved_tensors
        wshape, pshape, pad_word, pad_pos = ctx.cfg
        want_w, want_p = ctx.needs_input_grad[2], (pshape is not None and ctx.needs_input_grad[3])
        if g is None or not (want_w or want_p):
            return None, None, None, None, None, None
        g = g.float().contiguous()
        rows, D = g.shape
        dword = torch.zeros(wshape, dtype=torch.float32, device=g.device) if want_w else None
        dpos = torch.zeros(pshape, dtype=torch.float32, device=g.device) if want_p else None
        STATS.call('embed_bwd', 1, nat.lib().lamp_embed_bwd,
                   (g.data_ptr(), seq.data_ptr(), pos.data_ptr(), rows, D, pad_word, pad_pos, nat.ptr(dword), nat.ptr(dpos),
                    nat.stream()), nbytes=rows * D * 4 * (1 + int(want_w) + int(want_p)))
        return None, None, dword, dpos, None, None


def embed_train(seq: torch.Tensor, pos: Optional[torch.Tensor], word_emb, pos_emb) -> torch.Tensor:
    """``word_emb(seq) (+ pos_emb(pos))`` for flat int64 id vectors through :class:`EmbedTrainFunction`
    (``word_emb`` / ``pos_emb``: ``nn.Embedding`` modules, fp32 tables); the operand planes of the result are stashed
    on the returned [rows, D] tensor."""
    def pad_of(m):
        return -1 if m is None or m.padding_idx is None else int(m.padding_idx)
    seq = seq.contiguous().long()
    pos = None if pos_emb is None else pos.contiguous().long()
    out, hi, lo = EmbedTrainFunction.apply(seq, pos, word_emb.weight, None if pos_emb is None else pos_emb.weight,
                                           pad_of(word_emb), pad_of(pos_emb))
    out._lamp_planes = (hi, lo, out._version, nat.PREC_FP32)
    return out


def _train_seed() -> int:
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def ffn_train(x: torch.Tensor, mod) -> torch.Tensor:
    """PositionwiseFeedForward.forward in training mode through :class:`FFNTrainFunction`; the operand planes of the
    result are stashed on the returned tensor for the next lamp_b200 sub-layer."""
    p = float(mod.dropout.p) if mod.training else 0.0
    prec = nat.PREC_FP32
    x_hi, x_lo = planes_of(x, prec)
    ln = mod.layer_norm
    out, hi, lo = FFNTrainFunction.apply(x, x_hi, x_lo, mod.w_1.weight, mod.w_1.bias, mod.w_2.weight, mod.w_2.bias,
                                         ln.weight, ln.bias, ln.eps, p, _train_seed() if p > 0 else 0)
    out._lamp_planes = (hi, lo, out._version, prec)
    return out


def mha_train(q: torch.Tensor, kv: Optional[torch.Tensor], mask, mod, want_attn: bool = True):
    """MultiHeadAttention.forward in training mode through :class:`MHATrainFunction` (``kv`` None: self-attention).
    -> (out [B, Lq, D] with stashed planes, attn [H*B, Lq, Lk] after dropout as the reference returns it -- or None
    with ``want_attn=False``, in which case no probability tensor is written at all and the backward recomputes P)."""
    prec = nat.PREC_FP32
    p_attn = float(mod.attention.dropout.p) if mod.training else 0.0
    p_out = float(mod.dropout.p) if mod.training else 0.0
    q_hi, q_lo = planes_of(q, prec)
    kv_map = None
    if kv is not None:
        pk = getattr(kv, '_lamp_train_packed', None)
        if pk is not None and pk['version'] == kv._version and PACKED_TRAINING:
            # padding-aware encoder output (Encoders.GraphEncoder, training): project K|V on the packed rows only
            kv_map = (pk['src_row'], pk['idx'], kv.shape[1])
            kv = pk['packed']
    kv_hi, kv_lo = (None, None) if kv is None else planes_of(kv, prec)
    ln = mod.layer_norm
    out, hi, lo, attn = MHATrainFunction.apply(
        q, q_hi, q_lo, kv, kv_hi, kv_lo, mod.w_qs.weight, mod.w_ks.weight, mod.w_vs.weight, mod.fc.weight, ln.weight,
        ln.bias, ln.eps, mask, mod.n_head, mod.d_k, mod.attention.temperature, p_attn, p_out,
        _train_seed() if p_attn > 0 else 0, _train_seed() if p_out > 0 else 0, bool(want_attn), kv_map)
    out._lamp_planes = (hi, lo, out._version, prec)
    return out, (attn if want_attn else None)


def gold_binary(gold: torch.Tensor, n_labels: int, skip: int = 4) -> torch.Tensor:
    """Device-side ``utils.get_gold_binary`` (utils/utils.py:205-216, called every step from train.py:34 / test.py:47):
    ``gold`` [B, W] int64 label-id rows (ids offset by the 4 special tokens, EOS-terminated, PAD-padded) -> multi-hot
    targets [B, n_labels] fp32 on the device, without the reference's per-row Python loop on the host."""
    nat.require_cuda(gold)
    gold = gold.contiguous().long()
    B, W = gold.shape
    out = torch.empty((B, n_labels), dtype=torch.float32, device=gold.device)
    STATS.call('gold_binary', 1, nat.lib().lamp_gold_binary,
               (gold.data_ptr(), B, W, n_labels, skip, out.data_ptr(), nat.stream()), nbytes=B * (W * 8 + n_labels * 4))
    return out


_BCE_WS: Dict[int, torch.Tensor] = {}  # per device: ticket + block partials, zeroed once (the kernel re-arms it)


class BCEWithLogitsFunction(torch.autograd.Function):
    """``F.binary_cross_entropy_with_logits(logits, target, reduction='mean')`` (train.py:38) with the gradient produced
    in the same pass (``lamp_bce_logits``): one kernel instead of torch's forward + backward element-wise chains."""

    @staticmethod
    def forward(ctx, logits, target):
        nat.require_cuda(logits, target)
        x = logits.contiguous().float()
        y = target.contiguous().float()
        dev = x.device
        ws = _BCE_WS.get(dev.index)
        if ws is None:
            ws = _BCE_WS[dev.index] = torch.zeros(nat.lib().lamp_bce_logits_workspace_bytes() // 4, dtype=torch.float32,
                                                  device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        STATS.call('bce_logits', 1, nat.lib().lamp_bce_logits,
                   (x.data_ptr(), y.data_ptr(), x.numel(), loss.data_ptr(), nat.ptr(dx), ws.data_ptr(), ws.numel() * 4,
                    nat.stream()), nbytes=x.numel() * 12)
        ctx.save_for_backward(dx if dx is not None else x.new_empty(0))
        ctx.shape = tuple(logits.shape)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return (dx * g).view(ctx.shape) if dx.numel() else None, None


def bce_with_logits(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Mean BCE-with-logits on the native kernel (loss + gradient in one pass); same value as torch's."""
    return BCEWithLogitsFunction.apply(logits, target)


_WARNED: set = set()


def warn_torch_fallback(what: str, why: str) -> None:
    """Nothing falls back silently: the first time a stage runs as torch ops (CUDA, eager) instead of the native
    kernels -- a training stage while the native training path is enabled, or an eval call whose shape the kernels do
    not cover -- say so (once per stage / reason).  There is never a CPU fallback."""
    if (what, why) not in _WARNED:
        _WARNED.add((what, why))
        import warnings
        warnings.warn(f'lamp_b200: {what} runs as torch ops, not on the native kernels ({why})', RuntimeWarning, stacklevel=3)


def linear_train(x: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor], prec: Optional[int] = None) -> torch.Tensor:
    """Differentiable ``x W^T (+b)`` of the training path: native when the shape allows, torch otherwise."""
    w2 = W.reshape(W.shape[0], -1)
    N, K = w2.shape
    if NATIVE_TRAINING and x.is_cuda and x.dtype == torch.float32 and N % 8 == 0 and K % 8 == 0 and x.numel() > 0:
        return LinearFunction.apply(x, W, b, default_precision() if prec is None else prec)
    if NATIVE_TRAINING and x.numel() > 0:
        warn_torch_fallback('a linear projection of the training path',
                            f'needs fp32 CUDA input and feature counts that are multiples of 8; got {x.dtype}, N={N}, K={K}')
    return torch.nn.functional.linear(x, w2, b)


def layernorm_train(x: torch.Tensor, ln: torch.nn.LayerNorm, prec: Optional[int] = None) -> torch.Tensor:
    D = x.shape[-1]
    if NATIVE_TRAINING and x.is_cuda and x.dtype == torch.float32 and D % 4 == 0 and D <= 4096 and x.numel() > 0 \
            and ln.elementwise_affine and tuple(ln.normalized_shape) == (D,):
        return LayerNormFunction.apply(x, ln.weight, ln.bias, ln.eps, default_precision() if prec is None else prec)
    if NATIVE_TRAINING and x.numel() > 0:
        warn_torch_fallback('a LayerNorm of the training path', f'needs fp32 CUDA input with D % 4 == 0, D <= 4096; got {x.dtype}, D={D}')
    return ln(x)


def embed(seq: torch.Tensor, pos: Optional[torch.Tensor], word_emb: torch.Tensor, pos_emb: Optional[torch.Tensor],
          prec: int, want_f32: bool = True, row_index=None, m_dev=None) -> Act:
    nat.require_cuda(seq, word_emb)
    seq = seq.contiguous().long()
    rows = seq.numel()
    D = word_emb.shape[1]
    out = torch.empty((rows, D), dtype=torch.float32, device=seq.device) if want_f32 else None
    hi, lo = _empty_planes(rows, D, prec, seq.device)
    if pos_emb is not None:
        pos = pos.contiguous().long()
    el_bytes = 4 + (4 if want_f32 else 0) + (4 if lo is not None else 2)
    STATS.call('embed', 1, nat.lib().lamp_embed,
               (seq.data_ptr(), nat.ptr(pos) if pos_emb is not None else None, word_emb.data_ptr(), nat.ptr(pos_emb),
                rows, D, nat.ptr(out), hi.data_ptr(), nat.ptr(lo), nat.ptr(row_index), nat.ptr(m_dev), nat.stream()),
               nbytes=lambda m: (rows if m is None else min(rows, m)) * D * el_bytes, rows_dev=m_dev)
    return Act(out, hi, lo, rows, D, 0, m_dev)


def gather_rows(src: torch.Tensor, index: torch.Tensor, D: int) -> torch.Tensor:
    """out[r, :] = src[index[r], :] (fp32) -- dense API tensor from a packed activation."""
    rows = index.numel()
    out = torch.empty((rows, D), dtype=torch.float32, device=src.device)
    STATS.call('gather_rows', 1, nat.lib().lamp_gather_rows,
               (src.data_ptr(), index.data_ptr(), rows, D, out.data_ptr(), nat.stream()), nbytes=rows * D * 8)
    return out


def zero_guard_rows(a: Act, nguard: int = 128) -> None:
    """Zero the rows that follow the rows in use of a packed Act (see zero_guard_rows_kernel)."""
    if a.m_dev is None:
        return
    STATS.call('zero_guard_rows', 1, nat.lib().lamp_zero_guard_rows,
               (a.hi.data_ptr(), nat.ptr(a.lo), a.cols, a.cols, a.m_dev.data_ptr(), a.rows, nguard, nat.stream()))


def diag_proj_act(a: Act, B: int, L: int, W: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """Label projection of the decoder output given as an Act ([B*L, D], possibly a deferred LayerNorm)."""
    if a.ln is None:
        return diag_proj(act_f32(a).view(B, L, a.cols), W, bias)
    ln = a.ln
    D = a.cols
    out = torch.empty((B, L), dtype=torch.float32, device=a.hi.device)
    STATS.call('diag_proj', 1, nat.lib().lamp_diag_proj_ln,
               (a.hi.data_ptr(), nat.ptr(a.lo), ln.stats.data_ptr(), ln.nparts, ln.gamma.data_ptr(), ln.beta.data_ptr(),
                float(ln.eps), W.data_ptr(), nat.ptr(bias), B, L, D, out.data_ptr(), nat.stream()),
               flops=2.0 * B * L * D, nbytes=B * L * D * (2 if a.lo is None else 4))
    return out


def diag_proj(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """x [B, L, D], W [L, D] -> logits [B, L]."""
    nat.require_cuda(x, W)
    B, L, D = x.shape
    x = x.contiguous()
    out = torch.empty((B, L), dtype=torch.float32, device=x.device)
    STATS.call('diag_proj', 1, nat.lib().lamp_diag_proj,
               (x.data_ptr(), W.data_ptr(), nat.ptr(bias), B, L, D, out.data_ptr(), nat.stream()),
               flops=2.0 * B * L * D, nbytes=B * L * D * 4)
    return out
