"""Drop-in replacements for ``lamp/SubLayers.py`` -- same class names, constructor signatures, ``forward``
signatures / return tuples and ``state_dict`` keys -- with the arithmetic executed by the sm_100a kernels.

* inference (``torch.no_grad()`` or ``module.eval()``): the fused kernel path
  (``lamp_b200.ops`` -> ``liblamp_b200.so``);
* training (``module.train()`` with grad enabled): the attention core runs on the native kernels in both directions
  (``ops.SDPAFunction``: forward with in-kernel dropout, ``lamp_attn_core_bwd``), and so do the projections, both FFN
  contractions and the LayerNorms (``ops.LinearFunction`` / ``ops.LayerNormFunction``: tcgen05 GEMM for y and dx,
  ``lamp_gemm_tn_acc`` for dW / db, ``lamp_layernorm_bwd``); head split/merge, ReLU, residual adds and the two
  element-wise dropouts are torch glue.

There is no CPU path: CPU tensors raise.  Reference rough edges fixed at the boundary: masks may be
``bool`` or ``uint8``; nothing calls ``.cuda()`` unconditionally.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init

from . import _native as nat
from . import ops


def _needs_autograd(module: nn.Module, *tensors) -> bool:
    """The fused kernels are forward-only.  They serve every call made under ``torch.no_grad()`` or in ``eval()``
    mode (outputs are then not attached to the autograd graph); a module in ``train()`` mode with grad enabled takes
    the differentiable composed path (which also applies dropout)."""
    return torch.is_grad_enabled() and module.training


class XavierLinear(nn.Module):
    """``nn.Linear`` with xavier-normal weights (lamp/SubLayers.py:7-13); keeps the ``.linear.weight`` key."""

    def __init__(self, d_in, d_out, bias=True):
        super().__init__()
        self.linear = nn.Linear(d_in, d_out, bias=bias)
        init.xavier_normal_(self.linear.weight)

    def forward(self, x):
        return self.linear(x)


class ScaledDotProductAttention(nn.Module):
    """softmax(mask(q k^T / temperature)) v  -- lamp/SubLayers.py:16-43.

    q ``[N, Lq, d]``, k/v ``[N, Lk, d]``, attn_mask ``[N, Lq, Lk]`` (True / non-zero = masked).  Returns
    ``(output, attn)``.  ``attn_type='sigmoid'`` (unreachable from the reference's layers) and autograd runs use
    the composed torch path.
    """

    def __init__(self, temperature, dropout=0.1, attn_type='softmax'):
        super().__init__()
        self.temperature = float(temperature)
        self.dropout = nn.Dropout(dropout)
        self.attn_kind = attn_type
        self.precision = None  # None -> ops.default_precision()

    def _composed(self, q, k, v, attn_mask):
        attn = torch.bmm(q, k.transpose(1, 2)) / self.temperature
        if attn_mask is not None:
            attn = attn.masked_fill(attn_mask.bool(), float('-inf'))
        attn = torch.softmax(attn, dim=2) if self.attn_kind == 'softmax' else torch.sigmoid(attn)
        attn = self.dropout(attn)
        return torch.bmm(attn, v), attn

    def _kernel_ok(self, q, k, v) -> bool:
        d = q.shape[-1]
        return (self.attn_kind == 'softmax' and d % 16 == 0 and d <= 128 and k.shape[-1] == d and v.shape[-1] == d)

    def train_core(self, q, k, v, attn_mask):
        """Differentiable attention core for the training path: native forward (dropout on the probabilities inside
        the kernel, seeded from torch's CPU generator) and native backward (``ops.SDPAFunction``); the composed torch
        ops remain for shapes the kernels do not cover and when ``ops.NATIVE_ATTENTION_BACKWARD`` is off."""
        if not (ops.NATIVE_ATTENTION_BACKWARD and self._kernel_ok(q, k, v) and q.dtype == torch.float32):
            if ops.NATIVE_ATTENTION_BACKWARD:
                ops.warn_torch_fallback('the attention core of the training path',
                                        f'needs softmax attention, fp32, head width % 16 == 0 and <= 128; got '
                                        f'{self.attn_kind}, {q.dtype}, d={q.shape[-1]}')
            return self._composed(q, k, v, attn_mask)
        p = float(self.dropout.p) if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
        prec = ops.default_precision() if self.precision is None else self.precision
        return ops.SDPAFunction.apply(q, k, v, attn_mask, self.temperature, prec, p, seed)

    def forward(self, q, k, v, attn_mask=None, stop_sig=False):
        nat.require_cuda(q, k, v, attn_mask)
        if _needs_autograd(self, q, k, v):
            return self.train_core(q, k, v, attn_mask)
        if not self._kernel_ok(q, k, v):
            ops.warn_torch_fallback('ScaledDotProductAttention (eval)',
                                    f'the kernels need softmax attention and a head width that is a multiple of 16 and <= 128; '
                                    f'got {self.attn_kind}, d={q.shape[-1]}')
            return self._composed(q, k, v, attn_mask)
        prec = ops.default_precision() if self.precision is None else self.precision
        return ops.sdpa(q, k, v, attn_mask, self.temperature, prec, want_attn=True)


class MultiHeadAttention(nn.Module):
    """Masked multi-head attention over label nodes -- lamp/SubLayers.py:46-121.

    ``forward(q, k, v, attn_mask=None, dec_self=False) -> (LayerNorm(fc(heads) + q), attn [H*B, Lq, Lk])``.
    The fused path runs: one projection GEMM (Q|K|V concatenated for self-attention, Q and K|V for
    label<-input attention) -> masked softmax attention kernel (mask read in place, never tiled per head) ->
    fc GEMM with the residual added in its epilogue -> LayerNorm that also emits the operand planes of the
    next layer.  ``return_attn=False`` (internal callers) skips the probability output.
    """

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1, dropout2=False, attn_type='softmax'):
        super().__init__()
        self.n_head = n_head
        self.d_k = d_k
        self.d_v = d_v
        self.d_model = d_model
        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(d_model, n_head * d_v, bias=False)
        nn.init.normal_(self.w_qs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_ks.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))
        # lamp/SubLayers.py:61-65: `dropout2` (when truthy) replaces the attention-probability dropout rate
        self.attention = ScaledDotProductAttention(temperature=np.power(d_k, 0.5), attn_type=attn_type,
                                                   dropout=dropout2 if dropout2 else dropout)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(d_model)
        if n_head > 1:
            self.fc = nn.Linear(n_head * d_v, d_model, bias=False)
            nn.init.xavier_normal_(self.fc.weight)
        self.attn_kind = attn_type
        self.precision = None
        self._wp = ops.WeightPlanes()

    # ------------------------------------------------------------------ composed (autograd) path
    def _fused_train_ok(self, q, k, v) -> bool:
        return (ops.FUSED_TRAINING and ops.NATIVE_TRAINING and ops.NATIVE_ATTENTION_BACKWARD and k is v and self.fused_ok()
                and self.n_head > 1 and q.is_cuda and q.dtype == torch.float32 and k.dtype == torch.float32
                and q.dim() == 3 and self._prec() == nat.PREC_FP32 and q.numel() > 0 and k.numel() > 0)

    def _composed(self, q, k, v, attn_mask, want_attn=True):
        if self._fused_train_ok(q, k, v):
            # one autograd node for the whole sub-layer, everything in the tensor-core operand layouts (ops.MHATrainFunction)
            return ops.mha_train(q, None if k is q else k, attn_mask, self, want_attn=want_attn)
        d_k, d_v, n_head = self.d_k, self.d_v, self.n_head
        sz_b, len_q, _ = q.size()
        len_k, len_v = k.size(1), v.size(1)
        residual = q
        lin = ops.linear_train  # native GEMMs (forward, dx, dW) when the shape allows, torch otherwise
        if q is k and k is v:
            # self-attention: one [3*H*d, D] contraction for Q | K | V (one GEMM forward, one dx and one dW backward)
            qkv = lin(q, torch.cat((self.w_qs.weight, self.w_ks.weight, self.w_vs.weight), dim=0), None)
            qp, kp, vp = qkv.split((n_head * d_k, n_head * d_k, n_head * d_v), dim=-1)
        else:
            qp = lin(q, self.w_qs.weight, None)
            if k is v:
                kv = lin(k, torch.cat((self.w_ks.weight, self.w_vs.weight), dim=0), None)
                kp, vp = kv.split((n_head * d_k, n_head * d_v), dim=-1)
            else:
                kp, vp = lin(k, self.w_ks.weight, None), lin(v, self.w_vs.weight, None)
        qh = qp.reshape(sz_b, len_q, n_head, d_k).permute(2, 0, 1, 3).reshape(-1, len_q, d_k)
        kh = kp.reshape(sz_b, len_k, n_head, d_k).permute(2, 0, 1, 3).reshape(-1, len_k, d_k)
        vh = vp.reshape(sz_b, len_v, n_head, d_v).permute(2, 0, 1, 3).reshape(-1, len_v, d_v)
        if attn_mask is not None:
            attn_mask = attn_mask.bool().repeat(n_head, 1, 1)
        out, attn = self.attention.train_core(qh, kh, vh, attn_mask)  # native forward + backward of the core
        out = out.view(n_head, sz_b, len_q, d_v).permute(1, 2, 0, 3).reshape(sz_b, len_q, -1)
        if hasattr(self, 'fc'):
            out = lin(out, self.fc.weight, None)
        out = self.dropout(out)
        return ops.layernorm_train(out + residual, self.layer_norm), attn

    # ------------------------------------------------------------------ fused path on Act objects
    def fused_ok(self) -> bool:
        d = self.d_k
        return (self.attn_kind == 'softmax' and self.d_k == self.d_v and d % 16 == 0 and d <= 128
                and self.d_model % 8 == 0 and (self.n_head > 1 or d == self.d_model))

    def _prec(self) -> int:
        return ops.default_precision() if self.precision is None else self.precision

    def forward_act(self, q: ops.Act, kv, B: int, Lq: int, Lk: int, attn_mask, want_attn: bool,
                    kv_proj=None, want_f32: bool = True, kv_ranges=None):
        """q: Act [B*Lq (or Lq, broadcast), D]; kv: None (self-attention) or Act [B*Lk, D].
        ``kv_proj``: optional precomputed ``(Act [B*Lk, ld], k_col0, v_col0)`` K|V projection (GraphDecoder batches
        the label<-input K|V projections of all its layers into one GEMM).  -> (Act out, probs or None)."""
        prec = self._prec()
        H, d, D = self.n_head, self.d_k, self.d_model
        hd = H * d
        if kv is None and kv_proj is None:
            qkv = ops.project(q, self._wp, 'qkv', (self.w_qs.weight, self.w_ks.weight, self.w_vs.weight), 3 * hd, prec)
            qp, q_col0, kvp, k_col0, v_col0 = qkv, 0, qkv, hd, 2 * hd
        else:
            qp, q_col0 = ops.project(q, self._wp, 'q', (self.w_qs.weight,), hd, prec), 0
            if kv_proj is not None:
                kvp, k_col0, v_col0 = kv_proj
            else:
                kvp = ops.project(kv, self._wp, 'kv', (self.w_ks.weight, self.w_vs.weight), 2 * hd, prec)
                k_col0, v_col0 = 0, hd
        kv_start, kv_len = kv_ranges if kv_ranges is not None else (None, None)  # packed (padding-aware) keys
        o, probs = ops.attention(qp, q_col0, kvp, k_col0, v_col0, B, H, Lq, Lk, d, prec, attn_mask, want_attn,
                                 out_f32=(H == 1), kv_start=kv_start, kv_len=kv_len)
        ln = self.layer_norm
        if H > 1:
            f_hi, f_lo = self._wp.get('fc', (self.fc.weight,), prec)
            out = ops.linear_residual_ln(o, f_hi, f_lo, D, prec, q, ln.weight, ln.bias, ln.eps, want_f32=want_f32)
        else:
            out = ops.layernorm(o.f32, ln.weight, ln.bias, ln.eps, prec, add=q, want_f32=want_f32)
        return out, probs

    def forward(self, q, k, v, attn_mask=None, dec_self=False, return_attn=True):
        nat.require_cuda(q, k, v, attn_mask)
        if _needs_autograd(self, q, k, v):
            return self._composed(q, k, v, attn_mask, want_attn=return_attn)
        if not self.fused_ok():
            ops.warn_torch_fallback('MultiHeadAttention (eval)',
                                    f'the kernels need softmax attention, d_k == d_v, d_k % 16 == 0, d_k <= 128 and d_model % 8 == 0; '
                                    f'got {self.attn_kind}, d_k={self.d_k}, d_v={self.d_v}, d_model={self.d_model}')
            return self._composed(q, k, v, attn_mask)
        B, Lq, D = q.shape
        Lk = k.shape[1]
        prec = self._prec()
        qa = ops.act_from_tensor(q, prec)
        if k is v:
            kva = None if (k is q) else ops.act_from_tensor(k, prec)
            out, probs = self.forward_act(qa, kva, B, Lq, Lk, attn_mask, return_attn)
        else:
            # distinct key / value sources (never produced by the reference's layers): project each into one [K|V] buffer
            hd = self.n_head * self.d_k
            ka, va = ops.act_from_tensor(k, prec), ops.act_from_tensor(v, prec)
            hi, lo = ops._empty_planes(B * Lk, 2 * hd, prec, q.device)
            for src, w, off in ((ka, self.w_ks.weight, 0), (va, self.w_vs.weight, hd)):
                w_hi, w_lo = self._wp.get('k' if off == 0 else 'v', (w,), prec)
                ops.gemm(src.hi, src.lo, D, w_hi, w_lo, D, B * Lk, hd, D, prec, out_hi=hi[:, off:],
                         out_lo=None if lo is None else lo[:, off:], ldp=2 * hd)
            out, probs = self.forward_act(qa, None, B, Lq, Lk, attn_mask, return_attn,
                                          kv_proj=(ops.Act(None, hi, lo, B * Lk, 2 * hd), 0, hd))
        res = out.f32.view(B, Lq, D)
        ops.stash_planes(res, out, prec)
        return res, probs


class PositionwiseFeedForward(nn.Module):
    """Conv1d(k=1) -> ReLU -> Conv1d(k=1) -> dropout -> +residual -> LayerNorm  -- lamp/SubLayers.py:125-142.
    Fused path: GEMM(+bias, ReLU, planes out) -> GEMM(+bias, +residual) -> LayerNorm(+planes)."""

    def __init__(self, d_in, d_hid, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, 1)
        self.w_2 = nn.Conv1d(d_hid, d_in, 1)
        self.layer_norm = nn.LayerNorm(d_in)
        self.dropout = nn.Dropout(dropout)
        self.precision = None
        self._wp = ops.WeightPlanes()

    def _composed(self, x):
        prec = ops.default_precision() if self.precision is None else self.precision
        if (ops.FUSED_TRAINING and ops.NATIVE_TRAINING and self.fused_ok() and x.is_cuda and x.dtype == torch.float32
                and x.numel() > 0 and prec == nat.PREC_FP32):
            return ops.ffn_train(x, self)   # one autograd node for the whole sub-layer (ops.FFNTrainFunction)
        # Conv1d(k=1) == per-position linear map; F.linear keeps the training path in true fp32 (cuDNN convolutions
        # default to TF32, which would put 1e-3-level noise into the gradients)
        h = F.relu(ops.linear_train(x, self.w_1.weight, self.w_1.bias))
        out = ops.linear_train(h, self.w_2.weight, self.w_2.bias)
        return ops.layernorm_train(self.dropout(out) + x, self.layer_norm)

    def fused_ok(self) -> bool:
        return self.w_1.in_channels % 8 == 0 and self.w_1.out_channels % 8 == 0

    def forward_act(self, x: ops.Act, want_planes: bool = True, want_f32: bool = True) -> ops.Act:
        prec = ops.default_precision() if self.precision is None else self.precision
        D, dh = self.w_1.in_channels, self.w_1.out_channels
        w2_hi, w2_lo = self._wp.get('w2', (self.w_2.weight,), prec)
        h = ops.project(x, self._wp, 'w1', (self.w_1.weight,), dh, prec, bias=self.w_1.bias, relu=True)
        ln = self.layer_norm
        return ops.linear_residual_ln(h, w2_hi, w2_lo, D, prec, x, ln.weight, ln.bias, ln.eps, bias=self.w_2.bias,
                                      want_planes=want_planes, want_f32=want_f32)

    def forward(self, x):
        nat.require_cuda(x)
        if _needs_autograd(self, x):
            return self._composed(x)
        if not self.fused_ok():
            ops.warn_torch_fallback('PositionwiseFeedForward (eval)', 'the kernels need d_in and d_hid to be multiples of 8')
            return self._composed(x)
        prec = ops.default_precision() if self.precision is None else self.precision
        out = self.forward_act(ops.act_from_tensor(x, prec))
        res = out.f32.view(x.shape)
        ops.stash_planes(res, out, prec)
        return res
