"""Drop-in replacements for ``lamp/Layers.py``: ``EncoderLayer`` and ``DecoderLayer`` (the label-graph layer)."""
import torch
import torch.nn as nn

from . import _native as nat
from . import ops
from .SubLayers import MultiHeadAttention, PositionwiseFeedForward, _needs_autograd


class EncoderLayer(nn.Module):
    """lamp/Layers.py:9-20.  The reference computes token self-attention and then DISCARDS its output
    (``:18`` feeds ``enc_input`` to the FFN), so ``enc_output == pos_ffn(enc_input)``.  The fused path therefore
    runs the attention only when its probabilities are requested (``return_attn=True``, the reference's return
    value); the ``slf_attn`` parameters are kept for state-dict / optimizer compatibility."""

    def __init__(self, d_model, d_inner_hid, n_head, d_k, d_v, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner_hid, dropout=dropout)

    def forward_act(self, x: ops.Act, B: int, T: int, slf_attn_mask, want_attn: bool, want_f32: bool = True):
        attn = None
        if want_attn:
            _, attn = self.slf_attn.forward_act(x, None, B, T, T, slf_attn_mask, True, want_f32=False)
        return self.pos_ffn.forward_act(x, want_f32=want_f32), attn

    def forward(self, enc_input, slf_attn_mask=None, return_attn=True):
        nat.require_cuda(enc_input, slf_attn_mask)
        if _needs_autograd(self, enc_input) or not (self.slf_attn.fused_ok() and self.pos_ffn.fused_ok()):
            _, attn = self.slf_attn(enc_input, enc_input, enc_input, attn_mask=slf_attn_mask)
            return self.pos_ffn(enc_input), attn
        B, T, D = enc_input.shape
        prec = self.pos_ffn.precision if self.pos_ffn.precision is not None else ops.default_precision()
        out, attn = self.forward_act(ops.act_from_tensor(enc_input, prec), B, T, slf_attn_mask, return_attn)
        res = out.f32.view(B, T, D)
        ops.stash_planes(res, out, prec)
        return res, attn


class DecoderLayer(nn.Module):
    """One round of label message passing -- lamp/Layers.py:22-48:
    label<-input attention, FFN, label<-label attention under the label-graph mask, FFN.
    Returns ``(dec_output, dec_output_int, dec_slf_attn, dec_enc_attn)``."""

    def __init__(self, d_model, d_inner_hid, n_head, n_head2, d_k, d_v, dropout=0.1, dropout2=False,
                 no_dec_self_att=False, ffn=True, attn_type='softmax'):
        super().__init__()
        self.enc_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.pos_ffn1 = PositionwiseFeedForward(d_model, d_inner_hid, dropout=dropout)
        if not no_dec_self_att:
            self.slf_attn = MultiHeadAttention(n_head2, d_model, d_k, d_v, dropout=dropout, dropout2=dropout2)
        self.pos_ffn2 = PositionwiseFeedForward(d_model, d_inner_hid, dropout=dropout)

    def fused_ok(self) -> bool:
        ok = self.enc_attn.fused_ok() and self.pos_ffn1.fused_ok() and self.pos_ffn2.fused_ok()
        return ok and (not hasattr(self, 'slf_attn') or self.slf_attn.fused_ok())

    def forward_act(self, x: ops.Act, enc: ops.Act, B: int, L: int, T: int, slf_attn_mask, dec_enc_attn_mask,
                    want_attn: bool, kv_proj=None, last: bool = False, want_int_f32: bool = True,
                    want_out_f32: bool = True, kv_ranges=None, defer_out: bool = False):
        """Inside the layer activations exist as tensor-core operand planes only (with ``ops.DEFER_LAYERNORM`` as
        PRE-LayerNorm planes + row statistics, normalised by their consumers); fp32 copies are written just for the
        tensors the caller asked for (layer output, intermediate output).  ``defer_out``: the caller accepts the layer
        output as a deferred-LayerNorm Act (GraphDecoder -> LAMP's fused label projection)."""
        out, enc_attn = self.enc_attn.forward_act(x, enc, B, L, T, dec_enc_attn_mask, want_attn, kv_proj=kv_proj,
                                                  want_f32=False, kv_ranges=kv_ranges)
        has_slf = hasattr(self, 'slf_attn')
        out = self.pos_ffn1.forward_act(out, want_f32=want_int_f32 and has_slf)
        out_int, slf_attn = None, None
        if has_slf:
            out_int = out
            out, slf_attn = self.slf_attn.forward_act(out, None, B, L, L, slf_attn_mask, want_attn, want_f32=False)
        out = self.pos_ffn2.forward_act(out, want_planes=not last, want_f32=(want_out_f32 or last) and not defer_out)
        return out, out_int, slf_attn, enc_attn

    def forward(self, dec_input, enc_output, slf_attn_mask=None, dec_enc_attn_mask=None, return_attns=True):
        nat.require_cuda(dec_input, enc_output, slf_attn_mask, dec_enc_attn_mask)
        if _needs_autograd(self, dec_input, enc_output) or not self.fused_ok():
            # training: the attention maps are only materialised when the caller wants them (the reference always
            # returns them; GraphDecoder reads them for return_attns only) -- see ops.MHATrainFunction
            ra = return_attns or not self.training
            out, enc_attn = self.enc_attn(dec_input, enc_output, enc_output, attn_mask=dec_enc_attn_mask, return_attn=ra)
            out = self.pos_ffn1(out)
            if hasattr(self, 'slf_attn'):
                out_int = out
                out, slf_attn = self.slf_attn(out, out, out, attn_mask=slf_attn_mask, dec_self=True, return_attn=ra)
            else:
                out_int, slf_attn = None, None
            return self.pos_ffn2(out), out_int, slf_attn, enc_attn
        B, L, D = dec_input.shape
        T = enc_output.shape[1]
        prec = self.enc_attn._prec()
        out, out_int, slf_attn, enc_attn = self.forward_act(
            ops.act_from_tensor(dec_input, prec), ops.act_from_tensor(enc_output, prec), B, L, T, slf_attn_mask,
            dec_enc_attn_mask, return_attns)
        res = out.f32.view(B, L, D)
        ops.stash_planes(res, out, prec)
        return res, (None if out_int is None else out_int.f32.view(B, L, D)), slf_attn, enc_attn
