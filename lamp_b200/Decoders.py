"""Drop-in replacement for ``lamp/Decoders.py:GraphDecoder`` -- the driver of the label-graph message passing."""
import numpy as np
import torch
import torch.nn as nn

from . import Constants
from . import _native as nat
from . import ops
from . import utils
from .Layers import DecoderLayer
from .SubLayers import _needs_autograd


class GraphDecoder(nn.Module):
    """lamp/Decoders.py:96-163.  Label nodes (one embedding per label) are updated by ``n_layers`` rounds of
    label<-input attention and label<-label attention under the label-graph mask.

    Differences in mechanics (not in results) on the fused path:
      * the label mask is ONE ``[L, L]`` device-resident bool buffer read in place by the attention kernel
        (the reference tiles it to ``[B, L, L]`` on the host and copies it to the device every call, ``:141``,
        and again per head, lamp/SubLayers.py:102);
      * the label embeddings are not tiled over the batch (``:132-134``): the first layer's Q projection runs once
        on ``[L, D]`` and is broadcast by the attention kernel;
      * the K|V projections of the encoder output for ALL layers run as one GEMM (the encoder output is read once).
    """

    def __init__(self, n_tgt_vocab, n_max_seq, n_layers=6, n_head=8, n_head2=8, d_k=64, d_v=64,
                 d_word_vec=512, d_model=512, d_inner_hid=1024, dropout=0.1, dropout2=0.1,
                 no_dec_self_att=False, label_adj_matrix=None, label_mask=None,
                 enc_vec=True, graph_conv=False, attn_type='softmax'):
        super().__init__()
        self.enc_vec = enc_vec
        self.n_tgt_vocab = n_tgt_vocab
        self.dropout = nn.Dropout(dropout)
        self.constant_input = torch.arange(n_tgt_vocab).view(-1, 1)
        self.tgt_word_emb = nn.Embedding(n_tgt_vocab, d_word_vec)

        # ---- label-graph mask (1 / True = NO edge = masked), lamp/Decoders.py:108-118
        if label_adj_matrix is not None:
            adj = label_adj_matrix.clone()  # the reference patches the caller's tensor in place; we do not
            empty_rows = adj.sum(dim=1) < 1
            idx = torch.nonzero(empty_rows).flatten()
            adj[idx, idx] = 1  # a label with no edge at all attends to itself (avoids an all -inf row)
            self.label_mask = utils.swap_0_1(adj, 1, 0).unsqueeze(0)
        elif label_mask == 'inveye':
            self.label_mask = 1 - torch.eye(n_tgt_vocab)
        else:
            # 'none' -> fully connected label graph.  (The reference leaves the attribute undefined for unknown
            # strings -- a bare `NotImplementedError` expression at :120 -- and then fails in forward.)
            self.label_mask = None
        if self.label_mask is not None:
            self.register_buffer('_label_mask_dev', self.label_mask.reshape(n_tgt_vocab, n_tgt_vocab) != 0,
                                 persistent=False)
        else:
            self._label_mask_dev = None

        self.layer_stack = nn.ModuleList([
            DecoderLayer(d_model, d_inner_hid, n_head, n_head2, d_k, d_v, dropout=dropout, dropout2=dropout2,
                         no_dec_self_att=no_dec_self_att, attn_type=attn_type)
            for _ in range(n_layers)])
        self._wp = ops.WeightPlanes()

    def fused_ok(self) -> bool:
        return len(self.layer_stack) > 0 and all(l.fused_ok() for l in self.layer_stack)

    # ------------------------------------------------------------------ composed (autograd) path
    def _forward_composed(self, src_seq, enc_output, return_attns, int_preds):
        B = src_seq.size(0)
        dev = enc_output.device
        L = self.n_tgt_vocab
        tgt_seq = torch.arange(L, device=dev).unsqueeze(0).expand(B, L)
        # every sample embeds the same ids 0..L-1 (lamp/Decoders.py:131-134): the table itself, broadcast over the
        # batch -- its gradient is then a sum over the batch dimension instead of a sort-based embedding backward
        dec_input = self.tgt_word_emb.weight[:L].unsqueeze(0).expand(B, L, self.tgt_word_emb.weight.shape[1])
        pad_mask = None
        if not self.enc_vec:
            pad_mask = utils.get_attn_padding_mask(tgt_seq, src_seq[:, 0:enc_output.size(1)])
        slf_mask = None
        if self._label_mask_dev is not None:
            slf_mask = self._label_mask_dev.unsqueeze(0).expand(B, L, L)
        int_outs, slf_attns, enc_attns = [], [], []
        out = dec_input
        for layer in self.layer_stack:
            out, out_int, slf_attn, enc_attn = layer(out, enc_output, slf_attn_mask=slf_mask,
                                                     dec_enc_attn_mask=pad_mask, return_attns=return_attns)
            if int_preds:
                if out_int is not None:
                    int_outs.append(out_int)
                int_outs.append(out)
            if return_attns:
                slf_attns.append(slf_attn)
                enc_attns.append(enc_attn)
        return out, int_outs, slf_attns, enc_attns

    # ------------------------------------------------------------------ fused path
    def _forward_fused(self, src_seq, enc_output, return_attns, int_preds, defer_out=False):
        B, T, D = enc_output.shape
        L = self.n_tgt_vocab
        prec = self.layer_stack[0].enc_attn._prec()
        # Padding-aware keys: when the encoder output comes from lamp_b200's GraphEncoder it carries the PACKED
        # non-PAD token rows (operand planes) and each sample's row range; the K|V projection and the label<-input
        # attention then never touch PAD tokens.  Same result as the key-padding mask of lamp/Decoders.py:137-138.
        packed = getattr(enc_output, '_lamp_packed', None)
        kv_ranges = None
        if (packed is not None and ops.PADDING_AWARE and not return_attns and not self.enc_vec
                and packed['version'] == enc_output._version and packed['prec'] == prec and packed['shape'] == (B, T, D)):
            enc = packed['act']
            kv_ranges = (packed['kv_start'], packed['kv_len'])
            pad_mask = packed['key_is_pad']  # one byte per packed row (all zero unless a PAD token carries a position)
        else:
            enc = ops.act_from_tensor(enc_output, prec)
            pad_mask = None
            if not self.enc_vec:
                pad_mask = src_seq[:, 0:T].eq(Constants.PAD).unsqueeze(1)  # [B, 1, T] -> query stride 0
        emb = self.tgt_word_emb.weight
        e_hi, e_lo = self._wp.get('label_emb', (emb,), prec)
        x = ops.Act(emb.detach(), e_hi, e_lo, L, D, bcast_rows=B * L)  # shared by every sample
        slf_mask = None if self._label_mask_dev is None else self._label_mask_dev.unsqueeze(0)  # [1, L, L]
        # one GEMM for the K|V projections of every layer
        kv_params = []
        for layer in self.layer_stack:
            kv_params += [layer.enc_attn.w_ks.weight, layer.enc_attn.w_vs.weight]
        hd = self.layer_stack[0].enc_attn.n_head * self.layer_stack[0].enc_attn.d_k
        kv_all = ops.project(enc, self._wp, 'kv_all', tuple(kv_params), 2 * hd * len(self.layer_stack), prec,
                             ld_pad=ops.KV_LD_PAD)
        if kv_ranges is not None:
            ops.zero_guard_rows(kv_all)  # rows a KV tile may read past the packed keys must be finite
        int_outs, slf_attns, enc_attns = [], [], []
        n = len(self.layer_stack)
        for i, layer in enumerate(self.layer_stack):
            x, x_int, slf_attn, enc_attn = layer.forward_act(
                x, enc, B, L, T, slf_mask, pad_mask, return_attns, kv_proj=(kv_all, 2 * hd * i, 2 * hd * i + hd),
                last=(i == n - 1), want_int_f32=int_preds, want_out_f32=int_preds or i == n - 1, kv_ranges=kv_ranges,
                defer_out=defer_out and i == n - 1 and not int_preds)
            if int_preds:
                if x_int is not None:
                    int_outs.append(x_int.f32.view(B, L, D))
                int_outs.append(x.f32.view(B, L, D))
            if return_attns:
                slf_attns.append(slf_attn)
                enc_attns.append(enc_attn)
        if x.ln is not None and defer_out:
            return x, int_outs, slf_attns, enc_attns  # deferred-LayerNorm Act [B*L, D] (see LAMP.forward)
        return ops.act_f32(x).view(B, L, D), int_outs, slf_attns, enc_attns

    def forward(self, tgt, src_seq, enc_output, return_attns=False, int_preds=False, _defer_out=False):
        """``tgt`` is unused (as in the reference).  Returns ``(dec_output, None)`` | ``(dec_output, int_outs)``
        | ``(dec_output, dec_slf_attns, dec_enc_attns)`` exactly like lamp/Decoders.py:158-163.
        ``_defer_out`` (internal, used by ``LAMP.forward``): ``dec_output`` may come back as an ``ops.Act`` whose final
        LayerNorm is still pending, to be applied inside the label-projection kernel."""
        nat.require_cuda(src_seq, enc_output)
        if _needs_autograd(self, enc_output) or not self.fused_ok():
            if not _needs_autograd(self, enc_output):
                ops.warn_torch_fallback('GraphDecoder (eval)', 'a layer has a shape the kernels do not cover (see the '
                                        'MultiHeadAttention / PositionwiseFeedForward conditions)')
            out, int_outs, slf_attns, enc_attns = self._forward_composed(src_seq, enc_output, return_attns, int_preds)
        else:
            out, int_outs, slf_attns, enc_attns = self._forward_fused(src_seq, enc_output, return_attns, int_preds,
                                                                      defer_out=_defer_out)
        if int_preds:
            return out, int_outs
        if return_attns:
            return out, slf_attns, enc_attns
        return out, None
