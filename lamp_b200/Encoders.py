"""Drop-in replacement for ``lamp/Encoders.py:GraphEncoder``."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import Constants
from . import _native as nat
from . import ops
from . import utils
from .Layers import EncoderLayer
from .SubLayers import _needs_autograd


class GraphEncoder(nn.Module):
    """lamp/Encoders.py:31-110: word (+ frozen sinusoid position) embeddings, ``n_layers`` x ``EncoderLayer``,
    optional pooling (``enc_transform``).

    Fused path: one gather kernel writes ``emb + pos`` as fp32 and as tensor-core operand planes; each layer is the
    fused FFN (the token self-attention of the reference does not influence ``enc_output`` -- see
    ``Layers.EncoderLayer`` -- and is evaluated only for ``return_attns=True``).  The genomics front-end
    (``onehot=True``) and the per-sample ``adj`` masks of the ``sider`` dataset are outside the label-graph path and
    run on the composed torch path.
    """

    def __init__(self, n_src_vocab, n_max_seq, n_layers=6, n_head=8, d_k=64, d_v=64,
                 d_word_vec=512, d_model=512, d_inner_hid=1024, onehot=False, enc_transform='',
                 dropout=0.1, no_enc_pos_embedding=False):
        super().__init__()
        n_position = n_max_seq + 1
        self.n_max_seq = n_max_seq
        self.d_model = d_model
        self.onehot = onehot
        self.enc_transform = enc_transform
        self.dropout = nn.Dropout(dropout)
        if onehot:
            self.src_word_emb = nn.Embedding(n_src_vocab, n_src_vocab, padding_idx=Constants.PAD)
            with torch.no_grad():
                self.src_word_emb.weight.zero_()
                self.src_word_emb.weight[1:, 1:] = torch.eye(n_src_vocab - 1)
            self.conv1 = nn.Conv1d(9, d_model, 16, stride=1, padding=8)
            self.conv2 = nn.Conv1d(d_model, d_model, 16, stride=1, padding=8)
        else:
            self.src_word_emb = nn.Embedding(n_src_vocab, d_word_vec, padding_idx=Constants.PAD)
        if no_enc_pos_embedding is False:
            self.position_enc = nn.Embedding(n_position, d_word_vec, padding_idx=Constants.PAD)
            self.position_enc.weight.data = utils.position_encoding_init(n_position, d_word_vec)
        self.layer_stack = nn.ModuleList([
            EncoderLayer(d_model, d_inner_hid, n_head, d_k, d_v, dropout=dropout) for _ in range(n_layers)])

    def fused_ok(self, adj) -> bool:
        return (not self.onehot and not adj and len(self.layer_stack) > 0 and self.d_model % 8 == 0
                and all(l.pos_ffn.fused_ok() and l.slf_attn.fused_ok() for l in self.layer_stack))

    def _pool(self, enc_output, src_seq, batch_size):
        t = self.enc_transform
        if t == '':
            return enc_output
        if t == 'max':
            pooled = enc_output.max(dim=1).values  # (the reference's 'max' branch references an undefined name)
        elif t == 'sum':
            pooled = enc_output.sum(1)
        elif t == 'mean':
            pooled = enc_output.sum(1) / ((src_seq > 0).sum(dim=1).float().view(-1, 1))
        elif t == 'flatten':
            pooled = enc_output.reshape(batch_size, -1).float()
        else:
            raise NotImplementedError(t)
        return pooled.view(batch_size, 1, -1)

    def _forward_composed(self, src_seq, adj, src_pos, return_attns):
        enc_input = self.src_word_emb(src_seq)
        if self.onehot:
            x = F.relu(self.dropout(self.conv1(enc_input.transpose(1, 2))))[:, :, 0:-1]
            x = F.max_pool1d(x, 2, 2)
            enc_input = F.relu(self.conv2(x).transpose(1, 2))[:, 0:-1, :]
            enc_input = enc_input + self.position_enc(src_pos[:, 0:enc_input.size(1)])
            src_seq = src_seq[:, 0:enc_input.size(1)]
        elif hasattr(self, 'position_enc'):
            enc_input = enc_input + self.position_enc(src_pos)
        mask = utils.get_attn_padding_mask(src_seq, src_seq)
        if adj:
            mask = mask.clone()
            for i, a in enumerate(adj):
                n = a.size(0)
                mask[i, 0:n, 0:n] = (a == 0)
        attns = []
        out = enc_input
        for layer in self.layer_stack:
            if return_attns or not ops.ELIDE_DEAD_ENCODER_ATTENTION:
                out, attn = layer(out, slf_attn_mask=mask)
            else:
                # lamp/Layers.py:16-18: the layer output is pos_ffn(enc_input); the self-attention result is discarded
                # and its parameters never receive gradients, so it is only evaluated when its map is asked for
                out, attn = layer.pos_ffn(out), None
            attns.append(attn)
        return self._pool(out, src_seq, src_seq.size(0)), attns

    def _train_packed_ok(self, src_seq, adj, return_attns) -> bool:
        return (ops.PACKED_TRAINING and ops.FUSED_TRAINING and ops.NATIVE_TRAINING and ops.ELIDE_DEAD_ENCODER_ATTENTION
                and not return_attns and not adj and not self.onehot and self.enc_transform == '' and src_seq.is_cuda
                and len(self.layer_stack) > 0 and ops.default_precision() == nat.PREC_FP32
                and all(l.pos_ffn.fused_ok() and l.pos_ffn.precision is None for l in self.layer_stack)
                and self.src_word_emb.weight.dtype == torch.float32 and not torch.cuda.is_current_stream_capturing())

    def _forward_train_packed(self, src_seq, src_pos):
        """Training forward on the PACKED non-PAD token rows.  The encoder is a per-row map (embedding, FFN, LayerNorm;
        the token self-attention never reaches the output, lamp/Layers.py:16-18) and the decoder masks PAD keys
        (lamp/Decoders.py:137-138), so PAD rows neither influence the logits nor receive gradients: the FFN stack runs
        on the n non-PAD rows plus ONE PAD representative (so that the returned enc_output still holds the reference's
        values at PAD positions).  One host read (the row count) per step."""
        B, T = src_seq.shape
        seq_flat, pos_flat = src_seq.reshape(-1), src_pos.reshape(-1)
        keep = seq_flat.ne(Constants.PAD)
        idx = torch.nonzero(keep).squeeze(1)                      # [n] packed row -> dense position (host sync)
        n = idx.numel()
        src_row = torch.where(keep, torch.cumsum(keep.to(torch.int64), 0) - 1, torch.full_like(seq_flat, n))
        pad_tok = torch.zeros((1,), dtype=seq_flat.dtype, device=seq_flat.device)  # (PAD, position 0): the representative
        seq_p = torch.cat((seq_flat.index_select(0, idx), pad_tok))
        pos_emb = self.position_enc if hasattr(self, 'position_enc') else None
        pos_p = torch.cat((pos_flat.index_select(0, idx), pad_tok)) if pos_emb is not None else None
        x = ops.embed_train(seq_p, pos_p, self.src_word_emb, pos_emb)   # lamp_embed / lamp_embed_bwd, planes stashed
        out = x.unsqueeze(0)                                      # [1, n + 1, D]
        out._lamp_planes = x._lamp_planes[:2] + (out._version,) + x._lamp_planes[3:]
        for layer in self.layer_stack:
            out = layer.pos_ffn(out)                              # ops.FFNTrainFunction, planes stashed on the result
        dense = out[0].index_select(0, src_row).view(B, T, self.d_model)   # API tensor (PAD rows = the representative)
        dense._lamp_train_packed = dict(packed=out, src_row=src_row, idx=idx, version=dense._version)
        return dense

    def _forward_fused(self, src_seq, src_pos, return_attns):
        B, T = src_seq.shape
        prec = ops.default_precision() if self.layer_stack[0].pos_ffn.precision is None \
            else self.layer_stack[0].pos_ffn.precision
        pos_w = self.position_enc.weight if hasattr(self, 'position_enc') else None
        n = len(self.layer_stack)
        if ops.PADDING_AWARE and not return_attns:
            return self._forward_packed(src_seq, src_pos, pos_w, prec), [None] * n
        x = ops.embed(src_seq, src_pos, self.src_word_emb.weight, pos_w, prec, want_f32=False)
        mask = src_seq.eq(Constants.PAD).unsqueeze(1) if return_attns else None  # [B, 1, T]
        attns = []
        for i, layer in enumerate(self.layer_stack):
            # only the last layer's output is an API tensor (enc_output); in between activations stay in operand form
            x, attn = layer.forward_act(x, B, T, mask, return_attns, want_f32=(i == n - 1))
            attns.append(attn)
        out = x.f32.view(B, T, self.d_model)
        if self.enc_transform == '':
            ops.stash_planes(out, x, prec)
        return self._pool(out, src_seq, B), attns

    def _forward_packed(self, src_seq, src_pos, pos_w, prec):
        """Padding-aware encoder.  Every PAD position holds the same token (id 0) and -- as produced by the
        reference loader, utils/data_loader.py:270 -- the same position id 0, so all those rows of every layer are
        IDENTICAL; the encoder is a per-row map (embedding, FFN, LayerNorm: the token self-attention never reaches
        the output, lamp/Layers.py:16-18).  We therefore run it on the packed list of distinct rows -- all rows that
        are not (PAD, position 0), plus ONE representative of those -- and un-pack into the dense [B, T, D] API tensor
        at the end.  The row count stays on the device (``m_dev``): no host synchronisation.  The packed planes and
        each sample's row range travel with the returned tensor so that GraphDecoder can skip PAD keys altogether."""
        B, T = src_seq.shape
        R = B * T
        seq_flat = src_seq.reshape(-1)
        is_pad = seq_flat.eq(Constants.PAD)
        rep = is_pad & src_pos.reshape(-1).eq(0) if pos_w is not None else is_pad  # rows equal to the representative
        order = torch.argsort(rep.to(torch.uint8), stable=True)         # distinct rows first, in original order
        n_keep = R - rep.sum()                                          # device scalar
        m_dev = torch.clamp(n_keep + 1, max=R).to(torch.int32).reshape(1)  # + the representative (if any)
        rank = torch.cumsum((~rep).to(torch.int64), 0) - 1
        src_row = torch.where(rep, n_keep.to(torch.int64), rank)        # dense row -> packed row
        x = ops.embed(src_seq, src_pos, self.src_word_emb.weight, pos_w, prec, want_f32=False, row_index=order,
                      m_dev=m_dev)
        n = len(self.layer_stack)
        for i, layer in enumerate(self.layer_stack):
            x, _ = layer.forward_act(x, B, T, None, False, want_f32=(i == n - 1) and not ops.DEFER_LAYERNORM)
        unpack = None
        if x.ln is not None:
            # the last LayerNorm is still pending: apply it while un-packing into the dense API tensor; the decoder's
            # K|V projection consumes the deferred packed activation directly
            out = ops.materialize(x, prec, want_f32=True, want_planes=False, index=src_row).f32
        elif ops.DEFER_UNPACK and self.enc_transform == '':
            # forward being captured by graphs.EvalGraphCache: the dense [B, T, D] API tensor is written AFTER each
            # replay, straight into a fresh tensor (no copy out of a static buffer); inside the graph enc_output is
            # only a shape carrier -- the decoder consumes the packed rows
            unpack = (ops.act_f32(x), src_row)
            out = unpack[0].new_empty(1).expand(R, self.d_model)
        else:
            out = ops.gather_rows(ops.act_f32(x), src_row, self.d_model)
        out = out.view(B, T, self.d_model)
        if self.enc_transform == '':
            kv_len = (~rep).view(B, T).sum(dim=1)
            kv_start = torch.cumsum(kv_len, 0) - kv_len
            out._lamp_packed = dict(act=x, kv_start=kv_start.to(torch.int32), kv_len=kv_len.to(torch.int32),
                                    key_is_pad=is_pad[order].to(torch.uint8), version=out._version, prec=prec,
                                    shape=(B, T, self.d_model), unpack=unpack)
        return self._pool(out, src_seq, B)

    def forward(self, src_seq, adj, src_pos, return_attns=False):
        nat.require_cuda(src_seq, src_pos)
        if _needs_autograd(self) and self._train_packed_ok(src_seq, adj, return_attns):
            out, attns = self._forward_train_packed(src_seq, src_pos), [None] * len(self.layer_stack)
        elif _needs_autograd(self) or not self.fused_ok(adj):
            if not _needs_autograd(self):
                ops.warn_torch_fallback('GraphEncoder (eval)', 'the one-hot (genomics) front-end, per-sample adjacency masks '
                                        '(sider) and feature counts that are not multiples of 8 run as torch ops')
            out, attns = self._forward_composed(src_seq, adj, src_pos, return_attns)
        else:
            out, attns = self._forward_fused(src_seq, src_pos, return_attns)
        return (out, attns) if return_attns else (out, None)
