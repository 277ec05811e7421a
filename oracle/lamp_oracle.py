"""CPU oracle for the LaMP label-graph attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lamp_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do, and there only as the checker / the timed CPU baseline -- never as the product path.

This is a *functional restatement* (plain torch CPU ops, no ``nn.Module``, no reference import) of
the reference algorithm.  All arithmetic in the reference is ATen floating point, so the
restatement is written against ``torch`` (fp32 by default, fp64 on request for arbitration) rather
than numpy/C; every function cites the reference ``file:line`` it follows (paths relative to
``/root/reference``).

Parity pinning: the reference ships no golden vectors or tests, so the
oracle is pinned against *outputs of the reference itself*: ``tests/golden/make_golden.py`` imports
the unmodified reference classes in the build container, runs them on seeded inputs and commits
the results under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures (bit-exact on the generating machine, <=2e-6 elsewhere because MKL kernels
may differ between hosts).

Parameters are passed as a flat mapping ``name -> tensor`` using the reference's own
``state_dict`` key names, so a reference checkpoint can be fed directly.
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

PAD = 0  # lamp/Constants.py:1

Params = Mapping[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# mask producers (row A6)
# --------------------------------------------------------------------------------------
def position_encoding_init(n_position: int, d_pos_vec: int) -> torch.Tensor:
    """Sinusoid table, row 0 = zeros.  lamp/utils.py:9-19."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_pos_vec, dtype=np.float64)[None, :]
    table = pos / np.power(10000.0, 2.0 * np.floor(j / 2.0) / d_pos_vec)
    table[0, :] = 0.0
    table[1:, 0::2] = np.sin(table[1:, 0::2])
    table[1:, 1::2] = np.cos(table[1:, 1::2])
    return torch.from_numpy(table).type(torch.FloatTensor)


def padding_mask(seq_q: torch.Tensor, seq_k: torch.Tensor) -> torch.Tensor:
    """``[B, Lq, Lk]`` bool, True where the KEY token is PAD.  lamp/utils.py:26-34."""
    assert seq_q.dim() == 2 and seq_k.dim() == 2
    b, len_q = seq_q.shape
    _, len_k = seq_k.shape
    return seq_k.eq(PAD).unsqueeze(1).expand(b, len_q, len_k)


def prior_adjacency(train_tgt: Sequence[Sequence[int]], n_tgt_dict: int) -> torch.Tensor:
    """Label co-occurrence adjacency from training label sets.  utils/data_loader.py:37-47.

    ``train_tgt`` rows are ``[BOS, l1, ..., lk, EOS]`` with label ids offset by the 4 special
    tokens; the matrix is ``eye(L)`` plus a symmetric 1 for every pair of distinct labels that
    co-occur in a row (first and last element of the row are skipped).
    """
    n = n_tgt_dict - 4
    adj = torch.eye(n)
    for sample in train_tgt:
        inner = sample[1:-1]
        for i, idx1 in enumerate(inner):
            for idx2 in sample[i + 1:-1]:
                if idx1 != idx2:
                    adj[idx1 - 4, idx2 - 4] = 1
                    adj[idx2 - 4, idx1 - 4] = 1
    return adj


def label_mask_from(n_labels: int, label_adj_matrix: Optional[torch.Tensor] = None,
                    label_mask: Optional[str] = None) -> Optional[torch.Tensor]:
    """``[L, L]`` bool mask, True = masked (NO edge).  lamp/Decoders.py:108-118 + lamp/utils.py:46-50.

    prior : rows of the adjacency that are entirely empty get a forced self edge (``:109-112``),
            then 0/non-0 are swapped (``:113``).
    inveye: ``1 - eye`` (``:115-116``), i.e. every label attends only to itself.
    none  : no mask (``:117-118``).
    """
    if label_adj_matrix is not None:
        adj = label_adj_matrix.clone()
        for i in range(adj.size(0)):
            if adj[i].sum().item() < 1:
                adj[i, i] = 1
        return adj == 0
    if label_mask == 'inveye':
        return (1 - torch.eye(n_labels)) != 0
    if label_mask == 'none' or label_mask is None:
        return None
    raise NotImplementedError(label_mask)


# --------------------------------------------------------------------------------------
# A1: ScaledDotProductAttention.forward  (lamp/SubLayers.py:27-43)
# --------------------------------------------------------------------------------------
def sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, attn_mask: Optional[torch.Tensor],
         temperature: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """q ``[N, Lq, d]``, k/v ``[N, Lk, d]``, mask ``[N, Lq, Lk]`` bool (True = masked).

    ``:28`` bmm, ``:29`` divide by temperature, ``:32`` masked_fill(-inf), ``:39`` softmax(dim=2),
    ``:40`` dropout (identity in eval), ``:41`` bmm.  Returns ``(output, attn)``.
    """
    attn = torch.bmm(q, k.transpose(1, 2))
    attn = attn / temperature
    if attn_mask is not None:
        attn = attn.masked_fill(attn_mask, -np.inf)
    attn = torch.softmax(attn, dim=2)
    output = torch.bmm(attn, v)
    return output, attn


# --------------------------------------------------------------------------------------
# A2: MultiHeadAttention.forward  (lamp/SubLayers.py:77-121)
# --------------------------------------------------------------------------------------
def mha(p: Params, prefix: str, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
        attn_mask: Optional[torch.Tensor], n_head: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode MHA.  Returns ``(out [B,Lq,D], attn [H*B,Lq,Lk])`` (head-major batch index h*B+b).

    ``:85`` residual = un-projected q; ``:91-93`` bias-free projections; ``:96-98`` head split
    (head-major); ``:102`` mask repeated per head; ``:104`` sdpa with temperature sqrt(d_k)
    (``:63/:65``, ``np.power(d_k, 0.5)``); ``:106-107`` head merge; ``:109-110`` bias-free ``fc`` only
    if n_head > 1 (``:72-74``); ``:117/119`` LayerNorm(out + residual), eps 1e-5.
    """
    w_q, w_k, w_v = p[prefix + 'w_qs.weight'], p[prefix + 'w_ks.weight'], p[prefix + 'w_vs.weight']
    d_k = w_q.shape[0] // n_head
    d_v = w_v.shape[0] // n_head
    sz_b, len_q, d_model = q.shape
    len_k = k.shape[1]
    residual = q
    qh = F.linear(q, w_q).view(sz_b, len_q, n_head, d_k)
    kh = F.linear(k, w_k).view(sz_b, len_k, n_head, d_k)
    vh = F.linear(v, w_v).view(sz_b, len_k, n_head, d_v)
    qh = qh.permute(2, 0, 1, 3).contiguous().view(-1, len_q, d_k)
    kh = kh.permute(2, 0, 1, 3).contiguous().view(-1, len_k, d_k)
    vh = vh.permute(2, 0, 1, 3).contiguous().view(-1, len_k, d_v)
    if attn_mask is not None:
        attn_mask = attn_mask.repeat(n_head, 1, 1)
    out, attn = sdpa(qh, kh, vh, attn_mask, float(np.power(d_k, 0.5)))
    out = out.view(n_head, sz_b, len_q, d_v).permute(1, 2, 0, 3).contiguous().view(sz_b, len_q, -1)
    if n_head > 1:
        out = F.linear(out, p[prefix + 'fc.weight'])
    out = F.layer_norm(out + residual, (d_model,), p[prefix + 'layer_norm.weight'],
                       p[prefix + 'layer_norm.bias'], 1e-5)
    return out, attn


# --------------------------------------------------------------------------------------
# PositionwiseFeedForward.forward  (lamp/SubLayers.py:125-142), eval mode
# --------------------------------------------------------------------------------------
def ffn(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Conv1d(k=1) -> ReLU -> Conv1d(k=1) -> +residual -> LayerNorm.  Weights are ``[out, in, 1]``."""
    w1, b1 = p[prefix + 'w_1.weight'], p[prefix + 'w_1.bias']
    w2, b2 = p[prefix + 'w_2.weight'], p[prefix + 'w_2.bias']
    out = x.transpose(1, 2)
    out = F.conv1d(F.relu(F.conv1d(out, w1, b1)), w2, b2)
    out = out.transpose(1, 2)
    return F.layer_norm(out + x, (x.shape[-1],), p[prefix + 'layer_norm.weight'],
                        p[prefix + 'layer_norm.bias'], 1e-5)


# --------------------------------------------------------------------------------------
# A3: DecoderLayer.forward  (lamp/Layers.py:34-48)
# --------------------------------------------------------------------------------------
def decoder_layer(p: Params, prefix: str, dec_input: torch.Tensor, enc_output: torch.Tensor,
                  slf_attn_mask: Optional[torch.Tensor], dec_enc_attn_mask: Optional[torch.Tensor],
                  n_head: int, n_head2: int, no_dec_self_att: bool = False):
    """label<-input MHA, FFN, label<-label MHA under the label mask, FFN.

    Returns ``(dec_output, dec_output_int, dec_slf_attn, dec_enc_attn)`` like ``:48``.
    """
    out, enc_attn = mha(p, prefix + 'enc_attn.', dec_input, enc_output, enc_output, dec_enc_attn_mask, n_head)
    out = ffn(p, prefix + 'pos_ffn1.', out)
    if not no_dec_self_att:
        out_int = out
        out, slf_attn = mha(p, prefix + 'slf_attn.', out, out, out, slf_attn_mask, n_head2)
    else:
        out_int, slf_attn = None, None
    out = ffn(p, prefix + 'pos_ffn2.', out)
    return out, out_int, slf_attn, enc_attn


# --------------------------------------------------------------------------------------
# A5: EncoderLayer.forward  (lamp/Layers.py:15-20) -- note the attention result is discarded
# --------------------------------------------------------------------------------------
def encoder_layer(p: Params, prefix: str, enc_input: torch.Tensor, slf_attn_mask: Optional[torch.Tensor],
                  n_head: int, compute_dead_attention: bool = True):
    """``:16`` computes self-attention, ``:18`` overwrites its output with ``pos_ffn(enc_input)``.

    ``compute_dead_attention=True`` reproduces the reference's work (used when timing the CPU
    baseline and when the attention probabilities are requested); the returned ``enc_output`` does
    not depend on it.
    """
    attn = None
    if compute_dead_attention:
        _, attn = mha(p, prefix + 'slf_attn.', enc_input, enc_input, enc_input, slf_attn_mask, n_head)
    return ffn(p, prefix + 'pos_ffn.', enc_input), attn


# --------------------------------------------------------------------------------------
# A5: GraphEncoder.forward  (lamp/Encoders.py:64-110), text path (no onehot, no per-sample adj)
# --------------------------------------------------------------------------------------
def graph_encoder(p: Params, prefix: str, src_seq: torch.Tensor, src_pos: torch.Tensor, n_layers: int,
                  n_head: int, return_attns: bool = False, compute_dead_attention: bool = True,
                  enc_transform: str = ''):
    """``:66`` word embedding (PAD row is whatever the table holds), ``:75`` += frozen sinusoid
    positions when the table exists, ``:82`` key-padding mask, ``:91-92`` layer loop,
    ``:96-105`` optional pooling (``sum`` / ``mean`` / ``flatten``)."""
    enc_input = F.embedding(src_seq, p[prefix + 'src_word_emb.weight'])
    if prefix + 'position_enc.weight' in p:
        enc_input = enc_input + F.embedding(src_pos, p[prefix + 'position_enc.weight'])
    mask = padding_mask(src_seq, src_seq)
    attns = []
    enc_output = enc_input
    for i in range(n_layers):
        enc_output, attn = encoder_layer(p, f'{prefix}layer_stack.{i}.', enc_output, mask, n_head,
                                         compute_dead_attention or return_attns)
        attns.append(attn)
    if enc_transform != '':
        b = src_seq.shape[0]
        if enc_transform == 'sum':
            enc_output = enc_output.sum(1)
        elif enc_transform == 'mean':
            enc_output = enc_output.sum(1) / ((src_seq > 0).sum(dim=1).float().view(-1, 1))
        elif enc_transform == 'flatten':
            enc_output = enc_output.reshape(b, -1).float()
        else:
            raise NotImplementedError(enc_transform)  # 'max' references an undefined name at :98
        enc_output = enc_output.view(b, 1, -1)
    return (enc_output, attns) if return_attns else (enc_output, None)


# --------------------------------------------------------------------------------------
# A4: GraphDecoder.forward  (lamp/Decoders.py:127-163)
# --------------------------------------------------------------------------------------
def graph_decoder(p: Params, prefix: str, src_seq: torch.Tensor, enc_output: torch.Tensor,
                  label_mask: Optional[torch.Tensor], n_layers: int, n_head: int, n_head2: int,
                  enc_vec: bool = False, no_dec_self_att: bool = False,
                  return_attns: bool = False, int_preds: bool = False):
    """``:132-134`` label embeddings of arange(L) tiled over the batch; ``:137-138`` key-padding mask
    over the encoder tokens unless ``enc_vec``; ``:141`` label mask tiled over the batch;
    ``:146-147`` layer loop; ``:149-163`` optional intermediate outputs / attention maps."""
    b = src_seq.shape[0]
    emb = p[prefix + 'tgt_word_emb.weight']
    n_labels = emb.shape[0]
    tgt_seq = torch.arange(n_labels).view(1, -1).repeat(b, 1)
    dec_input = F.embedding(tgt_seq, emb)
    pad_mask = None
    if not enc_vec:
        pad_mask = padding_mask(tgt_seq, src_seq[:, 0:enc_output.size(1)])
    slf_mask = None
    if label_mask is not None:
        slf_mask = label_mask.view(1, n_labels, n_labels).repeat(b, 1, 1)
    int_outs: List[torch.Tensor] = []
    slf_attns, enc_attns = [], []
    out = dec_input
    for i in range(n_layers):
        out, out_int, slf_attn, enc_attn = decoder_layer(
            p, f'{prefix}layer_stack.{i}.', out, enc_output, slf_mask, pad_mask, n_head, n_head2, no_dec_self_att)
        if int_preds:
            if out_int is not None:
                int_outs.append(out_int)
            int_outs.append(out)
        if return_attns:
            slf_attns.append(slf_attn)
            enc_attns.append(enc_attn)
    if int_preds:
        return out, int_outs
    if return_attns:
        return out, slf_attns, enc_attns
    return out, None


# --------------------------------------------------------------------------------------
# LAMP.forward  (lamp/Models.py:110-137), -encoder graph -decoder graph
# --------------------------------------------------------------------------------------
def lamp_forward(p: Params, cfg: Mapping, src_seq: torch.Tensor, src_pos: torch.Tensor,
                 label_mask: Optional[torch.Tensor], compute_dead_attention: bool = True):
    """Returns ``(logits [B, L], enc_output)``.

    ``:116`` encoder, ``:117`` decoder, ``:124`` full ``[B, L, L]`` projection by
    ``tgt_word_proj.linear.weight`` (bias-free for the graph decoder, ``:79-90``) and ``:126`` its
    diagonal.  ``cfg`` keys: n_layers_enc, n_layers_dec, n_head, n_head2, enc_transform (opt).
    """
    enc_transform = cfg.get('enc_transform', '')
    enc_output, _ = graph_encoder(p, 'encoder.', src_seq, src_pos, cfg['n_layers_enc'], cfg['n_head'],
                                  compute_dead_attention=compute_dead_attention, enc_transform=enc_transform)
    dec_output, _ = graph_decoder(p, 'decoder.', src_seq, enc_output, label_mask, cfg['n_layers_dec'],
                                  cfg['n_head'], cfg.get('n_head2', cfg['n_head']),
                                  enc_vec=(enc_transform != ''),
                                  no_dec_self_att=cfg.get('no_dec_self_att', False))
    seq_logit = F.linear(dec_output, p['tgt_word_proj.linear.weight'], p.get('tgt_word_proj.linear.bias'))
    seq_logit = torch.diagonal(seq_logit, 0, 1, 2)
    return seq_logit.reshape(-1, seq_logit.size(-1)), enc_output


def to_dtype(p: Params, dtype: torch.dtype) -> Dict[str, torch.Tensor]:
    """Cast every floating tensor of a parameter mapping (fp64 arbitration runs)."""
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in p.items()}
