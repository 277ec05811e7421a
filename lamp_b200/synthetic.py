"""Deterministic synthetic inputs for the label-graph path.

The reference's dataset tarball is not available offline, so every test / bench configuration runs
on synthetic data.  Everything here is generated with ``numpy.random.RandomState`` (bit-stable
across numpy/torch versions and across the build container and the GPU box) and returned as CPU
torch tensors.  Shapes and value conventions follow the reference loaders:

* token rows: ids in ``[4, V+4)`` followed by PAD=0 tails; position ids ``1..len`` with 0 on PAD
  (``utils/data_loader.py:261-279``);
* label sets: rows ``[BOS, l+4, ..., EOS]`` as written by ``utils/preprocess.py:200-232`` and consumed
  by the prior-adjacency builder (``utils/data_loader.py:37-47``);
* weights: the reference initialisers (``lamp/SubLayers.py:57-59,74``; torch defaults elsewhere),
  keyed by the reference ``state_dict`` names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

PAD, UNK, BOS, EOS = 0, 1, 2, 3


def make_tokens(batch: int, max_len: int, vocab: int, seed: int, min_len: int = 20,
                full_length_first: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(src_seq, src_pos)`` int64 ``[batch, max_len]``; lengths ~ U{min_len..max_len}, PAD tails."""
    rs = np.random.RandomState(seed)
    lo = min(min_len, max_len)
    lens = rs.randint(lo, max_len + 1, size=batch)
    if full_length_first and batch > 0:
        lens[0] = max_len  # pad_to_longest pads to the longest row: keep T fixed
    seq = rs.randint(4, vocab + 4, size=(batch, max_len)).astype(np.int64)
    pos = np.tile(np.arange(1, max_len + 1, dtype=np.int64), (batch, 1))
    dead = np.arange(max_len)[None, :] >= lens[:, None]
    seq[dead] = PAD
    pos[dead] = 0
    return torch.from_numpy(seq), torch.from_numpy(pos)


def make_label_sets(n_labels: int, n_docs: Optional[int] = None, seed: int = 0) -> List[List[int]]:
    """Synthetic training label sets: doc j carries label ``j mod L`` (so every label occurs, which
    the reference loader requires) plus k ~ U{0..4} further labels drawn Zipf(s=1) over the labels."""
    rs = np.random.RandomState(seed)
    n_docs = 10 * n_labels if n_docs is None else n_docs
    w = 1.0 / np.arange(1, n_labels + 1)
    w /= w.sum()
    rows = []
    for j in range(n_docs):
        k = rs.randint(0, 5)
        extra = rs.choice(n_labels, size=k, replace=True, p=w).tolist() if k else []
        labs = sorted(set([j % n_labels] + extra))
        rows.append([BOS] + [l + 4 for l in labs] + [EOS])
    return rows


def prior_adjacency(train_tgt: List[List[int]], n_labels: int) -> torch.Tensor:
    """Vectorised equivalent of the reference's O(docs * k^2) Python loop
    (``utils/data_loader.py:37-47``): ``eye(L)`` plus symmetric co-occurrence edges."""
    adj = np.eye(n_labels, dtype=np.float32)
    for row in train_tgt:
        labs = np.asarray(row[1:-1], dtype=np.int64) - 4
        if labs.size > 1:
            adj[np.ix_(labs, labs)] = 1.0
    return torch.from_numpy(adj)


def bernoulli_adjacency(n_labels: int, p: float, seed: int) -> torch.Tensor:
    rs = np.random.RandomState(seed)
    a = (rs.rand(n_labels, n_labels) < p)
    a = np.logical_or(a, a.T)
    np.fill_diagonal(a, True)
    return torch.from_numpy(a.astype(np.float32))


def _normal(rs, shape, std):
    return torch.from_numpy((rs.standard_normal(size=shape) * std).astype(np.float32))


def _uniform(rs, shape, bound):
    return torch.from_numpy(rs.uniform(-bound, bound, size=shape).astype(np.float32))


def mha_params(rs, prefix: str, n_head: int, d_model: int, d_k: int, d_v: int,
               random_ln: bool = False) -> Dict[str, torch.Tensor]:
    """lamp/SubLayers.py:54-59 (projections), :72-74 (fc, xavier normal), :69 (LayerNorm)."""
    p = {
        prefix + 'w_qs.weight': _normal(rs, (n_head * d_k, d_model), math.sqrt(2.0 / (d_model + d_k))),
        prefix + 'w_ks.weight': _normal(rs, (n_head * d_k, d_model), math.sqrt(2.0 / (d_model + d_k))),
        prefix + 'w_vs.weight': _normal(rs, (n_head * d_v, d_model), math.sqrt(2.0 / (d_model + d_v))),
    }
    if n_head > 1:
        p[prefix + 'fc.weight'] = _normal(rs, (d_model, n_head * d_v), math.sqrt(2.0 / (d_model + n_head * d_v)))
    p.update(ln_params(rs, prefix + 'layer_norm.', d_model, random_ln))
    return p


def ln_params(rs, prefix: str, d: int, random_ln: bool) -> Dict[str, torch.Tensor]:
    if random_ln:
        return {prefix + 'weight': torch.from_numpy((1.0 + 0.2 * rs.standard_normal(d)).astype(np.float32)),
                prefix + 'bias': torch.from_numpy((0.1 * rs.standard_normal(d)).astype(np.float32))}
    return {prefix + 'weight': torch.ones(d), prefix + 'bias': torch.zeros(d)}


def ffn_params(rs, prefix: str, d_in: int, d_hid: int, random_ln: bool = False) -> Dict[str, torch.Tensor]:
    """lamp/SubLayers.py:128-130: two Conv1d(k=1) (torch default init: U(+-1/sqrt(fan_in))) + LayerNorm."""
    b1, b2 = 1.0 / math.sqrt(d_in), 1.0 / math.sqrt(d_hid)
    p = {
        prefix + 'w_1.weight': _uniform(rs, (d_hid, d_in, 1), b1), prefix + 'w_1.bias': _uniform(rs, (d_hid,), b1),
        prefix + 'w_2.weight': _uniform(rs, (d_in, d_hid, 1), b2), prefix + 'w_2.bias': _uniform(rs, (d_in,), b2),
    }
    p.update(ln_params(rs, prefix + 'layer_norm.', d_in, random_ln))
    return p


def lamp_params(n_src_vocab: int, n_labels: int, n_max_seq: int, d_model: int, d_inner: int, n_head: int,
                n_layers_enc: int, n_layers_dec: int, seed: int = 0, n_head2: Optional[int] = None,
                random_ln: bool = False, pos_enc: bool = True) -> Dict[str, torch.Tensor]:
    """A full ``LAMP(encoder='graph', decoder='graph')`` state dict (SURVEY.md section 8b key set)."""
    from .utils import position_encoding_init
    rs = np.random.RandomState(seed)
    n_head2 = n_head if not n_head2 else n_head2
    d_k = d_model // n_head
    p: Dict[str, torch.Tensor] = {}
    emb = _normal(rs, (n_src_vocab, d_model), 1.0)
    emb[PAD] = 0.0  # nn.Embedding(padding_idx=PAD)
    p['encoder.src_word_emb.weight'] = emb
    if pos_enc:
        p['encoder.position_enc.weight'] = position_encoding_init(n_max_seq + 1, d_model)
    for i in range(n_layers_enc):
        p.update(mha_params(rs, f'encoder.layer_stack.{i}.slf_attn.', n_head, d_model, d_k, d_k, random_ln))
        p.update(ffn_params(rs, f'encoder.layer_stack.{i}.pos_ffn.', d_model, d_inner, random_ln))
    p['decoder.tgt_word_emb.weight'] = _normal(rs, (n_labels, d_model), 1.0)
    for i in range(n_layers_dec):
        p.update(mha_params(rs, f'decoder.layer_stack.{i}.enc_attn.', n_head, d_model, d_k, d_k, random_ln))
        p.update(ffn_params(rs, f'decoder.layer_stack.{i}.pos_ffn1.', d_model, d_inner, random_ln))
        d_k2 = d_k  # the reference passes the same d_k/d_v to both attentions (lamp/Layers.py:26,30)
        p.update(mha_params(rs, f'decoder.layer_stack.{i}.slf_attn.', n_head2, d_model, d_k2, d_k2, random_ln))
        p.update(ffn_params(rs, f'decoder.layer_stack.{i}.pos_ffn2.', d_model, d_inner, random_ln))
    p['tgt_word_proj.weight'] = p['decoder.tgt_word_emb.weight']  # alias, unused by forward (Models.py:88-90)
    p['tgt_word_proj.linear.weight'] = _normal(rs, (n_labels, d_model), math.sqrt(2.0 / (d_model + n_labels)))
    return p


def make_dataset_dict(n_labels: int, vocab: int, n_train: int, n_valid: int, n_test: int, max_len: int = 40,
                      seed: int = 0):
    """A ``train_valid_test.pt``-style dict in the reference's on-disk format (utils/preprocess.py:200-232):
    word/label <-> index dicts with the 4 special tokens first, ``src`` rows ``[BOS, w..., EOS]`` and ``tgt`` rows
    ``[BOS, l..., EOS]``; every label occurs in ``train`` (required by utils/data_loader.py:60-73)."""
    import argparse
    rs = np.random.RandomState(seed)
    special = {'<blank>': PAD, '<unk>': UNK, '<s>': BOS, '</s>': EOS}
    src_dict = dict(special)
    src_dict.update({f'w{i}': i + 4 for i in range(vocab)})
    tgt_dict = dict(special)
    tgt_dict.update({f'l{i}': i + 4 for i in range(n_labels)})

    def split(n, s):
        tg = make_label_sets(n_labels, n_docs=n, seed=s)
        src = []
        for _ in range(n):
            ln = rs.randint(5, max_len + 1)
            src.append([BOS] + rs.randint(4, vocab + 4, size=ln).tolist() + [EOS])
        return {'src': src, 'tgt': tg}
    settings = argparse.Namespace(max_seq_len=max_len + 2, max_tgt_len=n_labels)
    return {'settings': settings, 'dict': {'src': src_dict, 'tgt': tgt_dict},
            'train': split(max(n_train, n_labels), seed + 1), 'valid': split(n_valid, seed + 2),
            'test': split(n_test, seed + 3)}
