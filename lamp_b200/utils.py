"""Host-side mask / table helpers of the label-graph path (mirror of lamp/utils.py).

These run once per model or once per batch on the host or with a couple of tiny torch ops; the
per-sample work (mask application, softmax, aggregation) happens inside the CUDA kernels.
"""
import numpy as np
import torch

from . import Constants


def position_encoding_init(n_position, d_pos_vec):
    """Frozen sinusoid position table, row 0 (PAD position) all zeros -- lamp/utils.py:9-19."""
    pos = np.arange(n_position, dtype=np.float64).reshape(-1, 1)
    dim = np.arange(d_pos_vec, dtype=np.float64).reshape(1, -1)
    angle = pos / np.power(10000.0, 2.0 * np.floor(dim / 2.0) / d_pos_vec)
    table = np.zeros_like(angle)
    table[1:, 0::2] = np.sin(angle[1:, 0::2])
    table[1:, 1::2] = np.cos(angle[1:, 1::2])
    return torch.from_numpy(table).type(torch.FloatTensor)


def get_attn_padding_mask(seq_q, seq_k, unsqueeze=True):
    """Bool ``[B, Lq, Lk]`` (a stride-0 expansion over Lq), True where the key token is PAD --
    lamp/utils.py:26-34.  The fused attention kernel consumes the expanded view without ever
    materialising it (query stride 0)."""
    assert seq_q.dim() == 2 and seq_k.dim() == 2
    mb_size, len_q = seq_q.size()
    _, len_k = seq_k.size()
    pad_attn_mask = seq_k.eq(Constants.PAD).unsqueeze(1)
    if unsqueeze:
        pad_attn_mask = pad_attn_mask.expand(mb_size, len_q, len_k)
    return pad_attn_mask


def get_attn_subsequent_mask(seq):
    """Strict upper-triangular (future) mask -- lamp/utils.py:36-44.  Not used by the graph decoder."""
    assert seq.dim() == 2
    b, n = seq.size()
    return torch.triu(torch.ones(b, n, n, dtype=torch.uint8, device=seq.device), diagonal=1)


def swap_0_1(tensor, on_zero, on_non_zero):
    """Elementwise: zeros -> ``on_zero``, everything else -> ``on_non_zero`` -- lamp/utils.py:46-50."""
    return torch.where(tensor == 0, torch.full_like(tensor, on_zero), torch.full_like(tensor, on_non_zero))


def get_gold_binary(gold, tgt_vocab_size):
    """Drop-in for ``utils.utils.get_gold_binary`` (utils/utils.py:205-216; train.py:34, test.py:47) that builds the
    multi-hot target matrix ON THE DEVICE with one kernel instead of a per-row Python loop on the host.  ``gold`` may be
    the CPU copy the reference's loops pass (``gold.data.cpu()``) or a CUDA tensor; the result is a CUDA tensor, so the
    ``.cuda()`` the reference appends is a no-op.  Without a CUDA device the reference semantics run on the host."""
    import torch
    if torch.cuda.is_available():
        from . import ops
        g = gold if gold.is_cuda else gold.cuda(non_blocking=True)
        return ops.gold_binary(g, tgt_vocab_size)
    out = torch.zeros(gold.size(0), tgt_vocab_size + 4)
    for i in range(gold.size(0)):
        idx = gold[i][gold[i] > 0][0:-1]
        if len(idx) > 0:
            out[i].index_fill_(0, idx, 1)
    return out[:, 4:]
