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
