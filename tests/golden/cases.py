"""Case registry shared by ``make_golden.py`` (which runs the REFERENCE) and the tests (which run
the oracle and the CUDA path).  A case is fully described by small integers + seeds; inputs and
weights are regenerated from them with ``lamp_b200.synthetic`` (numpy RandomState), so the
committed fixtures only hold the reference OUTPUTS.
"""
from __future__ import annotations

import numpy as np
import torch

from lamp_b200 import synthetic as syn

# ---- MultiHeadAttention cases (reference: lamp/SubLayers.py:46-121) ---------------------------
#   name: (B, Lq, Lk, D, H, mask_kind, seed, keep_rows_stride)
MHA_CASES = {
    'self_L103_H4_prior':   dict(B=2, Lq=103, Lk=103, D=512, H=4, mask='prior', seed=11, self_attn=True),
    'self_L103_H4_inveye':  dict(B=2, Lq=103, Lk=103, D=512, H=4, mask='inveye', seed=12, self_attn=True),
    'self_L103_H4_bern02':  dict(B=2, Lq=103, Lk=103, D=512, H=4, mask='bern0.02', seed=13, self_attn=True,
                                 random_ln=True),
    'self_L159_H8_none':    dict(B=2, Lq=159, Lk=159, D=512, H=8, mask='none', seed=14, self_attn=True,
                                 attn_row_stride=6),
    'enc_L103_T300_H4_pad': dict(B=3, Lq=103, Lk=300, D=512, H=4, mask='pad', seed=15, self_attn=False,
                                 attn_row_stride=4),
    'enc_L103_T1_H4_vec':   dict(B=3, Lq=103, Lk=1, D=512, H=4, mask='none', seed=16, self_attn=False),
    'self_L40_H1_nofc':     dict(B=2, Lq=40, Lk=40, D=64, H=1, mask='prior', seed=17, self_attn=True),
    'self_L50_H2_diagonly': dict(B=2, Lq=50, Lk=50, D=128, H=2, mask='diagrow', seed=18, self_attn=True),
    'self_L983_H4_prior':   dict(B=1, Lq=983, Lk=983, D=512, H=4, mask='prior', seed=19, self_attn=True,
                                 row_stride=8, keep_attn=False),
    'self_L300_H16_d64':    dict(B=1, Lq=300, Lk=300, D=1024, H=16, mask='none', seed=20, self_attn=True,
                                 row_stride=4, keep_attn=False),
}


def label_adj(kind: str, n: int, seed: int):
    """Adjacency (1 = edge) for a mask kind, or None."""
    if kind == 'prior':
        return syn.prior_adjacency(syn.make_label_sets(n, seed=seed), n)
    if kind.startswith('bern'):
        return syn.bernoulli_adjacency(n, float(kind[4:]), seed)
    if kind == 'diagrow':
        # rows 3 and 7 have NO edges at all -> GraphDecoder forces the self edge (Decoders.py:109-112)
        a = syn.bernoulli_adjacency(n, 0.1, seed)
        for r in (3, 7):
            a[r, :] = 0
            a[:, r] = 0
        return a
    return None


def mha_inputs(c: dict):
    """-> (params dict with '' prefix, q, kv, mask bool [B,Lq,Lk] or None)."""
    rs = np.random.RandomState(c['seed'])
    B, Lq, Lk, D, H = c['B'], c['Lq'], c['Lk'], c['D'], c['H']
    d = D // H
    p = syn.mha_params(rs, '', H, D, d, d, random_ln=c.get('random_ln', False))
    q = torch.from_numpy(rs.standard_normal((B, Lq, D)).astype(np.float32))
    kv = q if c['self_attn'] else torch.from_numpy(rs.standard_normal((B, Lk, D)).astype(np.float32))
    kind = c['mask']
    mask = None
    if kind == 'pad':
        lens = rs.randint(1, Lk + 1, size=B)
        lens[0] = Lk
        keypad = torch.from_numpy(np.arange(Lk)[None, :] >= lens[:, None])
        mask = keypad.unsqueeze(1).expand(B, Lq, Lk)
    elif kind == 'inveye':
        mask = ((1 - torch.eye(Lq)) != 0).unsqueeze(0).expand(B, Lq, Lk)
    elif kind != 'none':
        adj = label_adj(kind, Lq, c['seed'])
        for i in range(Lq):  # Decoders.py:109-112
            if adj[i].sum().item() < 1:
                adj[i, i] = 1
        mask = (adj == 0).unsqueeze(0).expand(B, Lq, Lk)
    return p, q, kv, mask


# ---- full-model cases (reference: lamp/Models.py LAMP, encoder='graph', decoder='graph') ----
MODEL_CASES = {
    # cfg-1 dims (README flags) at a tiny batch
    'lamp_L103_prior': dict(B=2, T=300, V=500, L=103, D=512, d_inner=512, H=4, n_enc=2, n_dec=2,
                            mask='prior', seed=31),
    'lamp_L37_none':   dict(B=3, T=64, V=200, L=37, D=128, d_inner=256, H=8, n_enc=1, n_dec=3,
                            mask='none', seed=32, random_ln=True),
    'lamp_L37_inveye': dict(B=3, T=64, V=200, L=37, D=128, d_inner=256, H=2, n_enc=1, n_dec=2,
                            mask='inveye', seed=33, random_ln=True, pos_enc=False),
    'lamp_L20_meanvec': dict(B=4, T=48, V=100, L=20, D=64, d_inner=128, H=4, n_enc=1, n_dec=1,
                             mask='prior', seed=34, enc_transform='mean'),
}


def model_inputs(c: dict):
    """-> (state dict, cfg dict, src_seq, src_pos, adjacency-or-None)."""
    p = syn.lamp_params(c['V'] + 4, c['L'], c['T'], c['D'], c['d_inner'], c['H'], c['n_enc'], c['n_dec'],
                        seed=c['seed'], random_ln=c.get('random_ln', False), pos_enc=c.get('pos_enc', True))
    src_seq, src_pos = syn.make_tokens(c['B'], c['T'], c['V'], c['seed'] + 1000, min_len=min(20, c['T'] // 2))
    adj = label_adj(c['mask'], c['L'], c['seed']) if c['mask'] not in ('none', 'inveye') else None
    cfg = dict(n_layers_enc=c['n_enc'], n_layers_dec=c['n_dec'], n_head=c['H'], n_head2=c['H'],
               enc_transform=c.get('enc_transform', ''), label_mask=c['mask'])
    return p, cfg, src_seq, src_pos, adj
