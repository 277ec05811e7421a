"""The CPU oracle (oracle/lamp_oracle.py) against the fixtures produced by the REFERENCE itself
(tests/golden/make_golden.py).  Pins the oracle; runs without a GPU."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import lamp_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
TOL = dict(rtol=2e-5, atol=2e-6)  # MKL kernels may differ between hosts; bit-exact on the generating host


def load(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, name + '.npz')).items()}


def test_sdpa_golden():
    g = load('sdpa_small')
    rs = np.random.RandomState(5)
    n, lq, lk, d = 6, 33, 47, 32
    q = torch.from_numpy(rs.standard_normal((n, lq, d)).astype(np.float32))
    k = torch.from_numpy(rs.standard_normal((n, lk, d)).astype(np.float32))
    v = torch.from_numpy(rs.standard_normal((n, lk, d)).astype(np.float32))
    mask = torch.from_numpy(rs.rand(n, lq, lk) < 0.3)
    mask[:, :, 0] = False
    out, attn = orc.sdpa(q, k, v, mask, float(np.power(d, 0.5)))
    torch.testing.assert_close(out, g['out'], **TOL)
    torch.testing.assert_close(attn, g['attn'], **TOL)


@pytest.mark.parametrize('name', list(cases.MHA_CASES))
def test_mha_golden(name):
    c = cases.MHA_CASES[name]
    g = load('mha_' + name)
    p, q, kv, mask = cases.mha_inputs(c)
    out, attn = orc.mha(p, '', q, kv, kv, mask, c['H'])
    torch.testing.assert_close(out[:, ::c.get('row_stride', 1)], g['out'], **TOL)
    if 'attn' in g:
        torch.testing.assert_close(attn[:, ::c.get('attn_row_stride', 1)], g['attn'], **TOL)
    else:
        torch.testing.assert_close(attn.sum(-1), g['attn_rowsum'], **TOL)
    assert attn.shape == (c['H'] * c['B'], c['Lq'], c['Lk'])  # head-major batch


@pytest.mark.parametrize('name', list(cases.MODEL_CASES))
def test_model_golden(name):
    c = cases.MODEL_CASES[name]
    g = load('model_' + name)
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    lm = orc.label_mask_from(c['L'], adj, c['mask'])
    logits, enc_out = orc.lamp_forward(p, cfg, src_seq, src_pos, lm)
    torch.testing.assert_close(logits, g['logits'], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(enc_out[:, ::7], g['enc_output'], rtol=1e-4, atol=1e-5)
    # the dead encoder attention does not influence the result (lamp/Layers.py:16-18)
    logits2, _ = orc.lamp_forward(p, cfg, src_seq, src_pos, lm, compute_dead_attention=False)
    assert torch.equal(logits, logits2)
    # attention maps and intermediate predictions
    enc_o, enc_attns = orc.graph_encoder(p, 'encoder.', src_seq, src_pos, c['n_enc'], c['H'], return_attns=True,
                                         enc_transform=c.get('enc_transform', ''))
    torch.testing.assert_close(enc_attns[0][:, ::11, ::3], g['enc_slf_attn0'], rtol=1e-4, atol=1e-6)
    dec_o, slf, enc = orc.graph_decoder(p, 'decoder.', src_seq, enc_o, lm, c['n_dec'], c['H'], c['H'],
                                        enc_vec=c.get('enc_transform', '') != '', return_attns=True)
    torch.testing.assert_close(slf[0], g['dec_slf_attn0'], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(enc[-1][:, ::5], g['dec_enc_attn_last'], rtol=1e-4, atol=1e-6)
    _, int_outs = orc.graph_decoder(p, 'decoder.', src_seq, enc_o, lm, c['n_dec'], c['H'], c['H'],
                                    enc_vec=c.get('enc_transform', '') != '', int_preds=True)
    assert len(int_outs) - 1 == int(g['n_int_preds'])  # Models.py:129 drops the last one
    w = p['tgt_word_proj.linear.weight']
    ip0 = torch.einsum('bld,ld->bl', int_outs[0], w)
    torch.testing.assert_close(ip0, g['int_pred0'], rtol=1e-4, atol=1e-5)


def test_state_dict_keys_match_reference():
    """synthetic.lamp_params produces exactly the reference's 77-key state dict."""
    c = cases.MODEL_CASES['lamp_L103_prior']
    p, *_ = cases.model_inputs(c)
    want = {}
    with open(os.path.join(GOLD, 'state_keys_lamp_L103_prior.txt')) as f:
        for line in f:
            k, shape = line.strip().split(' ', 1)
            want[k] = eval(shape)
    assert len(want) == 77
    assert {k: tuple(v.shape) for k, v in p.items()} == want


def test_prior_adjacency_matches_reference_loop():
    from lamp_b200 import synthetic as syn
    rows = syn.make_label_sets(23, n_docs=60, seed=3)
    a = orc.prior_adjacency(rows, 23 + 4)  # literal restatement of data_loader.py:37-47
    b = syn.prior_adjacency(rows, 23)      # vectorised product-side builder
    assert torch.equal(a, b)
    assert torch.equal(a, a.T) and bool((a.diagonal() == 1).all())


def test_label_mask_polarity_and_forced_diagonal():
    adj = torch.zeros(4, 4)
    adj[0, 1] = adj[1, 0] = 1
    m = orc.label_mask_from(4, adj, 'prior')
    assert m.dtype == torch.bool
    assert not m[0, 1] and m[0, 0]          # edge -> not masked; no self edge on a non-empty row
    assert not m[2, 2] and m[2].sum() == 3  # empty row -> forced self edge only
    assert adj[2, 2] == 0                   # oracle does not mutate the caller's tensor
    assert torch.equal(orc.label_mask_from(3, None, 'inveye'), ~torch.eye(3, dtype=torch.bool))
    assert orc.label_mask_from(3, None, 'none') is None
