"""GPU tests of the fused training sub-layers (round 2; SURVEY.md 8f N4 second pass): ``ops.FFNTrainFunction`` and
``ops.MHATrainFunction`` against fp64 autograd through the CPU oracle's sub-layers (lamp/SubLayers.py:77-142), the
planes-native attention backward (``lamp_attn_bwd_planes``) and the counter-hash dropout kernels."""
import numpy as np
import pytest
import torch

import cases
from lamp_b200 import _native as nat
from lamp_b200 import ops
from lamp_b200 import synthetic as syn
from lamp_b200.SubLayers import MultiHeadAttention, PositionwiseFeedForward
from oracle import lamp_oracle as orc

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def set_dropout(mod, p):
    for m in mod.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = p


@pytest.mark.parametrize('B,L,D,dh', [(3, 37, 128, 256), (2, 103, 512, 512), (1, 300, 64, 128)])
def test_ffn_train_function_matches_fp64_autograd(B, L, D, dh):
    rs = np.random.RandomState(B * 1000 + L)
    p = syn.ffn_params(rs, '', D, dh, random_ln=True)
    mod = PositionwiseFeedForward(D, dh, dropout=0.0)
    mod.load_state_dict(p, strict=True)
    mod = mod.to(DEV).train()
    x = torch.from_numpy(rs.standard_normal((B, L, D)).astype(np.float32))
    w = torch.from_numpy(rs.standard_normal((B, L, D)).astype(np.float32))      # loss = <out, w>
    xg = x.to(DEV).requires_grad_(True)
    ops.STATS.reset()
    out = mod(xg)
    assert ops.STATS.by_kernel.get('gemm_planes', 0) == 2 and 'split_planes' in ops.STATS.by_kernel
    assert getattr(out, '_lamp_planes', None) is not None          # planes travel with the activation
    (out * w.to(DEV)).sum().backward()
    assert ops.STATS.by_kernel.get('relu_mask', 0) == 1 and ops.STATS.by_kernel.get('gemm_tn', 0) == 4
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    x64 = x.double().requires_grad_(True)
    ref = orc.ffn(p64, '', x64)
    (ref * w.double()).sum().backward()
    assert rel_err(out, ref) < 2e-5
    assert rel_err(xg.grad, x64.grad) < 5e-5
    for name, q in mod.named_parameters():
        assert rel_err(q.grad, p64[name].grad) < 5e-5, name


@pytest.mark.parametrize('case', ['self_prior', 'self_none', 'enc_pad', 'self_L159_H8'])
def test_mha_train_function_matches_fp64_autograd(case):
    cfg = dict(self_prior=dict(B=2, Lq=103, Lk=103, D=512, H=4, mask='prior', self_attn=True),
               self_none=dict(B=3, Lq=37, Lk=37, D=128, H=8, mask='none', self_attn=True),
               enc_pad=dict(B=3, Lq=53, Lk=90, D=256, H=4, mask='pad', self_attn=False),
               self_L159_H8=dict(B=2, Lq=159, Lk=159, D=512, H=8, mask='none', self_attn=True))[case]
    c = dict(cfg, seed=dict(self_prior=101, self_none=102, enc_pad=103, self_L159_H8=104)[case])
    p, q, kv, mask = cases.mha_inputs(c)
    rs = np.random.RandomState(7)
    for k_ in p:
        if 'layer_norm' in k_:   # non-trivial affine so that dgamma / dbeta are exercised
            p[k_] = p[k_] + torch.from_numpy(0.1 * rs.standard_normal(p[k_].shape).astype(np.float32))
    d = c['D'] // c['H']
    mod = MultiHeadAttention(c['H'], c['D'], d, d, dropout=0.0)
    mod.load_state_dict(p, strict=True)
    mod = mod.to(DEV).train()
    w = torch.from_numpy(rs.standard_normal(q.shape).astype(np.float32))
    qg = q.to(DEV).requires_grad_(True)
    kvg = qg if c['self_attn'] else kv.to(DEV).requires_grad_(True)
    m_dev = None if mask is None else mask.to(DEV)
    ops.STATS.reset()
    out, attn = mod(qg, kvg, kvg, attn_mask=m_dev)
    assert ops.STATS.by_kernel.get('attn_core_train', 0) == 2
    (out * w.to(DEV)).sum().backward()
    assert ops.STATS.by_kernel.get('attn_core_bwd', 0) == 5
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    q64 = q.double().requires_grad_(True)
    kv64 = q64 if c['self_attn'] else kv.double().requires_grad_(True)
    ref, ref_attn = orc.mha(p64, '', q64, kv64, kv64, mask, c['H'])
    (ref * w.double()).sum().backward()
    assert rel_err(out, ref) < 2e-5 and rel_err(attn, ref_attn) < 2e-5
    assert rel_err(qg.grad, q64.grad) < 5e-5
    if not c['self_attn']:
        assert rel_err(kvg.grad, kv64.grad) < 5e-5
    for name, prm in mod.named_parameters():
        assert rel_err(prm.grad, p64[name].grad) < 5e-5, name


def test_dropout_kernels_share_one_mask_and_scale():
    rows, D, p = 517, 256, 0.3
    y0 = torch.randn(rows, D, device=DEV) + 3.0          # no exact zeros
    zero = torch.zeros(rows, D, device=DEV)
    y = ops.dropout_add(y0, zero, p, seed=1234)
    kept = y != 0
    frac = float(kept.float().mean())
    assert abs(frac - (1 - p)) < 0.01, frac
    assert torch.allclose(y[kept], y0[kept] / (1 - p), rtol=1e-6)
    hi, lo = ops.dropout_split(torch.ones(rows, D, device=DEV), p, seed=1234)
    g = hi.float() + lo.float()
    assert torch.equal(g != 0, kept) and torch.allclose(g[kept], torch.full_like(g[kept], 1 / (1 - p)), rtol=1e-5)
    y2 = ops.dropout_add(y0, zero, p, seed=1235)
    assert not torch.equal(y2 != 0, kept)
    # p = 0: identity + residual / plain split
    x = torch.randn(rows, D, device=DEV)
    assert torch.equal(ops.dropout_add(y0, x, 0.0, 0), y0 + x)
    hi, lo = ops.dropout_split(y0, 0.0, 0)
    h2, l2 = ops.split(y0, nat.PREC_FP32)
    assert torch.equal(hi, h2) and torch.equal(lo, l2)


def test_fused_sublayers_with_dropout_are_consistent_between_forward_and_backward():
    """With dropout the backward recomputes the masks from the seeds: check the FFN gradient against torch autograd
    through an explicit re-statement that uses the SAME masks (read back from the forward's effect)."""
    B, L, D, dh, p = 2, 40, 128, 256, 0.25
    rs = np.random.RandomState(3)
    prm = syn.ffn_params(rs, '', D, dh, random_ln=True)
    mod = PositionwiseFeedForward(D, dh, dropout=p)
    mod.load_state_dict(prm, strict=True)
    mod = mod.to(DEV).train()
    x = torch.randn(B, L, D, device=DEV)
    w = torch.randn(B, L, D, device=DEV)
    torch.manual_seed(11)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())     # the seed ffn_train is about to draw
    torch.manual_seed(11)
    xg = x.clone().requires_grad_(True)
    out = mod(xg)
    (out * w).sum().backward()
    # explicit restatement with the same mask
    hi, lo = ops.dropout_split(torch.ones(B * L, D, device=DEV), p, seed)
    keep = (hi.float() + lo.float()).view(B, L, D).double()
    p64 = {k: v.double().to(DEV).requires_grad_(True) for k, v in prm.items()}
    x64 = x.double().requires_grad_(True)
    h = torch.relu(torch.nn.functional.linear(x64, p64['w_1.weight'][:, :, 0], p64['w_1.bias']))
    y = torch.nn.functional.linear(h, p64['w_2.weight'][:, :, 0], p64['w_2.bias']) * keep + x64
    ref = torch.nn.functional.layer_norm(y, (D,), p64['layer_norm.weight'], p64['layer_norm.bias'], 1e-5)
    (ref * w.double()).sum().backward()
    assert rel_err(out, ref) < 2e-5
    assert rel_err(xg.grad, x64.grad) < 5e-5
    for name, q in mod.named_parameters():
        assert rel_err(q.grad, p64[name].grad) < 5e-5, name


@pytest.mark.parametrize('case', ['self_prior', 'enc_pad', 'self_L300_prior'])
def test_mha_train_recompute_form_matches_fp64_autograd_and_the_stored_form(case):
    """return_attn=False in training: the forward writes no probability tensor, the backward rebuilds P from the saved
    row statistics, the mask and the dropout hash (lamp_attn_bwd_planes, recompute form).  (i) without dropout: fp64
    autograd of the oracle; (ii) with attention dropout 0.2 and the same seeds: identical to the stored-P form."""
    cfg = dict(self_prior=dict(B=2, Lq=103, Lk=103, D=512, H=4, mask='prior', self_attn=True, seed=201),
               enc_pad=dict(B=3, Lq=53, Lk=90, D=256, H=4, mask='pad', self_attn=False, seed=203),
               # multi-tile label graph (L > 128): the forward core reads the mask as packed bits, the recompute
               # backward as bytes
               self_L300_prior=dict(B=2, Lq=300, Lk=300, D=256, H=2, mask='prior', self_attn=True, seed=205))[case]
    c = dict(cfg)
    p, q, kv, mask = cases.mha_inputs(c)
    d = c['D'] // c['H']
    rs = np.random.RandomState(9)
    w = torch.from_numpy(rs.standard_normal(q.shape).astype(np.float32)).to(DEV)
    m_dev = None if mask is None else mask.to(DEV)

    def run(p_drop, return_attn, seed):
        mod = MultiHeadAttention(c['H'], c['D'], d, d, dropout=p_drop)
        mod.load_state_dict(p, strict=True)
        mod = mod.to(DEV).train()
        qg = q.to(DEV).requires_grad_(True)
        kvg = qg if c['self_attn'] else kv.to(DEV).requires_grad_(True)
        torch.manual_seed(seed)
        ops.STATS.reset()
        out, attn = mod(qg, kvg, kvg, attn_mask=m_dev, return_attn=return_attn)
        assert (attn is None) == (not return_attn)
        assert ops.STATS.by_kernel.get('attn_core_train', 0) == (2 if return_attn else 1)
        (out * w).sum().backward()
        grads = {n: x.grad.clone() for n, x in mod.named_parameters()}
        grads['q'] = qg.grad.clone()
        if not c['self_attn']:
            grads['kv'] = kvg.grad.clone()
        return out.detach(), grads

    out, grads = run(0.0, False, 1)
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    q64 = q.double().requires_grad_(True)
    kv64 = q64 if c['self_attn'] else kv.double().requires_grad_(True)
    ref, _ = orc.mha(p64, '', q64, kv64, kv64, mask, c['H'])
    (ref * w.double().cpu()).sum().backward()
    assert rel_err(out, ref) < 2e-5 and rel_err(grads['q'], q64.grad) < 5e-5
    for name in p:
        assert rel_err(grads[name], p64[name].grad) < 5e-5, name
    # with dropout: stored-P form and recompute form see the same masks (same seeds) -> same results
    out_a, g_a = run(0.2, True, 77)
    out_b, g_b = run(0.2, False, 77)
    assert rel_err(out_b, out_a) < 1e-6
    for name in g_a:
        assert rel_err(g_b[name], g_a[name]) < 2e-5, name


@pytest.mark.parametrize('rows,V,P,D,with_pos', [(4000, 500, 61, 128, True), (777, 50, 301, 512, True),
                                                  (1000, 20, 0, 64, False)])
def test_embed_train_matches_torch_embedding_forward_and_backward(rows, V, P, D, with_pos):
    """``ops.embed_train`` (lamp_embed + lamp_embed_bwd) against ``nn.Embedding(padding_idx=0)`` sums in fp64: heavy id
    collisions (V << rows), PAD ids present, the padding rows of both tables get no gradient."""
    g = torch.Generator().manual_seed(rows)
    word = torch.nn.Embedding(V, D, padding_idx=0)
    posm = torch.nn.Embedding(P, D, padding_idx=0) if with_pos else None
    with torch.no_grad():
        word.weight[0].normal_()          # a state dict may hold a non-zero padding row: it is read, never updated
    seq = torch.randint(0, V, (rows,), generator=g)
    pos = torch.randint(0, P, (rows,), generator=g) if with_pos else None
    w = torch.randn(rows, D, generator=g)
    ref = word.weight.double()[seq] + (posm.weight.double()[pos] if with_pos else 0)
    gw = torch.zeros(V, D, dtype=torch.float64).index_add_(0, seq, w.double())
    gw[0] = 0
    if with_pos:
        gp = torch.zeros(P, D, dtype=torch.float64).index_add_(0, pos, w.double())
        gp[0] = 0
    word, posm = word.to(DEV), (posm.to(DEV) if with_pos else None)
    ops.STATS.reset()
    out = ops.embed_train(seq.to(DEV), pos.to(DEV) if with_pos else None, word, posm)
    assert out.shape == (rows, D) and getattr(out, '_lamp_planes', None) is not None
    hi, lo = out._lamp_planes[:2]
    assert rel_err(hi.float() + lo.float(), ref) < 2e-5
    assert rel_err(out, ref) < 1e-6
    (out * w.to(DEV)).sum().backward()
    assert ops.STATS.by_kernel.get('embed', 0) == 1 and ops.STATS.by_kernel.get('embed_bwd', 0) == 1
    assert rel_err(word.weight.grad, gw) < 1e-5 and float(word.weight.grad[0].abs().max()) == 0.0
    if with_pos:
        assert rel_err(posm.weight.grad, gp) < 1e-5 and float(posm.weight.grad[0].abs().max()) == 0.0


def test_weight_and_bias_gradient_kernels_on_ragged_row_counts():
    """``lamp_gemm_tn_acc`` with the vectorised column-sum kernel: row counts that are not multiples of the 32-row
    step, N = 8 ... 1024, accumulation into a pre-loaded db."""
    rs = np.random.RandomState(11)
    for M, N, K in ((1, 8, 8), (31, 264, 64), (33, 512, 128), (26368, 512, 512), (4191, 1024, 64)):
        dy = torch.from_numpy(rs.standard_normal((M, N)).astype(np.float32)).to(DEV)
        x = torch.from_numpy(rs.standard_normal((M, K)).astype(np.float32)).to(DEV)
        d_hi, d_lo = ops.split(dy, nat.PREC_FP32)
        x_hi, x_lo = ops.split(x, nat.PREC_FP32)
        dW, db = ops._gemm_tn(d_hi, d_lo, N, x_hi, x_lo, K, M, True)
        assert rel_err(db, dy.double().sum(0)) < 2e-5, (M, N, K)
        assert rel_err(dW, dy.double().t() @ x.double()) < 2e-5, (M, N, K)


def test_diag_proj_backward_batch_slices():
    """dW / dbias of the diagonal label projection are reduced over batch slices: ragged batch sizes."""
    rs = np.random.RandomState(12)
    for B, L, D in ((1, 5, 64), (7, 103, 512), (256, 103, 512), (33, 983, 128)):
        g = torch.from_numpy(rs.standard_normal((B, L)).astype(np.float32)).to(DEV)
        x = torch.from_numpy(rs.standard_normal((B, L, D)).astype(np.float32)).to(DEV)
        W = torch.from_numpy(rs.standard_normal((L, D)).astype(np.float32)).to(DEV)
        dx, dW, dbias = (torch.empty_like(x), torch.full((L, D), 7.0, device=DEV), torch.full((L,), 7.0, device=DEV))
        nat.check(nat.lib().lamp_diag_proj_bwd(g.data_ptr(), x.data_ptr(), W.data_ptr(), B, L, D, dx.data_ptr(),
                                               dW.data_ptr(), dbias.data_ptr(), nat.stream()), 'diag_proj_bwd')
        assert rel_err(dW, torch.einsum('bl,bld->ld', g.double(), x.double())) < 1e-5, (B, L, D)
        assert rel_err(dbias, g.double().sum(0)) < 1e-5
        assert rel_err(dx, g.double().unsqueeze(-1) * W.double().unsqueeze(0)) < 1e-6


def test_training_weight_planes_cache_builds_w_and_wt_and_follows_updates():
    """``ops._wplanes`` through ``TRAIN_WEIGHTS`` (``lamp_split_planes_multi``): stacked parameters, transposed form,
    Conv1d-shaped weights, ragged shapes, in-place updates (optimizer steps) and ``refresh_all``."""
    g = torch.Generator().manual_seed(21)
    mk = lambda *shape: torch.nn.Parameter(torch.randn(*shape, generator=g).to(DEV))
    Wq, Wk, Wv, Wc, Wr = mk(96, 200), mk(96, 200), mk(40, 200), mk(72, 136, 1), mk(33, 50)

    def check(ws, transpose):
        hi, lo = ops._wplanes(ws if len(ws) > 1 else ws[0], nat.PREC_FP32, transpose=transpose)
        want = torch.cat([w.detach().reshape(w.shape[0], -1) for w in ws], dim=0)
        want = want.t() if transpose else want
        assert hi.shape == want.shape and hi.dtype == torch.bfloat16
        assert torch.equal(hi, want.to(torch.bfloat16))                       # hi = bf16_rn(w)
        assert rel_err(hi.float() + lo.float(), want) < 2e-5
        return hi

    for tr in (False, True):
        for ws in ((Wq, Wk, Wv), (Wq,), (Wk, Wv), (Wc,), (Wr,)):
            check(ws, tr)
    ops.STATS.reset()
    h1 = check((Wq, Wk, Wv), True)
    assert ops.STATS.launches == 0                                           # unchanged weights: cache hit
    with torch.no_grad():
        Wk.mul_(1.5)
        Wr.add_(1.0)
    ops.TRAIN_WEIGHTS.refresh_all(torch.device(DEV, torch.cuda.current_device()))
    n = ops.STATS.launches
    assert n >= 1
    h2 = check((Wq, Wk, Wv), True)
    check((Wk, Wv), False)
    check((Wr,), True)
    assert ops.STATS.launches == n                                           # all refreshed by the one call
    assert h2.data_ptr() == h1.data_ptr()                                    # same buffers (CUDA-graph safe)
    # non-leaf weights (DataParallel replicas) take the direct path
    hi, lo = ops._wplanes((Wq * 1.0, Wk * 1.0), nat.PREC_FP32)
    assert rel_err(hi.float() + lo.float(), torch.cat((Wq, Wk)).detach()) < 2e-5


@pytest.mark.parametrize('M,N,K,bias', [(300, 512, 512, True), (26368, 512, 512, False), (1000, 128, 256, True),
                                        (77, 264, 64, True)])
def test_gemm_dropout_epilogue_equals_gemm_then_dropout_add(M, N, K, bias):
    """``lamp_gemm_planes_drop`` (dropout + residual inside the GEMM epilogue) against the two-kernel sequence
    ``lamp_gemm_planes`` -> ``lamp_dropout_add`` with the same seed: the same elements are dropped (they equal the
    residual exactly), the kept ones agree to the last bits."""
    rs = np.random.RandomState(M + N)
    a = torch.from_numpy(rs.standard_normal((M, K)).astype(np.float32)).to(DEV)
    w = torch.from_numpy((rs.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)).to(DEV)
    x = torch.from_numpy(rs.standard_normal((M, N)).astype(np.float32)).to(DEV)
    b = torch.from_numpy(rs.standard_normal((N,)).astype(np.float32)).to(DEV) if bias else None
    a_hi, a_lo = ops.split(a, nat.PREC_FP32)
    w_hi, w_lo = ops.split(w, nat.PREC_FP32)
    p, seed = 0.3, 123456789
    y0 = torch.empty((M, N), dtype=torch.float32, device=DEV)
    ops.gemm(a_hi, a_lo, K, w_hi, w_lo, K, M, N, K, nat.PREC_FP32, bias=b, out_f32=y0, ldo=N)
    want = ops.dropout_add(y0, x, p, seed)
    got = torch.empty_like(want)
    ops.gemm_drop(a_hi, a_lo, K, w_hi, w_lo, K, M, N, K, bias=b, p_drop=p, seed=seed, residual=x, ldr=N, out_f32=got, ldo=N)
    dropped_want, dropped_got = (want == x), (got == x)
    assert torch.equal(dropped_want, dropped_got)
    frac = float(dropped_got.float().mean())
    assert abs(frac - p) < 0.02, frac
    assert rel_err(got, want) < 1e-6
    ref = (a.double() @ w.double().t() + (b.double() if bias else 0)) / (1 - p)
    kept = ~dropped_got
    assert rel_err((got.double() - x.double())[kept], ref[kept]) < 2e-5


@pytest.mark.parametrize('rows,D,p', [(1000, 512, 0.2), (333, 128, 0.0), (26368, 512, 0.5), (50, 1024, 0.1)])
def test_layernorm_backward_with_fused_dropout_planes(rows, D, p):
    """``lamp_layernorm_bwd_drop``: same dx / dgamma / dbeta as ``lamp_layernorm_bwd`` and bit-identical planes to a
    ``lamp_dropout_split`` pass over dx with the same seed."""
    g = torch.Generator().manual_seed(rows)
    y = torch.randn(rows, D, generator=g).to(DEV)
    gy = torch.randn(rows, D, generator=g).to(DEV)
    gamma = (torch.rand(D, generator=g) + 0.5).to(DEV)
    seed = 987654321
    dy0, dg0, db0 = ops._layernorm_bwd(y, gy, gamma, 1e-5)
    dy1, dg1, db1, (hi, lo) = ops._layernorm_bwd(y, gy, gamma, 1e-5, drop=(p, seed))
    assert torch.equal(dy0, dy1)
    assert rel_err(dg1, dg0) < 1e-5 and rel_err(db1, db0) < 1e-5      # (atomics: order of the partial sums differs)
    h2, l2 = ops.dropout_split(dy0, p, seed)
    assert torch.equal(hi, h2) and torch.equal(lo, l2)
