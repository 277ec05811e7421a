"""GPU parity of the drop-in modules (lamp_b200.SubLayers / Layers / Decoders / Encoders / Models) against the
reference-generated fixtures and the CPU oracle.  Tolerance: 1e-3 relative (north_star), fp32 path."""
import os

import numpy as np
import pytest
import torch

import cases
import lamp_b200
from lamp_b200.Models import LAMP
from lamp_b200.SubLayers import MultiHeadAttention, PositionwiseFeedForward, ScaledDotProductAttention
from lamp_b200.Layers import DecoderLayer, EncoderLayer
from oracle import lamp_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
DEV = 'cuda'
TOL = 1e-3


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, name + '.npz')).items()}


def build_model(c, p, adj):
    d = c['D'] // c['H']
    m = LAMP(c['V'] + 4, c['L'], c['T'], c['L'], n_layers_enc=c['n_enc'], n_layers_dec=c['n_dec'], n_head=c['H'],
             n_head2=c['H'], d_word_vec=c['D'], d_model=c['D'], d_inner_hid=c['d_inner'], d_k=d, d_v=d, dropout=0.2,
             dec_dropout=0.2, dec_dropout2=False, proj_share_weight=True, encoder='graph', decoder='graph',
             enc_transform=c.get('enc_transform', ''), no_enc_pos_embedding=not c.get('pos_enc', True),
             label_adj_matrix=adj, label_mask=c['mask'])
    m.load_state_dict(p, strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize('name', list(cases.MODEL_CASES))
def test_lamp_forward_golden(name):
    c = cases.MODEL_CASES[name]
    g = load('model_' + name)
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    adj_before = None if adj is None else adj.clone()
    model = build_model(c, p, adj)
    assert adj is None or torch.equal(adj, adj_before)  # caller's adjacency is not mutated
    src = (src_seq.to(DEV), src_pos.to(DEV))
    with torch.no_grad():
        logits, enc_out, none = model(src, None, None, None)
    assert none is None and logits.shape == (c['B'], c['L'])
    e = rel_err(logits, g['logits'])
    e_enc = rel_err(enc_out[:, ::7], g['enc_output'])
    print(f'{name}: logits vs reference {e:.2e}, enc_output {e_enc:.2e}')
    assert e < TOL and e_enc < TOL
    # attention maps: layout [H*B, Lq, Lk] head-major, returned as in the reference
    with torch.no_grad():
        logits2, _, enc_attns, dec_rest = model(src, None, None, None, return_attns=True)
    dec_slf, dec_enc = dec_rest
    assert rel_err(logits2, g['logits']) < TOL
    assert rel_err(dec_slf[0], g['dec_slf_attn0']) < TOL
    assert rel_err(dec_enc[-1][:, ::5], g['dec_enc_attn_last']) < TOL
    assert rel_err(enc_attns[0][0][:, ::11, ::3], g['enc_slf_attn0']) < TOL
    with torch.no_grad():
        logits3, _, int_preds = model(src, None, None, None, int_preds=True)
    assert len(int_preds) == int(g['n_int_preds'])
    assert rel_err(int_preds[0], g['int_pred0']) < TOL
    assert rel_err(logits3, g['logits']) < TOL


def test_eval_mode_without_no_grad_uses_fused_path_and_matches():
    c = cases.MODEL_CASES['lamp_L37_none']
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    model = build_model(c, p, adj)
    src = (src_seq.to(DEV), src_pos.to(DEV))
    logits, _, _ = model(src, None, None, None)  # grad enabled, eval mode
    assert not logits.requires_grad
    g = load('model_lamp_L37_none')
    assert rel_err(logits, g['logits']) < TOL


def test_training_path_is_differentiable_and_matches_oracle_without_dropout():
    c = dict(cases.MODEL_CASES['lamp_L37_inveye'])
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    model = build_model(c, p, adj)
    model.train()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    src = (src_seq.to(DEV), src_pos.to(DEV))
    logits, _, _ = model(src, None, None, None)
    assert logits.requires_grad
    lm = orc.label_mask_from(c['L'], adj, c['mask'])
    ref, _ = orc.lamp_forward(p, cfg, src_seq, src_pos, lm)
    assert rel_err(logits, ref) < TOL
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, torch.zeros_like(logits))
    loss.backward()
    got = [n for n, q in model.named_parameters() if q.grad is not None]
    dead = [n for n, q in model.named_parameters() if q.grad is None and q.requires_grad]
    assert any('decoder.layer_stack.0.slf_attn.w_qs' in n for n in got)
    # the encoder self-attention never receives gradients (lamp/Layers.py:16-18), nor does the alias / pos table
    assert all(('encoder.layer_stack' in n and 'slf_attn' in n) or n == 'encoder.position_enc.weight' for n in dead), dead


def test_training_gradients_native_attention_core_equals_composed_path():
    """Training step with the attention core, the projections / FFN contractions and the LayerNorms native in both
    directions (ops.SDPAFunction, LinearFunction, LayerNormFunction) vs the all-torch composed path: same logits and
    the same gradient for every parameter (dropout off so that both are deterministic)."""
    from lamp_b200 import ops
    c = dict(cases.MODEL_CASES['lamp_L37_none'])
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    src = (src_seq.to(DEV), src_pos.to(DEV))
    res = {}
    for native in (True, False):
        ops.NATIVE_ATTENTION_BACKWARD = ops.NATIVE_TRAINING = native
        try:
            model = build_model(c, p, adj)
            model.train()
            for mod in model.modules():
                if isinstance(mod, torch.nn.Dropout):
                    mod.p = 0.0
            ops.STATS.reset()
            logits, _, _ = model(src, None, None, None)
            tgt = (torch.arange(logits.numel(), device=DEV).view_as(logits) % 3 == 0).float()
            torch.nn.functional.binary_cross_entropy_with_logits(logits, tgt).backward()
            torch.cuda.synchronize()
            res[native] = (logits.detach(), {n: q.grad.clone() for n, q in model.named_parameters() if q.grad is not None},
                           dict(ops.STATS.by_kernel))
        finally:
            ops.NATIVE_ATTENTION_BACKWARD = ops.NATIVE_TRAINING = True
    assert res[True][2].get('attn_core_bwd', 0) > 0 and res[False][2].get('attn_core_bwd', 0) == 0
    assert res[True][2].get('gemm_tn', 0) > 0 and res[True][2].get('layernorm_bwd', 0) > 0
    assert res[False][2].get('gemm_tn', 0) == 0 and res[False][2].get('layernorm_bwd', 0) == 0
    assert rel_err(res[True][0], res[False][0]) < 1e-4
    assert res[True][1].keys() == res[False][1].keys()
    worst = max((rel_err(res[True][1][n], res[False][1][n]), n) for n in res[True][1])
    print('worst gradient difference native vs composed:', worst)
    assert worst[0] < 1e-3, worst


@pytest.mark.parametrize('name', ['lamp_L37_none', 'lamp_L103_prior'])
def test_training_gradients_match_oracle_autograd_fp64(name):
    """Parameter gradients of one native training step (dropout off) against autograd through the CPU oracle in fp64
    -- the oracle being pinned to the reference, this anchors the backward to the reference's own arithmetic.
    The model has ReLU kinks: a pre-activation that is 0 to within the forward's rounding error (a handful out of 1e5)
    can land on the other side in fp32, which changes the gradient of that FFN's w_1 / b_1 (and, diluted, everything
    upstream) by a rank-1 term -- 0.9 % in max-norm for `lamp_L103_prior`, independent of the kernels
    (scripts/probes/grad_debug3.py: every tensor before that ReLU agrees to 1e-5).  Bars: `lamp_L37_none` (no such
    pre-activation) 1e-3 in max-norm for every gradient; `lamp_L103_prior` 5e-3 in Frobenius norm for every gradient,
    median below 1e-3 and 90 % below 5e-3 in max-norm."""
    from lamp_b200 import ops
    c = dict(cases.MODEL_CASES[name])
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    B = src_seq.shape[0]
    tgt = (torch.arange(B * c['L']).view(B, c['L']) % 5 == 0).float()
    model = build_model(c, p, adj)
    model.train()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    ops.STATS.reset()
    logits, _, _ = model((src_seq.to(DEV), src_pos.to(DEV)), None, None, None)
    torch.nn.functional.binary_cross_entropy_with_logits(logits, tgt.to(DEV)).backward()
    assert ops.STATS.by_kernel.get('attn_core_bwd', 0) > 0 and ops.STATS.by_kernel.get('gemm_tn', 0) > 0
    pd = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in p.items()}
    lm = orc.label_mask_from(c['L'], adj, c['mask'])
    ref_logits, _ = orc.lamp_forward(pd, cfg, src_seq, src_pos, lm)
    torch.nn.functional.binary_cross_entropy_with_logits(ref_logits, tgt.double()).backward()
    assert rel_err(logits.detach(), ref_logits.detach()) < TOL
    mx, fro = [], []
    for n, q in model.named_parameters():
        ref = pd[n].grad if n in pd else None
        if q.grad is None or ref is None or float(ref.abs().max()) == 0.0:
            continue
        g = q.grad.double().cpu()
        mx.append((rel_err(g, ref), n))
        fro.append((float((g - ref).norm() / ref.norm()), n))
    assert len(mx) > 40
    mx.sort()
    print(f'{name}: {len(mx)} gradients vs oracle fp64 autograd: max-norm median {mx[len(mx) // 2][0]:.1e}, '
          f'90th pct {mx[int(len(mx) * 0.9)][0]:.1e}, worst {mx[-1]}; Frobenius worst {max(fro)}')
    if name == 'lamp_L37_none':       # no pre-activation within rounding error of 0 in this case: tight everywhere
        assert mx[-1][0] < TOL, mx[-5:]
    else:                             # one kink flip in the last decoder layer's first FFN (see the docstring)
        assert max(fro)[0] < 5e-3, max(fro)
        assert mx[int(len(mx) * 0.9)][0] < 5e-3 and mx[len(mx) // 2][0] < TOL, mx[-5:]


def test_training_with_dropout_runs_native_core_and_is_seeded():
    """model.train() with the reference's dropout rates: the native core applies dropout to the probabilities, two
    steps with the same torch seed give identical losses and gradients, another seed gives different ones."""
    from lamp_b200 import ops
    c = dict(cases.MODEL_CASES['lamp_L37_none'])
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    src = (src_seq.to(DEV), src_pos.to(DEV))

    def step(seed):
        torch.manual_seed(seed)
        model = build_model(c, p, adj)
        model.train()
        logits, _, _ = model(src, None, None, None)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, torch.zeros_like(logits))
        loss.backward()
        return float(loss), model.decoder.layer_stack[0].slf_attn.w_qs.weight.grad.clone()

    ops.STATS.reset()
    l1, g1 = step(5)
    assert ops.STATS.by_kernel.get('attn_core_bwd', 0) > 0
    l2, g2 = step(5)
    l3, g3 = step(6)
    # same seed -> same dropout masks: identical loss; the gradients agree to fp32 summation order (the split-K weight
    # gradient reduces its partial tiles with fp32 red.global.add, whose order is not fixed)
    assert l1 == l2 and rel_err(g2, g1) < 1e-5
    assert l1 != l3 and rel_err(g3, g1) > 1e-3
    assert torch.isfinite(g1).all()


@pytest.mark.parametrize('name', ['self_L103_H4_prior', 'enc_L103_T300_H4_pad', 'self_L40_H1_nofc',
                                  'self_L983_H4_prior'])
def test_mha_module(name):
    c = cases.MHA_CASES[name]
    g = load('mha_' + name)
    p, q, kv, mask = cases.mha_inputs(c)
    d = c['D'] // c['H']
    m = MultiHeadAttention(c['H'], c['D'], d, d, dropout=0.1)
    m.load_state_dict(p, strict=True)
    m = m.to(DEV).eval()
    qd = q.to(DEV)
    kvd = qd if kv is q else kv.to(DEV)
    md = None if mask is None else mask.to(DEV)
    with torch.no_grad():
        out, attn = m(qd, kvd, kvd, attn_mask=md)
    assert attn.shape == (c['H'] * c['B'], c['Lq'], c['Lk'])
    assert rel_err(out[:, ::c.get('row_stride', 1)], g['out']) < TOL
    # uint8 masks (the reference's .byte()) are accepted too
    if md is not None:
        with torch.no_grad():
            out2, _ = m(qd, kvd, kvd, attn_mask=md.contiguous().to(torch.uint8), return_attn=False)
        assert torch.equal(out, out2)
    # distinct k / v tensors
    with torch.no_grad():
        out3, _ = m(qd, kvd, kvd.clone(), attn_mask=md, return_attn=False)
    assert rel_err(out3, out) < 1e-5


def test_sdpa_module_and_layers():
    rs = np.random.RandomState(0)
    q = torch.from_numpy(rs.standard_normal((6, 50, 64)).astype(np.float32))
    k = torch.from_numpy(rs.standard_normal((6, 70, 64)).astype(np.float32))
    v = torch.from_numpy(rs.standard_normal((6, 70, 64)).astype(np.float32))
    mask = torch.from_numpy(rs.rand(6, 50, 70) < 0.5)
    mask[:, :, 3] = False
    sd = ScaledDotProductAttention(8.0).to(DEV).eval()
    with torch.no_grad():
        out, attn = sd(q.to(DEV), k.to(DEV), v.to(DEV), attn_mask=mask.to(DEV))
    ro, ra = orc.sdpa(q, k, v, mask, 8.0)
    assert rel_err(out, ro) < TOL and rel_err(attn, ra) < TOL
    # DecoderLayer / EncoderLayer / FFN modules against the oracle
    from lamp_b200 import synthetic as syn
    D, H, dh, B, L, T = 128, 4, 256, 3, 37, 50
    p = {}
    p.update(syn.mha_params(rs, 'enc_attn.', H, D, D // H, D // H, True))
    p.update(syn.ffn_params(rs, 'pos_ffn1.', D, dh, True))
    p.update(syn.mha_params(rs, 'slf_attn.', H, D, D // H, D // H, True))
    p.update(syn.ffn_params(rs, 'pos_ffn2.', D, dh, True))
    layer = DecoderLayer(D, dh, H, H, D // H, D // H)
    layer.load_state_dict(p, strict=True)
    layer = layer.to(DEV).eval()
    x = torch.from_numpy(rs.standard_normal((B, L, D)).astype(np.float32))
    enc = torch.from_numpy(rs.standard_normal((B, T, D)).astype(np.float32))
    lm = torch.from_numpy(rs.rand(L, L) < 0.6)
    lm[torch.arange(L), torch.arange(L)] = False
    slf = lm.unsqueeze(0).expand(B, L, L)
    with torch.no_grad():
        out, out_int, sa, ea = layer(x.to(DEV), enc.to(DEV), slf_attn_mask=slf.to(DEV))
    r_out, r_int, r_sa, r_ea = orc.decoder_layer(p, '', x, enc, slf, None, H, H)
    for a, b in ((out, r_out), (out_int, r_int), (sa, r_sa), (ea, r_ea)):
        assert rel_err(a, b) < TOL
    pe = {}
    pe.update(syn.mha_params(rs, 'slf_attn.', H, D, D // H, D // H, True))
    pe.update(syn.ffn_params(rs, 'pos_ffn.', D, dh, True))
    el = EncoderLayer(D, dh, H, D // H, D // H)
    el.load_state_dict(pe, strict=True)
    el = el.to(DEV).eval()
    with torch.no_grad():
        eo, eattn = el(enc.to(DEV))
    r_eo, r_eattn = orc.encoder_layer(pe, '', enc, None, H)
    assert rel_err(eo, r_eo) < TOL and rel_err(eattn, r_eattn) < TOL
    ff = PositionwiseFeedForward(D, dh)
    ff.load_state_dict({k[len('pos_ffn.'):]: v for k, v in pe.items() if k.startswith('pos_ffn.')})
    ff = ff.to(DEV).eval()
    with torch.no_grad():
        fo = ff(enc.to(DEV))
    assert rel_err(fo, orc.ffn(pe, 'pos_ffn.', enc)) < TOL


def test_cpu_tensors_raise():
    m = MultiHeadAttention(2, 32, 16, 16).eval()
    x = torch.zeros(1, 4, 32)
    with pytest.raises(RuntimeError, match='no CPU path'):
        m(x, x, x)


def test_full_size_properties_cfg2():
    """BASELINE cfg-2 shape at a bench-size batch, through size-independent properties: batch-permutation
    equivariance, independence of samples from one another, probabilities sum to one."""
    c = cases.MHA_CASES['self_L103_H4_prior']
    p, q, kv, mask = cases.mha_inputs(c)
    d = c['D'] // c['H']
    m = MultiHeadAttention(c['H'], c['D'], d, d)
    m.load_state_dict(p, strict=True)
    m = m.to(DEV).eval()
    B = 1024
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(B, 103, 512, device=DEV, generator=g)
    md = mask[:1].to(DEV).expand(B, 103, 103)
    with torch.no_grad():
        out, _ = m(x, x, x, attn_mask=md, return_attn=False)
        perm = torch.randperm(B, device=DEV, generator=g)
        out_p, _ = m(x[perm], x[perm], x[perm], attn_mask=md, return_attn=False)
        out_s, attn_s = m(x[:3], x[:3], x[:3], attn_mask=md[:3])
    assert torch.isfinite(out).all()
    assert torch.equal(out[perm], out_p)            # deterministic, sample-independent
    assert torch.equal(out[:3], out_s)
    assert rel_err(attn_s.sum(-1), torch.ones_like(attn_s.sum(-1))) < 1e-5
    ref, _ = orc.mha(p, '', x[:2].cpu(), x[:2].cpu(), x[:2].cpu(), mask[:2], c['H'])
    assert rel_err(out[:2], ref) < TOL


def test_padding_aware_path_equals_dense_path_and_oracle():
    """The packed (PAD-skipping) encoder / K|V projection / label<-input attention gives the same enc_output (PAD rows
    included) and logits as the dense path and the oracle -- also for inputs the reference loader never produces
    (PAD in the middle of a row, a PAD token carrying a non-zero position id, a row that is entirely PAD-free)."""
    from lamp_b200 import ops
    c = dict(cases.MODEL_CASES['lamp_L37_none'], B=6)
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    src_seq, src_pos = src_seq.clone(), src_pos.clone()
    src_seq[1, 5] = 0
    src_pos[1, 5] = 0           # PAD in the middle of a row
    src_seq[2, 7] = 0           # PAD token that keeps its position id (not equal to the representative row)
    src_seq[3] = src_seq[0]
    src_pos[3] = src_pos[0]     # no padding at all (row 0 is full length by construction)
    model = build_model(c, p, adj)
    src = (src_seq.to(DEV), src_pos.to(DEV))
    outs = {}
    for aware in (True, False):
        ops.PADDING_AWARE = aware
        try:
            with torch.no_grad():
                outs[aware] = model(src, None, None, None)
        finally:
            ops.PADDING_AWARE = True
    lm = orc.label_mask_from(c['L'], adj, c['mask'])
    ref_logits, ref_enc = orc.lamp_forward(p, cfg, src_seq, src_pos, lm)
    for aware in (True, False):
        logits, enc_out, _ = outs[aware]
        assert rel_err(logits, ref_logits) < TOL, aware
        assert rel_err(enc_out, ref_enc) < TOL, aware
    assert rel_err(outs[True][0], outs[False][0]) < 1e-4
    assert rel_err(outs[True][1], outs[False][1]) < 1e-5
    # int_preds / return_attns still work (they fall back to dense keys where the layout demands it)
    with torch.no_grad():
        l2, _, enc_attns, dec_rest = model(src, None, None, None, return_attns=True)
    assert rel_err(l2, ref_logits) < TOL


@pytest.mark.parametrize('name', ['lamp_L103_prior', 'lamp_L37_none', 'lamp_L20_meanvec'])
def test_deferred_layernorm_equals_explicit_layernorm_and_oracle(name):
    """ops.DEFER_LAYERNORM (LayerNorm folded into the consumers' epilogues, no LayerNorm kernels inside the stack) vs
    the explicit GEMM + LayerNorm kernels vs the oracle, incl. int_preds / return_attns (materialised tensors)."""
    from lamp_b200 import ops
    c = cases.MODEL_CASES[name]
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    # non-trivial gamma / beta everywhere so that the folding is exercised
    g = torch.Generator().manual_seed(7)
    p = {k: (v + 0.2 * torch.randn(v.shape, generator=g) if 'layer_norm' in k else v) for k, v in p.items()}
    model = build_model(c, p, adj)
    src = (src_seq.to(DEV), src_pos.to(DEV))
    lm = orc.label_mask_from(c['L'], adj, c['mask'])
    ref_logits, ref_enc = orc.lamp_forward(p, cfg, src_seq, src_pos, lm)
    outs = {}
    default = ops.DEFER_LAYERNORM
    for defer in (True, False):
        ops.DEFER_LAYERNORM = defer
        try:
            ops.STATS.reset()
            with torch.no_grad():
                outs[defer] = model(src, None, None, None)
                ints = model(src, None, None, None, int_preds=True)
                attn = model(src, None, None, None, return_attns=True)
            torch.cuda.synchronize()
            kern = dict(ops.STATS.by_kernel)
        finally:
            ops.DEFER_LAYERNORM = default
        logits, enc_out, _ = outs[defer]
        e1, e2 = rel_err(logits, ref_logits), rel_err(enc_out, ref_enc)
        print(f'{name} defer={defer}: logits {e1:.2e} enc {e2:.2e} kernels {kern}')
        assert e1 < TOL and e2 < TOL
        assert rel_err(ints[0], ref_logits) < TOL and rel_err(attn[0], ref_logits) < TOL
    assert rel_err(outs[True][0], outs[False][0]) < 1e-4
    assert rel_err(outs[True][1], outs[False][1]) < 1e-4


def test_graphed_train_step_equals_eager_and_redraws_dropout():
    """lamp_b200.GraphedTrainStep: (i) without dropout a replay produces the eager step's loss and gradients for new
    inputs, (ii) with dropout successive replays on the same batch draw different attention-dropout masks (the
    device-side seed counter advances inside the graph) and stay finite."""
    from lamp_b200 import ops
    c = dict(cases.MODEL_CASES['lamp_L37_none'])
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    B, T = src_seq.shape
    tgt = (torch.arange(B * c['L']).view(B, c['L']) % 4 == 0).float()
    loss_fn = torch.nn.functional.binary_cross_entropy_with_logits
    try:
        model = build_model(c, p, adj)
        model.train()
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        step = lamp_b200.GraphedTrainStep(model, loss_fn, B, T)
        assert step.kernels_per_replay > 50
        # new inputs after capture: a permutation of the batch
        perm = torch.arange(B - 1, -1, -1)
        loss = step(src_seq[perm], src_pos[perm], tgt[perm]).clone()
        g_graph = {n: q.grad.clone() for n, q in model.named_parameters() if q.grad is not None}
        ops.TRAIN_SEED_DEV = None
        model.zero_grad(set_to_none=True)
        logits, _, _ = model((src_seq[perm].to(DEV), src_pos[perm].to(DEV)), None, None, None)
        l2 = loss_fn(logits, tgt[perm].to(DEV))
        l2.backward()
        assert abs(float(loss) - float(l2)) < 1e-6
        for n, q in model.named_parameters():
            if q.grad is not None:
                assert rel_err(g_graph[n], q.grad) < 1e-5, n
        # dropout: replays differ
        model2 = build_model(c, p, adj)
        model2.train()
        step2 = lamp_b200.GraphedTrainStep(model2, loss_fn, B, T, example=(src_seq, src_pos, tgt))
        losses = [float(step2(src_seq, src_pos, tgt)) for _ in range(4)]
        assert len(set(losses)) == 4, losses
        assert all(torch.isfinite(q.grad).all() for q in model2.parameters() if q.grad is not None)
    finally:
        ops.TRAIN_SEED_DEV = None


def test_graphed_train_step_with_capturable_optimizer_learns():
    """A whole training iteration (zero_grad, forward, loss, backward, Adam step) as one graph replay: the loss on a
    fixed batch goes down and the parameters move."""
    from lamp_b200 import ops
    c = dict(cases.MODEL_CASES['lamp_L37_none'])
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    B, T = src_seq.shape
    tgt = (torch.arange(B * c['L']).view(B, c['L']) % 4 == 0).float()
    try:
        model = build_model(c, p, adj)
        model.train()
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        w0 = model.decoder.layer_stack[0].slf_attn.w_qs.weight.detach().clone()
        opt = torch.optim.Adam(model.get_trainable_parameters(), lr=1e-3, betas=(0.9, 0.98), capturable=True)
        step = lamp_b200.GraphedTrainStep(model, torch.nn.functional.binary_cross_entropy_with_logits, B, T,
                                          example=(src_seq, src_pos, tgt), optimizer=opt)
        losses = [float(step(src_seq, src_pos, tgt)) for _ in range(12)]
        print('graphed train+Adam losses', [round(x, 4) for x in losses])
        assert losses[-1] < 0.8 * losses[0]
        assert not torch.equal(w0, model.decoder.layer_stack[0].slf_attn.w_qs.weight.detach())
    finally:
        ops.TRAIN_SEED_DEV = None


def test_graphed_forward_replays_equal_eager_for_new_batches():
    """GraphedForward (CUDA-graph replay of LAMP.forward): replaying with NEW token ids -- different padding, hence a
    different device-side packed row count -- gives bit-identical logits / enc_output to the eager call."""
    import lamp_b200
    c = dict(cases.MODEL_CASES['lamp_L37_none'], B=6)
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    model = build_model(c, p, adj)
    runner = lamp_b200.GraphedForward(model, batch=src_seq.shape[0], seq_len=src_seq.shape[1])
    assert runner.kernels_per_replay > 10
    g = torch.Generator().manual_seed(3)
    for trial in range(3):
        seq, pos = src_seq.clone(), src_pos.clone()
        if trial:
            for b in range(seq.shape[0]):  # new lengths and ids
                n = int(torch.randint(1, seq.shape[1] + 1, (1,), generator=g))
                seq[b, :n] = torch.randint(4, c['V'] + 4, (n,), generator=g)
                pos[b, :n] = torch.arange(1, n + 1)
                seq[b, n:] = 0
                pos[b, n:] = 0
        with torch.no_grad():
            ref_logits, ref_enc, _ = model((seq.to(DEV), pos.to(DEV)), None, None, None)
        src = (seq.pin_memory(), pos.pin_memory()) if trial == 1 else (seq.to(DEV), pos.to(DEV))
        logits, enc = runner(*src)
        torch.cuda.synchronize()
        assert torch.equal(logits, ref_logits), trial
        assert torch.equal(enc, ref_enc), trial
    with pytest.raises(RuntimeError):
        runner(src_seq[:2].to(DEV), src_pos[:2].to(DEV))


def test_graphed_train_step_follows_eager_optimizer_steps_between_replays():
    """The documented loop ``loss = step(...); optimizer.step()``: the weights change BETWEEN replays, on the host
    side, so the replay itself must rebuild the weight operand planes (W and W^T of ``ops.TRAIN_WEIGHTS``): its loss
    and gradients equal an eager step on the updated weights."""
    from lamp_b200 import ops
    c = dict(cases.MODEL_CASES['lamp_L37_none'])
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    B, T = src_seq.shape
    tgt = (torch.arange(B * c['L']).view(B, c['L']) % 3 == 0).float()
    loss_fn = torch.nn.functional.binary_cross_entropy_with_logits
    try:
        model = build_model(c, p, adj)
        model.train()
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        # an eager step first, so that the cache entries exist (and are fresh) when the capture starts
        loss_fn(model((src_seq.to(DEV), src_pos.to(DEV)), None, None, None)[0], tgt.to(DEV)).backward()
        step = lamp_b200.GraphedTrainStep(model, loss_fn, B, T, example=(src_seq, src_pos, tgt))
        opt = torch.optim.SGD(model.get_trainable_parameters(), lr=0.5)
        seen = []
        for _ in range(3):
            loss = float(step(src_seq, src_pos, tgt))
            g_graph = {n: q.grad.clone() for n, q in model.named_parameters() if q.grad is not None}
            opt.step()                                   # eager, outside the graph
            seen.append(loss)
        assert seen[2] < seen[1] < seen[0], seen         # the replays saw the new weights
        loss = float(step(src_seq, src_pos, tgt))
        g_graph = {n: q.grad.clone() for n, q in model.named_parameters() if q.grad is not None}
        ops.TRAIN_SEED_DEV = None
        model.zero_grad(set_to_none=True)
        l2 = loss_fn(model((src_seq.to(DEV), src_pos.to(DEV)), None, None, None)[0], tgt.to(DEV))
        l2.backward()
        assert abs(loss - float(l2)) < 1e-6 * max(1.0, abs(loss))
        for n, q in model.named_parameters():
            if q.grad is not None:
                assert rel_err(g_graph[n], q.grad) < 1e-5, n
    finally:
        ops.TRAIN_SEED_DEV = None
