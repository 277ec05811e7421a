"""GPU tests of the drop-in boundary as the reference drives it (round 2): the eval CUDA-graph cache inside
``LAMP.forward``, weight-plane freshness across optimizer steps and captured graphs, the flat-buffer gradient reducer,
the reference's own ``main.py`` epoch through ``lamp_b200.compat`` and (on >= 2 GPUs) ``nn.DataParallel`` replicas."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases
import lamp_b200
from lamp_b200 import distributed as lds
from lamp_b200 import graphs, ops
from lamp_b200.Models import LAMP
from oracle import lamp_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
DEV = 'cuda'
TOL = 1e-3


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build_model(c, p, adj, dev=DEV, dropout=0.2):
    d = c['D'] // c['H']
    m = LAMP(c['V'] + 4, c['L'], c['T'], c['L'], n_layers_enc=c['n_enc'], n_layers_dec=c['n_dec'], n_head=c['H'],
             n_head2=c['H'], d_word_vec=c['D'], d_model=c['D'], d_inner_hid=c['d_inner'], d_k=d, d_v=d, dropout=dropout,
             dec_dropout=dropout, dec_dropout2=False, proj_share_weight=True, encoder='graph', decoder='graph',
             enc_transform=c.get('enc_transform', ''), no_enc_pos_embedding=not c.get('pos_enc', True),
             label_adj_matrix=adj, label_mask=c['mask'])
    m.load_state_dict(p, strict=True)
    return m.to(dev).eval()


def oracle_logits(model, c, cfg, src_seq, src_pos, adj):
    p = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    lm = orc.label_mask_from(c['L'], adj, c['mask'])
    return orc.lamp_forward(p, cfg, src_seq, src_pos, lm, compute_dead_attention=False)[0]


def test_eval_graph_cache_replays_equal_eager_and_track_shapes():
    """test.py:41 calls model(src, adj, None, None) batch after batch with the length of the longest document of each
    batch: the first call of a (batch, length bucket) is eager, the second captures, the rest replay; logits and
    enc_output equal the plain launch sequence for every length in the bucket."""
    c = cases.MODEL_CASES['lamp_L37_none']
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    model = build_model(c, p, adj)
    assert graphs.EVAL_GRAPHS
    for T in (64, 50, 64, 40, 33, 64, 17, 20, 30):
        seq, pos = src_seq[:, :T].clone(), src_pos[:, :T].clone()
        seq[0, T - 1], pos[0, T - 1] = 5, T        # one full-length row, as pad_to_longest guarantees
        src = (seq.to(DEV), pos.to(DEV))
        with torch.no_grad():
            want, want_enc, _ = model._forward_impl(src, None, None, None)
        got, got_enc, none = model(src, None, None, None)      # grad enabled + eval, exactly as test.py runs it
        assert none is None and got.shape == want.shape and got_enc.shape == want_enc.shape == (c['B'], T, c['D'])
        assert rel_err(got, want) < 1e-6 and rel_err(got_enc, want_enc) < 1e-6, T
        lm = orc.label_mask_from(c['L'], adj, c['mask'])
        ref, _ = orc.lamp_forward(p, cfg, seq, pos, lm, compute_dead_attention=False)
        assert rel_err(got, ref) < TOL
    cache = model.__dict__['_eval_graphs']
    assert cache.captures == 2 and cache.replays == 7 and cache.eager_calls == 2, (cache.captures, cache.replays, cache.eager_calls)   # length buckets 64 and 32
    # outputs are copies: a later replay must not overwrite what an earlier call returned
    a, _, _ = model((src_seq.to(DEV), src_pos.to(DEV)), None, None, None)
    keep = a.clone()
    seq2 = src_seq.clone()
    seq2[seq2 > 0] = 7
    model((seq2.to(DEV), src_pos.to(DEV)), None, None, None)
    assert torch.equal(a, keep)


def test_eval_forward_sees_new_weights_after_training_eager_and_captured():
    """the weight planes are cached; a captured eval forward and a captured Adam step must not let
    an eval after training run on old planes.  Sequence: eval (captures) -> eager Adam steps -> eval -> graphed
    train+Adam replays (no tensor version bumps) -> eval; every eval is checked against the oracle on the CURRENT
    state dict, and differs from the previous one."""
    c = dict(cases.MODEL_CASES['lamp_L37_none'])
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    model = build_model(c, p, adj, dropout=0.0)
    src = (src_seq.to(DEV), src_pos.to(DEV))
    gold = (torch.rand(c['B'], c['L'], generator=torch.Generator().manual_seed(3)) < 0.3).float().to(DEV)
    loss_fn = torch.nn.functional.binary_cross_entropy_with_logits

    def eval_logits():
        model.eval()
        out = [model(src, None, None, None)[0] for _ in range(3)]   # eager, capture, replay
        assert torch.equal(out[0], out[1]) and torch.equal(out[1], out[2])
        assert rel_err(out[2], oracle_logits(model, c, cfg, src_seq, src_pos, adj)) < TOL
        return out[2]

    l0 = eval_logits()
    # (1) eager optimizer steps bump tensor versions
    model.train()
    opt = torch.optim.Adam(model.get_trainable_parameters(), lr=1e-2)
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        loss_fn(model(src, None, None, gold)[0], gold).backward()
        opt.step()
    l1 = eval_logits()
    assert rel_err(l1, l0) > 1e-2
    # (2) a capturable optimizer inside the training graph: replays change the weights without version bumps
    model.train()
    opt2 = torch.optim.Adam(model.get_trainable_parameters(), lr=1e-2, capturable=True)
    step = lamp_b200.GraphedTrainStep(model, loss_fn, c['B'], c['T'], example=(src[0], src[1], gold), optimizer=opt2)
    for _ in range(3):
        step(src[0], src[1], gold)
    l2 = eval_logits()
    assert rel_err(l2, l1) > 1e-3
    for _ in range(2):
        step(src[0], src[1], gold)
    l3 = eval_logits()
    assert rel_err(l3, l2) > 1e-4
    ops.TRAIN_SEED_DEV = None
    # (3) the explicit serving graph follows load_state_dict as well
    model.eval()
    runner = lamp_b200.GraphedForward(model, c['B'], c['T'], example=src)
    model.load_state_dict({k: v.to(DEV) for k, v in p.items()}, strict=True)
    back, _ = runner(src[0], src[1])
    assert rel_err(back, l0) < 1e-6


def test_gradient_reducer_single_gpu_flat_views_and_fused_adam():
    """world = 1: gradients accumulate into the flat buffer through views, dead parameters keep grad None, and a fused
    Adam step on the views equals the same step on ordinary gradients."""
    c = dict(cases.MODEL_CASES['lamp_L37_none'])
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    src = (src_seq.to(DEV), src_pos.to(DEV))
    gold = (torch.rand(c['B'], c['L'], generator=torch.Generator().manual_seed(3)) < 0.3).float().to(DEV)
    loss_fn = torch.nn.functional.binary_cross_entropy_with_logits
    models = [build_model(c, p, adj, dropout=0.0).train() for _ in range(2)]
    opts = [torch.optim.Adam(m.get_trainable_parameters(), lr=1e-3, fused=True) for m in models]
    red = lds.GradientReducer(models[0].get_trainable_parameters(), world=1)
    managed = {id(q) for q in red.params}   # the frozen sinusoid table is not a trainable parameter (Models.py:97-107)
    for step in range(3):
        red.zero_grad()
        loss_fn(models[0](src, None, None, gold)[0], gold).backward()
        red.finish()
        opts[1].zero_grad(set_to_none=True)
        loss_fn(models[1](src, None, None, gold)[0], gold).backward()
        for (n, a), (_, b) in zip(models[0].named_parameters(), models[1].named_parameters()):
            assert (a.grad is None) == (b.grad is None), n
            if a.grad is not None:
                if id(a) in managed:
                    assert a.grad.untyped_storage().data_ptr() == red.flat.untyped_storage().data_ptr(), n
                assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-6), n   # (fp32 `red` order differs run to run)
                b.grad.copy_(a.grad)   # Adam's first steps turn last-bit gradient differences into +-lr: same inputs
        opts[0].step()
        opts[1].step()
        for (n, a), (_, b) in zip(models[0].named_parameters(), models[1].named_parameters()):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-8), n
    dead = [n for n, q in models[0].named_parameters() if q.grad is None and q.requires_grad]
    assert dead and all('encoder.layer_stack' in n and 'slf_attn' in n for n in dead), dead
    assert red.stats['elements'] == sum(q.numel() for q in models[0].get_trainable_parameters() if q.grad is not None)


def _run_main_epoch(tmp_path, arms, gpus=None):
    if not os.path.exists(os.path.join(ROOT, 'baseline', '_ref', 'main.py')):
        pytest.skip('baseline/_ref did not travel with this snapshot (run __graft_entry__.build() in the build container)')
    cmd = [sys.executable, os.path.join(ROOT, 'scripts', 'run_reference_main.py'), '--tiny', '--arms', arms, '--out',
           str(tmp_path / 'epoch')] + (['--gpus', gpus] if gpus is not None else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads(lines[-1]), r


def test_reference_main_py_runs_one_epoch_on_the_gpu_through_compat(tmp_path):
    """The reference's UNMODIFIED main.py (train epoch + validation + test + metrics + checkpoint) with the label-graph
    classes rebound to lamp_b200, next to the same main.py as stock PyTorch: both finish, losses are finite and close
    (different dropout streams, same data / init seeds are not shared -> only a loose agreement is asserted)."""
    s, r = _run_main_epoch(tmp_path, 'dropin,reference', gpus='0')
    assert s['dropin']['returncode'] == 0, s['dropin'].get('error_tail')
    assert s['reference']['returncode'] == 0, s['reference'].get('error_tail')
    for arm in ('dropin', 'reference'):
        for k in ('training_bce_per_doc', 'validation_bce_per_doc', 'testing_bce_per_doc'):
            assert np.isfinite(s[arm][k]) and 0 < s[arm][k] < 1.0, (arm, k, s[arm])
    assert abs(s['dropin']['testing_bce_per_doc'] - s['reference']['testing_bce_per_doc']) < 0.05
    log = open(tmp_path / 'epoch' / 'dropin.log').read()
    assert 'using prior mask' in log


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (nn.DataParallel replicas)')
def test_data_parallel_replicas_match_single_gpu(tmp_path):
    """config_args.py:82 forces multi_gpu=True, so on a multi-GPU node main.py wraps the model in nn.DataParallel
    (main.py:106-108): replicas are shallow copies sharing the weight-plane caches.  Eval logits and training gradients
    through DataParallel must equal the single-GPU ones, repeatedly (cache thrash / races would show up as drift)."""
    c = dict(cases.MODEL_CASES['lamp_L37_none'], B=6)
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    single = build_model(c, p, adj, dev='cuda:0', dropout=0.0)
    dp = torch.nn.DataParallel(build_model(c, p, adj, dev='cuda:0', dropout=0.0), device_ids=[0, 1])
    src = (src_seq.to('cuda:0'), src_pos.to('cuda:0'))
    gold = (torch.rand(c['B'], c['L'], generator=torch.Generator().manual_seed(3)) < 0.3).float().to('cuda:0')
    for _ in range(3):
        with torch.no_grad():
            want = single._forward_impl(src, None, None, None)[0]
        got, enc, none = dp(src, None, None, None)
        assert none is None and got.shape == want.shape
        assert rel_err(got, want) < 1e-6
    single.train()
    dp.train()
    loss_fn = torch.nn.functional.binary_cross_entropy_with_logits
    for it in range(2):
        single.zero_grad(set_to_none=True)
        dp.zero_grad(set_to_none=True)
        loss_fn(single(src, None, None, gold)[0], gold).backward()
        loss_fn(dp(src, None, None, gold)[0], gold).backward()
        for (n, a), (_, b) in zip(dp.module.named_parameters(), single.named_parameters()):
            if b.grad is None:
                # dead parameters: DataParallel's Broadcast backward hands back zeros where a single GPU leaves None
                assert a.grad is None or float(a.grad.abs().max()) == 0.0, n
            else:
                assert rel_err(a.grad, b.grad) < 1e-4, (it, n)
    # ... and the whole main.py epoch through DataParallel
    s, r = _run_main_epoch(tmp_path, 'dropin', gpus='0,1')
    assert s['dropin']['returncode'] == 0 and s['dropin']['data_parallel'], s['dropin'].get('error_tail')
    assert np.isfinite(s['dropin']['testing_bce_per_doc'])


def _gold_rows(B, W, L, seed):
    """Rows as the reference loader builds them after `tgt[:, 1:]`: label ids + 4, then EOS = 3, then PAD = 0."""
    rs = np.random.RandomState(seed)
    g = np.zeros((B, W), dtype=np.int64)
    for b in range(B):
        k = rs.randint(0, min(W - 1, 6) + 1)
        labs = rs.choice(L, size=k, replace=False) + 4
        g[b, :k] = labs
        g[b, k] = 3
    g[0, :] = 0                       # a row with nothing at all
    g[1, :] = 0
    g[1, 0] = 3                       # only the EOS
    g[2, :3] = [7, 7, 3]              # a repeated label
    return torch.from_numpy(g)


def test_gold_binary_kernel_matches_reference_loop():
    """utils/utils.py:205-216 restated inline (entries > 0, minus the last one, index_fill, drop 4 columns)."""
    for (B, W, L) in ((37, 9, 103), (5, 40, 12), (300, 64, 983)):
        gold = _gold_rows(B, W, L, seed=B)
        want = torch.zeros(B, L + 4)
        for i in range(B):
            idx = gold[i][gold[i] > 0][0:-1]
            if len(idx) > 0:
                want[i].index_fill_(0, idx, 1)
        want = want[:, 4:]
        got = ops.gold_binary(gold.to(DEV), L)
        assert got.shape == (B, L) and torch.equal(got.cpu(), want), (B, W, L)
        assert torch.equal(lamp_b200.utils.get_gold_binary(gold, L).cpu(), want)   # the drop-in entry point (CPU input)


def test_bce_with_logits_kernel_matches_torch_value_and_gradient():
    g = torch.Generator().manual_seed(5)
    for shape in ((32, 103), (256, 103), (7, 983), (1100, 103)):
        x = (torch.randn(shape, generator=g) * 4).to(DEV).requires_grad_(True)
        y = (torch.rand(shape, generator=g) < 0.1).float().to(DEV)
        want = torch.nn.functional.binary_cross_entropy_with_logits(x.double(), y.double())
        (gw,) = torch.autograd.grad(want, x)
        for _ in range(2):   # second call: the ticket counter was re-armed by the kernel
            got = ops.bce_with_logits(x, y)
            (gg,) = torch.autograd.grad(got * 3.0, x)
            g_val, w_val = float(got.detach()), float(want.detach())
            assert abs(g_val - w_val) < 1e-6 * max(1.0, abs(w_val)), shape
            assert rel_err(gg, gw * 3.0) < 1e-5, shape


def test_deepcopy_after_graph_capture_gives_an_independent_working_model():
    """train.py:45 deep-copies the model mid-training; with captured eval graphs and filled weight-plane caches the copy
    must work on its own (fresh caches) and follow ITS parameters."""
    import copy
    c = cases.MODEL_CASES['lamp_L37_none']
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    model = build_model(c, p, adj)
    src = (src_seq.to(DEV), src_pos.to(DEV))
    a = [model(src, None, None, None)[0] for _ in range(3)][-1]       # eager, capture, replay
    twin = copy.deepcopy(model)
    assert '_eval_graphs' not in twin.__dict__
    b = [twin(src, None, None, None)[0] for _ in range(3)][-1]
    assert torch.equal(a, b)
    with torch.no_grad():
        for q in twin.parameters():
            q.mul_(1.01)
    b2 = twin(src, None, None, None)[0]
    a2 = model(src, None, None, None)[0]
    assert torch.equal(a2, a) and rel_err(b2, a) > 1e-4
