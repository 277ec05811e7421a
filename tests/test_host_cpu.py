"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares, the host-side mirror of
the reference interface (signatures, state-dict keys, mask construction) and the 'no CPU path' contract."""
import ctypes
import inspect
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases
import lamp_b200
from lamp_b200 import _native as nat
from lamp_b200 import synthetic as syn
from lamp_b200 import utils as lutils
from lamp_b200.Decoders import GraphDecoder
from lamp_b200.Encoders import GraphEncoder
from lamp_b200.Layers import DecoderLayer, EncoderLayer
from lamp_b200.Models import LAMP
from lamp_b200.SubLayers import MultiHeadAttention, PositionwiseFeedForward, ScaledDotProductAttention
from oracle import lamp_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    from lamp_b200 import build
    return build.build()


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'lamp_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(lamp_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_header_symbol(built):
    syms = header_symbols()
    assert len(syms) >= 15
    lib = ctypes.CDLL(built)
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/lamp_b200.h but not exported'
    assert set(syms) == set(nat.EXPORTED_SYMBOLS)  # the ctypes binding covers the whole header
    assert nat.lib().lamp_version() >= 100


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/lamp_b200.h against the ctypes argtypes of lamp_b200/_native.py: same argument
    count, and pointer / integer / float / size_t arguments in the same positions (a mismatch here only shows up as a
    crash or garbage on the GPU box)."""
    import ctypes as C
    text = open(os.path.join(ROOT, 'include', 'lamp_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = re.findall(r'\b(?:int|size_t|const char\*)\s+(lamp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S)
    assert len(protos) == len(nat.EXPORTED_SYMBOLS)

    def kind_of_c(arg: str) -> str:
        arg = ' '.join(arg.split())
        if '*' in arg:
            return 'ptr'
        if arg.startswith('float'):
            return 'float'
        if arg.startswith('size_t'):
            return 'size'
        if arg.startswith('uint64_t') or arg.startswith('int64_t'):
            return 'i64'
        if arg.startswith('int'):
            return 'int'
        raise AssertionError(f'unrecognised C argument: {arg}')

    def kind_of_ct(t) -> str:
        return {C.c_void_p: 'ptr', C.c_char_p: 'ptr', C.c_float: 'float', C.c_size_t: 'size', C.c_int64: 'i64',
                C.c_uint64: 'i64', C.c_int: 'int'}[t]

    for name, args in protos:
        args = args.strip()
        c_kinds = [] if args in ('', 'void') else [kind_of_c(a) for a in args.split(',')]
        ct_kinds = [kind_of_ct(t) for t in nat._SIGNATURES[name][0]]
        # size_t and 64-bit integers share a register class on this ABI; everything else must agree exactly
        norm = lambda ks: ['i64' if k == 'size' else k for k in ks]
        assert norm(c_kinds) == norm(ct_kinds), f'{name}: header {c_kinds} vs ctypes {ct_kinds}'


def test_library_is_built_for_sm100a_with_tcgen05_and_tma(built):
    out = subprocess.run(['cuobjdump', '-lelf', built], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    sass = subprocess.run(['cuobjdump', '-sass', built], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'LDTM'):
        assert mnemonic in sass, mnemonic


def test_no_gpu_means_error_code_not_crash(built):
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert nat.lib().lamp_device_check() < 0
    assert nat.lib().lamp_last_error()


# constructor / forward parameter names of the reference (lamp/SubLayers.py:17,27,47,77,126,135; lamp/Layers.py:10,15,
# 23,34; lamp/Decoders.py:97-101,127; lamp/Encoders.py:32-35,64; lamp/Models.py:19-25,110)
REF_SIGNATURES = {
    (ScaledDotProductAttention, '__init__'): ['temperature', 'dropout', 'attn_type'],
    (ScaledDotProductAttention, 'forward'): ['q', 'k', 'v', 'attn_mask', 'stop_sig'],
    (MultiHeadAttention, '__init__'): ['n_head', 'd_model', 'd_k', 'd_v', 'dropout', 'dropout2', 'attn_type'],
    (MultiHeadAttention, 'forward'): ['q', 'k', 'v', 'attn_mask', 'dec_self'],
    (PositionwiseFeedForward, '__init__'): ['d_in', 'd_hid', 'dropout'],
    (PositionwiseFeedForward, 'forward'): ['x'],
    (EncoderLayer, '__init__'): ['d_model', 'd_inner_hid', 'n_head', 'd_k', 'd_v', 'dropout'],
    (EncoderLayer, 'forward'): ['enc_input', 'slf_attn_mask'],
    (DecoderLayer, '__init__'): ['d_model', 'd_inner_hid', 'n_head', 'n_head2', 'd_k', 'd_v', 'dropout', 'dropout2',
                                 'no_dec_self_att', 'ffn', 'attn_type'],
    (DecoderLayer, 'forward'): ['dec_input', 'enc_output', 'slf_attn_mask', 'dec_enc_attn_mask'],
    (GraphDecoder, '__init__'): ['n_tgt_vocab', 'n_max_seq', 'n_layers', 'n_head', 'n_head2', 'd_k', 'd_v',
                                 'd_word_vec', 'd_model', 'd_inner_hid', 'dropout', 'dropout2', 'no_dec_self_att',
                                 'label_adj_matrix', 'label_mask', 'enc_vec', 'graph_conv', 'attn_type'],
    (GraphDecoder, 'forward'): ['tgt', 'src_seq', 'enc_output', 'return_attns', 'int_preds'],
    (GraphEncoder, '__init__'): ['n_src_vocab', 'n_max_seq', 'n_layers', 'n_head', 'd_k', 'd_v', 'd_word_vec',
                                 'd_model', 'd_inner_hid', 'onehot', 'enc_transform', 'dropout',
                                 'no_enc_pos_embedding'],
    (GraphEncoder, 'forward'): ['src_seq', 'adj', 'src_pos', 'return_attns'],
    (LAMP, '__init__'): ['n_src_vocab', 'n_tgt_vocab', 'n_max_seq_e', 'n_max_seq_d', 'n_layers_enc', 'n_layers_dec',
                         'n_head', 'n_head2', 'd_word_vec', 'd_model', 'd_inner_hid', 'd_k', 'd_v', 'dropout',
                         'dec_dropout', 'dec_dropout2', 'proj_share_weight', 'embs_share_weight', 'encoder',
                         'decoder', 'enc_transform', 'onehot', 'no_enc_pos_embedding', 'no_dec_self_att', 'loss',
                         'label_adj_matrix', 'label_mask', 'matching_mlp', 'graph_conv', 'attn_type', 'int_preds'],
    (LAMP, 'forward'): ['src', 'adj', 'tgt_seq', 'binary_tgt', 'return_attns', 'int_preds'],
}


@pytest.mark.parametrize('key', list(REF_SIGNATURES), ids=lambda k: f'{k[0].__name__}.{k[1]}')
def test_signatures_mirror_reference(key):
    cls, meth = key
    names = [n for n in inspect.signature(getattr(cls, meth)).parameters if n != 'self']
    want = REF_SIGNATURES[key]
    assert names[:len(want)] == want       # same names, same order
    extra = names[len(want):]              # our additions must be optional keyword flags
    params = inspect.signature(getattr(cls, meth)).parameters
    assert all(params[n].default is not inspect.Parameter.empty for n in extra)


def test_signatures_match_the_reference_source_when_available():
    ref = '/root/reference'
    if not os.path.isdir(ref):
        pytest.skip('reference checkout not present (GPU box)')
    sys.path.insert(0, ref)
    try:
        import lamp.SubLayers as rs, lamp.Layers as rl, lamp.Decoders as rd, lamp.Encoders as re_, lamp.Models as rm
        pairs = [(rs.MultiHeadAttention, MultiHeadAttention), (rs.ScaledDotProductAttention, ScaledDotProductAttention),
                 (rs.PositionwiseFeedForward, PositionwiseFeedForward), (rl.DecoderLayer, DecoderLayer),
                 (rl.EncoderLayer, EncoderLayer), (rd.GraphDecoder, GraphDecoder), (re_.GraphEncoder, GraphEncoder),
                 (rm.LAMP, LAMP)]
        for rcls, ncls in pairs:
            for meth in ('__init__', 'forward'):
                rp = inspect.signature(getattr(rcls, meth)).parameters
                np_ = inspect.signature(getattr(ncls, meth)).parameters
                rnames = list(rp)
                assert list(np_)[:len(rnames)] == rnames, (rcls, meth)
                for n in rnames:
                    rdft, ndft = rp[n].default, np_[n].default
                    assert (rdft is inspect.Parameter.empty) == (ndft is inspect.Parameter.empty), (rcls, meth, n)
                    if rdft is not inspect.Parameter.empty:
                        assert rdft == ndft, (rcls.__name__, meth, n, rdft, ndft)
    finally:
        sys.path.remove(ref)
        for m in [m for m in sys.modules if m == 'lamp' or m.startswith('lamp.')]:
            del sys.modules[m]


def make_model(c, p, adj):
    d = c['D'] // c['H']
    m = LAMP(c['V'] + 4, c['L'], c['T'], c['L'], n_layers_enc=c['n_enc'], n_layers_dec=c['n_dec'], n_head=c['H'],
             n_head2=c['H'], d_word_vec=c['D'], d_model=c['D'], d_inner_hid=c['d_inner'], d_k=d, d_v=d,
             encoder='graph', decoder='graph', enc_transform=c.get('enc_transform', ''),
             no_enc_pos_embedding=not c.get('pos_enc', True), label_adj_matrix=adj, label_mask=c['mask'])
    return m


def test_state_dict_keys_and_strict_load():
    c = cases.MODEL_CASES['lamp_L103_prior']
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    m = make_model(c, p, adj)
    want = {}
    with open(os.path.join(ROOT, 'tests', 'golden', 'state_keys_lamp_L103_prior.txt')) as f:
        for line in f:
            k, shape = line.strip().split(' ', 1)
            want[k] = eval(shape)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == want
    m.load_state_dict(p, strict=True)
    # the alias parameter is the label-embedding parameter (lamp/Models.py:89); forward uses .linear.weight
    assert m.tgt_word_proj.weight is m.decoder.tgt_word_emb.weight
    frozen = {id(q) for q in m.encoder.position_enc.parameters()}
    assert all(id(q) not in frozen for q in m.get_trainable_parameters())
    with pytest.raises(NotImplementedError):
        LAMP(10, 5, 8, 5, encoder='rnn', decoder='graph')
    with pytest.raises(NotImplementedError):
        LAMP(10, 5, 8, 5, encoder='graph', decoder='sa_m')


@pytest.mark.parametrize('kind', ['prior', 'inveye', 'none', 'diagrow'])
def test_label_mask_matches_oracle(kind):
    L = 23
    adj = cases.label_adj(kind, L, 5)
    adj_copy = None if adj is None else adj.clone()
    dec = GraphDecoder(L, L, n_layers=1, n_head=2, n_head2=2, d_k=8, d_v=8, d_word_vec=16, d_model=16,
                       d_inner_hid=32, label_adj_matrix=adj, label_mask=kind if adj is None else 'prior',
                       enc_vec=False)
    want = orc.label_mask_from(L, adj_copy, kind)
    if want is None:
        assert dec.label_mask is None and dec._label_mask_dev is None
        return
    assert torch.equal(dec._label_mask_dev, want)
    assert dec.label_mask.reshape(L, L).ne(0).equal(want)      # reference-style float attribute kept
    assert '_label_mask_dev' not in dec.state_dict()           # plain attribute in the reference: not a state key
    if adj is not None:
        assert torch.equal(adj, adj_copy)                      # caller's adjacency untouched
        if kind == 'diagrow':
            assert not want[3, 3] and want[3].sum() == L - 1   # empty row -> forced self edge (Decoders.py:109-112)


def test_utils_match_oracle():
    a = lutils.position_encoding_init(41, 64)
    b = orc.position_encoding_init(41, 64)
    assert torch.equal(a, b) and a.dtype == torch.float32 and bool((a[0] == 0).all())
    seq_q = torch.tensor([[5, 6, 0], [7, 0, 0]])
    seq_k = torch.tensor([[5, 0, 0, 9], [0, 1, 2, 0]])
    m = lutils.get_attn_padding_mask(seq_q, seq_k)
    assert torch.equal(m, orc.padding_mask(seq_q, seq_k)) and m.stride(1) == 0
    t = torch.tensor([[0., 2.], [3., 0.]])
    assert torch.equal(lutils.swap_0_1(t, 1, 0), torch.tensor([[1., 0.], [0., 1.]]))


def test_no_cpu_path():
    x = torch.zeros(2, 5, 32)
    for mod, args in ((MultiHeadAttention(2, 32, 16, 16).eval(), (x, x, x)),
                      (PositionwiseFeedForward(32, 64).eval(), (x,)),
                      (ScaledDotProductAttention(4.0).eval(), (x, x, x))):
        with pytest.raises(RuntimeError, match='no CPU path'):
            mod(*args)


def test_package_does_not_import_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, 'lamp_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f


def test_synthetic_generators_are_deterministic():
    a = syn.make_tokens(4, 30, 100, seed=3)
    b = syn.make_tokens(4, 30, 100, seed=3)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert int(a[0][0].ne(0).sum()) == 30 and bool(((a[0] == 0) == (a[1] == 0)).all())
    rows = syn.make_label_sets(11, seed=1)
    assert {l - 4 for r in rows for l in r[1:-1]} == set(range(11))     # every label occurs
    adj = syn.prior_adjacency(rows, 11)
    assert torch.equal(adj, adj.T)


def test_reference_main_py_reaches_the_fused_forward(tmp_path):
    """Drop-in check of the boundary: the reference's UNMODIFIED main.py, with lamp_b200 rebound into it
    (lamp_b200.compat), loads a synthetic dataset, builds the model through its own 30-kwarg call and enters the
    train loop; on this GPU-less box the first forward must stop with the 'no CPU path' error."""
    ref = '/root/reference'
    if not os.path.isdir(ref):
        pytest.skip('reference checkout not present (GPU box)')
    if torch.cuda.is_available():
        pytest.skip('meant for the GPU-less build container')
    data = syn.make_dataset_dict(n_labels=12, vocab=50, n_train=40, n_valid=8, n_test=8, max_len=20)
    os.makedirs(tmp_path / 'data' / 'synth')
    torch.save(data, tmp_path / 'data' / 'synth' / 'train_valid_test.pt')
    env = dict(os.environ, PYTHONPATH=ROOT, LAMP_B200_NO_GPU_SHIM='1')
    cmd = [sys.executable, '-m', 'lamp_b200.run_main', ref, '-dataroot', str(tmp_path / 'data') + '/', '-dataset', 'synth',
           '-results_dir', str(tmp_path / 'results') + '/', '-batch_size', '8', '-d_model', '32', '-d_inner_hid', '32',
           '-n_layers_enc', '1', '-n_layers_dec', '1', '-n_head', '2', '-epoch', '1', '-encoder', 'graph', '-decoder',
           'graph', '-label_mask', 'prior', '-no_cuda', '-overwrite']
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    out = r.stdout + r.stderr
    assert 'using prior mask' in out, out[-2000:]
    assert 'no CPU path' in out, out[-3000:]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) runs without a GPU and prints exactly
    one JSON line with the contract's keys."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                        '--cpu-batch', '2'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['value'] > 0
    have_ref = os.path.exists(os.path.join(ROOT, 'baseline', '_ref', 'lamp', 'Models.py')) or os.path.isdir('/root/reference')
    assert d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port')
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_get_gold_binary_host_semantics_equal_the_reference():
    """lamp_b200.utils.get_gold_binary (device kernel on a GPU box; this host restatement elsewhere) against the
    reference's own utils.utils.get_gold_binary (utils/utils.py:205-216) on rows incl. empty / EOS-only ones."""
    import importlib
    import numpy as np
    ref = None
    for d in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if os.path.exists(os.path.join(d, 'utils', 'utils.py')):
            ref = d
            break
    if ref is None:
        pytest.skip('reference tree not present')
    if torch.cuda.is_available():
        pytest.skip('host semantics are exercised on the GPU-less container; the GPU suite checks the kernel')
    sys.path.insert(0, ref)
    try:
        for m in [k for k in sys.modules if k == 'utils' or k.startswith('utils.')]:
            del sys.modules[m]
        ru = importlib.import_module('utils.utils')
        from lamp_b200 import utils as lu
        rs = np.random.RandomState(0)
        B, W, L = 24, 9, 30
        g = np.zeros((B, W), dtype=np.int64)
        for b in range(B):
            k = rs.randint(0, 6)
            g[b, :k] = rs.choice(L, size=k, replace=False) + 4
            g[b, k] = 3
        g[0, :] = 0
        g[1, :] = 0
        g[1, 0] = 3
        g = torch.from_numpy(g)
        assert torch.equal(ru.get_gold_binary(g, L), lu.get_gold_binary(g, L))
    finally:
        sys.path.remove(ref)
        for m in [k for k in sys.modules if k == 'utils' or k.startswith('utils.')]:
            del sys.modules[m]


def test_model_deepcopy_and_pickle_drop_runtime_caches():
    """train.py:45 runs copy.deepcopy(model) on every step of epoch `thresh1`: run-time caches (weight-plane caches with
    their locks, the eval CUDA-graph cache) must not break it, and the copy must start with empty caches."""
    import copy
    import io
    from lamp_b200 import ops
    m = LAMP(54, 12, 20, 12, n_layers_enc=1, n_layers_dec=1, n_head=2, n_head2=2, d_word_vec=32, d_model=32,
             d_inner_hid=32, d_k=16, d_v=16, encoder='graph', decoder='graph', label_mask='inveye')
    m.__dict__['_eval_graphs'] = object()
    m.__dict__['_lamp_planes_state'] = dict(stamp=1)
    c = copy.deepcopy(m)
    assert '_eval_graphs' not in c.__dict__ and '_lamp_planes_state' not in c.__dict__
    assert isinstance(c.decoder._wp, ops.WeightPlanes) and c.decoder._wp is not m.decoder._wp
    assert c.tgt_word_proj.weight is c.decoder.tgt_word_emb.weight            # the reference's alias survives the copy
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), c.state_dict().values()))
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    assert list(r.state_dict()) == list(m.state_dict())


def test_zero_pool_hands_out_aligned_disjoint_zero_views():
    """ops._ZeroPool (one zero fill per backward node): views have the requested shapes, start on 16-byte boundaries,
    do not overlap, and must be taken in the declared order."""
    from lamp_b200 import ops
    shapes = [(5,), (5,), (3, 7), (3,), (7, 3), (1,)]
    zp = ops._ZeroPool(torch.device('cpu'), *shapes)
    views = [zp.take(sh) for sh in shapes]
    base = views[0].data_ptr()
    spans = []
    for v, sh in zip(views, shapes):
        assert tuple(v.shape) == sh and v.dtype == torch.float32 and float(v.abs().sum()) == 0.0
        assert (v.data_ptr() - base) % 16 == 0
        spans.append((v.data_ptr(), v.data_ptr() + v.numel() * 4))
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    views[2].fill_(1.0)
    assert float(views[1].sum()) == 0.0 and float(views[3].sum()) == 0.0
    zp2 = ops._ZeroPool(torch.device('cpu'), (4,), (2, 2))
    with pytest.raises(AssertionError):
        zp2.take((2, 2))
