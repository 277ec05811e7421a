"""GPU parity tests of the native kernels, called through the C ABI (ctypes), against the CPU oracle and the
reference-generated golden fixtures.  Tolerance for the fp32 path: 1e-3 relative (north_star); the observed
error is reported and is expected to sit two orders of magnitude below that."""
import os

import numpy as np
import pytest
import torch

import cases
from lamp_b200 import _native as nat
from oracle import lamp_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
DEV = 'cuda'


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def ws(nbytes):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=DEV)


def planes(x: torch.Tensor, three=True):
    x = x.contiguous()
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    hi = torch.empty(rows, cols, dtype=torch.bfloat16, device=DEV)
    lo = torch.empty_like(hi) if three else None
    nat.check(nat.lib().lamp_split_planes(x.data_ptr(), rows, cols, cols, hi.data_ptr(), nat.ptr(lo), cols,
                                          nat.stream()), 'split')
    return hi, lo


def test_library_loads_on_gpu():
    assert nat.lib().lamp_device_check() == 0
    assert nat.lib().lamp_sm_count() > 0


def test_split_planes_roundtrip():
    x = torch.randn(37, 64, device=DEV) * 3
    hi, lo = planes(x)
    ref_hi = x.to(torch.bfloat16)
    assert torch.equal(hi, ref_hi)
    ref_lo = (x - ref_hi.float()).to(torch.bfloat16)
    assert torch.equal(lo, ref_lo)
    assert rel_err(hi.float() + lo.float(), x) < 2 ** -15


@pytest.mark.parametrize('M,N,K,prec', [
    (128, 128, 64, 0), (300, 128, 64, 0), (1000, 512, 512, 0), (257, 1536, 512, 0), (103, 64, 128, 0),
    (4096, 1024, 512, 0), (515, 192, 72, 0), (1000, 512, 512, 1), (300, 128, 64, 1), (5000, 256, 1024, 0),
])
@pytest.mark.parametrize('bk,pair', [(0, 1), (32, 1), (64, 1), (32, 0), (64, 0)])
def test_gemm_planes(M, N, K, prec, bk, pair):
    nat.check(nat.lib().lamp_set_tuning(1, bk), 'tune')
    nat.check(nat.lib().lamp_set_tuning(2, pair), 'tune')
    try:
        _gemm_case(M, N, K, prec)
    finally:
        nat.check(nat.lib().lamp_set_tuning(1, 0), 'tune')
        nat.check(nat.lib().lamp_set_tuning(2, 1), 'tune')


def _gemm_case(M, N, K, prec):
    g = torch.Generator(device='cpu').manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    three = prec == 0
    a_hi, a_lo = planes(a, three)
    w_hi, w_lo = planes(w, three)
    out = torch.full((M, N), float('nan'), device=DEV)
    o_hi = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    o_lo = torch.empty_like(o_hi) if three else None
    L = nat.lib()
    # plain product
    nat.check(L.lamp_gemm_planes(a_hi.data_ptr(), nat.ptr(a_lo), K, w_hi.data_ptr(), nat.ptr(w_lo), K, M, N, K, prec,
                                 None, 0, None, 0, 0, out.data_ptr(), N, None, None, 0, None, nat.stream()), 'gemm')
    torch.cuda.synchronize()
    if three:
        ref = a.double() @ w.double().T
        tol = 2e-5
    else:
        ref = a.to(torch.bfloat16).double() @ w.to(torch.bfloat16).double().T
        tol = 1e-5
    e = rel_err(out, ref)
    print(f'gemm {M}x{N}x{K} prec={prec}: rel err {e:.2e}')
    assert e < tol
    # bias + relu + residual, fp32 and plane outputs
    nat.check(L.lamp_gemm_planes(a_hi.data_ptr(), nat.ptr(a_lo), K, w_hi.data_ptr(), nat.ptr(w_lo), K, M, N, K, prec,
                                 bias.data_ptr(), 1, res.data_ptr(), N, 0, out.data_ptr(), N, o_hi.data_ptr(),
                                 nat.ptr(o_lo), N, None, nat.stream()), 'gemm-epi')
    torch.cuda.synchronize()
    ref2 = torch.relu(ref + bias.double()) + res.double()
    assert rel_err(out, ref2) < tol
    assert torch.equal(o_hi, out.to(torch.bfloat16))
    if three:
        assert torch.equal(o_lo, (out - o_hi.float()).to(torch.bfloat16))
    # residual broadcast (row % resid_mod)
    mod = 7
    nat.check(L.lamp_gemm_planes(a_hi.data_ptr(), nat.ptr(a_lo), K, w_hi.data_ptr(), nat.ptr(w_lo), K, M, N, K, prec,
                                 None, 0, res.data_ptr(), N, mod, out.data_ptr(), N, None, None, 0, None, nat.stream()),
              'gemm-mod')
    torch.cuda.synchronize()
    ref3 = ref + res.double()[torch.arange(M, device=DEV) % mod]
    assert rel_err(out, ref3) < tol


@pytest.mark.parametrize('M,N,K,pair', [(1000, 512, 512, 1), (103, 512, 128, 1), (777, 384, 256, 1), (2048, 512, 64, 0),
                                         (130, 264, 512, 1), (5000, 512, 1024, 1)])
def test_gemm_fused_layernorm(M, N, K, pair):
    """lamp_gemm_ln_planes: LayerNorm(A W^T + bias + residual) on chip, vs fp64."""
    nat.check(nat.lib().lamp_set_tuning(2, pair), 'tune')
    try:
        g = torch.Generator(device='cpu').manual_seed(M + N + K)
        a = torch.randn(M, K, generator=g).to(DEV)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
        bias = torch.randn(N, generator=g).to(DEV)
        gam = (1 + 0.3 * torch.randn(N, generator=g)).to(DEV)
        bet = torch.randn(N, generator=g).to(DEV)
        L = nat.lib()
        a_hi, a_lo = planes(a)
        w_hi, w_lo = planes(w)
        for mod in (0, 9):
            res = torch.randn(mod if mod else M, N, generator=g).to(DEV) + 3.0   # non-zero mean rows
            out = torch.full((M, N), float('nan'), device=DEV)
            o_hi = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
            o_lo = torch.empty_like(o_hi)
            nat.check(L.lamp_gemm_ln_planes(a_hi.data_ptr(), a_lo.data_ptr(), K, w_hi.data_ptr(), w_lo.data_ptr(), K, M, N,
                                            K, 0, bias.data_ptr(), res.data_ptr(), N, mod, gam.data_ptr(), bet.data_ptr(),
                                            1e-5, out.data_ptr(), N, o_hi.data_ptr(), o_lo.data_ptr(), N, nat.stream()),
                      'gemm_ln')
            torch.cuda.synchronize()
            rr = res.double()[torch.arange(M, device=DEV) % mod] if mod else res.double()
            ref = torch.nn.functional.layer_norm(a.double() @ w.double().T + bias.double() + rr, (N,), gam.double(),
                                                 bet.double(), 1e-5)
            e = rel_err(out, ref)
            print(f'gemm_ln {M}x{N}x{K} pair={pair} mod={mod}: {e:.2e}')
            assert e < 3e-5
            assert torch.equal(o_hi, out.to(torch.bfloat16))
            assert torch.equal(o_lo, (out - o_hi.float()).to(torch.bfloat16))
        rc = L.lamp_gemm_ln_planes(a_hi.data_ptr(), a_lo.data_ptr(), K, w_hi.data_ptr(), w_lo.data_ptr(), K, M, 128, K, 0,
                                   None, None, 0, 0, gam.data_ptr(), bet.data_ptr(), 1e-5, out.data_ptr(), 128, None, None,
                                   0, nat.stream())
        assert rc == -1   # row does not span two accumulator stages -> caller must use the unfused pair
    finally:
        nat.check(nat.lib().lamp_set_tuning(2, 1), 'tune')


@pytest.mark.parametrize('M,N,K,N2,pair', [(1000, 512, 512, 1536, 1), (103, 512, 128, 512, 1), (777, 384, 256, 64, 1),
                                            (2048, 512, 64, 2048, 0), (130, 264, 512, 128, 1), (5000, 128, 1024, 512, 1),
                                            (300, 1024, 256, 256, 1)])
def test_gemm_deferred_layernorm_chain(M, N, K, N2, pair):
    """Deferred LayerNorm: producer GEMM (pre-norm planes + row stats, plain / broadcast / deferred residual) ->
    consumers (A-operand GEMM with folded gamma, lamp_ln_apply incl. gather, lamp_diag_proj_ln), each vs fp64."""
    nat.check(nat.lib().lamp_set_tuning(2, pair), 'tune')
    try:
        g = torch.Generator(device='cpu').manual_seed(M + N + K + N2)
        L = nat.lib()
        a = torch.randn(M, K, generator=g).to(DEV)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
        bias = torch.randn(N, generator=g).to(DEV)
        gam = (1 + 0.3 * torch.randn(N, generator=g)).to(DEV)
        bet = torch.randn(N, generator=g).to(DEV)
        a_hi, a_lo = planes(a)
        w_hi, w_lo = planes(w)
        np_ = L.lamp_gemm_stats_parts(N)
        assert np_ == 2 * ((N + 255) // 256 if N > 128 else 1)

        def producer(res32=None, mod=0, res_planes=None, rstats=None, rgam=None, rbet=None):
            y_hi = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
            y_lo = torch.empty_like(y_hi)
            st = torch.full((M, np_, 2), float('nan'), device=DEV)
            rh, rl = res_planes if res_planes is not None else (None, None)
            nat.check(L.lamp_gemm_planes_rstats(
                a_hi.data_ptr(), a_lo.data_ptr(), K, w_hi.data_ptr(), w_lo.data_ptr(), K, M, N, K, 0, bias.data_ptr(),
                nat.ptr(res32), nat.ptr(rh), nat.ptr(rl), N, mod, nat.ptr(rstats), np_ if rstats is not None else 0, 1e-5,
                nat.ptr(rgam), nat.ptr(rbet), y_hi.data_ptr(), y_lo.data_ptr(), N, st.data_ptr(), None, nat.stream()),
                'rstats')
            torch.cuda.synchronize()
            return y_hi, y_lo, st

        base = a.double() @ w.double().T + bias.double()
        # 1. plain fp32 residual, full and broadcast (row % mod)
        res = torch.randn(M, N, generator=g).to(DEV) + 2.0
        y_hi, y_lo, st = producer(res32=res)
        y_ref = base + res.double()
        assert rel_err(y_hi.float() + y_lo.float(), y_ref) < 3e-5
        s = st.double().sum(1)
        assert rel_err(s[:, 0], y_ref.sum(1)) < 1e-5 and rel_err(s[:, 1], (y_ref ** 2).sum(1)) < 1e-5
        res9 = torch.randn(9, N, generator=g).to(DEV)
        yb_hi, yb_lo, _ = producer(res32=res9, mod=9)
        assert rel_err(yb_hi.float() + yb_lo.float(), base + res9.double()[torch.arange(M, device=DEV) % 9]) < 3e-5
        # 2. plain planes residual
        yp_hi, yp_lo, _ = producer(res_planes=planes(res))
        assert rel_err(yp_hi.float() + yp_lo.float(), y_ref) < 3e-5
        # 3. deferred residual: residual = LayerNorm(y) of a previous producer
        ln_ref = torch.nn.functional.layer_norm(y_ref, (N,), gam.double(), bet.double(), 1e-5)
        y2_hi, y2_lo, st2 = producer(res_planes=(y_hi, y_lo), rstats=st, rgam=gam, rbet=bet)
        e = rel_err(y2_hi.float() + y2_lo.float(), base + ln_ref)
        print(f'deferred residual {M}x{N}x{K}: {e:.2e}')
        assert e < 3e-5
        # 4. materialise (plain and through a gather index)
        out = torch.full((M, N), float('nan'), device=DEV)
        o_hi = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        o_lo = torch.empty_like(o_hi)
        nat.check(L.lamp_ln_apply(y_hi.data_ptr(), y_lo.data_ptr(), st.data_ptr(), np_, gam.data_ptr(), bet.data_ptr(),
                                  1e-5, M, N, None, out.data_ptr(), o_hi.data_ptr(), o_lo.data_ptr(), None, nat.stream()),
                  'ln_apply')
        torch.cuda.synchronize()
        assert rel_err(out, ln_ref) < 3e-5
        assert torch.equal(o_hi, out.to(torch.bfloat16)) and torch.equal(o_lo, (out - o_hi.float()).to(torch.bfloat16))
        idx = torch.randint(0, M, (2 * M + 3,), generator=g).to(DEV)
        outg = torch.empty(idx.numel(), N, device=DEV)
        nat.check(L.lamp_ln_apply(y_hi.data_ptr(), y_lo.data_ptr(), st.data_ptr(), np_, gam.data_ptr(), bet.data_ptr(),
                                  1e-5, idx.numel(), N, idx.data_ptr(), outg.data_ptr(), None, None, None, nat.stream()),
                  'ln_apply-gather')
        torch.cuda.synchronize()
        assert torch.equal(outg, out[idx])
        # 5. consumer GEMM: LN(y) W2^T + b2 (ReLU) with gamma folded into the weight planes
        w2 = (torch.randn(N2, N, generator=g) / N ** 0.5).to(DEV)
        b2 = torch.randn(N2, generator=g).to(DEV)
        wg_hi, wg_lo = planes((w2 * gam.unsqueeze(0)).contiguous())
        colsum = (wg_hi.float() + wg_lo.float()).sum(1).contiguous()
        biasf = (w2.double() @ bet.double() + b2.double()).float().contiguous()
        for relu in (0, 1):
            c_hi = torch.empty(M, N2, dtype=torch.bfloat16, device=DEV)
            c_lo = torch.empty_like(c_hi)
            nat.check(L.lamp_gemm_planes_dln(y_hi.data_ptr(), y_lo.data_ptr(), N, st.data_ptr(), np_, 1e-5,
                                             wg_hi.data_ptr(), wg_lo.data_ptr(), N, colsum.data_ptr(), biasf.data_ptr(), M,
                                             N2, N, 0, relu, c_hi.data_ptr(), c_lo.data_ptr(), N2, None, nat.stream()),
                      'dln')
            torch.cuda.synchronize()
            ref = ln_ref @ w2.double().T + b2.double()
            if relu:
                ref = ref.clamp_min(0)
            e = rel_err(c_hi.float() + c_lo.float(), ref)
            print(f'deferred A {M}x{N2}x{N} relu={relu}: {e:.2e}')
            assert e < 3e-5
        # 6. label projection on the deferred tensor (rows = B*Lb)
        Lb = 7 if M % 7 == 0 else (5 if M % 5 == 0 else 1)
        Wl = torch.randn(Lb, N, generator=g).to(DEV)
        bl = torch.randn(Lb, generator=g).to(DEV)
        logits = torch.empty(M // Lb, Lb, device=DEV)
        nat.check(L.lamp_diag_proj_ln(y_hi.data_ptr(), y_lo.data_ptr(), st.data_ptr(), np_, gam.data_ptr(), bet.data_ptr(),
                                      1e-5, Wl.data_ptr(), bl.data_ptr(), M // Lb, Lb, N, logits.data_ptr(), nat.stream()),
                  'diag_ln')
        torch.cuda.synchronize()
        ref = (ln_ref.view(M // Lb, Lb, N) * Wl.double()).sum(-1) + bl.double()
        assert rel_err(logits, ref) < 3e-5
    finally:
        nat.check(nat.lib().lamp_set_tuning(2, 1), 'tune')


def run_sdpa(q, k, v, mask, temperature, prec=0, want_attn=True):
    N, Lq, d = q.shape
    Lk = k.shape[1]
    L = nat.lib()
    out = torch.full((N, Lq, d), float('nan'), device=DEV)
    attn = torch.full((N, Lq, Lk), float('nan'), device=DEV) if want_attn else None
    w = ws(L.lamp_sdpa_workspace_bytes(N, Lq, Lk, d))
    if mask is not None:
        m8 = mask.to(DEV).to(torch.uint8)  # may be an expanded (stride-0) view
        sb, sq, sk = m8.stride()
        mp = m8.data_ptr()
    else:
        m8, mp, sb, sq, sk = None, None, 0, 0, 0
    nat.check(L.lamp_sdpa_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), mp, sb, sq, sk, out.data_ptr(), nat.ptr(attn),
                              N, Lq, Lk, d, float(temperature), prec, w.data_ptr(), w.numel(), nat.stream()), 'sdpa')
    torch.cuda.synchronize()
    return out, attn


def test_sdpa_golden_small():
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, 'sdpa_small.npz')).items()}
    rs = np.random.RandomState(5)
    n, lq, lk, d = 6, 33, 47, 32
    q = torch.from_numpy(rs.standard_normal((n, lq, d)).astype(np.float32))
    k = torch.from_numpy(rs.standard_normal((n, lk, d)).astype(np.float32))
    v = torch.from_numpy(rs.standard_normal((n, lk, d)).astype(np.float32))
    mask = torch.from_numpy(rs.rand(n, lq, lk) < 0.3)
    mask[:, :, 0] = False
    out, attn = run_sdpa(q.to(DEV), k.to(DEV), v.to(DEV), mask, np.power(d, 0.5))
    e_o, e_a = rel_err(out, g['out']), rel_err(attn, g['attn'])
    print(f'sdpa golden: out {e_o:.2e} attn {e_a:.2e}')
    assert e_o < 1e-4 and e_a < 1e-4


@pytest.mark.parametrize('N,Lq,Lk,d,maskkind', [
    (4, 103, 103, 128, 'rand'),   # single KV tile, P aliases Q
    (3, 103, 300, 128, 'pad'),    # multi-tile, BLOCK_KV = 64, key padding (query stride 0)
    (5, 159, 159, 64, 'none'),    # two KV tiles of 128, d = 64
    (2, 300, 300, 64, 'label'),   # 3x3 tiles, shared [L,L] mask (batch stride 0)
    (3, 40, 1, 16, 'none'),       # enc_vec path: one key
    (2, 130, 257, 96, 'rand'),    # ragged everything
    (300, 103, 103, 128, 'label'),  # more work items than SMs -> persistent loop
])
def test_sdpa_vs_oracle(N, Lq, Lk, d, maskkind):
    g = torch.Generator().manual_seed(N * 1000 + Lq + Lk + d)
    q = torch.randn(N, Lq, d, generator=g)
    k = torch.randn(N, Lk, d, generator=g)
    v = torch.randn(N, Lk, d, generator=g)
    mask = None
    if maskkind == 'rand':
        mask = torch.rand(N, Lq, Lk, generator=g) < 0.4
        mask[:, :, 0] = False
    elif maskkind == 'pad':
        lens = torch.randint(1, Lk + 1, (N,), generator=g)
        mask = (torch.arange(Lk)[None, :] >= lens[:, None]).unsqueeze(1).expand(N, Lq, Lk)
    elif maskkind == 'label':
        m = torch.rand(Lq, Lk, generator=g) < 0.7
        m[torch.arange(Lq), torch.arange(Lq) % Lk] = False
        mask = m.unsqueeze(0).expand(N, Lq, Lk)
    temp = float(np.power(d, 0.5))
    ref_o, ref_a = orc.sdpa(q.double(), k.double(), v.double(), mask, temp)
    out, attn = run_sdpa(q.to(DEV), k.to(DEV), v.to(DEV), mask, temp)
    e_o, e_a = rel_err(out, ref_o), rel_err(attn, ref_a)
    print(f'sdpa N={N} Lq={Lq} Lk={Lk} d={d} {maskkind}: out {e_o:.2e} attn {e_a:.2e}')
    assert e_o < 1e-4 and e_a < 1e-4
    assert torch.isfinite(out).all()


def test_sdpa_fully_masked_row_is_nan_like_reference():
    q = torch.randn(2, 20, 32)
    k = torch.randn(2, 24, 32)
    v = torch.randn(2, 24, 32)
    mask = torch.zeros(2, 20, 24, dtype=torch.bool)
    mask[1, 5, :] = True
    ref_o, _ = orc.sdpa(q, k, v, mask, 32 ** 0.5)
    out, attn = run_sdpa(q.to(DEV), k.to(DEV), v.to(DEV), mask, 32 ** 0.5)
    assert torch.isnan(ref_o[1, 5]).all() and torch.isnan(out[1, 5]).all()
    keep = torch.ones(2, 20, dtype=torch.bool)
    keep[1, 5] = False
    assert rel_err(out[keep.to(DEV)], ref_o[keep]) < 1e-4


def run_mha(p, q, kv, mask, H, want_attn=True, prec=0):
    B, Lq, D = q.shape
    self_attn = kv is q
    Lk = kv.shape[1]
    d = p['w_qs.weight'].shape[0] // H
    L = nat.lib()
    dev = {k: v.to(DEV).contiguous() for k, v in p.items()}
    qd = q.to(DEV).contiguous()
    kvd = qd if self_attn else kv.to(DEV).contiguous()
    out = torch.full((B, Lq, D), float('nan'), device=DEV)
    attn = torch.full((H * B, Lq, Lk), float('nan'), device=DEV) if want_attn else None
    w = ws(L.lamp_mha_workspace_bytes(B, Lq, Lk, D, H, d, int(self_attn), int(want_attn)))
    if mask is not None:
        m8 = mask.to(DEV).to(torch.uint8)
        sb, sq, sk = m8.stride()
        mp = m8.data_ptr()
    else:
        m8, mp, sb, sq, sk = None, None, 0, 0, 0
    fc = dev.get('fc.weight')
    nat.check(L.lamp_mha_fwd(qd.data_ptr(), None if self_attn else kvd.data_ptr(), dev['w_qs.weight'].data_ptr(),
                             dev['w_ks.weight'].data_ptr(), dev['w_vs.weight'].data_ptr(), nat.ptr(fc),
                             dev['layer_norm.weight'].data_ptr(), dev['layer_norm.bias'].data_ptr(), mp, sb, sq, sk,
                             out.data_ptr(), nat.ptr(attn), B, Lq, Lk, D, H, d, prec, 1e-5, w.data_ptr(), w.numel(),
                             nat.stream()), 'mha')
    torch.cuda.synchronize()
    return out, attn


@pytest.mark.parametrize('name', list(cases.MHA_CASES))
def test_mha_golden(name):
    """lamp_mha_fwd against the REFERENCE outputs (tests/golden) and the fp64 oracle."""
    c = cases.MHA_CASES[name]
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, 'mha_' + name + '.npz')).items()}
    p, q, kv, mask = cases.mha_inputs(c)
    out, attn = run_mha(p, q, kv, mask, c['H'])
    rs = c.get('row_stride', 1)
    e_ref = rel_err(out[:, ::rs], g['out'])
    p64 = orc.to_dtype(p, torch.float64)
    o64, a64 = orc.mha(p64, '', q.double(), kv.double(), kv.double(), mask, c['H'])
    e_64 = rel_err(out, o64)
    e_a = rel_err(attn, a64)
    print(f'mha {name}: vs reference {e_ref:.2e}  vs fp64 oracle {e_64:.2e}  attn {e_a:.2e}')
    assert e_ref < 1e-3 and e_64 < 1e-3 and e_a < 1e-3   # north_star tolerance
    assert e_64 < 5e-5                                    # what the 3-term path is expected to deliver
    if 'attn' in g:
        assert rel_err(attn[:, ::c.get('attn_row_stride', 1)], g['attn']) < 1e-3
    assert attn.shape == (c['H'] * c['B'], c['Lq'], c['Lk'])


def test_mha_bf16_mode():
    """cfg-3 style run (bf16 operands): looser, stated tolerance 2e-2 vs the fp32 reference."""
    c = cases.MHA_CASES['self_L159_H8_none']
    p, q, kv, mask = cases.mha_inputs(c)
    out, _ = run_mha(p, q, kv, mask, c['H'], want_attn=False, prec=1)
    o32, _ = orc.mha(p, '', q, kv, kv, mask, c['H'])
    e = rel_err(out, o32)
    print(f'mha bf16 mode rel err {e:.2e}')
    assert e < 2e-2


def test_ffn_vs_oracle():
    rs = np.random.RandomState(3)
    from lamp_b200 import synthetic as syn
    for rows, D, dh in [(2 * 103, 512, 512), (77, 64, 128), (1000, 512, 1024)]:
        p = syn.ffn_params(rs, '', D, dh, random_ln=True)
        x = torch.from_numpy(rs.standard_normal((1, rows, D)).astype(np.float32))
        ref = orc.ffn(orc.to_dtype(p, torch.float64), '', x.double())
        dev = {k: v.to(DEV).contiguous() for k, v in p.items()}
        xd = x.to(DEV)
        out = torch.full((rows, D), float('nan'), device=DEV)
        L = nat.lib()
        w = ws(L.lamp_ffn_workspace_bytes(rows, D, dh))
        nat.check(L.lamp_ffn_fwd(xd.data_ptr(), dev['w_1.weight'].data_ptr(), dev['w_1.bias'].data_ptr(),
                                 dev['w_2.weight'].data_ptr(), dev['w_2.bias'].data_ptr(),
                                 dev['layer_norm.weight'].data_ptr(), dev['layer_norm.bias'].data_ptr(),
                                 out.data_ptr(), rows, D, dh, 0, 1e-5, w.data_ptr(), w.numel(), nat.stream()), 'ffn')
        torch.cuda.synchronize()
        e = rel_err(out, ref[0])
        print(f'ffn rows={rows} D={D} dh={dh}: {e:.2e}')
        assert e < 5e-5


def test_layernorm_embed_diag():
    L = nat.lib()
    g = torch.Generator().manual_seed(1)
    for rows, D in [(5, 64), (1000, 512), (33, 1024), (9, 2048)]:
        y = torch.randn(rows, D, generator=g).to(DEV)
        add = torch.randn(4, D, generator=g).to(DEV)
        gam = torch.randn(D, generator=g).to(DEV)
        bet = torch.randn(D, generator=g).to(DEV)
        out = torch.empty(rows, D, device=DEV)
        hi = torch.empty(rows, D, dtype=torch.bfloat16, device=DEV)
        lo = torch.empty_like(hi)
        nat.check(L.lamp_layernorm(y.data_ptr(), add.data_ptr(), 4, gam.data_ptr(), bet.data_ptr(), 1e-5, rows, D,
                                   out.data_ptr(), hi.data_ptr(), lo.data_ptr(), None, nat.stream()), 'ln')
        z = y.double() + add.double()[torch.arange(rows, device=DEV) % 4]
        ref = torch.nn.functional.layer_norm(z, (D,), gam.double(), bet.double(), 1e-5)
        assert rel_err(out, ref) < 1e-5
        assert torch.equal(hi, out.to(torch.bfloat16))
        assert torch.equal(lo, (out - hi.float()).to(torch.bfloat16))
    V, P, D, rows = 50, 20, 128, 333
    we = torch.randn(V, D, generator=g).to(DEV)
    pe = torch.randn(P, D, generator=g).to(DEV)
    seq = torch.randint(0, V, (rows,), generator=g).to(DEV)
    pos = torch.randint(0, P, (rows,), generator=g).to(DEV)
    out = torch.empty(rows, D, device=DEV)
    nat.check(L.lamp_embed(seq.data_ptr(), pos.data_ptr(), we.data_ptr(), pe.data_ptr(), rows, D, out.data_ptr(), None,
                           None, None, None, nat.stream()), 'embed')
    assert torch.equal(out, we[seq] + pe[pos])
    B, Ln = 7, 37
    x = torch.randn(B, Ln, D, generator=g).to(DEV)
    W = torch.randn(Ln, D, generator=g).to(DEV)
    lg = torch.empty(B, Ln, device=DEV)
    nat.check(L.lamp_diag_proj(x.data_ptr(), W.data_ptr(), None, B, Ln, D, lg.data_ptr(), nat.stream()), 'diag')
    ref = torch.einsum('bld,ld->bl', x.double(), W.double())
    assert rel_err(lg, ref) < 1e-5


def test_error_codes_not_exceptions():
    L = nat.lib()
    x = torch.zeros(4, 6, device=DEV)
    rc = L.lamp_split_planes(x.data_ptr(), 4, 6, 6, x.data_ptr(), None, 6, nat.stream())
    assert rc == -1 and b'multiples of 4' in L.lamp_last_error()
    rc = L.lamp_mha_fwd(x.data_ptr(), None, x.data_ptr(), x.data_ptr(), x.data_ptr(), None, x.data_ptr(), x.data_ptr(),
                        None, 0, 0, 0, x.data_ptr(), None, 1, 4, 4, 8, 2, 4, 0, 1e-5, None, 0, nat.stream())
    assert rc == -1  # fc weight missing with n_head > 1


@pytest.mark.parametrize('B,H,Lq,Lk,d,maskkind', [
    (5, 4, 103, 103, 128, 'label'),   # bench shape: single KV tile, staged TMA-store epilogue
    (3, 2, 103, 300, 128, 'pad'),     # label<-input: 64-key tiles through a 3-slot ring
    (2, 8, 159, 159, 64, 'none'),     # cfg-3 dims: two q tiles, two KV tiles of 128
    (2, 2, 300, 520, 64, 'label'),    # long rows: the lazy reference maximum is exercised over 5 tiles
    (40, 4, 103, 103, 128, 'label'),  # more items than SMs (persistent loop, O / staging buffers reused)
])
@pytest.mark.parametrize('growth', [0.0, 3.0])
def test_attn_core_planes_epilogues_and_lazy_rescale(B, H, Lq, Lk, d, maskkind, growth):
    """lamp_attn_core_planes (A1, lamp/SubLayers.py:27-43) on plane operands: the staged (TMA store) and the direct
    epilogue agree with an fp64 softmax(QK^T/temperature)V of the same operands, also when the scores grow along the
    key axis so that the running maximum moves by far more (and by far less) than the lazy-rescale threshold."""
    from lamp_b200 import ops
    g = torch.Generator().manual_seed(B * 7 + Lq + Lk + d + int(growth))
    hd = H * d
    q = torch.randn(B * Lq, hd, generator=g)
    kv = torch.randn(B * Lk, 2 * hd, generator=g)
    if growth:
        ramp = 1.0 + growth * torch.arange(Lk, dtype=torch.float32) / Lk   # later keys score higher
        kv[:, :hd] *= ramp.repeat(B)[:, None]
    mask = None
    if maskkind == 'label':
        m = torch.rand(Lq, Lk, generator=g) < 0.6
        m[torch.arange(Lq), torch.arange(Lq) % Lk] = False
        mask = m.unsqueeze(0)
    elif maskkind == 'pad':
        lens = torch.randint(1, Lk + 1, (B,), generator=g)
        mask = (torch.arange(Lk)[None, :] >= lens[:, None]).unsqueeze(1)
    qa = ops.Act(None, *ops.split(q.to(DEV), 0), B * Lq, hd)
    kva = ops.Act(None, *ops.split(kv.to(DEV), 0), B * Lk, 2 * hd)
    # fp64 reference on the operands as the kernel sees them (hi + lo)
    qd = (qa.hi.double() + qa.lo.double()).view(B, Lq, H, d).permute(0, 2, 1, 3)
    kd = (kva.hi.double() + kva.lo.double())[:, :hd].reshape(B, Lk, H, d).permute(0, 2, 1, 3)
    vd = (kva.hi.double() + kva.lo.double())[:, hd:].reshape(B, Lk, H, d).permute(0, 2, 1, 3)
    s = qd @ kd.transpose(-1, -2) / float(np.power(d, 0.5))
    if mask is not None:
        s = s.masked_fill(mask.to(DEV)[:, None].expand(B, H, Lq, Lk) if mask.shape[0] == B
                          else mask.to(DEV)[None].expand(B, H, Lq, Lk), float('-inf'))
    ref = (torch.softmax(s, -1) @ vd).permute(0, 2, 1, 3).reshape(B * Lq, hd)
    outs = []
    try:
        for stage in (1, 0):
            nat.check(nat.lib().lamp_set_tuning(4, stage), 'tune')
            o, _ = ops.attention(qa, 0, kva, 0, hd, B, H, Lq, Lk, d, 0, None if mask is None else mask.to(DEV), False)
            torch.cuda.synchronize()
            out = o.hi.float() + o.lo.float()
            e = rel_err(out, ref)
            print(f'attn planes B={B} H={H} Lq={Lq} Lk={Lk} d={d} {maskkind} growth={growth} stage={stage}: {e:.2e}')
            assert e < 2e-5
            outs.append(out)
    finally:
        nat.check(nat.lib().lamp_set_tuning(4, 1), 'tune')
    assert torch.equal(outs[0], outs[1])   # same arithmetic, only the way out of the SM differs


@pytest.mark.parametrize('B,H,Lq,Lk,d,per_sample', [(2, 4, 983, 983, 128, False), (3, 2, 300, 520, 64, False),
                                                      (3, 2, 103, 103, 128, True), (2, 8, 159, 159, 64, True),
                                                      (2, 2, 70, 333, 32, True)])
def test_attn_core_packed_mask_equals_byte_mask(B, H, Lq, Lk, d, per_sample):
    """lamp_pack_mask_bits + lamp_attn_core_planes_mbits (one mask word per thread and KV tile) give bit-identical
    output to the byte-mask path of lamp_attn_core_planes, for the shared [Lq, Lk] label mask (cfg-4 dims incl.) and
    for per-sample [B, Lq, Lk] masks; the packed words equal a host-side packing."""
    from lamp_b200 import ops
    g = torch.Generator().manual_seed(B + H + Lq + Lk + d)
    hd = H * d
    qa = ops.Act(None, *ops.split(torch.randn(B * Lq, hd, generator=g).to(DEV), 0), B * Lq, hd)
    kva = ops.Act(None, *ops.split(torch.randn(B * Lk, 2 * hd, generator=g).to(DEV), 0), B * Lk, 2 * hd)
    m = torch.rand(B if per_sample else 1, Lq, Lk, generator=g) < 0.7
    m[:, torch.arange(Lq), torch.arange(Lq) % Lk] = False
    m8 = m.to(DEV).view(torch.uint8).expand(B, Lq, Lk)
    # packing kernel vs host packing
    words, mbb, mbq = ops.mask_bits(m8, B, Lq, Lk)
    W = (Lk + 31) // 32
    pad = torch.zeros(m.shape[0], Lq, W * 32, dtype=torch.int64)
    pad[:, :, :Lk] = m.long()
    ref_words = (pad.view(m.shape[0], Lq, W, 32) << torch.arange(32)).sum(-1)
    ref_words = torch.where(ref_words >= 2 ** 31, ref_words - 2 ** 32, ref_words).to(torch.int32)
    assert torch.equal(words.cpu(), ref_words)
    assert mbq == W and mbb == (Lq * W if per_sample else 0)
    L = nat.lib()
    outs = []
    for packed in (True, False):
        o_hi = torch.full((B * Lq, hd), float('nan'), dtype=torch.bfloat16, device=DEV)
        o_lo = torch.full_like(o_hi, float('nan'))
        sb, sq, sk = m8.stride()
        if packed:
            nat.check(L.lamp_attn_core_planes_mbits(qa.hi.data_ptr(), qa.lo.data_ptr(), hd, 0, 0, kva.hi.data_ptr(),
                                                    kva.lo.data_ptr(), 2 * hd, 0, hd, B, H, Lq, Lk, d, float(d ** 0.5), 0,
                                                    words.data_ptr(), mbb, mbq, o_hi.data_ptr(), o_lo.data_ptr(), hd,
                                                    None, 0, nat.stream()), 'mbits')
        else:
            nat.check(L.lamp_attn_core_planes(qa.hi.data_ptr(), qa.lo.data_ptr(), hd, 0, 0, kva.hi.data_ptr(),
                                              kva.lo.data_ptr(), 2 * hd, 0, hd, B, H, Lq, Lk, d, float(d ** 0.5), 0,
                                              m8.data_ptr(), sb, sq, sk, o_hi.data_ptr(), o_lo.data_ptr(), hd, None, 0,
                                              None, None, None, None, None, 0, nat.stream()), 'bytes')
        torch.cuda.synchronize()
        outs.append((o_hi.clone(), o_lo.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert not torch.isnan(outs[0][0].float()).any()


@pytest.mark.parametrize('N,Lq,Lk,d,maskkind', [(6, 103, 103, 128, 'label'), (4, 103, 300, 128, 'pad'),
                                                 (5, 159, 159, 64, 'none'), (3, 70, 45, 32, 'label'),
                                                 (2, 200, 130, 48, 'pad'), (2, 33, 17, 16, 'none'),
                                                 (2, 260, 257, 112, 'label')])
@pytest.mark.parametrize('bwd_tc', [1, 0])
def test_attn_core_backward_vs_autograd(N, Lq, Lk, d, maskkind, bwd_tc):
    """lamp_attn_core_bwd (dq, dk, dv of softmax(mask(q k^T / T)) v) against fp64 torch autograd of the same function,
    through ops.SDPAFunction (native forward + native backward): the batched tcgen05 version and the warp-MMA
    version (LAMP_TUNE_ATTN_BWD_TC = 0)."""
    nat.check(nat.lib().lamp_set_tuning(7, bwd_tc), 'tune')
    try:
        _attn_core_backward_vs_autograd(N, Lq, Lk, d, maskkind)
    finally:
        nat.check(nat.lib().lamp_set_tuning(7, 1), 'tune')


def _attn_core_backward_vs_autograd(N, Lq, Lk, d, maskkind):
    from lamp_b200 import ops
    g = torch.Generator().manual_seed(N + Lq + Lk + d)
    q = torch.randn(N, Lq, d, generator=g).to(DEV).requires_grad_(True)
    k = torch.randn(N, Lk, d, generator=g).to(DEV).requires_grad_(True)
    v = torch.randn(N, Lk, d, generator=g).to(DEV).requires_grad_(True)
    go = torch.randn(N, Lq, d, generator=g).to(DEV)
    mask = None
    if maskkind == 'label':
        m = torch.rand(Lq, Lk, generator=g) < 0.6
        m[torch.arange(Lq), torch.arange(Lq) % Lk] = False
        mask = m.unsqueeze(0).expand(N, Lq, Lk).to(DEV)
    elif maskkind == 'pad':
        lens = torch.randint(1, Lk + 1, (N,), generator=g)
        mask = (torch.arange(Lk)[None, :] >= lens[:, None]).unsqueeze(1).expand(N, Lq, Lk).to(DEV)
    T = float(np.power(d, 0.5))
    out, attn = ops.SDPAFunction.apply(q, k, v, mask, T, 0)
    out.backward(go)
    qd, kd, vd = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    s = qd @ kd.transpose(1, 2) / T
    if mask is not None:
        s = s.masked_fill(mask, float('-inf'))
    ref = torch.softmax(s, -1) @ vd
    ref.backward(go.double())
    assert rel_err(out, ref) < 2e-5
    for name, a, b in (('dq', q.grad, qd.grad), ('dk', k.grad, kd.grad), ('dv', v.grad, vd.grad)):
        e = rel_err(a, b)
        print(f'attn bwd N={N} Lq={Lq} Lk={Lk} d={d} {maskkind} {name}: {e:.2e}')
        assert e < 5e-5, name


@pytest.mark.parametrize('N,Lq,Lk,d,pdrop', [(8, 103, 103, 128, 0.2), (4, 103, 300, 128, 0.1), (4, 159, 159, 64, 0.5),
                                              (3, 260, 257, 32, 0.3)])
def test_attn_core_training_forward_dropout_and_backward(N, Lq, Lk, d, pdrop):
    """lamp_sdpa_fwd_train: dropout on the probabilities inside the attention kernel (lamp/SubLayers.py:40).
    (i) probs_pre is the softmax, attn == probs_pre * keep / (1 - p) for a kept set of the right size that changes
    with the seed and not with anything else, (ii) out == attn @ v, i.e. the PV product consumed exactly the kept set
    the probability kernel reports, (iii) the backward equals fp64 autograd of the same function with that kept set."""
    from lamp_b200 import ops
    g = torch.Generator().manual_seed(N + Lq + Lk + d)
    q = torch.randn(N, Lq, d, generator=g).to(DEV).requires_grad_(True)
    k = torch.randn(N, Lk, d, generator=g).to(DEV).requires_grad_(True)
    v = torch.randn(N, Lk, d, generator=g).to(DEV).requires_grad_(True)
    go = torch.randn(N, Lq, d, generator=g).to(DEV)
    m = torch.rand(Lq, Lk, generator=g) < 0.3
    m[torch.arange(Lq), torch.arange(Lq) % Lk] = False
    mask = m.unsqueeze(0).expand(N, Lq, Lk).to(DEV)
    T = float(np.power(d, 0.5))
    out, attn, pre = ops.sdpa_train(q.detach(), k.detach(), v.detach(), mask, T, 0, pdrop, 1234)
    out2, attn2, _ = ops.sdpa_train(q.detach(), k.detach(), v.detach(), mask, T, 0, pdrop, 1234)
    out3, attn3, _ = ops.sdpa_train(q.detach(), k.detach(), v.detach(), mask, T, 0, pdrop, 99)
    torch.cuda.synchronize()
    assert torch.equal(attn, attn2) and torch.equal(out, out2)          # same seed -> same kept set
    assert not torch.equal(attn != 0, attn3 != 0)                        # another seed -> another kept set
    s = (q.detach().double() @ k.detach().double().transpose(1, 2) / T).masked_fill(mask, float('-inf'))
    P = torch.softmax(s, -1)
    assert rel_err(pre, P) < 2e-5
    keep = attn != 0
    live = P > 1e-30
    frac = float(keep[live].double().mean())
    assert abs(frac - (1 - pdrop)) < 0.01, frac
    assert rel_err(attn, P * keep / (1 - pdrop)) < 2e-5
    assert rel_err(out, (P * keep / (1 - pdrop)) @ v.detach().double()) < 2e-5
    # backward through the Function with the same seed
    o, a = ops.SDPAFunction.apply(q, k, v, mask, T, 0, pdrop, 1234)
    assert torch.equal(a, attn)
    o.backward(go)
    qd, kd, vd = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    sd = (qd @ kd.transpose(1, 2) / T).masked_fill(mask, float('-inf'))
    ref = (torch.softmax(sd, -1) * keep / (1 - pdrop)) @ vd
    ref.backward(go.double())
    for name, x, y in (('dq', q.grad, qd.grad), ('dk', k.grad, kd.grad), ('dv', v.grad, vd.grad)):
        e = rel_err(x.detach(), y)
        print(f'attn train p={pdrop} N={N} Lq={Lq} Lk={Lk} d={d} {name}: {e:.2e}')
        assert e < 5e-5, name


@pytest.mark.parametrize('M,N,K,bias', [(3296, 512, 512, True), (1000, 1536, 512, False), (777, 264, 72, True),
                                         (130, 64, 1024, True), (9600, 512, 512, False)])
@pytest.mark.parametrize('tn_tc', [1, 0])
def test_linear_function_grads(M, N, K, bias, tn_tc):
    """ops.LinearFunction: y, dx on the tcgen05 GEMM, dW / db on lamp_gemm_tn_acc (tcgen05 kernel with MN-major
    operands, and the warp-MMA version behind LAMP_TUNE_GEMM_TN_TC=0) -- vs fp64 autograd."""
    from lamp_b200 import ops
    nat.check(nat.lib().lamp_set_tuning(6, tn_tc), 'tune')
    try:
        _linear_function_grads(M, N, K, bias)
    finally:
        nat.check(nat.lib().lamp_set_tuning(6, 1), 'tune')


def _linear_function_grads(M, N, K, bias):
    from lamp_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(DEV).requires_grad_(True)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(N, generator=g).to(DEV).requires_grad_(True) if bias else None
    go = torch.randn(M, N, generator=g).to(DEV)
    y = ops.LinearFunction.apply(x, W, b, 0)
    y.backward(go)
    xd, Wd = x.detach().double().requires_grad_(True), W.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    ref = torch.nn.functional.linear(xd, Wd, bd)
    ref.backward(go.double())
    pairs = [('y', y.detach(), ref.detach()), ('dx', x.grad, xd.grad), ('dW', W.grad, Wd.grad)]
    if bias:
        pairs.append(('db', b.grad, bd.grad))
    for name, a, r in pairs:
        e = rel_err(a, r)
        print(f'linear {M}x{N}x{K} {name}: {e:.2e}')
        assert e < 3e-5, name


@pytest.mark.parametrize('rows,D', [(3296, 512), (1000, 1024), (37, 64), (5000, 2048), (1, 512)])
def test_layernorm_function_grads(rows, D):
    from lamp_b200 import ops
    g = torch.Generator().manual_seed(rows + D)
    x = (torch.randn(rows, D, generator=g) * 2 + 0.5).to(DEV).requires_grad_(True)
    gam = (1 + 0.3 * torch.randn(D, generator=g)).to(DEV).requires_grad_(True)
    bet = torch.randn(D, generator=g).to(DEV).requires_grad_(True)
    go = torch.randn(rows, D, generator=g).to(DEV)
    y = ops.LayerNormFunction.apply(x, gam, bet, 1e-5, 0)
    y.backward(go)
    xd, gd, bd = (t.detach().double().requires_grad_(True) for t in (x, gam, bet))
    ref = torch.nn.functional.layer_norm(xd, (D,), gd, bd, 1e-5)
    ref.backward(go.double())
    for name, a, r in (('y', y.detach(), ref.detach()), ('dx', x.grad, xd.grad), ('dgamma', gam.grad, gd.grad),
                       ('dbeta', bet.grad, bd.grad)):
        e = rel_err(a, r)
        print(f'layernorm {rows}x{D} {name}: {e:.2e}')
        assert e < 2e-5, name


@pytest.mark.parametrize('B,L,D,bias', [(32, 103, 512, False), (5, 983, 512, True), (3, 7, 64, True)])
def test_diag_proj_function_grads(B, L, D, bias):
    from lamp_b200 import ops
    g = torch.Generator().manual_seed(B + L + D)
    x = torch.randn(B, L, D, generator=g).to(DEV).requires_grad_(True)
    W = torch.randn(L, D, generator=g).to(DEV).requires_grad_(True)
    b = torch.randn(L, generator=g).to(DEV).requires_grad_(True) if bias else None
    go = torch.randn(B, L, generator=g).to(DEV)
    y = ops.DiagProjFunction.apply(x, W, b)
    y.backward(go)
    xd, Wd = x.detach().double().requires_grad_(True), W.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    ref = torch.diagonal(torch.nn.functional.linear(xd, Wd, bd), 0, 1, 2)   # lamp/Models.py:124-126
    ref.backward(go.double())
    pairs = [('y', y.detach(), ref.detach()), ('dx', x.grad, xd.grad), ('dW', W.grad, Wd.grad)]
    if bias:
        pairs.append(('db', b.grad, bd.grad))
    for name, a, r in pairs:
        assert rel_err(a, r) < 1e-5, name
