#!/usr/bin/env python
"""Kernel micro-benchmarks on the bench shapes (CUDA events, inputs >> L2).  Prints one line per case:
algorithmic TFLOP/s and GB/s.  usage: python scripts/bench_kernels.py [gemm] [attn] [ln]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lamp_b200 import _native as nat  # noqa: E402
from lamp_b200 import ops  # noqa: E402

DEV = 'cuda'


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def bench_gemm(prec=0):
    B = 1024
    shapes = [('ffn1 enc', B * 300, 512, 512, 'planes'), ('ffn2 enc', B * 300, 512, 512, 'f32res'),
              ('kv_all', B * 300, 2048, 512, 'planes'), ('qkv', B * 103, 1536, 512, 'planes'),
              ('fc', B * 103, 512, 512, 'f32res'), ('ffn1 dec', B * 103, 512, 512, 'planes')]
    for name, M, N, K, mode in shapes:
        a = torch.randn(M, K, device=DEV)
        w = torch.randn(N, K, device=DEV) / K ** 0.5
        a_hi, a_lo = ops.split(a, prec)
        w_hi, w_lo = ops.split(w, prec)
        res = torch.randn(M, N, device=DEV) if mode == 'f32res' else None
        out = torch.empty(M, N, device=DEV) if mode == 'f32res' else None
        o_hi, o_lo = ops._empty_planes(M, N, prec, DEV) if mode == 'planes' else (None, None)
        for bk, pair in ((32, 1), (32, 0), (64, 1)):
            nat.check(nat.lib().lamp_set_tuning(1, bk), 'tune')
            nat.check(nat.lib().lamp_set_tuning(2, pair), 'tune')

            def run():
                ops.gemm(a_hi, a_lo, K, w_hi, w_lo, K, M, N, K, prec, residual=res, ldr=N, out_f32=out, ldo=N,
                         out_hi=o_hi, out_lo=o_lo, ldp=N)
            t = timeit(run)
            print(f'gemm {name:9s} M={M} N={N} K={K} {mode:7s} prec={prec} BK={bk} pair={pair}: {t * 1e6:8.1f} us  '
                  f'{2.0 * M * N * K / t / 1e12:7.1f} TFLOP/s alg', flush=True)
        nat.check(nat.lib().lamp_set_tuning(1, 0), 'tune')
        nat.check(nat.lib().lamp_set_tuning(2, 1), 'tune')


def bench_gemm_ln(prec=0):
    B = 1024
    for name, M in (('enc', B * 300), ('dec', B * 103)):
        N = K = 512
        a = ops.Act(None, *ops.split(torch.randn(M, K, device=DEV), prec), M, K)
        w_hi, w_lo = ops.split(torch.randn(N, K, device=DEV) / K ** 0.5, prec)
        res = ops.Act(torch.randn(M, N, device=DEV), None, None, M, N)
        g = torch.ones(N, device=DEV)
        b = torch.zeros(N, device=DEV)
        bias = torch.zeros(N, device=DEV)
        for fuse in (True, False):
            ops.FUSE_LAYERNORM = fuse
            ops.DEFER_LAYERNORM = False
            t = timeit(lambda: ops.linear_residual_ln(a, w_hi, w_lo, N, prec, res, g, b, 1e-5, bias=bias))
            print(f'gemm+LN {name} M={M} fused={fuse}: {t * 1e6:8.1f} us  {2.0 * M * N * K / t / 1e12:7.1f} TFLOP/s alg',
                  flush=True)
        ops.FUSE_LAYERNORM = False


def bench_gemm_pres(prec=0):
    B = 1024
    for name, M in (('enc', B * 300), ('dec', B * 103)):
        N = K = 512
        a = ops.Act(None, *ops.split(torch.randn(M, K, device=DEV), prec), M, K)
        w_hi, w_lo = ops.split(torch.randn(N, K, device=DEV) / K ** 0.5, prec)
        rf = torch.randn(M, N, device=DEV)
        res_f = ops.Act(rf, None, None, M, N)
        res_p = ops.Act(None, *ops.split(rf, prec), M, N)
        for nm, r in (('fp32 residual', res_f), ('planes residual', res_p)):
            t = timeit(lambda: ops.linear_residual_f32(a, w_hi, w_lo, N, prec, r))
            print(f'gemm f32out {name} M={M} {nm}: {t * 1e6:8.1f} us  {2.0 * M * N * K / t / 1e12:7.1f} TFLOP/s alg',
                  flush=True)


def bench_dln(prec=0):
    """Deferred-LayerNorm GEMM flavours next to their plain counterparts (same shapes as the bench step)."""
    B = 1024
    for name, M in (('enc', B * 160), ('dec', B * 103)):
        N = K = 512
        a = ops.Act(None, *ops.split(torch.randn(M, K, device=DEV), prec), M, K)
        w_hi, w_lo = ops.split(torch.randn(N, K, device=DEV) / K ** 0.5, prec)
        rf = torch.randn(M, N, device=DEV)
        g = torch.ones(N, device=DEV)
        b = torch.zeros(N, device=DEV)
        bias = torch.zeros(N, device=DEV)
        res_f = ops.Act(rf, None, None, M, N)
        res_p = ops.Act(None, *ops.split(rf, prec), M, N)
        res_d = ops.linear_residual_deferred(a, w_hi, w_lo, N, prec, res_p, g, b, 1e-5, bias=bias)
        for nm, r in (('fp32 residual', res_f), ('planes residual', res_p)):
            t = timeit(lambda: ops.linear_residual_f32(a, w_hi, w_lo, N, prec, r, bias=bias))
            print(f'gemm f32out  {name} M={M} {nm:18s}: {t * 1e6:8.1f} us  {2.0 * M * N * K / t / 1e12:7.1f} TFLOP/s alg',
                  flush=True)
        for nm, r in (('fp32 residual', res_f), ('planes residual', res_p), ('deferred residual', res_d)):
            t = timeit(lambda: ops.linear_residual_deferred(a, w_hi, w_lo, N, prec, r, g, b, 1e-5, bias=bias))
            print(f'gemm rstats  {name} M={M} {nm:18s}: {t * 1e6:8.1f} us  {2.0 * M * N * K / t / 1e12:7.1f} TFLOP/s alg',
                  flush=True)
        wp = ops.WeightPlanes()
        for N2 in (512, 1536, 2048):
            w2 = torch.nn.Parameter(torch.randn(N2, K, device=DEV) / K ** 0.5)
            b2 = torch.nn.Parameter(torch.zeros(N2, device=DEV))
            for nm, x in (('plain A', res_p), ('deferred A', res_d)):
                t = timeit(lambda: ops.project(x, wp, f'w{N2}', (w2,), N2, prec, bias=b2, relu=True))
                print(f'gemm planes  {name} M={M} N={N2} {nm:11s}: {t * 1e6:8.1f} us  '
                      f'{2.0 * M * N2 * K / t / 1e12:7.1f} TFLOP/s alg', flush=True)
        t = timeit(lambda: ops.materialize(res_d, prec, want_f32=True, want_planes=False))
        print(f'ln_apply     {name} M={M}: {t * 1e6:8.1f} us  {M * N * 8 / t / 1e9:7.1f} GB/s', flush=True)


def bench_attn(prec=0):
    B, H, d = 1024, 4, 128
    hd = H * d
    for name, Lq, Lk, masked in [('self L=103', 103, 103, True), ('enc T=300', 103, 300, True)]:
        q = ops.Act(None, *ops.split(torch.randn(B * Lq, hd, device=DEV), prec), B * Lq, hd)
        kv = ops.Act(None, *ops.split(torch.randn(B * Lk, 2 * hd, device=DEV), prec), B * Lk, 2 * hd)
        mask = (torch.rand(1, Lq, Lk, device=DEV) < 0.5)
        mask[:, :, 0] = False
        if Lk == 300:
            mask = (torch.rand(B, 1, Lk, device=DEV) < 0.3)
            mask[:, :, 0] = False

        def run():
            ops.attention(q, 0, kv, 0, hd, B, H, Lq, Lk, d, prec, mask if masked else None, False)
        nbytes = (B * Lq * hd * 2 + 2 * B * Lk * hd) * 4
        for rep in range(2):
            for compact in (1, 0):
                nat.check(nat.lib().lamp_set_tuning(3, compact), 'tune')
                t = timeit(run)
                print(f'attn {name:11s} B={B} H={H} d={d} compact={compact}: {t * 1e6:8.1f} us  {nbytes / t / 1e9:7.1f} GB/s alg  '
                      f'{4.0 * B * H * Lq * Lk * d / t / 1e12:6.1f} TFLOP/s alg', flush=True)
        nat.check(nat.lib().lamp_set_tuning(3, 1), 'tune')


def bench_ln(prec=0):
    for rows in (1100 * 160, 1100 * 103):
        y = torch.randn(rows, 512, device=DEV)
        g = torch.ones(512, device=DEV)
        b = torch.zeros(512, device=DEV)
        for want_f32 in (True, False):
            t = timeit(lambda: ops.layernorm(y, g, b, 1e-5, prec, want_f32=want_f32))
            nb = rows * 512 * (12 if want_f32 else 8)
            print(f'layernorm rows={rows} f32_out={int(want_f32)}: {t * 1e6:8.1f} us  {nb / t / 1e9:7.1f} GB/s', flush=True)


if __name__ == '__main__':
    what = sys.argv[1:] or ['gemm', 'attn', 'ln']
    torch.manual_seed(0)
    if 'gemm' in what:
        bench_gemm(0)
    if 'gemm1' in what:
        bench_gemm(1)
    if 'gemm_pres' in what:
        bench_gemm_pres(0)
    if 'gemm_ln' in what:
        bench_gemm_ln(0)
    if 'dln' in what:
        bench_dln(0)
    if 'attn' in what:
        bench_attn(0)
    if 'ln' in what:
        bench_ln(0)
