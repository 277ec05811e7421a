#!/usr/bin/env python
"""Precision / throughput frontier of the projection GEMMs.

Error side (runs on the CPU, no GPU needed): the oracle (fp64 arbitration copy of the reference algorithm) is re-run
with the OPERANDS of every dense contraction on the path (lamp/SubLayers.py:91-93,110,133-136: w_qs/w_ks/w_vs, fc,
w_1, w_2) rounded the way a tensor-core operand format would round them, products accumulated exactly (fp64: the
tensor core accumulates in fp32, whose error is 2-3 orders below every operand rounding studied here).  The attention
core (QK^T, PV) is HBM-bound and stays on the 3-term format in every mode, so only the GEMMs -- 64 % of the forward --
change.  Modes:

  bf16x3       hi/lo bf16 planes, hi*hi + hi*lo + lo*hi  (the shipped LAMP_PREC_FP32)          3 MMAs (bf16 rate)
  fp16x2w      A one fp16 plane, W two fp16 planes (hi + lo): a*wh + a*wl                       2 MMAs
  fp16x2a      A two fp16 planes, W one                                                          2 MMAs
  tf32         both operands rounded (RN) to 10 explicit mantissa bits, kind::tf32             1 MMA at half rate = 2
  tf32_trunc   tf32 as the tensor core reads raw fp32 (low 13 bits ignored)                      2
  fp16         both operands one fp16 plane                                                      1 MMA
  bf16         both operands one bf16 plane (LAMP_PREC_BF16)                                     1 MMA
  mixed_qk3    bf16x3 for the Q and K projections (their product is exponentiated), fp16 elsewhere

Metric: max |logits - ref| / max |ref| (the tests' rel_err) against the fp64 oracle, on the four committed model
fixtures and a cfg-4 shaped model (L=983, 4 decoder layers, prior mask).  Output: one JSON line per (case, mode) ->
profiles/r02_precision_frontier.jsonl.  Speed side: `--gpu` adds samples/s of the modes the library implements.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

from oracle import lamp_oracle as orc  # noqa: E402  (test infrastructure: this script is a checker, not product)
import cases  # noqa: E402

MMA_COST = dict(bf16x3=3, fp16x2w=2, fp16x2a=2, tf32=2, tf32_trunc=2, fp16=1, bf16=1, mixed_qk3=None, exact=None)


def rn(x64, dtype):
    return x64.to(torch.float32).to(dtype).to(torch.float64)


def tf32_rn(x64):
    x = x64.to(torch.float32).contiguous()
    i = x.view(torch.int32)
    i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32).to(torch.float64)


def tf32_trunc(x64):
    x = x64.to(torch.float32).contiguous()
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32).to(torch.float64)


def split2(x64, dtype):
    hi = rn(x64, dtype)
    lo = rn(x64.to(torch.float32).to(torch.float64) - hi, dtype)
    return hi, lo


def product(a, w, mode):
    """a [.., K] @ w[N, K]^T with operands rounded per `mode`; fp64 accumulation."""
    a32, w32 = a.to(torch.float32).to(torch.float64), w.to(torch.float32).to(torch.float64)  # HBM tensors are fp32
    if mode == 'exact':
        return a32 @ w32.t()
    if mode == 'bf16x3':
        ah, al = split2(a32, torch.bfloat16)
        wh, wl = split2(w32, torch.bfloat16)
        return ah @ wh.t() + ah @ wl.t() + al @ wh.t()
    if mode == 'fp16x2w':
        wh, wl = split2(w32, torch.float16)
        ar = rn(a32, torch.float16)
        return ar @ wh.t() + ar @ wl.t()
    if mode == 'fp16x2a':
        ah, al = split2(a32, torch.float16)
        wr = rn(w32, torch.float16)
        return ah @ wr.t() + al @ wr.t()
    if mode == 'tf32':
        return tf32_rn(a32) @ tf32_rn(w32).t()
    if mode == 'tf32_trunc':
        return tf32_trunc(a32) @ tf32_trunc(w32).t()
    if mode == 'fp16':
        return rn(a32, torch.float16) @ rn(w32, torch.float16).t()
    if mode == 'bf16':
        return rn(a32, torch.bfloat16) @ rn(w32, torch.bfloat16).t()
    raise KeyError(mode)


class Emulate:
    """Patch the oracle's dense contractions (F.linear / F.conv1d inside oracle.lamp_oracle) for one mode."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        mode = self.mode
        self._lin, self._conv = orc.F.linear, orc.F.conv1d
        qk_weights = self.qk_ids = getattr(self, 'qk_ids', set())

        def linear(x, w, b=None):
            m = mode
            if mode == 'mixed_qk3':
                m = 'bf16x3' if id(w) in qk_weights else 'fp16'
            y = product(x, w, m).to(x.dtype)
            return y if b is None else y + b

        def conv1d(x, w, b=None):  # k=1 convolution over [B, C, T]
            y = linear(x.transpose(1, 2), w.reshape(w.shape[0], -1), b)
            return y.transpose(1, 2)

        class _F:
            def __getattr__(self_inner, name):
                return getattr(F, name)
        f = _F()
        f.linear, f.conv1d = linear, conv1d
        self._F = orc.F
        orc.F = f
        return self

    def __exit__(self, *exc):
        orc.F = self._F
        return False


def run_case(name, p, cfg, src_seq, src_pos, lm, modes, out):
    p64 = orc.to_dtype(p, torch.float64)
    with Emulate('exact'):
        ref, _ = orc.lamp_forward(p64, cfg, src_seq, src_pos, lm, compute_dead_attention=False)
    qk_ids = {id(v) for k, v in p64.items() if k.endswith('w_qs.weight') or k.endswith('w_ks.weight')}
    for mode in modes:
        em = Emulate(mode)
        em.qk_ids = qk_ids
        with em:
            got, _ = orc.lamp_forward(p64, cfg, src_seq, src_pos, lm, compute_dead_attention=False)
        err = float((got - ref).abs().max() / ref.abs().max())
        rec = dict(case=name, mode=mode, rel_err=err, mma_per_kslice=MMA_COST[mode], layers=f"{cfg['n_layers_enc']}+{cfg['n_layers_dec']}")
        print(json.dumps(rec), flush=True)
        out.append(rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--modes', default='bf16x3,fp16x2w,fp16x2a,tf32,tf32_trunc,fp16,bf16,mixed_qk3')
    ap.add_argument('--cases', default='lamp_L103_prior,lamp_L37_none,lamp_L37_inveye,lamp_L20_meanvec,cfg4_L983')
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'r02_precision_frontier.jsonl'))
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    modes = args.modes.split(',')
    recs = []
    for name in args.cases.split(','):
        if name in cases.MODEL_CASES:
            c = cases.MODEL_CASES[name]
            p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
            lm = orc.label_mask_from(c['L'], adj, c['mask'])
        else:  # cfg-4 shaped: L=983, prior mask, 4 decoder layers (BASELINE.json configs[3]); B=2, T=120 keeps fp64 CPU time bounded
            from lamp_b200 import synthetic as syn
            L, T, V, D, H = 983, 120, 500, 512, 4
            p = syn.lamp_params(V + 4, L, T, D, 512, H, 2, 4, seed=41)
            src_seq, src_pos = syn.make_tokens(2, T, V, 1041)
            adj = syn.prior_adjacency(syn.make_label_sets(L, seed=41), L)
            cfg = dict(n_layers_enc=2, n_layers_dec=4, n_head=H, n_head2=H, enc_transform='', label_mask='prior')
            lm = orc.label_mask_from(L, adj, 'prior')
        run_case(name, p, cfg, src_seq, src_pos, lm, modes, recs)
    with open(args.out, 'w') as f:
        for r in recs:
            f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
