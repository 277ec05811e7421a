#!/usr/bin/env python
"""Throughput + roofline sweep over the BASELINE.json configurations that are not the bench line:

  cfg-2  L=103  D=512  H=4  prior mask  fp32      cfg-3  L=159  D=512  H=8  no mask  bf16
  cfg-4  L=983  D=512  H=4  prior mask  fp32 (4 decoder layers)
  cfg-5  L in {512,1024,2048,4096}  D=1024  H=16  dense label graph  bf16

For each: U1 = label<-label attention core (Q,K,V -> O), U2 = one self-attention MultiHeadAttention, U3 = the
GraphDecoder stack over T=300 token encodings (random enc_output, padded lengths U{T/3..T}).  Per unit: samples/s and
the algorithmic GB/s and TFLOP/s of SURVEY.md 8d next to the measured peaks, naming the binding roofline.
CUDA events, inputs resident, >= 3 warm-ups; one JSON line per (config, unit).
usage: python scripts/bench_configs.py [cfg2] [cfg3] [cfg4] [cfg5] [--iters N]
Multi-GPU (BASELINE cfg-4 "batch-sharded 4xB200", cfg-5 "8xB200 sweep"): launch under torchrun, one rank per GPU --
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_configs.py cfg5
every rank runs the same per-GPU batch on its own shard (weights and label graph replicated, no forward collective:
weak scaling), times are the max over ranks between barriers, rank 0 prints whole-job samples/s and per-GPU rooflines."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import cases  # noqa: E402
import lamp_b200  # noqa: E402
from lamp_b200 import _native as nat  # noqa: E402
from lamp_b200 import ops  # noqa: E402
from lamp_b200 import synthetic as syn  # noqa: E402
from lamp_b200.Decoders import GraphDecoder  # noqa: E402
from lamp_b200.SubLayers import MultiHeadAttention  # noqa: E402

RANK = int(os.environ.get('RANK', '0'))
WORLD = int(os.environ.get('WORLD_SIZE', '1'))
LOCAL_RANK = int(os.environ.get('LOCAL_RANK', '0'))
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
torch.cuda.set_device(LOCAL_RANK)
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', LOCAL_RANK))
DEV = 'cuda'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return p['hbm_gbs'], p.get('bf16_tflops', 1590.0), 'measured'
    return 6650.0, 1590.0, 'fallback'


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    if WORLD > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    if WORLD > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if WORLD > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=DEV)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / iters * 1e-3


def report(cfg, unit, B, sec, flops, nbytes, extra=None):
    hbm, tf, src = peaks()
    gbs, tfs = nbytes / sec / 1e9, flops / sec / 1e12
    t_hbm, t_tensor = nbytes / (hbm * 1e9), flops / (tf * 1e12)
    bound = 'hbm' if t_hbm >= t_tensor else 'tensor'
    line = dict(config=cfg, unit=unit, n_gpus=WORLD, batch_per_gpu=B, ms=sec * 1e3, samples_per_s=B * WORLD / sec,
                per_gpu=True, alg_gbs=gbs, alg_tflops=tfs,
                frac_hbm=gbs / hbm, frac_tensor_bf16=tfs / tf, binding=bound,
                frac_of_binding=(gbs / hbm if bound == 'hbm' else tfs / tf), peaks=dict(hbm_gbs=hbm, bf16_tflops=tf, src=src))
    if extra:
        line.update(extra)
    if RANK == 0:
        print(json.dumps(line), flush=True)


def run_config(name, L, D, H, n_layers, mask_kind, prec_name, B, T=300, d_inner=None, iters=10, units=('U1', 'U2', 'U3')):
    d = D // H
    d_inner = d_inner or D
    prec = nat.PREC_FP32 if prec_name == 'fp32' else nat.PREC_BF16
    e = 4 if prec_name == 'fp32' else 2   # bytes per element of the operand form (plane pair / bf16)
    lamp_b200.set_default_precision(prec_name)
    rs = np.random.RandomState(1 + RANK)
    adj = cases.label_adj('prior', L, 1) if mask_kind == 'prior' else None
    cfg = f'{name} L={L} D={D} H={H} mask={mask_kind} {prec_name}'
    try:
        dec = GraphDecoder(L, L, n_layers=n_layers, n_head=H, n_head2=H, d_k=d, d_v=d, d_word_vec=D, d_model=D,
                           d_inner_hid=d_inner, label_adj_matrix=adj, label_mask=mask_kind, enc_vec=False).to(DEV).eval()
        mask = None if dec._label_mask_dev is None else dec._label_mask_dev.unsqueeze(0)
        hd = H * d
        if 'U1' in units:
            qkv = ops.Act(None, *ops.split(torch.randn(B * L, 3 * hd, device=DEV), prec), B * L, 3 * hd)
            sec = timeit(lambda: ops.attention(qkv, 0, qkv, hd, 2 * hd, B, H, L, L, d, prec, mask, False), iters)
            report(cfg, 'U1 attention core (self)', B, sec, 4.0 * B * L * L * hd, 4.0 * B * L * hd * e)
        if 'U2' in units:
            mha = dec.layer_stack[0].slf_attn
            x = ops.Act(None, *ops.split(torch.randn(B * L, D, device=DEV), prec), B * L, D)
            with torch.no_grad():
                sec = timeit(lambda: mha.forward_act(x, None, B, L, L, mask, False, want_f32=False), iters)
            fl = B * (2.0 * D * D * 3 * L + 2.0 * D * D * L + 4.0 * L * L * D)
            by = B * 2.0 * L * D * e + 4.0 * D * D * e
            report(cfg, 'U2 MultiHeadAttention (self)', B, sec, fl, by)
        if 'U3' in units:
            src_seq, _ = syn.make_tokens(B, T, 1000, 2 + RANK, min_len=T // 3)
            src_seq = src_seq.to(DEV)
            enc = torch.randn(B, T, D, device=DEV)
            keys = int((src_seq != 0).sum().item())
            with torch.no_grad():
                sec = timeit(lambda: dec(None, src_seq, enc), iters)
            per_layer = (2.0 * D * D * 2 * L            # enc_attn: Q projection + fc
                         + 4.0 * L * (keys / B) * D      # enc_attn core over the non-PAD keys
                         + 2 * 2.0 * D * d_inner * L * 2  # two FFNs
                         + 2.0 * D * D * 4 * L + 4.0 * L * L * D)  # self-attention
            fl = B * n_layers * per_layer + 2.0 * B * T * D * 2 * hd * n_layers  # + K|V projection of enc_output
            report(cfg, f'U3 GraphDecoder x{n_layers} (T={T})', B, sec, fl, 0.0, dict(binding='tensor'))
    finally:
        lamp_b200.set_default_precision('fp32')
    del dec
    torch.cuda.empty_cache()


if __name__ == '__main__':
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    iters = int(sys.argv[sys.argv.index('--iters') + 1]) if '--iters' in sys.argv else 10
    what = args or ['cfg2', 'cfg3', 'cfg4', 'cfg5']
    torch.manual_seed(RANK)
    if 'cfg2' in what:
        for B in (32, 1024, 8192):
            run_config('cfg-2', 103, 512, 4, 2, 'prior', 'fp32', B, iters=iters, units=('U1', 'U2') if B == 8192 else ('U1', 'U2', 'U3'))
    if 'cfg3' in what:
        run_config('cfg-3', 159, 512, 8, 2, 'none', 'bf16', 1024, iters=iters)
    if 'cfg4' in what:
        run_config('cfg-4', 983, 512, 4, 4, 'prior', 'fp32', 128, iters=iters)
    if 'cfg5' in what:
        for L in (512, 1024, 2048, 4096):
            run_config('cfg-5', L, 1024, 16, 2, 'none', 'bf16', max(4, 16384 // L), iters=max(3, iters // 2))
    if WORLD > 1:
        dist.destroy_process_group()
