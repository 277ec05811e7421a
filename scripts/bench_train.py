#!/usr/bin/env python
"""Training-step timing of the label-graph model (forward + BCE loss + backward, the reference's train.py:28-48 minus
the optimizer) with the native training path (attention core, contractions, LayerNorm, label projection on lamp_b200
kernels in both directions) vs the all-torch composed path.  cfg-1 dims (L=103, T=300, D=512, H=4, 2+2 layers),
dropout 0.2 as in the reference's README command.  usage: python scripts/bench_train.py [--batch 32] [--steps 10]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lamp_b200 import ops  # noqa: E402
from lamp_b200 import synthetic as syn  # noqa: E402
from lamp_b200.Models import LAMP  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, nargs='+', default=[32, 256])
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--native-only', action='store_true', help='only the native eager step (profiling runs)')
ap.add_argument('--ddp', action='store_true',
                help='under torchrun: one process per GPU, per-rank batch, flat NCCL gradient all-reduce every step '
                     '(lamp_b200.distributed.allreduce_gradients); prints whole-job samples/s from rank 0')
args = ap.parse_args()
if args.ddp:
    import torch.distributed as dist
    from lamp_b200 import distributed as lds
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    rank, world, local_rank = lds.init_from_env()
    torch.cuda.set_device(local_rank)
    c0 = dict(L=103, T=300, V=20000, D=512, d_inner=512, H=4, n_enc=2, n_dec=2)
    d0 = c0['D'] // c0['H']
    for B in args.batch:
        params = syn.lamp_params(c0['V'] + 4, c0['L'], c0['T'], c0['D'], c0['d_inner'], c0['H'], 2, 2, seed=0)
        adj = syn.prior_adjacency(syn.make_label_sets(c0['L'], seed=0), c0['L'])
        src_seq, src_pos = syn.make_tokens(B, c0['T'], c0['V'], 1 + rank)
        src = (src_seq.cuda(), src_pos.cuda())
        tgt = (torch.rand(B, c0['L'], device='cuda') < 0.05).float()
        model = LAMP(c0['V'] + 4, c0['L'], c0['T'], c0['L'], n_layers_enc=2, n_layers_dec=2, n_head=4, n_head2=4,
                     d_word_vec=512, d_model=512, d_inner_hid=512, d_k=d0, d_v=d0, dropout=0.2, dec_dropout=0.2,
                     dec_dropout2=False, proj_share_weight=True, encoder='graph', decoder='graph', label_adj_matrix=adj,
                     label_mask='prior')
        model.load_state_dict(params, strict=True)
        model = model.cuda().train()
        plist = list(model.get_trainable_parameters())

        def ddp_step():
            model.zero_grad(set_to_none=True)
            logits, _, _ = model(src, None, None, None)
            torch.nn.functional.binary_cross_entropy_with_logits(logits, tgt).backward()
            return lds.allreduce_gradients(plist, world)

        for _ in range(3):
            n = ddp_step()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ddp_step()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(t.item())
            print(json.dumps(dict(what='data-parallel train step (fwd + BCE + bwd + flat NCCL gradient all-reduce)',
                                  n_gpus=world, batch_per_gpu=B, ms_per_step=ms, samples_per_s=B * world / ms * 1e3,
                                  allreduce_elements=n)), flush=True)
    dist.destroy_process_group()
    sys.exit(0)
c = dict(L=103, T=300, V=20000, D=512, d_inner=512, H=4, n_enc=2, n_dec=2)
dev = 'cuda'
for B in args.batch:
    params = syn.lamp_params(c['V'] + 4, c['L'], c['T'], c['D'], c['d_inner'], c['H'], c['n_enc'], c['n_dec'], seed=0)
    adj = syn.prior_adjacency(syn.make_label_sets(c['L'], seed=0), c['L'])
    src_seq, src_pos = syn.make_tokens(B, c['T'], c['V'], 1)
    src = (src_seq.to(dev), src_pos.to(dev))
    tgt = (torch.rand(B, c['L'], device=dev) < 0.05).float()
    for native in ((True,) if args.native_only else (True, False)):
        ops.NATIVE_ATTENTION_BACKWARD = ops.NATIVE_TRAINING = native
        d = c['D'] // c['H']
        model = LAMP(c['V'] + 4, c['L'], c['T'], c['L'], n_layers_enc=c['n_enc'], n_layers_dec=c['n_dec'], n_head=c['H'],
                     n_head2=c['H'], d_word_vec=c['D'], d_model=c['D'], d_inner_hid=c['d_inner'], d_k=d, d_v=d, dropout=0.2,
                     dec_dropout=0.2, dec_dropout2=False, proj_share_weight=True, encoder='graph', decoder='graph',
                     label_adj_matrix=adj, label_mask='prior')
        model.load_state_dict(params, strict=True)
        model = model.to(dev).train()

        def step():
            model.zero_grad(set_to_none=True)
            logits, _, _ = model(src, None, None, None)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, tgt)
            loss.backward()
            return loss

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ops.STATS.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        per_kernel = {}
        if native:  # one more step with a CUDA-event pair around every native call
            with ops.STATS.timed():
                step()
                per_kernel = {k: round(v['ms'], 2) for k, v in ops.STATS.stop_timing().items()}
        print(json.dumps(dict(what='train step (fwd + BCE + bwd)', native=native, batch=B, ms_per_step=ms,
                              native_kernel_ms=per_kernel,
                              samples_per_s=B / ms * 1e3, loss=float(loss.detach()),
                              native_launches_per_step=ops.STATS.launches / args.steps,
                              kernels={k: v // args.steps for k, v in ops.STATS.by_kernel.items()})), flush=True)
    ops.NATIVE_ATTENTION_BACKWARD = ops.NATIVE_TRAINING = True
    if args.native_only:
        continue
    # the same native step replayed as one CUDA graph (lamp_b200.GraphedTrainStep)
    import lamp_b200
    model = LAMP(c['V'] + 4, c['L'], c['T'], c['L'], n_layers_enc=c['n_enc'], n_layers_dec=c['n_dec'], n_head=c['H'],
                 n_head2=c['H'], d_word_vec=c['D'], d_model=c['D'], d_inner_hid=c['d_inner'], d_k=c['D'] // c['H'],
                 d_v=c['D'] // c['H'], dropout=0.2, dec_dropout=0.2, dec_dropout2=False, proj_share_weight=True,
                 encoder='graph', decoder='graph', label_adj_matrix=adj, label_mask='prior')
    model.load_state_dict(params, strict=True)
    model = model.to(dev).train()
    gstep = lamp_b200.GraphedTrainStep(model, torch.nn.functional.binary_cross_entropy_with_logits, B, c['T'],
                                       example=(src[0], src[1], tgt))
    for _ in range(3):
        gstep(src[0], src[1], tgt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = gstep(src[0], src[1], tgt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps(dict(what='train step (fwd + BCE + bwd), CUDA-graph replay', native=True, batch=B, ms_per_step=ms,
                          samples_per_s=B / ms * 1e3, loss=float(loss), native_launches_per_step=gstep.kernels_per_replay)),
          flush=True)
    ops.TRAIN_SEED_DEV = None
