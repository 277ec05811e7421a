#!/usr/bin/env python
"""Multi-GPU check of the batch-sharded path, run under torchrun with one rank per GPU (NCCL):
  1. fused eval forward on this rank's shard == the same rows of the full-batch forward (no forward collective);
  2. training step: shard-local BCE gradients + ONE flat NCCL all-reduce == single-process gradients on the full
     batch (the encoder self-attention parameters keep grad None on every rank and do not dead-lock the reduce).
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/ddp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import cases  # noqa: E402
from lamp_b200 import distributed as D  # noqa: E402
from lamp_b200.Models import LAMP  # noqa: E402


def build(c, p, adj, dev):
    d = c['D'] // c['H']
    m = LAMP(c['V'] + 4, c['L'], c['T'], c['L'], n_layers_enc=c['n_enc'], n_layers_dec=c['n_dec'], n_head=c['H'],
             n_head2=c['H'], d_word_vec=c['D'], d_model=c['D'], d_inner_hid=c['d_inner'], d_k=d, d_v=d, dropout=0.0,
             dec_dropout=0.0, dec_dropout2=False, encoder='graph', decoder='graph', label_adj_matrix=adj,
             label_mask=c['mask'])
    m.load_state_dict(p, strict=True)
    return m.to(dev)


def main():
    rank, world, local = D.init_from_env('nccl')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    c = dict(cases.MODEL_CASES['lamp_L37_none'], B=10)
    p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
    gold = (torch.rand(c['B'], c['L'], generator=torch.Generator().manual_seed(7)) < 0.2).float()
    model = build(c, p, adj, dev)

    # 1. forward: shard vs full
    model.eval()
    with torch.no_grad():
        full, _, _ = model((src_seq.to(dev), src_pos.to(dev)), None, None, None)
        s_seq, s_pos = D.shard_batch([src_seq, src_pos], rank, world)
        part, _, _ = model((s_seq.to(dev), s_pos.to(dev)), None, None, None)
    a, b = D.shard_range(c['B'], rank, world)
    fwd_equal = bool(torch.equal(part, full[a:b]))

    # 2. training step: sharded grads + all-reduce vs single-process grads
    model.train()
    s_gold = gold[a:b].to(dev)
    logits, _, _ = model((s_seq.to(dev), s_pos.to(dev)), None, None, None)
    torch.nn.functional.binary_cross_entropy_with_logits(logits, s_gold).backward()
    params = list(model.get_trainable_parameters())
    n_none = sum(1 for q in params if q.grad is None)
    nel = D.allreduce_gradients(params, local_weight=(b - a) * world / c['B'])
    ref = build(c, p, adj, dev).train()
    rl, _, _ = ref((src_seq.to(dev), src_pos.to(dev)), None, None, None)
    torch.nn.functional.binary_cross_entropy_with_logits(rl, gold.to(dev)).backward()
    worst = 0.0
    for q, r in zip(params, ref.get_trainable_parameters()):
        assert (q.grad is None) == (r.grad is None)
        if r.grad is not None:
            worst = max(worst, float((q.grad - r.grad).abs().max() / r.grad.abs().max().clamp_min(1e-20)))
    # 3. the bucketed, backward-overlapped reducer (flat buffer, p.grad views): three steps vs single-process gradients
    model2 = build(c, p, adj, dev).train()
    plist = list(model2.get_trainable_parameters())
    red = D.GradientReducer(plist, world=world, bucket_mb=0.25)
    red.local_weight = (b - a) * world / c['B']
    worst_red = 0.0
    for step in range(3):
        red.zero_grad()
        lg, _, _ = model2((s_seq.to(dev), s_pos.to(dev)), None, None, None)
        torch.nn.functional.binary_cross_entropy_with_logits(lg, s_gold).backward()
        red.finish()
        for q, r in zip(plist, ref.get_trainable_parameters()):
            assert (q.grad is None) == (r.grad is None)
            if r.grad is not None:
                worst_red = max(worst_red, float((q.grad - r.grad).abs().max() / r.grad.abs().max().clamp_min(1e-20)))
    worst = max(worst, worst_red)
    red_stats = dict(red.stats)
    res = torch.tensor([float(fwd_equal), worst], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN if False else dist.ReduceOp.MAX)
    ok = torch.tensor([float(fwd_equal)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'ddp_check world={world}: forward shard == full rows: {bool(ok.item())}; flat all-reduce of {nel} grads, '
              f'{n_none} params with grad None; max rel grad diff vs single process: {res[1].item():.2e}; '
              f'GradientReducer (3 steps): {red_stats}')
        assert ok.item() == 1.0 and res[1].item() < 1e-4
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
