#!/usr/bin/env python
"""Per-kernel CUDA-event breakdown of one training step (fwd + BCE + bwd) at the bench's cfg-1 or cfg-4 dims.
usage: python scripts/train_step_breakdown.py [cfg1|cfg4] [batch] [dropout rate override]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lamp_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
c = bench.CFG4 if which == 'cfg4' else bench.CFG
batch = int(sys.argv[2]) if len(sys.argv) > 2 else (32 if which == 'cfg4' else 256)
dev = torch.device('cuda', 0)
params, adj, src_seq, src_pos = bench.synth(batch, 500, c)
model = bench.build_model(c, params, adj, dev).train()
seq, pos = src_seq.to(dev), src_pos.to(dev)
if len(sys.argv) > 3:
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = float(sys.argv[3])
gold = (torch.rand(batch, c['L'], device=dev) < 0.05).float()


def step():
    model.zero_grad(set_to_none=True)
    logits, _, _ = model((seq, pos), None, None, gold)
    loss = ops.bce_with_logits(logits, gold)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
with ops.STATS.timed():
    step()
    per = ops.STATS.stop_timing()
print(json.dumps(dict(workload=which, batch=batch, dropout=(sys.argv[3] if len(sys.argv) > 3 else 'model default'), ms_per_step=ms,
                      native_kernel_ms={k: round(v['ms'], 3) for k, v in sorted(per.items(), key=lambda kv: -kv[1]['ms'])},
                      native_calls={k: v['calls'] for k, v in per.items()})))
