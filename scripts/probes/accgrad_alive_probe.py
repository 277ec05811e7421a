"""Do the AccumulateGrad nodes of an eager training step outlive it?  (run plain and under compute-sanitizer)"""
import gc, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/tests/golden')
import torch, lamp_b200
from lamp_b200 import ops
import cases
from test_gpu_dropin import build_model, DEV
c = dict(cases.MODEL_CASES['lamp_L37_none'])
p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
model = build_model(c, p, adj, dropout=0.0)
src = (src_seq.to(DEV), src_pos.to(DEV))
gold = (torch.rand(c['B'], c['L']) < 0.3).float().to(DEV)
loss_fn = torch.nn.functional.binary_cross_entropy_with_logits
model.train()
names = dict((id(q), n) for n, q in model.named_parameters())
params = [q for q in model.parameters() if q.requires_grad]

def acc(q):
    return q.view_as(q).grad_fn.next_functions[0][0]

def survivors(tag):
    alive = [names[id(q)] for q in params if acc(q).metadata.get('mark') == tag]
    return alive

mode = sys.argv[1] if len(sys.argv) > 1 else 'full'
logits = model(src, None, None, gold)[0]
for q in params:
    acc(q).metadata['mark'] = 'during_forward'        # the nodes the live graph references
loss = loss_fn(logits, gold)
print('kernels:', dict(ops.STATS.by_kernel))
seen, stack, kinds = set(), [loss.grad_fn], {}
while stack:
    n = stack.pop()
    if n is None or n in seen:
        continue
    seen.add(n)
    kinds[n.name()] = kinds.get(n.name(), 0) + 1
    stack.extend(f for f, _ in n.next_functions)
print('graph nodes:', kinds)
del seen, stack, n
if mode == 'full':
    loss.backward()
del loss, logits
print('alive after the step (graph dropped):', survivors('during_forward'))
gc.collect()
print('alive after gc.collect():', survivors('during_forward'))
