import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases
from lamp_b200 import ops
from test_gpu_model import build_model, rel_err
c = dict(cases.MODEL_CASES['lamp_L103_prior'])
p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
B = src_seq.shape[0]
tgt = (torch.arange(B * c['L']).view(B, c['L']) % 5 == 0).float()
orig = ops.linear_train
cap = {}
for mode in ('native', 'composed'):
    ops.NATIVE_ATTENTION_BACKWARD = ops.NATIVE_TRAINING = mode != 'composed'
    recs = []
    def wrapped(x, W, b, prec=None, recs=recs):
        rec = {'x': x.detach().clone(), 'shape': tuple(W.shape)}
        if x.requires_grad:
            x.register_hook(lambda g, rec=rec: rec.__setitem__('dx_total', g.detach().clone()))
        y = orig(x, W, b, prec)
        rec['y'] = y.detach().clone()
        y.register_hook(lambda g, rec=rec: rec.__setitem__('dy', g.detach().clone()))
        recs.append(rec)
        return y
    ops.linear_train = wrapped
    model = build_model(c, p, adj)
    model.train()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    logits, _, _ = model((src_seq.cuda(), src_pos.cuda()), None, None, None)
    torch.nn.functional.binary_cross_entropy_with_logits(logits, tgt.cuda()).backward()
    cap[mode] = recs
ops.linear_train = orig
for i, (a, b) in enumerate(zip(cap['native'], cap['composed'])):
    line = [f'{i:2d} W{a["shape"]} M={a["x"].numel() // a["x"].shape[-1]}']
    for k in ('x', 'y', 'dy', 'dx_total'):
        if k in a and k in b:
            line.append(f'{k} {rel_err(a[k], b[k]):.1e}')
    print(' '.join(line))
