"""Why does capturing the training step fail under compute-sanitizer only?  Variants of the test sequence
(tests/test_gpu_dropin.py::test_eval_forward_sees_new_weights...), one per process:  python <this> <variant>"""
import gc, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/tests/golden')
import torch, lamp_b200
from lamp_b200 import ops
import cases
from test_gpu_dropin import build_model, DEV
v = sys.argv[1]
c = dict(cases.MODEL_CASES['lamp_L37_none'])
p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
model = build_model(c, p, adj, dropout=0.0)
src = (src_seq.to(DEV), src_pos.to(DEV))
gold = (torch.rand(c['B'], c['L']) < 0.3).float().to(DEV)
loss_fn = torch.nn.functional.binary_cross_entropy_with_logits
if 'eval' in v:
    model.eval()
    for _ in range(3): model(src, None, None, None)
model.train()
if 'eager' in v:
    opt = torch.optim.SGD(model.get_trainable_parameters(), lr=1e-2) if 'sgd' in v else torch.optim.Adam(model.get_trainable_parameters(), lr=1e-2)
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        loss_fn(model(src, None, None, gold)[0], gold).backward()
        opt.step()
    if 'delopt' in v:
        del opt
    if 'sync' in v:
        torch.cuda.synchronize()
if 'eval2' in v:
    model.eval()
    for _ in range(3): model(src, None, None, None)
    model.train()
opt2 = torch.optim.Adam(model.get_trainable_parameters(), lr=1e-2, capturable=True) if 'inopt' in v else None
try:
    step = lamp_b200.GraphedTrainStep(model, loss_fn, c['B'], c['T'], example=(src[0], src[1], gold), optimizer=opt2)
    step(src[0], src[1], gold)
    torch.cuda.synchronize()
    print('VARIANT', v, 'ok')
except Exception as e:
    print('VARIANT', v, 'FAILED', str(e).splitlines()[0])
