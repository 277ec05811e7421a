"""Same question as accgrad_alive_probe.py without this package: plain torch ops and a python autograd.Function."""
import gc, torch
w = torch.nn.Parameter(torch.randn(64, 64, device='cuda'))
def acc(q):
    return q.view_as(q).grad_fn.next_functions[0][0]
class Sq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return x * x
    @staticmethod
    def backward(ctx, g):
        return 2 * ctx.saved_tensors[0] * g
for name, f in (('torch ops', lambda: (w * 2).sum()), ('python Function', lambda: Sq.apply(w).sum())):
    y = f()
    acc(w).metadata['mark'] = name
    del y
    print(name, ': AccumulateGrad alive after del:', acc(w).metadata.get('mark') == name)
