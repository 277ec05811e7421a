#!/usr/bin/env python
"""One launch of every GEMM epilogue flavour at the decoder shape (M = 1024*103, N = K = 512) -- a short command for
`ncu --set full -k regex:gemm_planes` (see profiles/README.md).  Launch order: f32out+fp32 res, f32out+planes res,
rstats+fp32 res, rstats+planes res, rstats+deferred res, planes plain A, planes deferred A."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lamp_b200 import ops  # noqa: E402

DEV = 'cuda'
prec = 0
M, N, K = 1024 * 103, 512, 512
torch.manual_seed(0)
a = ops.Act(None, *ops.split(torch.randn(M, K, device=DEV), prec), M, K)
w_hi, w_lo = ops.split(torch.randn(N, K, device=DEV) / K ** 0.5, prec)
rf = torch.randn(M, N, device=DEV)
g, b, bias = torch.ones(N, device=DEV), torch.zeros(N, device=DEV), torch.zeros(N, device=DEV)
res_f = ops.Act(rf, None, None, M, N)
res_p = ops.Act(None, *ops.split(rf, prec), M, N)
res_d = ops.linear_residual_deferred(a, w_hi, w_lo, N, prec, res_p, g, b, 1e-5, bias=bias)
torch.cuda.synchronize()
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    ops.linear_residual_f32(a, w_hi, w_lo, N, prec, res_f, bias=bias)
    ops.linear_residual_f32(a, w_hi, w_lo, N, prec, res_p, bias=bias)
    ops.linear_residual_deferred(a, w_hi, w_lo, N, prec, res_f, g, b, 1e-5, bias=bias)
    ops.linear_residual_deferred(a, w_hi, w_lo, N, prec, res_p, g, b, 1e-5, bias=bias)
    ops.linear_residual_deferred(a, w_hi, w_lo, N, prec, res_d, g, b, 1e-5, bias=bias)
    wp = ops.WeightPlanes()
    w2 = torch.nn.Parameter(torch.randn(N, K, device=DEV) / K ** 0.5)
    b2 = torch.nn.Parameter(torch.zeros(N, device=DEV))
    ops.project(res_p, wp, 'w', (w2,), N, prec, bias=b2, relu=True)
    ops.project(res_d, wp, 'w', (w2,), N, prec, bias=b2, relu=True)
torch.cuda.synchronize()
print('done')
