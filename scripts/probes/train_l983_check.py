"""Training step at cfg-4 dims (L = 983, prior mask): parameter gradients of the native training path and of the
all-torch fp32 composition, each against the same composition in fp64."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import cases
from lamp_b200 import ops, synthetic as syn
from lamp_b200.Models import LAMP
L, T, D, H, B = 983, 150, 512, 4, 8
params = syn.lamp_params(1004, L, T, D, 512, H, 1, 2, seed=0)
adj = cases.label_adj('prior', L, 1)
src_seq, src_pos = syn.make_tokens(B, T, 1000, 1)
res = {}
for mode in ('native', 'composed', 'fp64'):
    ops.NATIVE_ATTENTION_BACKWARD = ops.NATIVE_TRAINING = (mode == 'native')
    m = LAMP(1004, L, T, L, n_layers_enc=1, n_layers_dec=2, n_head=H, n_head2=H, d_word_vec=D, d_model=D, d_inner_hid=512,
             d_k=D // H, d_v=D // H, dropout=0.0, dec_dropout=0.0, dec_dropout2=False, proj_share_weight=True, encoder='graph',
             decoder='graph', label_adj_matrix=adj, label_mask='prior')
    m.load_state_dict(params, strict=True)
    m = m.cuda().train()
    if mode == 'fp64':
        m = m.double()
    logits, _, _ = m((src_seq.cuda(), src_pos.cuda()), None, None, None)
    torch.nn.functional.binary_cross_entropy_with_logits(logits, torch.zeros_like(logits)).backward()
    torch.cuda.synchronize()
    res[mode] = {n: p.grad.double().clone() for n, p in m.named_parameters() if p.grad is not None}
ops.NATIVE_ATTENTION_BACKWARD = ops.NATIVE_TRAINING = True
for mode in ('native', 'composed'):
    errs = sorted(((float((res[mode][n] - res['fp64'][n]).abs().max() / res['fp64'][n].abs().max().clamp_min(1e-30)), n)
                   for n in res['fp64']), reverse=True)
    print(mode, 'vs fp64: worst', errs[0], 'median %.2e' % errs[len(errs) // 2][0])
