import sys, os, cProfile, pstats, io, torch
sys.path.insert(0, '/root/repo')
from lamp_b200 import ops, synthetic as syn
from lamp_b200.Models import LAMP
c = dict(L=103, T=300, V=20000, D=512, d_inner=512, H=4, n_enc=2, n_dec=2)
B = 32
params = syn.lamp_params(c['V'] + 4, c['L'], c['T'], c['D'], c['d_inner'], c['H'], c['n_enc'], c['n_dec'], seed=0)
adj = syn.prior_adjacency(syn.make_label_sets(c['L'], seed=0), c['L'])
src_seq, src_pos = syn.make_tokens(B, c['T'], c['V'], 1)
src = (src_seq.cuda(), src_pos.cuda())
tgt = (torch.rand(B, c['L'], device='cuda') < 0.05).float()
d = c['D'] // c['H']
model = LAMP(c['V'] + 4, c['L'], c['T'], c['L'], n_layers_enc=2, n_layers_dec=2, n_head=4, n_head2=4, d_word_vec=512, d_model=512,
             d_inner_hid=512, d_k=d, d_v=d, dropout=0.2, dec_dropout=0.2, dec_dropout2=False, proj_share_weight=True,
             encoder='graph', decoder='graph', label_adj_matrix=adj, label_mask='prior')
model.load_state_dict(params, strict=True)
model = model.cuda().train()
def step():
    model.zero_grad(set_to_none=True)
    logits, _, _ = model(src, None, None, None)
    torch.nn.functional.binary_cross_entropy_with_logits(logits, tgt).backward()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:4500])
