import sys, os, torch
sys.path.insert(0, '/root/repo')
from lamp_b200 import synthetic as syn
B, T = 1100, 300
src_seq, src_pos = syn.make_tokens(B, T, 20000, 1)
src_seq, src_pos = src_seq.cuda(), src_pos.cuda()
def prep():
    R = B * T
    seq_flat = src_seq.reshape(-1)
    is_pad = seq_flat.eq(0)
    rep = is_pad & src_pos.reshape(-1).eq(0)
    order = torch.argsort(rep.to(torch.uint8), stable=True)
    n_keep = R - rep.sum()
    m_dev = torch.clamp(n_keep + 1, max=R).to(torch.int32).reshape(1)
    rank = torch.cumsum((~rep).to(torch.int64), 0) - 1
    src_row = torch.where(rep, n_keep.to(torch.int64), rank)
    kv_len = (~rep).view(B, T).sum(dim=1)
    kv_start = torch.cumsum(kv_len, 0) - kv_len
    a = kv_start.to(torch.int32); b = kv_len.to(torch.int32); c = is_pad[order].to(torch.uint8)
    return order, m_dev, src_row, a, b, c
for _ in range(5): prep()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = prep()
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(50): g.replay()
e1.record(); torch.cuda.synchronize()
print('index prep (graph replay): %.1f us' % (e0.elapsed_time(e1) / 50 * 1e3))
e0.record()
for _ in range(50): prep()
e1.record(); torch.cuda.synchronize()
print('index prep (eager): %.1f us' % (e0.elapsed_time(e1) / 50 * 1e3))
