import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from lamp_b200.Models import LAMP
from lamp_b200 import ops
import lamp_b200
dev = torch.device('cuda', 0)
B = 1100
params, adj, src_seq, src_pos = bench.synth(B, 100)
c = bench.CFG; d = c['D'] // c['H']
model = LAMP(c['V'] + 4, c['L'], c['T'], c['L'], n_layers_enc=2, n_layers_dec=2, n_head=4, n_head2=4, d_word_vec=512, d_model=512,
             d_inner_hid=512, d_k=d, d_v=d, encoder='graph', decoder='graph', label_adj_matrix=adj, label_mask='prior')
model.load_state_dict(params); model = model.to(dev).eval()
seq_d, pos_d = src_seq.to(dev), src_pos.to(dev)
seq_h, pos_h = src_seq.pin_memory(), src_pos.pin_memory()
out_h = torch.empty((B, 103)).pin_memory()
def run(mode, steps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps):
        if mode == 'dev':
            lg, _, _ = model((seq_d, pos_d), None, None, None)
        elif mode == 'h2d':
            lg, _, _ = model((seq_h.to(dev, non_blocking=True), pos_h.to(dev, non_blocking=True)), None, None, None)
        elif mode == 'd2h':
            lg, _, _ = model((seq_d, pos_d), None, None, None); out_h.copy_(lg, non_blocking=True)
        elif mode == 'both':
            lg, _, _ = model((seq_h.to(dev, non_blocking=True), pos_h.to(dev, non_blocking=True)), None, None, None); out_h.copy_(lg, non_blocking=True)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'{mode:5s} aware={ops.PADDING_AWARE}: launch {1e3*(t1-t0)/steps:.2f} ms/step, total {1e3*(t2-t0)/steps:.2f} ms/step', flush=True)
with torch.no_grad():
    for aware in (True, False):
        ops.PADDING_AWARE = aware
        for _ in range(3): model((seq_d, pos_d), None, None, None)
        for mode in ('dev', 'h2d', 'd2h', 'both', 'dev'):
            run(mode)
