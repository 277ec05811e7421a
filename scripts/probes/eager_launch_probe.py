"""Launch-path probe: eval forward at B=1100 launched eagerly / with a CUDA-event pair per native call / as a graph replay /
through the drop-in call, with allocator statistics (found the per-kernel timing mode holding activations alive).
usage: python scripts/probes/eager_launch_probe.py [batch]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
import lamp_b200
from lamp_b200 import ops, graphs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1100
dev = torch.device('cuda', 0)
params, adj, src_seq, src_pos = bench.synth(B, 100)
model = bench.build_model(bench.CFG, params, adj, dev).eval()
seq, pos = src_seq.to(dev), src_pos.to(dev)
eager = model._forward_impl

def t(label, fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter(); e0.record()
    hs = []
    for _ in range(n):
        a = time.perf_counter(); fn(); hs.append(time.perf_counter() - a)
    e1.record(); torch.cuda.synchronize()
    print(f'{label}: device {e0.elapsed_time(e1)/n:.3f} ms/step, host median {sorted(hs)[n//2]*1e3:.3f} max {max(hs)*1e3:.1f} ms, '
          f'mem alloc {torch.cuda.memory_allocated()/2**30:.1f} GiB reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB', flush=True)

with torch.no_grad():
    for _ in range(3): eager((seq, pos), None, None, None)
    t('eager fresh', lambda: eager((seq, pos), None, None, None))
    with ops.STATS.timed():
        t('eager fresh, event pairs', lambda: eager((seq, pos), None, None, None))
        ops.STATS.stop_timing()
    runner = lamp_b200.GraphedForward(model, B, 300, example=(seq, pos))
    t('graphed replay', lambda: runner.replay())
    t('eager after GraphedForward', lambda: eager((seq, pos), None, None, None))
    for _ in range(3): model((seq, pos), None, None, None)
    t('dropin', lambda: model((seq, pos), None, None, None))
    t('eager after dropin capture', lambda: eager((seq, pos), None, None, None))
    with ops.STATS.timed():
        t('eager after dropin capture, event pairs', lambda: eager((seq, pos), None, None, None))
        ops.STATS.stop_timing()
    import gc
    gc.collect(); gc.disable()
    t('eager, gc disabled', lambda: eager((seq, pos), None, None, None))
    with ops.STATS.timed():
        t('eager gc disabled, event pairs', lambda: eager((seq, pos), None, None, None), n=20)
        ops.STATS.stop_timing()
