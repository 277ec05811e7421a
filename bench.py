#!/usr/bin/env python
"""Benchmark of the label-graph forward (BASELINE.json metric: "label-graph forward samples/sec at L=103
d_model=512; HBM GB/s vs roofline").

A *step* is one ``LAMP.forward`` (graph encoder -> GraphDecoder label message passing -> diagonal label
projection; lamp/Models.py:110-137) in eval mode over one batch of synthetic documents at BASELINE cfg-1/2 dims:
L=103 labels, T=300 tokens, d_model=512, n_head=4, 2+2 layers, d_inner=512, prior label mask, fp32 in/out
(LAMP_PREC_FP32: 3-term split-bf16 tensor-core products, parity-checked at 1e-3 against the reference).

  value : samples/s, whole job, token ids already resident in HBM, CUDA-event timed, max over ranks.  Each step is one
          replay of the CUDA graph of LAMP.forward (lamp_b200.GraphedForward, the package's serving API; --no-graph
          launches kernel by kernel from Python instead)
  dropin: the same K steps through the reference-facing call itself -- ``model((src_seq, src_pos), None, None, None)``
          as test.py:41 makes it -- which LAMP.forward serves from its own shape-keyed CUDA-graph cache
  e2e   : the metric through that drop-in call with HOST (pinned) token ids: H2D copy + forward + D2H of the logits
          inside the timed region, host clock
  roofline      : dominant kernel (projection GEMM, tensor-bound) -- algorithmic FLOPs / measured kernel time.  The
                  per-kernel durations come from a second timed region of the same K steps, launched eagerly with a
                  CUDA-event pair around every native call (`eager_ms_per_step` is that region's step time)
  roofline_layernorm : the standalone LayerNorm launches of the step (streaming, 8 B/element): in-step HBM yardstick
  roofline_attn : masked label<-label attention core kernel under the label-graph mask (HBM-bound) -- algorithmic
                  bytes / measured kernel time; roofline_attn_enc: the same kernel on the label<-input shape (T=300)
  train         : the training step north_star's multi-GPU clause names -- forward + BCE + backward + bucketed,
                  backward-overlapped NCCL gradient all-reduce (lamp_b200.distributed.GradientReducer, replacing
                  nn.DataParallel of main.py:106-108) + Adam step (main.py:99) -- at cfg-1 dims and at cfg-4 dims
                  (L=983, 4 decoder layers), whole-job samples/s, with the all-reduce's own and exposed time
  cpu_baseline  : the UNMODIFIED reference (baseline/_ref, see baseline/reference.py) timed on the host cores on a
                  bounded sample ("kind": "reference"); the oracle port only if the reference tree did not travel
  torch_gpu_baseline : the unmodified reference modules as plain PyTorch on the same GPU (fp32, TF32 off)
  --impl reference : only the CPU arm, same config, "impl": "reference"

Launch: ``python bench.py --gpus N --steps K --warmup W`` (N > 1 under torchrun, one rank per GPU; the batch is
sharded, the label graph and weights replicated, no forward collective -> "scaling": "weak").
"""
import argparse
import gc
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's own banner / debug output (NCCL_DEBUG=VERSION on some boxes) goes to stderr
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':  # the version banner is a bare printf to stdout
    os.environ['NCCL_DEBUG'] = 'WARN'

CFG = dict(L=103, T=300, V=20000, D=512, d_inner=512, H=4, n_enc=2, n_dec=2, mask='prior', seed=0)
# BASELINE.json configs[3]: "delicious L=983 -label_mask prior d_model=512 n_layers_dec=4, batch-sharded ... NCCL grad allreduce"
CFG4 = dict(L=983, T=300, V=20000, D=512, d_inner=512, H=4, n_enc=2, n_dec=4, mask='prior', seed=4)
METRIC = 'label-graph forward samples/sec at L=103 d_model=512'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=1100,
                    help='samples per GPU per step (1100 x 103 label rows / x 300 token rows fill whole waves of the '
                         '74 CTA-pair tile scheduler)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'bf16'])
    ap.add_argument('--cpu-batch', type=int, default=32)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-torch-gpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step blocks')
    ap.add_argument('--train-batch', type=int, default=256, help='training samples per GPU per step, cfg-1 dims')
    ap.add_argument('--train-batch-cfg4', type=int, default=32, help='training samples per GPU per step, cfg-4 dims')
    ap.add_argument('--train-steps', type=int, default=8)
    ap.add_argument('--no-graph', action='store_true', help='launch every step kernel by kernel from Python')
    ap.add_argument('--tune', action='append', default=[], metavar='KEY=VALUE',
                    help='lamp_set_tuning knob (include/lamp_b200.h), for experiments only')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], tensor=p.get('bf16_tflops_sustained', p['bf16_tflops']), src='measured')
    return dict(hbm=6650.0, tensor=1400.0, src='fallback')


def lib_source_hash():
    """sha256 over the native sources the loaded library was built from (stable across rebuilds of the same tree);
    the ncu DRAM-traffic capture under profiles/ records the same hash, and a mismatch marks it stale."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'lamp_b200', 'csrc')
    for name in sorted(os.listdir(d)) + ['../../include/lamp_b200.h']:
        with open(os.path.join(d, name), 'rb') as f:
            h.update(name.encode() + b'\0' + f.read())
    return h.hexdigest()[:16]


def synth(batch, seed, c=CFG):
    from lamp_b200 import synthetic as syn
    params = syn.lamp_params(c['V'] + 4, c['L'], c['T'], c['D'], c['d_inner'], c['H'], c['n_enc'], c['n_dec'],
                             seed=c['seed'])
    adj = syn.prior_adjacency(syn.make_label_sets(c['L'], seed=c['seed']), c['L'])
    src_seq, src_pos = syn.make_tokens(batch, c['T'], c['V'], seed)
    return params, adj, src_seq, src_pos


def build_model(c, params, adj, dev):
    from lamp_b200.Models import LAMP
    d = c['D'] // c['H']
    model = LAMP(c['V'] + 4, c['L'], c['T'], c['L'], n_layers_enc=c['n_enc'], n_layers_dec=c['n_dec'], n_head=c['H'],
                 n_head2=c['H'], d_word_vec=c['D'], d_model=c['D'], d_inner_hid=c['d_inner'], d_k=d, d_v=d,
                 dropout=0.2, dec_dropout=0.2, dec_dropout2=False, proj_share_weight=True, encoder='graph',
                 decoder='graph', label_adj_matrix=adj, label_mask=c['mask'])
    model.load_state_dict(params, strict=True)
    return model.to(dev)


def cpu_reference_throughput(steps, warmup, batch):
    """The reference on the host cores, eval, fp32, all threads: the UNMODIFIED reference tree from baseline/_ref
    when it travelled with the snapshot (kind "reference"), the oracle port of it otherwise (kind "port")."""
    torch.set_num_threads(os.cpu_count() or 1)
    params, adj, src_seq, src_pos = synth(batch, 1234)
    from baseline import reference as ref
    if ref.ref_dir() is not None:
        import warnings
        warnings.filterwarnings('ignore')
        v, ms, _ = ref.forward_throughput(CFG, params, adj, src_seq, src_pos, 'cpu', steps, warmup)
        return v, ms, torch.get_num_threads(), 'reference', 'unmodified reference (baseline/_ref: lamp.Models.LAMP, torch CPU, MKL)'
    from oracle import lamp_oracle as orc
    lm = orc.label_mask_from(CFG['L'], adj, 'prior')
    cfg = dict(n_layers_enc=CFG['n_enc'], n_layers_dec=CFG['n_dec'], n_head=CFG['H'], n_head2=CFG['H'])
    with torch.no_grad():
        for _ in range(warmup):
            orc.lamp_forward(params, cfg, src_seq, src_pos, lm, compute_dead_attention=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.lamp_forward(params, cfg, src_seq, src_pos, lm, compute_dead_attention=True)
        dt = time.perf_counter() - t0
    return steps * batch / dt, dt / steps * 1e3, torch.get_num_threads(), 'port', 'oracle port of the reference (torch CPU, MKL)'


class ClockSampler:
    """SM clock and throttle reasons sampled in-process through NVML (nvidia_ml_py) every 10 ms while the timed region
    runs; falls back to an `nvidia-smi -lms` child when NVML cannot be loaded."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self._stop = [], None, None, threading.Event()
        self.smax = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if visible:
                ids = [x.strip() for x in visible.split(',') if x.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [('hw_slowdown', n.nvmlClocksEventReasonHwSlowdown if hasattr(n, 'nvmlClocksEventReasonHwSlowdown')
                 else n.nvmlClocksThrottleReasonHwSlowdown),
                ('hw_thermal_slowdown', getattr(n, 'nvmlClocksEventReasonHwThermalSlowdown',
                                                getattr(n, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40))),
                ('sw_thermal_slowdown', getattr(n, 'nvmlClocksEventReasonSwThermalSlowdown',
                                                getattr(n, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20))),
                ('sw_power_cap', getattr(n, 'nvmlClocksEventReasonSwPowerCap',
                                         getattr(n, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)))]
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.perf_counter(), sm, [name for name, bit in bits if mask & bit]))
            except Exception:
                pass
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1.0)
            inside = [(sm, rs) for ts, sm, rs in self.rows if t0 <= ts <= t1]
            if not inside:
                return dict(sm_mhz=None, sm_max_mhz=self.smax, reasons=[], samples=0, source='nvml')
            reasons = sorted({r for _, rs in inside for r in rs})
            return dict(sm_mhz=statistics.median(sm for sm, _ in inside), sm_max_mhz=self.smax, reasons=reasons,
                        samples=len(inside), source='nvml')
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.rows:
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 7 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(parts[0]))
                smax = max(smax, float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=smax or None, reasons=[], samples=0, source='nvidia-smi')
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm),
                    source='nvidia-smi')


def train_block(c, name, batch, steps, warmup, rank, world, dev, barrier, max_over_ranks):
    """Training step of the label-graph model (train.py:28-48 + main.py:99,106-108): zero_grad -> LAMP.forward (train
    mode, dropout 0.2) -> BCE-with-logits -> backward, gradients accumulated straight into the flat communication
    buffer and all-reduced bucket by bucket from backward hooks -> Adam(betas=(0.9, 0.98)).  Inputs and targets are
    resident.  Timed three ways: the full step, the same step with the collectives switched off (their exposed cost is
    the difference) and the all-reduce of the whole flat buffer alone."""
    import torch.distributed as dist
    import torch.nn.functional as F
    from lamp_b200 import distributed as lds
    from lamp_b200 import ops
    params, adj, src_seq, src_pos = synth(batch, 500 + rank, c)
    model = build_model(c, params, adj, dev).train()
    torch.manual_seed(1000 + rank)  # per-rank dropout streams
    seq, pos = src_seq.to(dev), src_pos.to(dev)
    # label-id rows as train.py receives them (tgt[:, 1:]: ids + 4, EOS, PAD): the multi-hot targets are built on the
    # device every step (ops.gold_binary, the kernel behind the drop-in utils.get_gold_binary), as train.py:34 does
    from lamp_b200 import synthetic as syn
    rows = syn.make_label_sets(c['L'], n_docs=batch, seed=700 + rank)
    width = max(len(r) for r in rows) - 1
    gold_rows = torch.zeros((batch, width), dtype=torch.int64)
    for i, r in enumerate(rows):
        gold_rows[i, :len(r) - 1] = torch.tensor(r[1:])
    gold_rows = gold_rows.to(dev)
    plist = list(model.get_trainable_parameters())
    opt = torch.optim.Adam(plist, betas=(0.9, 0.98), lr=2e-4, fused=True)
    red = lds.GradientReducer(plist, world=world)

    def step():
        red.zero_grad()
        gold = ops.gold_binary(gold_rows, c['L'])
        logits, _, _ = model((seq, pos), None, None, gold)
        loss = ops.bce_with_logits(logits, gold)   # value + gradient in one kernel (train.py:38)
        loss.backward()
        red.finish()
        opt.step()
        return loss

    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            loss = step()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / n, loss

    loss0 = float(step().detach())  # calibrates the reducer (dead-parameter layout, flat views)
    for _ in range(max(warmup, 2)):
        step()
    ops.STATS.reset()
    ms, loss = timed(steps)
    launches = ops.STATS.launches // steps
    out = dict(workload=name, metric='training samples/s (fwd + BCE + bwd + gradient all-reduce + Adam)',
               batch_per_gpu=batch, global_batch=batch * world, steps=steps, ms_per_step=ms,
               value=batch * world / ms * 1e3, unit='samples/s', loss_first=loss0, loss_last=float(loss.detach()),
               native_launches_per_step=launches, gradient_elements=red.stats['elements'],
               allreduce_bytes=red.stats['elements'] * 4, buckets=red.stats['buckets'], launch='eager')
    if world > 1:
        red.communicate = False
        ms_nocomm, _ = timed(steps)
        red.communicate = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(10):
            dist.all_reduce(red.flat)
        e1.record()
        barrier()
        ar = max_over_ranks(e0.elapsed_time(e1)) / 10
        exposed = max(ms - ms_nocomm, 0.0)
        out.update(ms_per_step_without_allreduce=ms_nocomm, allreduce_ms=ar, allreduce_exposed_ms=exposed,
                   allreduce_overlap_frac=max(0.0, 1.0 - exposed / ar) if ar > 0 else None,
                   allreduce_busbw_gbs=red.stats['elements'] * 4 * 2 * (world - 1) / world / (ar * 1e-3) / 1e9,
                   buckets_from_hooks=red.stats['launched_from_hooks'], buckets_at_finish=red.stats['launched_at_finish'])
    red.remove()
    del model, opt, red
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    config = dict(workload='LAMP.forward eval, synthetic cfg-1/2 dims: L=103 T=300 d_model=512 n_head=4 '
                           'n_layers 2+2 d_inner=512 prior label mask, fp32 I/O',
                  batch_per_gpu=args.batch, global_batch=args.batch * world, seq_len=CFG['T'], n_labels=CFG['L'],
                  precision=args.precision, parallelism=f'batch-sharded x{world}, replicated label graph',
                  l2='activations per step (GBs) exceed the 126 MB L2; no explicit flush')

    if args.impl == 'reference':
        if rank != 0:
            return
        v, ms, cores, kind, what = cpu_reference_throughput(args.steps, max(args.warmup, 1), args.cpu_batch)
        config['batch_per_gpu'] = config['global_batch'] = args.cpu_batch
        print(json.dumps(dict(
            metric=METRIC, value=v, unit='samples/s', n_gpus=args.gpus,
            steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling='weak',
            vs_baseline=None, dtype='f32', data='synthetic', impl='reference', config=config,
            cpu_baseline=dict(value=v, unit='samples/s', cores=cores, kind=kind,
                              sample=f'{args.steps} x LAMP.forward on B={args.cpu_batch} synthetic documents, '
                                     f'{what}, all host threads'),
            e2e=dict(value=v, unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import lamp_b200
    from lamp_b200 import graphs, ops
    lamp_b200.set_default_precision(args.precision)
    for kv in args.tune:
        k, v = kv.split('=')
        from lamp_b200 import _native as nat
        nat.check(nat.lib().lamp_set_tuning(int(k), int(v)), 'tune')
        config['tune'] = args.tune
    dev = torch.device('cuda', local_rank)

    params, adj, src_seq, src_pos = synth(args.batch, 100 + rank)
    c = CFG
    model = build_model(c, params, adj, dev).eval()
    seq_d, pos_d = src_seq.to(dev), src_pos.to(dev)
    seq_h, pos_h = src_seq.pin_memory(), src_pos.pin_memory()
    logits_h = torch.empty((args.batch, c['L']), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None  # polls from here on; the timed window is cut out later
    runner, launch_mode = None, 'eager'
    eager = model._forward_impl  # the plain launch sequence (no graph of any kind)
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            eager((seq_d, pos_d), None, None, None)
        if not args.no_graph:
            # the public serving API: one CUDA-graph launch per step (lamp_b200/graphs.py); same kernels as eager
            runner = lamp_b200.GraphedForward(model, args.batch, c['T'], example=(seq_d, pos_d))
            launch_mode = f'cuda-graph replay ({runner.kernels_per_replay} native kernels per step)'
            for _ in range(max(args.warmup, 3)):
                runner.replay()
        config['launch'] = launch_mode
        for _ in range(max(args.warmup, 3)):  # the drop-in call: first call eager, second captures, then replays
            model((seq_d, pos_d), None, None, None)
        gc.collect()
        gc.disable()  # a generation-2 collection inside a timed region costs a few hundred ms of launch stall
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # ---------------- (A) device-resident throughput: token ids already in HBM, K steps
        barrier()
        ops.STATS.reset()
        t_wall0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            if runner is not None:
                logits, _ = runner.replay()
            else:
                logits, _, _ = eager((seq_d, pos_d), None, None, None)
        e1.record()
        barrier()
        launches = ops.STATS.launches
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        # ---------------- (A2) the same K steps through the drop-in call the reference's loops make (test.py:41)
        barrier()
        e0.record()
        for _ in range(args.steps):
            logits, _, _ = model((seq_d, pos_d), None, None, None)
        e1.record()
        barrier()
        dropin_ms = max_over_ranks(e0.elapsed_time(e1))
        cache = model.__dict__.get('_eval_graphs')
        # (A') region A once more: the board is power-capped (SM clocks sag over the first few hundred ms of tensor
        # work), so the drop-in call is compared with a graph replay timed in the SAME thermal state
        replay_after_ms = None
        if runner is not None:
            barrier()
            e0.record()
            for _ in range(args.steps):
                runner.replay()
            e1.record()
            barrier()
            replay_after_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        # ---------------- (B) the same K steps launched kernel by kernel with a CUDA-event pair around every native
        #                  call: per-kernel durations for the roofline figures (and the eager-launch step time)
        barrier()
        with ops.STATS.timed():
            e0.record()
            host_t = [time.perf_counter()]
            for _ in range(args.steps):
                eager((seq_d, pos_d), None, None, None)
                host_t.append(time.perf_counter())
            e1.record()
            barrier()
            t_wall1 = time.perf_counter()
            per_kernel = ops.STATS.stop_timing()
        eager_ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
        # ---------------- (C) end to end through the drop-in call: pinned host ids -> H2D -> forward -> D2H logits
        def e2e_step():
            lg, _, _ = model((seq_h.to(dev, non_blocking=True), pos_h.to(dev, non_blocking=True)), None, None, None)
            logits_h.copy_(lg, non_blocking=True)
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)  # host clock: launch overheads and the final D2H included
        e2e_dev_ms = e0.elapsed_time(e1)
        gc.enable()

    total_samples = args.batch * world * args.steps
    value = total_samples / (ms_total * 1e-3)
    e2e_value = total_samples / e2e_s
    pk = peaks()

    # measured DRAM traffic per launch (dram__bytes_read + dram__bytes_write) from the committed ncu capture of this
    # same command (profiles/r02_traffic.json, written by scripts/ncu_traffic.py).  It is only quoted when it was taken
    # from the SAME native sources (hash), batch and precision as the library loaded now; otherwise null + stale flag.
    traffic, traffic_stale = {}, None
    tpath = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if os.path.exists(tpath) and world == 1:
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get('batch') == args.batch and tj.get('precision') == args.precision:
            if tj.get('lib_source_hash') == lib_source_hash():
                traffic = tj.get('avg_bytes_per_launch', {})
                traffic_stale = False
            else:
                traffic_stale = True

    def roof(name, bound):
        k = per_kernel.get(name)
        if not k or k['ms'] <= 0:
            return None
        sec = k['ms'] * 1e-3
        if bound == 'tensor':
            ach, peak, unit = k['flops'] / sec / 1e12, pk['tensor'], 'TFLOP/s'
        else:
            ach, peak, unit = k['bytes'] / sec / 1e9, pk['hbm'], 'GB/s'
        return dict(kernel=name, bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak,
                    traffic=traffic.get(name), traffic_stale=traffic_stale,
                    alg_bytes_per_launch=k['bytes'] / k['calls'],
                    peak_source=pk['src'], launches=k['calls'], avg_launch_ms=k['ms'] / k['calls'],
                    share_of_step=k['ms'] / sum(v['ms'] for v in per_kernel.values()),
                    alg_gbs=k['bytes'] / sec / 1e9, alg_tflops=k['flops'] / sec / 1e12)

    out = dict(
        metric=METRIC, value=value, unit='samples/s', n_gpus=world,
        steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_total / args.steps, higher_is_better=True,
        scaling='weak', vs_baseline=None, dtype='f32' if args.precision == 'fp32' else 'bf16', data='synthetic',
        config=config,
        e2e=dict(value=e2e_value, unit='samples/s',
                 h2d_bytes_per_step=int(seq_h.numel() * 8 + pos_h.numel() * 8) * world,
                 d2h_bytes_per_step=int(logits_h.numel() * 4) * world, ms_per_step=e2e_s / args.steps * 1e3,
                 device_ms_per_step=e2e_dev_ms / args.steps,
                 api='model((src_seq, src_pos), None, None, None) -- the reference-facing LAMP.forward call'),
        dropin=dict(value=total_samples / (dropin_ms * 1e-3), unit='samples/s', ms_per_step=dropin_ms / args.steps,
                    api='LAMP.forward (eval graph cache)', enabled=graphs.EVAL_GRAPHS,
                    graph_replay_ms_per_step_timed_right_after=replay_after_ms,
                    replays=None if cache is None else cache.replays,
                    captures=None if cache is None else cache.captures),
        gpu_launches=launches, clocks=clocks,
        eager_ms_per_step=eager_ms / args.steps,
        host_launch_ms=dict(median=statistics.median(b - a for a, b in zip(host_t, host_t[1:])) * 1e3,
                            max=max(b - a for a, b in zip(host_t, host_t[1:])) * 1e3),
        roofline=roof('gemm_planes', 'tensor'), roofline_attn=roof('attn_core_self', 'hbm'),
        roofline_attn_enc=roof('attn_core_enc', 'hbm'),
        roofline_layernorm=roof('layernorm', 'hbm'),   # pure streaming kernel (0.92-1.0 of the peak when timed alone): the
                                                       # HBM fraction a kernel can reach INSIDE this power-capped step
        kernels={k: dict(calls=v['calls'], ms=round(v['ms'], 3)) for k, v in per_kernel.items()},
        lib_source_hash=lib_source_hash())
    if args.precision != 'fp32':
        config['tolerance_note'] = 'bf16 operands: outside the 1e-3 fp32 contract (tests hold this mode to 3e-2)'

    # ---------------- training step (forward + loss + backward + gradient all-reduce + Adam), every N
    if not args.no_train:
        del runner
        model.__dict__.pop('_eval_graphs', None)
        torch.cuda.empty_cache()
        out['train'] = train_block(CFG, 'cfg-1 dims: L=103 T=300 d_model=512 n_head=4 n_layers 2+2, dropout 0.2',
                                   args.train_batch, args.train_steps, 2, rank, world, dev, barrier, max_over_ranks)
        out['train_cfg4'] = train_block(CFG4, 'cfg-4 dims: L=983 T=300 d_model=512 n_head=4 n_layers 2+4 prior mask, '
                                              'dropout 0.2', args.train_batch_cfg4, max(args.train_steps // 2, 3), 2,
                                        rank, world, dev, barrier, max_over_ranks)

    if rank == 0:
        if world == 1 and not args.no_torch_gpu_baseline:
            # the unmodified reference as plain PyTorch on this same GPU (shims 2+3 of SURVEY 8c; fp32, TF32 off)
            from baseline import reference as ref
            if ref.ref_dir() is not None:
                import warnings
                warnings.filterwarnings('ignore')
                b = args.batch
                while True:
                    try:
                        v, ms, ref_logits = ref.forward_throughput(CFG, params, adj, src_seq[:b], src_pos[:b], dev, 5, 2)
                        break
                    except torch.OutOfMemoryError:
                        torch.cuda.empty_cache()
                        b //= 2
                with torch.no_grad():
                    ours = eager((seq_d[:b], pos_d[:b]), None, None, None)[0].float().cpu()
                out['torch_gpu_baseline'] = dict(
                    value=v, unit='samples/s', ms_per_step=ms, batch=b,
                    what='unmodified reference lamp.Models.LAMP.forward, eval, fp32 (TF32 off), torch '
                         f'{torch.__version__} on the same GPU',
                    max_rel_diff_vs_lamp_b200=float((ours - ref_logits).abs().max() / ref_logits.abs().max()))
        if world == 1 and not args.no_cpu_baseline:
            v, ms, cores, kind, what = cpu_reference_throughput(10, 2, args.cpu_batch)
            out['cpu_baseline'] = dict(value=v, unit='samples/s', cores=cores, kind=kind,
                                       sample=f'10 x LAMP.forward on B={args.cpu_batch} synthetic documents '
                                              f'({ms:.0f} ms each), {what}, all host threads')
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
