"""The UNMODIFIED reference (QData/LaMP) as the baseline arm -- on the host cores and, as plain PyTorch, on the GPU.

The reference is pure Python with no ``setup.py``: "installing" it is placing its tree under ``baseline/_ref/``
(git-ignored, so no reference source enters the history; NOT gpurun-ignored, so it travels to the GPU box with the
snapshot).  ``install()`` does that from ``/root/reference`` when that checkout is present (the build container) and
is called by ``__graft_entry__.build()``.  Nothing here is imported by ``lamp_b200``: this is bench/test
infrastructure (``bench.py --impl reference``, ``bench.py``'s ``cpu_baseline`` / ``torch_gpu_baseline`` legs, tests).

The reference predates torch 1.0; three runtime shims, applied by monkey-patching torch around the
calls and undone afterwards -- no reference file is edited:
  1. ``Tensor.cuda`` -> identity              (CPU runs only: Decoders.py:132,141 call ``.cuda()`` unconditionally)
  2. ``Tensor.byte`` -> ``Tensor.bool``       (``masked_fill`` rejects uint8 masks on torch >= 2; Decoders.py:141)
  3. ``torch.load(weights_only=False)``       (main.py:23 loads a pickled dict; only needed for main.py runs)
"""
from __future__ import annotations

import contextlib
import importlib
import os
import shutil
import sys
import time
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
SOURCE = os.environ.get('LAMP_REFERENCE', '/root/reference')


def install(force: bool = False) -> Optional[str]:
    """Copy the reference checkout to ``baseline/_ref`` (build container only).  -> the path, or None if no source."""
    if os.path.isdir(REF_DIR) and not force and os.path.exists(os.path.join(REF_DIR, 'lamp', 'Models.py')):
        return REF_DIR
    if not os.path.isdir(SOURCE):
        return None
    if os.path.isdir(REF_DIR):
        shutil.rmtree(REF_DIR)
    shutil.copytree(SOURCE, REF_DIR, ignore=shutil.ignore_patterns('.git', '__pycache__', '*.png'))
    for root, dirs, files in os.walk(REF_DIR):  # the source mount is read-only: make the copy removable
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs or n.endswith('.py') else 0o644)
    return REF_DIR


def ref_dir() -> Optional[str]:
    """Where the reference can be imported from: the travelled copy, else the container's read-only checkout."""
    for d in (REF_DIR, SOURCE):
        if os.path.exists(os.path.join(d, 'lamp', 'Models.py')):
            return d
    return None


@contextlib.contextmanager
def shims(cpu: bool):
    saved_cuda, saved_byte = torch.Tensor.cuda, torch.Tensor.byte
    try:
        if cpu:
            torch.Tensor.cuda = lambda self, *a, **k: self
        torch.Tensor.byte = lambda self, *a, **k: self.bool()
        yield
    finally:
        torch.Tensor.cuda, torch.Tensor.byte = saved_cuda, saved_byte


def import_reference():
    """-> the reference's ``lamp.Models.LAMP`` class (imported from ``ref_dir()``)."""
    d = ref_dir()
    if d is None:
        raise RuntimeError('reference not available: neither baseline/_ref nor /root/reference exists')
    if d not in sys.path:
        sys.path.insert(0, d)
    mod = importlib.import_module('lamp.Models')
    if not os.path.abspath(mod.__file__).startswith(os.path.abspath(d)):
        raise RuntimeError(f'`lamp` resolved to {mod.__file__}, not to the reference under {d}')
    return mod.LAMP


def build_model(cfg: dict, params: dict, adj, device) -> torch.nn.Module:
    """The reference model at ``cfg`` (keys L T V D d_inner H n_enc n_dec mask) with the state dict ``params``."""
    LAMP = import_reference()
    d = cfg['D'] // cfg['H']
    import io
    with contextlib.redirect_stdout(io.StringIO()):  # the constructor prints ('using prior mask')
        m = LAMP(cfg['V'] + 4, cfg['L'], cfg['T'], cfg['L'], n_layers_enc=cfg['n_enc'], n_layers_dec=cfg['n_dec'],
                 n_head=cfg['H'], n_head2=cfg['H'], d_word_vec=cfg['D'], d_model=cfg['D'], d_inner_hid=cfg['d_inner'],
                 d_k=d, d_v=d, dropout=0.2, dec_dropout=0.2, dec_dropout2=False, proj_share_weight=True,
                 encoder='graph', decoder='graph', label_adj_matrix=None if adj is None else adj.clone(),
                 label_mask=cfg['mask'])
    m.load_state_dict(params, strict=True)
    return m.to(device).eval()


def forward_throughput(cfg: dict, params: dict, adj, src_seq, src_pos, device, steps: int, warmup: int):
    """Eval ``LAMP.forward`` of the unmodified reference (lamp/Models.py:110-137) -> (samples/s, ms/step, logits).
    CPU: host clock, all threads.  CUDA: fp32 with TF32 disabled (matmul and cuDNN), CUDA events."""
    dev = torch.device(device)
    cpu = dev.type == 'cpu'
    batch = src_seq.shape[0]
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False  # the FFN is Conv1d(k=1): cuDNN would otherwise use TF32
    try:
        with shims(cpu), torch.no_grad():
            model = build_model(cfg, params, adj, dev)
            seq, pos = src_seq.to(dev), src_pos.to(dev)
            for _ in range(warmup):
                out = model((seq, pos), None, None, None)
            if cpu:
                t0 = time.perf_counter()
                for _ in range(steps):
                    out = model((seq, pos), None, None, None)
                dt = time.perf_counter() - t0
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(dev)
                e0.record()
                for _ in range(steps):
                    out = model((seq, pos), None, None, None)
                e1.record()
                torch.cuda.synchronize(dev)
                dt = e0.elapsed_time(e1) * 1e-3
        return steps * batch / dt, dt / steps * 1e3, out[0].detach().float().cpu()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
