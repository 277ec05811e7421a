#!/usr/bin/env python
"""One epoch of the reference's own ``main.py`` (train + validation + test loops, metrics, checkpoint) on the GPU box,
(a) with the label-graph classes rebound to lamp_b200 (``python -m lamp_b200.run_main baseline/_ref ...``) and
(b) as stock PyTorch (``baseline/run_ref_main.py``, the unmodified reference with the torch>=2 shims only),
on the same synthetic ``train_valid_test.pt`` with BASELINE cfg-1's flags (README command: -batch_size 32 -d_model 512
-d_inner_hid 512 -n_layers_enc 2 -n_layers_dec 2 -n_head 4 -dropout 0.2 -dec_dropout 0.2 -lr 0.0002 -encoder graph
-decoder graph -label_mask prior).  Prints one JSON summary (epoch wall times and the B(CE) losses both arms print) and
keeps the raw logs.

usage: python scripts/run_reference_main.py [--out gpurun_out/r02_main_epoch] [--n-train 2048] [--arms dropin,reference]
       [--tiny]   (tiny dims for a quick functional check)   [--gpus 0,1]  (CUDA_VISIBLE_DEVICES for the child runs)
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lamp_b200 import synthetic as syn  # noqa: E402


def parse_log(text):
    """-> dict(train_min, valid_min, test_min, train_loss, valid_loss, test_loss) from the reference's prints."""
    out = {}
    for name in ('Training', 'Validation', 'Testing'):
        m = re.search(r'\(%s\) elapse: ([0-9.]+) min\s*\n\s*B : ([-0-9.e+naninf]+)' % name, text)
        if m:
            out[name.lower() + '_min'] = float(m.group(1))
            out[name.lower() + '_bce_per_doc'] = float(m.group(2))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'r02_main_epoch'))
    ap.add_argument('--n-train', type=int, default=2048)
    ap.add_argument('--n-eval', type=int, default=512)
    ap.add_argument('--epochs', type=int, default=1)
    ap.add_argument('--arms', default='dropin,reference')
    ap.add_argument('--tiny', action='store_true')
    ap.add_argument('--no-cuda', action='store_true', help='pass -no_cuda (CPU check of the reference arm)')
    ap.add_argument('--gpus', default=None, help='CUDA_VISIBLE_DEVICES for the child runs (default: inherit)')
    args = ap.parse_args()
    args.out = os.path.abspath(args.out)   # the child runs chdir into the reference tree
    os.makedirs(args.out, exist_ok=True)
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.exists(os.path.join(ref, 'main.py')):
        sys.exit('baseline/_ref is missing: run __graft_entry__.build() in the build container first')
    if args.tiny:
        dims = dict(L=12, V=50, T=20, batch=8, d=32, dh=32, layers=1, heads=2)
        n_train, n_eval = 64, 16
    else:
        dims = dict(L=103, V=20000, T=300, batch=32, d=512, dh=512, layers=2, heads=4)
        n_train, n_eval = args.n_train, args.n_eval
    dataroot = os.path.join(args.out, 'data')
    os.makedirs(os.path.join(dataroot, 'synth'), exist_ok=True)
    data = syn.make_dataset_dict(n_labels=dims['L'], vocab=dims['V'], n_train=n_train, n_valid=n_eval, n_test=n_eval,
                                 max_len=dims['T'], seed=0)
    torch.save(data, os.path.join(dataroot, 'synth', 'train_valid_test.pt'))
    flags = ['-dataroot', dataroot + '/', '-dataset', 'synth', '-batch_size', str(dims['batch']), '-d_model', str(dims['d']),
             '-d_inner_hid', str(dims['dh']), '-n_layers_enc', str(dims['layers']), '-n_layers_dec', str(dims['layers']),
             '-n_head', str(dims['heads']), '-epoch', str(args.epochs), '-dropout', '0.2', '-dec_dropout', '0.2',
             '-lr', '0.0002', '-encoder', 'graph', '-decoder', 'graph', '-label_mask', 'prior', '-overwrite',
             '-thresh1', '1']   # epoch == thresh1: train.py:45 deep-copies the model every step of that epoch + (['-no_cuda'] if args.no_cuda else [])
    env = dict(os.environ, PYTHONPATH=ROOT)
    if args.gpus is not None:
        env['CUDA_VISIBLE_DEVICES'] = args.gpus
    summary = dict(what='reference main.py, %d epoch(s), synthetic cfg-1 data' % args.epochs, dims=dims, n_train=n_train,
                   n_eval=n_eval, gpus=env.get('CUDA_VISIBLE_DEVICES', 'all visible'), n_visible=torch.cuda.device_count())
    for arm in args.arms.split(','):
        results = os.path.join(args.out, 'results_' + arm) + '/'
        if arm == 'dropin':
            cmd = [sys.executable, '-m', 'lamp_b200.run_main', ref] + flags + ['-results_dir', results]
        else:
            cmd = [sys.executable, os.path.join(ROOT, 'baseline', 'run_ref_main.py')] + flags + ['-results_dir', results]
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT)
        wall = time.time() - t0
        text = r.stdout + '\n--- stderr ---\n' + r.stderr
        with open(os.path.join(args.out, arm + '.log'), 'w') as f:
            f.write(' '.join(cmd) + '\n' + text)
        rec = parse_log(r.stdout)
        rec.update(returncode=r.returncode, wall_s=round(wall, 1), data_parallel='Using' in r.stdout and 'GPUs!' in r.stdout)
        if r.returncode != 0:
            rec['error_tail'] = (r.stderr or r.stdout)[-1500:]
        summary[arm] = rec
    a, b = summary.get('dropin', {}), summary.get('reference', {})
    if 'training_min' in a and 'training_min' in b and a['training_min'] > 0:
        summary['train_epoch_speedup'] = round(b['training_min'] / a['training_min'], 2)
    if a.get('testing_min') and b.get('testing_min'):
        summary['test_epoch_speedup'] = round(b['testing_min'] / a['testing_min'], 2)
    print(json.dumps(summary))
    with open(os.path.join(args.out, 'summary.json'), 'w') as f:
        json.dump(summary, f, indent=1)
    import shutil
    shutil.rmtree(dataroot, ignore_errors=True)          # the .pt and the checkpoints are not evidence
    for arm in args.arms.split(','):
        shutil.rmtree(os.path.join(args.out, 'results_' + arm), ignore_errors=True)
    ok = all(summary.get(arm, {}).get('returncode') == 0 for arm in args.arms.split(','))
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
