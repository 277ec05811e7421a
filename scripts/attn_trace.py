#!/usr/bin/env python
"""Pipeline timeline of the attention core (debug build with -DLAMP_ATTN_TRACE): clock64() stamps of CTA 0 for the
first 64 units of one launch, printed relative to the first stamp.  usage: python scripts/attn_trace.py [self|enc]

Events: 0 K load issued | 1 Q load issued | 2 V load issued | 4 S issued (operands landed, score buffer free) |
6 PV issued (P published, V landed, O buffer free) | 7 S complete seen by softmax | 8 row-max barrier passed |
9 P stored | 10 previous item's O complete (epilogue starts) | 11 epilogue done (tracer warp) |
12 P published seen by the MMA thread | 13 V landed seen | 14 K landed seen | 15 score buffer free seen."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lamp_b200 import build as b  # noqa: E402
from lamp_b200 import _native as nat  # noqa: E402

TRACE_LIB = os.path.join(ROOT, 'lamp_b200', 'liblamp_b200_trace.so')
if not os.path.exists(TRACE_LIB) or '--rebuild' in sys.argv:
    b.build(force=True, out=TRACE_LIB, defines=('LAMP_ATTN_TRACE',))
nat.LIB_PATH = TRACE_LIB
from lamp_b200 import ops  # noqa: E402

DEV = 'cuda'
NAMES = ['Kld', 'Qld', 'Vld', 'Send', 'Siss', 'PVend', 'PViss', 'Sdone', 'maxbar', 'Pst', 'Odone', 'epiend', 'Pseen', 'Vland',
         'Kland', 'Sfree']


def main():
    which = next((a for a in sys.argv[1:] if a in ('self', 'enc', 'encp', 'l159', 'l983', 'l4096')), 'self')
    B, H, d, prec = 1100, 4, 128, 0
    Lq, Lk = (103, 300) if which in ('enc', 'encp') else (103, 103)
    if which == 'l159':      # cfg-3: bibtex, H=8, bf16, fully connected
        B, H, d, prec, Lq, Lk = 1024, 8, 64, 1, 159, 159
    elif which == 'l983':    # cfg-4: delicious, prior mask
        B, H, d, prec, Lq, Lk = 128, 4, 128, 0, 983, 983
    elif which == 'l4096':   # cfg-5
        B, H, d, prec, Lq, Lk = 4, 16, 64, 1, 4096, 4096
    hd = H * d
    torch.manual_seed(0)
    kv_start = kv_len = None
    if which not in ('enc', 'encp'):
        qkv = ops.Act(None, *ops.split(torch.randn(B * Lq, 3 * hd, device=DEV), prec), B * Lq, 3 * hd)
        q = kv = qkv
        qc, kc, vc = 0, hd, 2 * hd
        mask = (torch.rand(1, Lq, Lk, device=DEV) < 0.5)
        mask[:, torch.arange(Lq), torch.arange(Lq)] = False
        if which in ('l159', 'l4096'):
            mask = None
    else:
        q = ops.Act(None, *ops.split(torch.randn(B * Lq, hd, device=DEV), prec), B * Lq, hd)
        kv = ops.Act(None, *ops.split(torch.randn(B * Lk, 2 * hd, device=DEV), prec), B * Lk, 2 * hd)
        qc, kc, vc = 0, 0, hd
        mask = (torch.rand(B, 1, Lk, device=DEV) < 0.3)
        mask[:, :, 0] = False
        if which == 'encp':   # packed keys as in the bench: per-sample key counts U{20..300}, K|V rows packed back to back
            lens = torch.randint(20, Lk + 1, (B,), device=DEV, dtype=torch.int32)
            lens[0] = Lk
            kv_len = lens
            kv_start = (torch.cumsum(lens, 0) - lens).to(torch.int32)
            mask = None
    for _ in range(300 if Lq < 200 else 30):  # long enough for the clocks to settle
        ops.attention(q, qc, kv, kc, vc, B, H, Lq, Lk, d, prec, mask, False, kv_start=kv_start, kv_len=kv_len)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.attention(q, qc, kv, kc, vc, B, H, Lq, Lk, d, prec, mask, False, kv_start=kv_start, kv_len=kv_len)
    e1.record()
    torch.cuda.synchronize()
    print(f'{which}: launch {e0.elapsed_time(e1) * 1e3:.1f} us')
    buf = (C.c_ulonglong * (64 * 64))()
    fn = nat.lib().lamp_debug_attn_trace
    fn.argtypes = [C.c_void_p]
    fn.restype = C.c_int
    assert fn(C.addressof(buf)) == 0
    t = [[buf[e * 64 + u] for u in range(64)] for e in range(64)]
    base = min(x for row in t[:16] for x in row if x)
    print('unit ' + ' '.join(f'{n:>7s}' for n in NAMES))
    for u in range(4, 28):
        print(f'{u:4d} ' + ' '.join(f'{(t[e][u] - base) if t[e][u] else -1:7d}' for e in range(16)))
    # steady-state per-unit period and stage gaps (cycles)
    us = range(8, 28)
    per = (t[6][27] - t[6][8]) / 19.0
    print(f'period (PV issue to PV issue): {per:.0f} cycles')
    def gap(a, b, shift=0):
        v = [t[b][u + shift] - t[a][u] for u in us if t[a][u] and t[b][u + shift]]
        return sum(v) / max(len(v), 1)
    print(f'K load issue -> S issue (same unit):    {gap(0, 4):.0f}')
    print(f'K load issue -> Q load issue:           {gap(0, 1):.0f}')
    print(f'V load issue -> PV issue (same unit):   {gap(2, 6):.0f}')
    print(f'V load issue(u) -> K load issue(u+1):   {gap(2, 0, 1):.0f}')
    print(f'S issue loop (issue -> end):            {gap(4, 3):.0f}')
    print(f'PV issue loop (issue -> end):           {gap(6, 5):.0f}')
    print(f'S issued -> S complete seen:            {gap(4, 7):.0f}')
    print(f'S complete -> max barrier:              {gap(7, 8):.0f}')
    print(f'max barrier -> P stored:                {gap(8, 9):.0f}')
    print(f'P stored -> PV issue:                   {gap(9, 6):.0f}')
    print(f'P stored -> prev O done seen:           {gap(9, 10):.0f}')
    print(f'epilogue (tracer warp):                 {gap(10, 11):.0f}')
    print(f'PV issue(u) -> O done seen (u+1):       {gap(6, 10, 1):.0f}')
    print(f'PV issue(u) -> V load issue (u+1):      {gap(6, 2, 1):.0f}')
    print(f'K load issue -> K landed seen:          {gap(0, 14):.0f}')
    print(f'V load issue -> V landed seen:          {gap(2, 13):.0f}')
    print(f'P stored (tracer) -> P published seen:  {gap(9, 12):.0f}')
    print(f'score buffer free seen -> S issue:      {gap(15, 4):.0f}')

    # per softmax warp: S seen / max barrier passed / P stored, relative to the tracer warp's S-seen stamp of the unit
    print('per-warp stamps relative to S-complete-seen of warp 4 (unit 20..23): Sseen maxbar Pst')
    for u in range(20, 24):
        ref = t[16][u]
        row = []
        for w in range(16):
            if t[16 + 3 * w][u]:
                row.append('w%d:%d/%d/%d' % (w + 5, t[16 + 3 * w][u] - ref, t[17 + 3 * w][u] - ref, t[18 + 3 * w][u] - ref))
        print(f'  u{u}: ' + ' '.join(row))
    main_done = True


if __name__ == '__main__':
    main()
