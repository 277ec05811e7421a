#!/usr/bin/env python
"""Top source lines / SASS instructions by warp-stall samples for each launch of an .ncu-rep captured with
--set full --import-source on (kernels compiled with -lineinfo).  usage: python scripts/ncu_hot_lines.py rep [N]"""
import csv
import io
import subprocess
import sys


def main(path, top=25):
    ids = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(ids.splitlines()))
    col = {h: i for i, h in enumerate(rows[0])}
    launches = [(r[col['ID']], r[col['Kernel Name']]) for r in rows[2:]]
    for k, (lid, name) in enumerate(launches):
        out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--launch-skip', str(k), '--launch-count', '1'], capture_output=True, text=True).stdout
        lines = [l for l in out.splitlines() if l.strip()]
        start = next((i for i, l in enumerate(lines) if l.startswith('"') and 'Sampling' in l), None)
        if start is None:
            print('no source page for', name[:80])
            continue
        rd = list(csv.reader(io.StringIO('\n'.join(lines[start:]))))
        hdr = {h: i for i, h in enumerate(rd[0])}
        samp = 'Warp Stall Sampling (All Samples)' if 'Warp Stall Sampling (All Samples)' in hdr else '# Samples'
        src = 'Source'
        tot = 0
        items = []
        for r in rd[1:]:
            if len(r) <= max(hdr[samp], hdr[src]):
                continue
            try:
                v = float(r[hdr[samp]].replace(',', '') or 0)
            except ValueError:
                continue
            tot += v
            items.append((v, r[hdr[src]] if src in hdr else r[1], r[0]))
        print('=' * 110)
        print(name[:100], ' total samples', int(tot))
        for v, text, addr in sorted(items, reverse=True)[:top]:
            print(f'  {100 * v / max(tot, 1):5.1f} %  {addr[:18]:18s} {text[:110]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
