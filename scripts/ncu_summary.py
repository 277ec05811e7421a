#!/usr/bin/env python
"""Compact per-launch summary of an .ncu-rep (run where ncu is installed; no GPU needed).
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor.sum', 'sm__inst_executed_pipe_uniform.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
    'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_bank_conflicts_pipe_lsu.sum',
    'smsp__cycles_active.avg', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('=' * 100)
        print(r[col['Kernel Name']][:110], ' grid', r[col['Grid Size']], ' block', r[col['Block Size']])
        for w in WANT:
            if w in col:
                print(f'  {w:78s} {r[col[w]]:>16s} {units[col[w]]}')
        # tensor-related extras
        for h in hdr:
            if ('tensor' in h and 'pct_of_peak_sustained_elapsed' in h and 'ops_path' not in h) or \
               ('stall' in h and h.endswith('.pct')):
                print(f'  {h:78s} {r[col[h]]:>16s} {units[col[h]]}')


if __name__ == '__main__':
    main(sys.argv[1])
