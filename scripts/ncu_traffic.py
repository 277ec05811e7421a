#!/usr/bin/env python
"""Turn an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch log of
bench.py into profiles/r02_traffic.json (average DRAM bytes per launch for each lamp_b200 kernel family).  The file
records the hash of the native sources it was captured with (bench.lib_source_hash); bench.py quotes it in
`roofline.traffic` only while that hash matches the library it runs.
usage: python scripts/ncu_traffic.py <launches.csv> <out.json> <batch> <precision>"""
import collections
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (lib_source_hash)

log, out, batch, precision = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = [r for r in csv.reader(open(log)) if len(r) > 5]
hdr, per = None, collections.OrderedDict()
for r in rows:
    if r[0] == 'ID':
        hdr = r
        continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        rec = per.setdefault(d['ID'], dict(name=d['Kernel Name']))
        val = float(d['Metric Value'].replace(',', ''))
        unit = d['Metric Unit']
        mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0}.get(unit, 1.0)
        rec[d['Metric Name']] = val * mult


def family(name):
    if 'gemm_planes_kernel' in name:
        return 'gemm_planes'
    if 'attn_core_kernel<128' in name:
        return 'attn_core_self'
    if 'attn_core_kernel<64' in name:
        return 'attn_core_enc'
    for k in ('layernorm', 'embed', 'diag_proj', 'split_planes'):
        if k in name:
            return k
    return None


agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for rec in per.values():
    fam = family(rec['name'])
    if fam is None:
        continue
    a = agg[fam]
    a[0] += 1
    a[1] += rec.get('dram__bytes_read.sum', 0.0) + rec.get('dram__bytes_write.sum', 0.0)
    a[2] += rec.get('gpu__time_duration.sum', 0.0)
res = dict(batch=batch, precision=precision, source=log, lib_source_hash=bench.lib_source_hash(),
           avg_bytes_per_launch={k: v[1] / v[0] for k, v in agg.items()},
           launches={k: v[0] for k, v in agg.items()},
           avg_duration_us={k: v[2] / v[0] * 1e6 for k, v in agg.items()})
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps(res, indent=1))
