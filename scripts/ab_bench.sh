#!/bin/bash
# Same-box A/B of two builds of the native library (LAMP_B200_LIB) and of environment variants, interleaved so that
# the board's power-capped clock drift hits every arm alike.  usage: scripts/ab_bench.sh <out-prefix> <rounds> name=ENV... 
# where each arm is "name:VAR=VALUE,VAR=VALUE" (empty list allowed: "name:").
out=$1; rounds=$2; shift 2
for r in $(seq 1 $rounds); do
  for arm in "$@"; do
    name=${arm%%:*}; envs=${arm#*:}
    ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
      python bench.py --no-train --no-cpu-baseline --no-torch-gpu-baseline --steps 20 > ${out}_${name}_r${r}.json 2> ${out}_${name}_r${r}.err )
  done
done
python - "$out" "$rounds" "$@" <<'PY'
import json, sys, statistics
out, rounds, arms = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
for arm in arms:
    name = arm.split(':')[0]
    rows = []
    for r in range(1, rounds + 1):
        try:
            rows.append(json.load(open(f'{out}_{name}_r{r}.json')))
        except Exception as e:
            print(name, r, 'ERR', e)
    if not rows:
        continue
    ms = [d['ms_per_step'] for d in rows]
    k = {n: round(statistics.median(d['kernels'][n]['ms'] / d['steps'] for d in rows if n in d['kernels']), 3) for n in rows[0]['kernels']}
    print(json.dumps(dict(arm=name, ms_per_step=[round(x, 3) for x in ms], best=round(min(ms), 3), median=round(statistics.median(ms), 3),
                          sm_mhz=[d['clocks']['sm_mhz'] for d in rows], kernels_ms_median=k,
                          gemm_frac=[round(d['roofline']['frac'], 3) for d in rows])))
PY
