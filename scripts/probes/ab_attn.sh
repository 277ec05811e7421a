for rep in 1 2; do
for lib in base new; do
  if [ $lib = base ]; then export LAMP_B200_LIB=$PWD/lamp_b200/liblamp_b200_base.so; else unset LAMP_B200_LIB; fi
  echo "== $lib"
  timeout 100 python scripts/bench_kernels.py attn | grep compact=1 | cut -c1-75
  timeout 200 python scripts/bench_configs.py cfg3 cfg4 --iters 6 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    if d['unit'].startswith('U1'): print(d['config'][:24], d['unit'][:12], '%.3f ms' % d['ms'])"
done; done
