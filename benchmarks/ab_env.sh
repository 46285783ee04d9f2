#!/bin/bash
# A/B helper: benchmarks/ab_env.sh VAR v1 v2 ... -> one short bench.py line per value of the env var
VAR=$1; shift
for v in "$@"; do
  export $VAR=$v
  python bench.py --steps ${STEPS:-6} --warmup 2 --no-cpu-baseline --e2e-steps 2 ${EXTRA} 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$VAR=$v', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['e2e']['matches_device_resident_run'], round(d['e2e']['value'],1))"
done
