# A/B of the L2-resident group schedule of the two-group sweep path (C3): group budget (MiB of statevectors per stream)
# x number of side streams.  QUMCMC_L2_GROUP_MIB=0 = the plain step-major schedule (every sweep streams all chains
# through HBM).  gpurun --timeout 900 -- 'bash benchmarks/l2_group_sweep.sh "0 1" "16 4" > gpurun_out/l2_group_sweep.txt 2>&1'
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-secondary"
if [ $# -eq 0 ]; then set -- "0 1" "8 8" "8 6" "16 4" "16 3" "24 3" "256 4" "1024 2"; fi
for cfg in "$@"; do
  set -- $cfg
  echo "== MIB=$1 STREAMS=$2 $3"
  env $3 QUMCMC_L2_DEBUG=1 QUMCMC_L2_GROUP_MIB=$1 QUMCMC_L2_STREAMS=$2 $B 2>gpurun_out/l2_dbg.err | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(round(d['value']), 'proposals/s', d['ms_per_step'], 'ms', 'frac', round(d['roofline']['frac'], 3), 'sm', d['clocks']['sm_mhz'], 'launches', d['gpu_launches'], 'acc', d['config'].get('acceptance_rate'))
"
  grep qumcmc gpurun_out/l2_dbg.err | tail -2
done
