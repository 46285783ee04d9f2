set -x
export QUMCMC_SWEEP_PF=222
CMD="python bench.py --chains 512 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/ncu_plain_r1e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 36 -c 6 -o gpurun_out/sweep_r1e $CMD > gpurun_out/ncu_full_r1e.log 2>&1
tail -3 gpurun_out/ncu_full_r1e.log
