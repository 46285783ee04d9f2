set -x
CMD="python bench.py --chains 512 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/ncu_plain_r1f.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1f.csv $CMD > gpurun_out/ncu_launch_r1f.log 2>&1
$CMD > gpurun_out/ncu_plain2_r1f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 36 -c 8 -o gpurun_out/sweep_r1f $CMD > gpurun_out/ncu_full_r1f.log 2>&1
tail -3 gpurun_out/ncu_full_r1f.log
