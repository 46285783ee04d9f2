set -x
CMD="python bench.py --chains 256 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mid_pipe_kernel -s 12 -c 4 -o gpurun_out/sweep_r1c $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
