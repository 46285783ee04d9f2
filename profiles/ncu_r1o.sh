# round 1, capture o: after hiq_sweep_kernel became the n=20 HI MID sweep -- bench line + full capture of the sweep kernels.
set -x
python bench.py > gpurun_out/r1o_bench.json 2> gpurun_out/r1o_bench.err
CMD="python bench.py --chains 512 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/ncu_plain_r1o.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1o_launches.csv $CMD > gpurun_out/ncu_launch_r1o.log 2>&1
$CMD > gpurun_out/ncu_plain2_r1o.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 36 -c 5 -o gpurun_out/sweep_r1o $CMD > gpurun_out/ncu_full_r1o.log 2>&1
python profiles/summarize_ncu.py gpurun_out/sweep_r1o.ncu-rep gpurun_out/r1o_sweep_ncu_full.json | grep -v stalls
rm -f gpurun_out/sweep_r1o.ncu-rep
cut -c1-160 gpurun_out/r1o_bench.json
