# round 1, capture g: bench.py launch list + full capture of the sweep kernels (C3), launch list with DRAM bytes of the
# C4 (n=28) generic path.  Every profiled command first runs to completion without ncu.  Reports are summarised on
# the box (profiles/summarize_ncu.py) because gpurun_out/ is capped at 64 MiB.
set -x
CMD="python bench.py --chains 512 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/ncu_plain_r1g.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1g.csv $CMD > gpurun_out/ncu_launch_r1g.log 2>&1
$CMD > gpurun_out/ncu_plain2_r1g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 36 -c 5 -o gpurun_out/sweep_r1g $CMD > gpurun_out/ncu_full_r1g.log 2>&1
python profiles/summarize_ncu.py gpurun_out/sweep_r1g.ncu-rep gpurun_out/r1g_sweep_ncu_full.json
C4="python benchmarks/configs.py c4"
$C4 > gpurun_out/c4_plain_r1g.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_r1g.csv $C4 > gpurun_out/ncu_c4_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hi8_sweep_kernel -s 4 -c 3 -o gpurun_out/hi8_r1g $C4 > gpurun_out/ncu_full_c4_r1g.log 2>&1
python profiles/summarize_ncu.py gpurun_out/hi8_r1g.ncu-rep gpurun_out/r1g_hi8_ncu_full.json
rm -f gpurun_out/hi8_r1g.ncu-rep
ls -la gpurun_out
