# round 1, capture n (end-of-round code): bench.py line, its launch list, full capture of the C3 sweep kernels,
# C4 launch list + full capture of the HI8 MID pass, secondary configs.  Reports summarised on the box.
set -x
python bench.py > gpurun_out/r1n_bench.json 2> gpurun_out/r1n_bench.err
CMD="python bench.py --chains 512 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/ncu_plain_r1n.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1n_launches.csv $CMD > gpurun_out/ncu_launch_r1n.log 2>&1
$CMD > gpurun_out/ncu_plain2_r1n.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 36 -c 5 -o gpurun_out/sweep_r1n $CMD > gpurun_out/ncu_full_r1n.log 2>&1
python profiles/summarize_ncu.py gpurun_out/sweep_r1n.ncu-rep gpurun_out/r1n_sweep_ncu_full.json > /dev/null
rm -f gpurun_out/sweep_r1n.ncu-rep
C4="python benchmarks/configs.py c4"
$C4 > gpurun_out/c4_plain_r1n.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1n_c4_launches.csv $C4 > gpurun_out/ncu_c4_r1n.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hi8_sweep_kernel -s 6 -c 2 -o gpurun_out/hi8_r1n $C4 > gpurun_out/ncu_full_c4_r1n.log 2>&1
python profiles/summarize_ncu.py gpurun_out/hi8_r1n.ncu-rep gpurun_out/r1n_hi8_ncu_full.json | tail -4
python benchmarks/configs.py c1 c2 c4 exact general > gpurun_out/r1n_secondary_configs.jsonl 2> gpurun_out/r1n_secondary.err
python benchmarks/c1_single_chain.py 10000 >> gpurun_out/r1n_secondary_configs.jsonl
cut -c1-200 gpurun_out/r1n_bench.json
