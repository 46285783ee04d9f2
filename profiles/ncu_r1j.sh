# round 1, capture j (final code of the round): C4 (n=28) launch list with DRAM bytes + full capture of the HI8 MID pass;
# reports summarised on the box (gpurun_out/ is capped at 64 MiB).  Every profiled command first runs without ncu.
set -x
C4="python benchmarks/configs.py c4"
$C4 > gpurun_out/c4_plain_r1j.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_r1j.csv $C4 > gpurun_out/ncu_c4_r1j.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hi8_sweep_kernel -s 6 -c 2 -o gpurun_out/hi8_r1j $C4 > gpurun_out/ncu_full_c4_r1j.log 2>&1
python profiles/summarize_ncu.py gpurun_out/hi8_r1j.ncu-rep gpurun_out/r1j_hi8_ncu_full.json
python benchmarks/configs.py c1 c2 c4 exact general > gpurun_out/r1j_secondary_configs.jsonl 2> gpurun_out/r1j_secondary.err
tail -3 gpurun_out/r1j_secondary_configs.jsonl
