# round 2, L2-resident group schedule: bench.py line (plain run), then the launch list of a 256-chain run with DRAM bytes
# per launch under `--cache-control none` (ncu must not flush the L2 between launches: residency across the sweeps of a
# group is the thing measured).  ncu serialises the side streams, so per-launch times are not the concurrent ones; the
# DRAM byte counts are what this capture is for.  gpurun --timeout 1500 -- 'bash profiles/ncu_r2l2.sh'
set -x
T=r2l2
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
CMD="python bench.py --chains 256 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-secondary"
$CMD > gpurun_out/ncu_plain_${T}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/ncu_launch_${T}.log 2>&1
python profiles/summarize_launches.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launches_summary.txt
QUMCMC_L2_GROUP_MIB=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -c 400 --csv --log-file gpurun_out/${T}_stepmajor_launches.csv $CMD > gpurun_out/ncu_launch_${T}_sm.log 2>&1
python profiles/summarize_launches.py gpurun_out/${T}_stepmajor_launches.csv > gpurun_out/${T}_stepmajor_launches_summary.txt
gzip -f gpurun_out/${T}_launches.csv
cat gpurun_out/${T}_launches_summary.txt gpurun_out/${T}_stepmajor_launches_summary.txt
cut -c1-400 gpurun_out/${T}_bench.json
