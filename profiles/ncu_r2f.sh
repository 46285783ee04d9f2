# round 2, final-code capture (L2-resident group schedule + XQ mode built in): bench.py line (plain), the launch list of a
# short run with DRAM bytes per launch under --cache-control none (group schedule and step-major), --set full summaries of
# the C3 sweep kernels under the group schedule (the statevectors stay in the L2 between the replays too) and of the XQ
# instantiations (two-body mixer, n = 20).  Numbers printed under ncu are never bench values.
# gpurun --timeout 1500 -- 'bash profiles/ncu_r2f.sh > gpurun_out/r2f_script.log 2>&1'
set -x
T=r2f
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
CMD="python bench.py --chains 96 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-secondary"
$CMD > gpurun_out/ncu_plain_${T}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -c 2500 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/ncu_launch_${T}.log 2>&1
python profiles/summarize_launches.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launches_summary.txt
QUMCMC_L2_GROUP_MIB=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -c 400 --csv --log-file gpurun_out/${T}_stepmajor_launches.csv $CMD > gpurun_out/ncu_launch_${T}_sm.log 2>&1
python profiles/summarize_launches.py gpurun_out/${T}_stepmajor_launches.csv > gpurun_out/${T}_stepmajor_launches_summary.txt
gzip -f gpurun_out/${T}_launches.csv
ncu --set full --cache-control none --clock-control none --import-source on -k regex:sweep_kernel -s 300 -c 8 -o gpurun_out/sweep_${T} $CMD > gpurun_out/ncu_full_${T}.log 2>&1
python profiles/summarize_ncu.py gpurun_out/sweep_${T}.ncu-rep gpurun_out/${T}_sweep_ncu_full.json | grep -v stalls
rm -f gpurun_out/sweep_${T}.ncu-rep
cat > /tmp/r2f_xq.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import qumcmc_b200 as q
m = q.random_ising_model(20, 1)
q.quantum_enhanced_mcmc(2, m, q.GenericMixer.TwoBodyMixer(20), (0.25, 0.6), temperature=1.0, n_chains=96, seed=4, dtype="fp32", record=False)
PY
python /tmp/r2f_xq.py > /dev/null 2>&1 &&
ncu --set full --cache-control none --clock-control none --import-source on -k regex:sweep_kernel -s 200 -c 6 -o gpurun_out/xq_${T} python /tmp/r2f_xq.py > /dev/null 2>&1
python profiles/summarize_ncu.py gpurun_out/xq_${T}.ncu-rep gpurun_out/${T}_xq_ncu_full.json | grep -v stalls
rm -f gpurun_out/xq_${T}.ncu-rep
python benchmarks/configs.py c1 c2 c4 exact general c128 > gpurun_out/${T}_secondary_configs.jsonl 2> gpurun_out/${T}_secondary.err
cat gpurun_out/${T}_launches_summary.txt gpurun_out/${T}_stepmajor_launches_summary.txt | grep -v "at::\|af_table\|chains_\|hop_prep\|schedule"
cut -c1-400 gpurun_out/${T}_bench.json
