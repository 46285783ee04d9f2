# round 2, final-code capture: bench.py line (plain), its launch list, --set full summaries of the hot kernels of every
# path that changed this round.  Every profiled command first runs to completion without ncu; numbers printed under ncu
# are never bench values.  gpurun --timeout 2400 -- 'bash profiles/ncu_r2.sh'
set -x
T=r2z
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
CMD="python bench.py --chains 512 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/ncu_plain_${T}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/ncu_launch_${T}.log 2>&1
python profiles/summarize_launches.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launches_summary.txt
# C3 sweep kernels: LO MID on 64-thread CTAs, HI MID (hiq), FIRST / FINAL roles
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 36 -c 6 -o gpurun_out/sweep_${T} $CMD > gpurun_out/ncu_full_${T}.log 2>&1
python profiles/summarize_ncu.py gpurun_out/sweep_${T}.ncu-rep gpurun_out/${T}_sweep_ncu_full.json | grep -v stalls
rm -f gpurun_out/sweep_${T}.ncu-rep
# complex128 sweeps (n = 20) and the Hadamard-basis general path (n = 20 two-body, complex64)
cat > /tmp/r2_small_cases.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import qumcmc_b200 as q
which = sys.argv[1]
m = q.random_ising_model(20, 1)
if which == "c128":
    q.quantum_enhanced_mcmc(2, m, q.GenericMixer.OneBodyMixer(20), (0.25, 0.6), temperature=1.0, n_chains=256, seed=4, dtype="fp64", record=False)
elif which == "general":
    q.quantum_enhanced_mcmc(2, m, q.GenericMixer.TwoBodyMixer(20), (0.25, 0.6), temperature=1.0, n_chains=256, seed=4, dtype="fp32", record=False)
elif which == "c2":
    from tests import models
    J, h, _ = models.bas3("both")
    mm = q.IsingEnergyFunction(J, h)
    q.quantum_enhanced_mcmc(500, mm, q.GenericMixer.OneBodyMixer(9), (0.25, 0.6), initial_state="111000111", temperature=1 / 1.5, n_chains=1024, seed=2, dtype="fp32", record=False)
PY
for W in c128 general c2; do python /tmp/r2_small_cases.py $W > gpurun_out/plain_${W}_${T}.log 2>&1; done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_c128_launches.csv python /tmp/r2_small_cases.py c128 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/${T}_c128_launches.csv > gpurun_out/${T}_c128_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:sweep128_kernel -s 12 -c 4 -o gpurun_out/c128_${T} python /tmp/r2_small_cases.py c128 > /dev/null 2>&1
python profiles/summarize_ncu.py gpurun_out/c128_${T}.ncu-rep gpurun_out/${T}_sweep128_ncu_full.json | grep -v stalls
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_general_launches.csv python /tmp/r2_small_cases.py general > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/${T}_general_launches.csv > gpurun_out/${T}_general_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:gen_pass_kernel -s 6 -c 4 -o gpurun_out/general_${T} python /tmp/r2_small_cases.py general > /dev/null 2>&1
python profiles/summarize_ncu.py gpurun_out/general_${T}.ncu-rep gpurun_out/${T}_general_ncu_full.json | grep -v stalls
ncu --set full --clock-control none --import-source on -k regex:warp_chain_kernel -c 1 -o gpurun_out/warp_${T} python /tmp/r2_small_cases.py c2 > /dev/null 2>&1
python profiles/summarize_ncu.py gpurun_out/warp_${T}.ncu-rep gpurun_out/${T}_warp_ncu_full.json | grep -v stalls
rm -f gpurun_out/c128_${T}.ncu-rep gpurun_out/general_${T}.ncu-rep gpurun_out/warp_${T}.ncu-rep
python benchmarks/configs.py c1 c2 c4 exact general c128 > gpurun_out/${T}_secondary_configs.jsonl 2> gpurun_out/${T}_secondary.err
python benchmarks/c1_single_chain.py 10000 >> gpurun_out/${T}_secondary_configs.jsonl
cut -c1-300 gpurun_out/${T}_bench.json
