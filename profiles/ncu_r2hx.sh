# round 2: --set full summaries of the on-chip kernels in Hadamard-basis mode (BAS 3x3 two-body on the warp kernel, n = 12
# three-body on the CTA kernel), each after its command ran to completion without ncu.
# gpurun --timeout 600 -- 'bash profiles/ncu_r2hx.sh > gpurun_out/r2hx_script.log 2>&1'
set -x
cat > /tmp/r2hx.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import qumcmc_b200 as q
from tests import models
which = sys.argv[1]
if which == "warp":
    J, h, _ = models.bas3("both")
    m = q.IsingEnergyFunction(J, h)
    q.quantum_enhanced_mcmc(300, m, q.GenericMixer.TwoBodyMixer(9), (0.25, 0.6), initial_state="111000111", temperature=1 / 1.5,
                            n_chains=1024, seed=2, dtype="fp32", record=False)
else:
    m = q.random_ising_model(12, 1)
    q.quantum_enhanced_mcmc(40, m, q.GenericMixer.ThreeBodyMixer(12), (0.25, 0.6), temperature=1.0, n_chains=1024, seed=2, dtype="fp32", record=False)
PY
for W in warp cta; do
  python /tmp/r2hx.py $W > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -o gpurun_out/hx_${W} python /tmp/r2hx.py $W > /dev/null 2>&1
  python profiles/summarize_ncu.py gpurun_out/hx_${W}.ncu-rep gpurun_out/r2hx_${W}_ncu_full.json
  rm -f gpurun_out/hx_${W}.ncu-rep
done
