"""n = 21..23 (three qubit groups, 16-64 MiB per statevector): r-major schedule vs one chain per side stream with the
statevector resident in the L2.  python benchmarks/l2_generic_sweep.py  (run once per QUMCMC_L2_GENERIC=0/1)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import qumcmc_b200 as q

for n, chains in ((21, 192), (22, 96), (23, 48)):
    m = q.random_ising_model(n, 1)
    mx = q.GenericMixer.OneBodyMixer(n)
    q.quantum_enhanced_mcmc(1, m, mx, (0.25, 0.6), temperature=1.0, n_chains=chains, seed=3, dtype="fp32", record=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    H = 4
    q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.0, n_chains=chains, seed=4, dtype="fp32", record=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"n": n, "chains": chains, "hops": H, "l2_generic": os.environ.get("QUMCMC_L2_GENERIC", "1"),
                      "streams": os.environ.get("QUMCMC_L2_STREAMS", "default"), "proposals_per_s": chains * H / dt}))
