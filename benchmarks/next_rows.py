#!/usr/bin/env python
"""One invocation of each next-row kernel at n=20, 4096 chains x 2000 hops: device
classical_mcmc, trajectory statistics (running KL), spin moments.  Used for the ncu launch list (profiles/ncu_r1l.sh)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import qumcmc_b200 as q  # noqa: E402
from qumcmc_b200.classical_mcmc_routines import run_classical_chains  # noqa: E402
from qumcmc_b200.training import spin_moments  # noqa: E402

m = q.random_ising_model(20, 1)
t0 = time.perf_counter()
b = run_classical_chains(m, q.UniformProposals(20), 2000, 4096, None, 1.0, 2)
t1 = time.perf_counter()
ex = q.Exact_Sampling(m, 1.0)
t2 = time.perf_counter()
st = q.batch_trajectory_statistics(m, b.initial, b.proposals, b.accepted, ex.probs)
t3 = time.perf_counter()
N, m1, m2 = spin_moments(m, np.concatenate([b.initial, b.proposals.reshape(-1)]), None)
t4 = time.perf_counter()
print("classical %.3f s, exact %.3f s, trajectory stats %.3f s, moments of %d states %.3f s, mean final KL %.4f"
      % (t1 - t0, t2 - t1, t3 - t2, N, t4 - t3, float(st["kldiv"][:, -1].mean())))
