#!/usr/bin/env python
"""One chain of a large model on the multi-group sweep path: python benchmarks/single_chain.py N [HOPS]."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import qumcmc_b200 as q  # noqa: E402

n = int(sys.argv[1])
H = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m = q.random_ising_model(n, 1)
mx = q.GenericMixer.OneBodyMixer(n)
q.quantum_enhanced_mcmc(1, m, mx, (0.25, 0.6), temperature=1.0, n_chains=1, seed=3, dtype="fp32")
torch.cuda.synchronize()
t0 = time.perf_counter()
b = q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.0, n_chains=1, seed=4, dtype="fp32")
torch.cuda.synchronize()
dt = time.perf_counter() - t0
from qumcmc_b200.rng import philox4x32  # noqa: E402
ts = [2 + ((int(philox4x32(4, np.array([0], dtype=np.uint64), hop, 0)[0, 2]) * 10) >> 32) for hop in range(H)]
rs = [max(1, int(np.floor(t / 0.8))) for t in ts]
alg = sum(2 * (1 << n) * 8 * r for r in rs)
print(json.dumps({"n": n, "hops": H, "seconds": dt, "proposals_per_s": H / dt, "layers": rs,
                  "one_pass_per_layer_GBps": alg / 1e9 / dt}))
