"""On-chip path with k-body mixers (BAS 3x3, n = 9, and n = 12): term loop vs Hadamard-basis mode.
QUMCMC_SMALL_HX=0 python benchmarks/small_kbody.py ; python benchmarks/small_kbody.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import qumcmc_b200 as q
from tests import models

J, h, _ = models.bas3("both")
cases = [(q.IsingEnergyFunction(J, h), 9, "111000111", 1024, 500), (q.random_ising_model(12, 1), 12, None, 1024, 100)]
for m, n, init, chains, H in cases:
    for name, mx in (("one", q.GenericMixer.OneBodyMixer(n)), ("two", q.GenericMixer.TwoBodyMixer(n)), ("three", q.GenericMixer.ThreeBodyMixer(n))):
        for dtype in ("fp32", "fp64"):
            kw = dict(temperature=1 / 1.5, n_chains=chains, seed=2, dtype=dtype, record=False)
            if init:
                kw["initial_state"] = init
            q.quantum_enhanced_mcmc(5, m, mx, (0.25, 0.6), **kw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), **kw)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(json.dumps({"n": n, "mixer": name + "-body", "terms": len(mx.lower()[1]), "dtype": dtype, "chains": chains, "hops": H,
                              "hx": os.environ.get("QUMCMC_SMALL_HX", "default"), "proposals_per_s": chains * H / dt}))
