#!/usr/bin/env python
"""CDTraining throughput: BAS 3x3 data (n = 9), quantum sampler, the reference's shape of
one chain x `mcmc_steps` hops per epoch (qumcmc/training.py:285-364), and the batched variant (64 chains per epoch).
Prints one JSON line per variant: epochs/s, hops/s end to end (sampling + moments + parameter update + exact KL)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import qumcmc_b200 as q  # noqa: E402
from qumcmc_b200.training import CDTraining  # noqa: E402

n = 9
ds = q.bas_dataset(grid_size=3)
states = sorted(set(ds.dataset))
data = q.DiscreteProbabilityDistribution({s: 1.0 / len(states) for s in states})
for label, kw, steps, epochs in (("1 chain x 1000 hops per epoch", {}, 1000, 30),
                                 ("64 chains x 1000 hops per epoch", {"n_chains": 64}, 1000, 30)):
    np.random.seed(0)
    tr = CDTraining(q.IsingEnergyFunction(np.zeros((n, n)), np.zeros(n)), 1.0, data)
    sm = q.QuantumMCMCSampler(steps, tr.model, q.GenericMixer.OneBodyMixer(n), (0.25, 0.6), temperature=1.0, **kw)
    tr.train(sm, mcmc_steps=steps, lr=0.02, epochs=2, verbose=False)          # warm-up (handles, allocations)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.train(sm, mcmc_steps=steps, lr=0.02, epochs=epochs, verbose=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    chains = kw.get("n_chains", 1)
    print(json.dumps({"config": "CDTraining BAS 3x3, quantum sampler, " + label, "epochs": epochs, "seconds": dt,
                      "epochs_per_s": epochs / dt, "hops_per_s": epochs * steps * chains / dt,
                      "kl_first": tr.training_history["kl_div"][2], "kl_last": tr.training_history["kl_div"][-1],
                      "reference_context": "238 hops/s end to end in mnist-data-proc.ipynb (SURVEY section 8 f1)"}))
