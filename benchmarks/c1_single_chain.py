#!/usr/bin/env python
"""C1 shape for profiling: README 10-spin model, one chain x 2000 hops (fp64, then fp32) -- two launches of smem_chain_kernel."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import qumcmc_b200 as q  # noqa: E402
from tests import models  # noqa: E402

J, h = models.readme_model()
m = q.IsingEnergyFunction(J, h)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for dtype in ("fp64", "fp32"):
    q.quantum_enhanced_mcmc(10, m, temperature=1 / 1.100209, seed=1, dtype=dtype)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ch = q.quantum_enhanced_mcmc(H, m, temperature=1 / 1.100209, seed=2, dtype=dtype)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(dtype, "%d hops in %.4f s = %.1f us per hop" % (H, dt, dt / H * 1e6))
