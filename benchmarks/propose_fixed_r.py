#!/usr/bin/env python
"""One proposal with a fixed number of mixer layers on one GPU (qmc_propose): python benchmarks/propose_fixed_r.py N R [REPS].
Weak-scaling reference for the sharded runs: n = 33 here has the per-GPU state of n = 36 on 8 GPUs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import qumcmc_b200 as q  # noqa: E402
from qumcmc_b200 import _lib  # noqa: E402

n, r = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
m = q.random_ising_model(n, 1)
dm = m.device_model(0)
dm.set_mixer(q.GenericMixer.OneBodyMixer(n))
dev = torch.device("cuda", 0)
t = lambda x: torch.from_numpy(x).to(dev)
s_in = t(np.array([0x5EED5EED5 & ((1 << n) - 1)], dtype=np.int64))
gam, rr, uu = t(np.array([0.45])), t(np.array([r], dtype=np.int32)), t(np.array([0.37]))
out = torch.empty(1, dtype=torch.int64, device=dev)
call = lambda: _lib.check(_lib.lib().qmc_propose(dm.handle, 1, s_in.data_ptr(), gam.data_ptr(), rr.data_ptr(), uu.data_ptr(),
                                                 None, 0.8, _lib.DTYPE_F32, out.data_ptr(), None))
call()
torch.cuda.synchronize()
best = 1e9
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e-3)
print(json.dumps({"config": "one GPU, one proposal", "n": n, "layers": r, "seconds_per_proposal": best,
                  "state_GiB": 8.0 * (1 << n) / 2 ** 30, "index": int(out.item())}))
