"""Three-body mixer at n = 20 / 16 in complex128: XT mode of the fp64 sweep kernels vs the general path (QUMCMC_NO_XT=1)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qumcmc_b200 as q
for n, chains, H in ((16, 256, 6), (20, 128, 4)):
    m = q.random_ising_model(n, 1)
    mx = q.GenericMixer.ThreeBodyMixer(n)
    q.quantum_enhanced_mcmc(1, m, mx, (0.25, 0.6), temperature=1.0, n_chains=chains, seed=3, dtype="fp64", record=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.0, n_chains=chains, seed=4, dtype="fp64", record=False)
    torch.cuda.synchronize()
    print(json.dumps({"n": n, "mixer": "three-body", "dtype": "fp64", "no_xt": os.environ.get("QUMCMC_NO_XT", "0"),
                      "proposals_per_s": chains * H / (time.perf_counter() - t0)}))
