#!/usr/bin/env python
"""Secondary measurements on the other BASELINE.json configs (the driver's headline is bench.py):
   C1 README 10-spin single chain x 10 000 hops (+ classical_mcmc + Exact_Sampling)
   C2 BAS 3x3, 1024 chains x 2 000 hops, gamma ~ U(0.25, 0.6) (shared-memory path, fp32 and fp64)
   C4 28-spin single chain (generic HBM sweep path) + Exact_Sampling log Z
Prints one JSON line per measurement.  Usage: python benchmarks/configs.py [c1 c2 c4 exact]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import qumcmc_b200 as q  # noqa: E402
from tests import models  # noqa: E402


def timed(fn, reps=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def c1():
    J, h = models.readme_model()
    m = q.IsingEnergyFunction(J, h, name="my_model")
    beta = 1.100209
    np.random.seed(1)
    q.quantum_enhanced_mcmc(n_hops=100, model=m, temperature=1 / beta)  # warm
    for dtype in ("fp64", "fp32"):
        dt, ch = timed(lambda: q.quantum_enhanced_mcmc(n_hops=10000, model=m, temperature=1 / beta, dtype=dtype))
        print(json.dumps({"config": "C1 README n=10, 1 chain x 10000 hops", "dtype": dtype, "seconds": dt,
                          "proposals_per_s": 10000 / dt, "acceptance": ch.acceptance_rate(),
                          "reference_published": "476-550 it/s (qulacs, authors' CPU; BASELINE.md)"}))
    dt, cl = timed(lambda: q.classical_mcmc(n_hops=10000, model=m, temperature=1 / beta))
    print(json.dumps({"config": "C1 classical_mcmc, 1 chain x 10000 hops (qmc_classical_chains)", "seconds": dt, "hops_per_s": 10000 / dt}))
    dt, cb = timed(lambda: q.classical_mcmc(n_hops=10000, model=m, temperature=1 / beta, n_chains=4096, seed=1))
    print(json.dumps({"config": "C1 classical_mcmc, 4096 chains x 10000 hops (incl. D2H of the recorded chains)", "seconds": dt,
                      "hops_per_s": 4096e4 / dt, "acceptance": cb.acceptance_rate()}))
    dt, ex = timed(lambda: q.Exact_Sampling(m, beta))
    print(json.dumps({"config": "C1 Exact_Sampling n=10 (incl. host arrays)", "seconds": dt, "logZ": ex.logZ}))


def c2():
    J, h, _ = models.bas3("both")
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.OneBodyMixer(9)
    for dtype in ("fp32", "fp64"):
        q.quantum_enhanced_mcmc(100, m, mx, (0.25, 0.6), initial_state="111000111", temperature=1 / 1.5, n_chains=1024,
                                seed=1, dtype=dtype)
        H = 2000
        dt, b = timed(lambda: q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), initial_state="111000111",
                                                      temperature=1 / 1.5, n_chains=1024, seed=2, dtype=dtype))
        print(json.dumps({"config": "C2 BAS 3x3, 1024 chains x %d hops, gamma~U(0.25,0.6)" % H, "dtype": dtype,
                          "seconds": dt, "proposals_per_s": 1024 * H / dt, "acceptance": b.acceptance_rate(),
                          "path": "shared-memory kernel (on-chip; no HBM roofline)",
                          "reference_published": "497-1045 it/s per chain (gamma-variation-v0.1.ipynb)"}))


def c4():
    n = 28
    m = q.random_ising_model(n, 1)
    mx = q.GenericMixer.OneBodyMixer(n)
    q.quantum_enhanced_mcmc(1, m, mx, (0.25, 0.6), temperature=1.0, n_chains=1, seed=3, dtype="fp32")  # warm + alloc
    H = 8
    dt, b = timed(lambda: q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.0, n_chains=1, seed=4, dtype="fp32"))
    # layers of these hops (host replica of the Philox draws)
    from qumcmc_b200.rng import philox4x32
    ts = [2 + ((int(philox4x32(4, np.array([0], dtype=np.uint64), hop, 0)[0, 2]) * 10) >> 32) for hop in range(H)]
    rs = [max(1, int(np.floor(t / 0.8))) for t in ts]
    alg = sum(2 * (1 << n) * 8 * r for r in rs)
    # bytes the passes actually move: K = 2 register groups above the LO qubits => 2 passes per layer, the first pass of a
    # proposal write-only and the last read-only: 16 B * 2^n * K * (r - 1) (qmc_profile_read_moved uses the same formula)
    moved = sum(16.0 * (1 << n) * 2 * (r - 1) for r in rs)
    print(json.dumps({"config": "C4 n=28 single chain x %d hops (2 GiB complex64 statevector)" % H, "seconds": dt,
                      "proposals_per_s": H / dt, "layers": rs, "algorithmic_GB": alg / 1e9,
                      "algorithmic_GBps": alg / 1e9 / dt, "frac_of_measured_hbm": alg / 1e9 / dt / 6547.8,
                      "passes_per_layer": 2, "moved_GB": moved / 1e9, "moved_GBps": moved / 1e9 / dt,
                      "per_pass_frac_of_measured_hbm": moved / 1e9 / dt / 6547.8,
                      "path": "generic sweeps: LO(12) + HI8 + HI8 qubit groups, 2 passes per layer; SURVEY 8d's B_alg counts one "
                              "pass per layer (k = 14 tile qubits), which strided 2^13-amplitude tiles cannot reach: "
                              "frac_of_measured_hbm is capped at 0.5 by construction, per_pass_frac is the kernels' own rate"}))


def general():
    """Mixers beyond one-body: two-body (n <= 20: XQ mode of the sweep kernels, 2 sweeps per layer; n > 20: Hadamard-basis
    general path) and three-body (general path at every n): proposals/s and the pass count."""
    import os
    for n, mixer_name, dtype, chains, H in [(16, "two", "fp32", 512, 8), (16, "two", "fp64", 512, 8), (20, "two", "fp32", 1024, 6),
                                            (20, "two", "fp64", 512, 6), (16, "three", "fp32", 512, 8), (20, "three", "fp32", 512, 6),
                                            (24, "two", "fp32", 32, 4), (26, "two", "fp32", 8, 4)]:
        m = q.random_ising_model(n, 1)
        mx = {"one": q.GenericMixer.OneBodyMixer, "two": q.GenericMixer.TwoBodyMixer, "three": q.GenericMixer.ThreeBodyMixer}[mixer_name](n)
        xq = n <= 20 and not os.environ.get("QUMCMC_NO_XQ") and (mixer_name == "two" or not os.environ.get("QUMCMC_NO_XT"))
        q.quantum_enhanced_mcmc(1, m, mx, (0.25, 0.6), temperature=1.0, n_chains=chains, seed=3, dtype=dtype)  # warm + alloc
        dt, b = timed(lambda: q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.0, n_chains=chains, seed=4, dtype=dtype))
        from qumcmc_b200.rng import philox4x32
        sum_r = 0
        for hop in range(H):
            w2 = philox4x32(4, np.arange(chains, dtype=np.uint64), hop, 0)[:, 2].astype(np.uint64)
            sum_r += int(np.maximum(1, np.floor((2 + ((w2 * np.uint64(10)) >> np.uint64(32))).astype(np.float64) / 0.8)).sum())
        K = max(1, -(-(n - 12) // 8))                        # groups above qubits 0..11
        esz = 8 if dtype == "fp32" else 16
        moved = 2.0 * esz * (1 << n) * (2 * K * sum_r - K * chains * H)   # FIRST write-only, 2K(r-1)+K-... passes read+write
        print(json.dumps({"config": "%s n=%d, %s-body mixer, %s, %d chains x %d hops" % (
                              ("sweep kernels (XQ mode)" if mixer_name == "two" else "sweep kernels (XT mode)") if xq else "general path", n, mixer_name, dtype, chains, H),
                          "seconds": dt, "proposals_per_s": chains * H / dt, "acceptance": b.acceptance_rate(),
                          "passes_per_layer": 2 * K, "approx_moved_GBps": moved / 1e9 / dt,
                          "reference_cost": "one statevector pass per gate: %d passes per layer" % (
                              n * (n - 1) // 2 + n + {"one": n, "two": n * (n - 1) // 2, "three": n * (n - 1) * (n - 2) // 6}[mixer_name])}))


def c128():
    """C3's model in the reference's own precision: n = 20, complex128 (16 MiB per chain), fp64 sweep kernels.
    Roofline (SURVEY 8d with b = 16): 2 * 2^20 * 16 B * r per proposal."""
    n = 20
    m = q.random_ising_model(n, 1)
    mx = q.GenericMixer.OneBodyMixer(n)
    for chains, H in ((2048, 6),):
        q.quantum_enhanced_mcmc(1, m, mx, (0.25, 0.6), temperature=1.0, n_chains=chains, seed=3, dtype="fp64", record=False)
        dt, b = timed(lambda: q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.0, n_chains=chains, seed=4,
                                                      dtype="fp64", record=False))
        from qumcmc_b200.rng import philox4x32
        sum_r = 0
        for hop in range(H):
            w2 = philox4x32(4, np.arange(chains, dtype=np.uint64), hop, 0)[:, 2].astype(np.uint64)
            sum_r += int(np.maximum(1, np.floor((2 + ((w2 * np.uint64(10)) >> np.uint64(32))).astype(np.float64) / 0.8)).sum())
        alg = 2.0 * (1 << n) * 16 * sum_r
        print(json.dumps({"config": "C3 model in complex128: n=20, %d chains x %d hops (fp64 sweep kernels)" % (chains, H),
                          "seconds": dt, "proposals_per_s": chains * H / dt, "acceptance": b.acceptance_rate(),
                          "algorithmic_GBps": alg / 1e9 / dt, "frac_of_measured_hbm": alg / 1e9 / dt / 6547.8,
                          "roofline_proposals_per_s": 6547.8e9 / (alg / (chains * H))}))


def exact():
    for n in (20, 24, 28):
        m = q.random_ising_model(n, 1)
        dm = m.device_model(0)
        from qumcmc_b200 import _device, _lib
        pair = torch.empty(2, dtype=torch.float64, device="cuda")
        f = lambda: _lib.check(_lib.lib().qmc_exact_partial(dm.handle, 1.0, 0, 1 << n, pair.data_ptr(), None,
                                                            _device.current_stream_ptr(0)))
        f()
        dt, _ = timed(f, reps=3)
        mm, ss = pair.cpu().numpy()
        print(json.dumps({"config": "Exact_Sampling log Z only, n=%d" % n, "seconds": dt, "Gstates_per_s": (1 << n) / dt / 1e9,
                          "logZ": float(mm + np.log(ss)), "bound": "fp64 ALU (bit-exact O(n^2) energy per state), 0 HBM bytes"}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c4", "exact", "general", "c128"]
    for w in which:
        {"c1": c1, "c2": c2, "c4": c4, "exact": exact, "general": general, "c128": c128}[w]()
