#!/usr/bin/env python
"""Small invocations of every kernel family, for compute-sanitizer (racecheck / memcheck / synccheck):

    compute-sanitizer --tool racecheck python benchmarks/sanitizer_cases.py [case ...]

Cases: small (smem_chain_kernel n = 9 fp64, n = 13 fp32), sweep (LO / HI / hiq sweeps at n = 14, 17, 20, FAST and
non-FAST), generic (n = 22: HI8 / HI6 passes, coarse sampler), general (gen_pass_kernel: two-body fp32 n = 14, one-body
complex128 n = 13), shard (emulated fused pass + exchange, chunked passes), next (classical chains, trajectory statistics,
state moments, exact sampler, visit histogram).  Sizes are tiny: the sanitizer serialises every launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def case_small(q):
    m = q.random_ising_model(9, 7)
    q.quantum_enhanced_mcmc(6, m, q.GenericMixer.OneBodyMixer(9), (0.25, 0.6), n_chains=3, seed=1, dtype="fp64")
    q.quantum_enhanced_mcmc(4, m, q.GenericMixer.TwoBodyMixer(9), (0.25, 0.6), n_chains=2, seed=1, dtype="fp32",
                            visit_histogram=True)
    m13 = q.random_ising_model(13, 7)
    q.quantum_enhanced_mcmc(3, m13, q.GenericMixer.OneBodyMixer(13), (0.25, 0.6), n_chains=2, seed=1, dtype="fp32")
    q.transition_probabilities(m13, q.GenericMixer.OneBodyMixer(13), 5, 0.4, 3, dtype="fp32")


def case_sweep(q):
    for n, C, H in ((14, 4, 3), (17, 2, 2), (20, 2, 2)):
        m = q.random_ising_model(n, 3)
        q.quantum_enhanced_mcmc(H, m, q.GenericMixer.OneBodyMixer(n), (0.25, 0.6), n_chains=C, seed=2, dtype="fp32",
                                visit_histogram=(n == 14))
        sub = q.CustomMixer(n, [[b] for b in range(n) if b % 3 != 1])            # non-FAST kernels
        q.transition_probabilities(m, sub, 9, 0.4, 3, dtype="fp32")
        q.transition_probabilities(m, sub, 9, 0.4, 4, dtype="fp32")


def case_generic(q):
    m = q.random_ising_model(22, 3)
    mx = q.GenericMixer.OneBodyMixer(22)
    q.quantum_enhanced_mcmc(1, m, mx, 0.4, n_chains=1, seed=2, dtype="fp32")
    q.transition_probabilities(m, mx, 9, 0.4, 3, dtype="fp32")
    m21 = q.random_ising_model(21, 3)
    q.transition_probabilities(m21, q.GenericMixer.OneBodyMixer(21), 9, 0.4, 3, dtype="fp32")   # HI6 with one active qubit


def case_general(q):
    m = q.random_ising_model(14, 4)
    q.quantum_enhanced_mcmc(2, m, q.GenericMixer.TwoBodyMixer(14), (0.25, 0.6), n_chains=2, seed=3, dtype="fp32")
    m13 = q.random_ising_model(13, 4)
    q.quantum_enhanced_mcmc(2, m13, q.GenericMixer.OneBodyMixer(13), (0.25, 0.6), n_chains=2, seed=3, dtype="fp64")


def case_shard(q):
    from tests.test_gpu_shard import _emulated_proposal
    _emulated_proposal(19, 1, 3, 12345, 0.4, us=[0.3], fused=True)
    _emulated_proposal(22, 3, 3, 54321, 0.4, us=[0.7], fused=True)
    _emulated_proposal(20, 2, 4, 777, 0.4, us=[0.5], fused=True, chunk_bits=3)
    _emulated_proposal(21, 1, 3, 99, 0.4, us=[0.5], fused=True, overlap=True) if "overlap" in _emulated_proposal.__code__.co_varnames else None


def case_next(q):
    m = q.random_ising_model(10, 5)
    b = q.classical_mcmc(50, m, temperature=1.0, n_chains=40, seed=4, visit_histogram=True)
    ex = q.Exact_Sampling(m, 1.0)
    q.batch_trajectory_statistics(m, b.initial, b.proposals, b.accepted, ex.probs)
    from qumcmc_b200.training import chain_moments
    chain_moments(m, b)
    q.Exact_Sampling(q.random_ising_model(18, 5), 1.0)


CASES = {"small": case_small, "sweep": case_sweep, "generic": case_generic, "general": case_general, "shard": case_shard,
         "next": case_next}

if __name__ == "__main__":
    import torch
    import qumcmc_b200 as q
    names = sys.argv[1:] or list(CASES)
    for name in names:
        CASES[name](q)
        torch.cuda.synchronize()
        print("case", name, "done", flush=True)
