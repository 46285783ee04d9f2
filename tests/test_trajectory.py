"""Trajectory statistics on the device against the numpy restatement of the reference's
trajectory_processing functions (oracle/np_oracle.py)."""
import numpy as np
import pytest

import oracle
from oracle import np_oracle
from tests import models


def test_oracle_running_kl_matches_the_dict_based_definition():
    """The vectorised running KL (trajectory_processing.py:137-151) equals kl_divergence over get_accepted_dict
    (prob_dist.py:130-147, the 'slow' variant kept as a comment at :126-134)."""
    rng = np.random.default_rng(0)
    p = rng.random(16)
    p /= p.sum()
    mc = rng.integers(0, 16, size=50)
    rk = np_oracle.running_kl(p, mc)
    for k in (1, 7, 50):
        vals, cnt = np.unique(mc[:k], return_counts=True)
        q = dict(zip(vals.tolist(), (cnt / k).tolist()))
        kl = sum(p[i] * np.log(p[i]) - p[i] * np.log(max(1e-8, q.get(i, 0.0))) for i in range(16))
        assert abs(rk[k - 1] - kl) < 1e-12


def _recorded_chains(n, C, H, seed, T=1.0):
    J, h = models.packaged_random_model(n, seed)
    kinds, params, probs = [0, 1], [0, 1], [0.3, 0.7]
    init = np.array([oracle.initial_state(seed, c, n) for c in range(C)], dtype=np.uint64)
    props = np.empty((C, H), dtype=np.uint64)
    acc = np.empty((C, H), dtype=np.uint8)
    for c in range(C):
        props[c], acc[c] = oracle.classical_chain(J, h, kinds, params, probs, int(init[c]), H, T, seed, chain_id=c)
    return J, h, init, props, acc


@pytest.mark.gpu
@pytest.mark.parametrize("n,C,H", [(6, 5, 300), (10, 3, 1500), (14, 2, 400)])
def test_device_trajectory_statistics_match_the_reference_definitions(n, C, H):
    import qumcmc_b200 as q
    J, h, init, props, acc = _recorded_chains(n, C, H, seed=n)
    m = q.IsingEnergyFunction(J, h)
    _, _, P = oracle.exact_boltzmann(J, h, 1.0)
    res = q.batch_trajectory_statistics(m, init, props, acc, P)
    for c in range(C):
        mc = np_oracle.markov_chain_from(int(init[c]), props[c], acc[c])
        assert np.array_equal(res["markov_chain"][c], mc)                               # bit-exact (indices)
        ediff, ham = np_oracle.trajectory_statistics(J, h, int(init[c]), props[c], acc[c])
        assert np.array_equal(res["energy"][c], ediff)                                  # fixed-order fp64 energies
        assert np.array_equal(res["hamming"][c], ham)
        assert np.allclose(res["magnetisation"][c], np_oracle.running_magnetisation(mc, n), rtol=1e-14, atol=1e-14)
        rk = np_oracle.running_kl(P, mc)
        assert np.abs(res["kldiv"][c] - rk).max() < 1e-10 * max(1.0, np.abs(rk).max())  # fp64 tolerance (north star)


@pytest.mark.gpu
def test_reference_named_functions_on_one_chain():
    import qumcmc_b200 as q
    from qumcmc_b200.trajectory_processing import (calculate_running_kl_divergence, calculate_runnning_magnetisation,
                                                   get_trajectory_statistics)
    n, H = 7, 500
    J, h, init, props, acc = _recorded_chains(n, 1, H, seed=3)
    m = q.IsingEnergyFunction(J, h)
    ex = q.Exact_Sampling(m, 1.0)
    chain = q.MCMCChain.from_arrays(n, int(init[0]), props[0], acc[0].astype(bool))
    mc = np_oracle.markov_chain_from(int(init[0]), props[0], acc[0])
    P = np.array([ex.boltzmann_pd[f"{i:0{n}b}"] for i in range(1 << n)])
    rk = calculate_running_kl_divergence(ex.boltzmann_pd, chain)
    assert len(rk) == H + 1 and np.abs(np.array(rk) - np_oracle.running_kl(P, mc)).max() < 1e-10
    assert np.allclose(calculate_runnning_magnetisation(chain), np_oracle.running_magnetisation(mc, n), atol=1e-13)
    st = get_trajectory_statistics(chain, ex)
    assert set(st) == {"acceptance_prob", "energy", "hamming", "kldiv"}
    assert sum(v["total"] for v in st["hamming"].values()) == H
    # acceptance_prob: the reference's dwell-time loop (trajectory_processing.py:266-277), restated
    cur, counter, want = chain.states[0].bitstring, 1, []
    for s in chain.states:
        if s.accepted and s.bitstring != cur:
            want.append(1 / counter)
            cur, counter = s.bitstring, 1
        else:
            counter += 1
    want.append(1 / counter)
    assert np.array_equal(st["acceptance_prob"], np.array(want))


@pytest.mark.gpu
def test_running_kl_of_quantum_chains_decreases_towards_zero():
    """End to end on the hot path: 64 quantum chains at n=8, pooled over time the running KL to the exact Boltzmann
    distribution falls by an order of magnitude (the figure the reference's notebooks plot)."""
    import qumcmc_b200 as q
    m = q.random_ising_model(8, 4)
    ex = q.Exact_Sampling(m, 1.0)
    P = np.array([ex.boltzmann_pd.get(f"{i:08b}", 0.0) for i in range(256)])
    b = q.quantum_enhanced_mcmc(4000, m, q.GenericMixer.OneBodyMixer(8), (0.25, 0.6), temperature=1.0, n_chains=64, seed=7)
    kl = q.batch_trajectory_statistics(m, b.initial, b.proposals, b.accepted, P, to_observe=("kldiv",))["kldiv"]
    assert kl.shape == (64, 4001)
    assert np.median(kl[:, -1]) < 0.05 * np.median(kl[:, 20]) and np.median(kl[:, -1]) < 0.5   # unvisited tail: p log(p/EPS)


@pytest.mark.gpu
def test_running_kl_batches_that_are_not_a_multiple_of_the_block_size(monkeypatch):
    """with a hash-table budget that forces several batches of chains (batch = 5, CTA = 64 threads) every
    output must equal the unbatched run -- threads past the batch must not touch the next batch's chains."""
    import qumcmc_b200 as q
    n, C, H = 8, 23, 120
    J, h, init, props, acc = _recorded_chains(n, C, H, seed=3)
    m = q.IsingEnergyFunction(J, h)
    _, _, P = oracle.exact_boltzmann(J, h, 1.0)
    full = q.batch_trajectory_statistics(m, init, props, acc, P)
    cap = 256  # power of two >= 2 (H + 1)
    monkeypatch.setenv("QUMCMC_TRAJ_BUDGET_BYTES", str(5 * cap * 12))
    small = q.batch_trajectory_statistics(m, init, props, acc, P)
    for key in ("markov_chain", "energy", "hamming", "magnetisation", "kldiv"):
        assert np.array_equal(np.asarray(full[key]), np.asarray(small[key])), key
    for c in range(C):
        mc = np_oracle.markov_chain_from(int(init[c]), props[c], acc[c])
        rk = np_oracle.running_kl(P, mc)
        assert np.abs(small["kldiv"][c] - rk).max() < 1e-10 * max(1.0, np.abs(rk).max())
        _, ham = np_oracle.trajectory_statistics(J, h, int(init[c]), props[c], acc[c])
        assert np.array_equal(small["hamming"][c], ham)
