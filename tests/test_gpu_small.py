"""GPU parity for the warp-per-chain kernel (small_n.cu: warp_chain_kernel, one-body mixers, 8 <= n <= 10) against the
oracle and against the CTA-per-chain kernel it replaces on those shapes (QUMCMC_NO_WARP_KERNEL=1)."""
import numpy as np
import pytest

import oracle
from oracle import np_oracle as npo
from tests import models

pytestmark = pytest.mark.gpu


def _api():
    import qumcmc_b200 as q
    return q


@pytest.mark.parametrize("n", [8, 9, 10])
@pytest.mark.parametrize("dtype,tol_p,tol_a", [("fp64", 1e-10, 1e-9), ("fp32", 1e-6, 2e-5)])
def test_warp_kernel_transition_probabilities(n, dtype, tol_p, tol_a):
    """r = 1 (no transposition back), odd and even r (the layout the last layer ends in differs), uniform and unequal
    one-body weights, the skip rule, gamma extremes."""
    q = _api()
    J, h = models.packaged_random_model(n, 7)
    m = q.IsingEnergyFunction(J, h)
    al = oracle.alpha(J, h)
    rng = np.random.default_rng(n)
    mixers = [q.GenericMixer.OneBodyMixer(n), q.CustomMixer(n, [[b] for b in range(n) if b % 3 != 1]),
              q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.CustomMixer(n, [[0], [n - 1]])], [0.7, 0.3])]
    for mx in mixers:
        _, mk, wk, _ = mx.lower()
        masks, w = [int(x) for x in mk], [float(x) for x in wk]
        for r in (1, 2, 3, 6, 13):
            s = int(rng.integers(0, 1 << n))
            gamma = float(rng.uniform(0.05, 0.9))
            p, a = q.transition_probabilities(m, mx, s, gamma, r, dtype=dtype, return_amplitudes=True)
            ref = oracle.evolve(J, h, al, gamma, 0.8, r, masks, w, s, gatewise=False)
            assert np.abs(p - np.abs(ref) ** 2).max() < tol_p, (n, r, dtype)
            assert np.linalg.norm(a - ref) < tol_a, (n, r, np.linalg.norm(a - ref))
    Jz = np.full((n, n), 4e-4); np.fill_diagonal(Jz, 0.0)
    hz = np.full(n, 4e-4)
    mz = q.IsingEnergyFunction(Jz, hz)
    mk, wk = models.one_body(n)
    p = q.transition_probabilities(mz, q.GenericMixer.OneBodyMixer(n), 5, 0.3, 6, dtype=dtype)
    ref = oracle.transition_probs(Jz, hz, oracle.alpha(Jz, hz), 0.3, 0.8, 6, mk, wk, 5, gatewise=False)
    assert np.abs(p - ref).max() < tol_p


@pytest.mark.parametrize("n,C,H", [(8, 37, 40), (9, 64, 60), (10, 21, 50)])
def test_warp_kernel_chains_equal_the_oracle_and_the_cta_kernel(monkeypatch, n, C, H):
    """fp64 chains: proposals, accept flags, counts, final states, Hamming and visit histograms equal oracle.run_chain and
    the CTA-per-chain kernel, hop for hop; chain counts that do not fill the last CTA; resumed runs (hop0)."""
    q = _api()
    J, h = models.packaged_random_model(n, 11)
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.OneBodyMixer(n)
    masks, w = models.one_body(n)
    seed = 2024
    b = q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.1, n_chains=C, seed=seed, dtype="fp64", chain_id0=3,
                                visit_histogram=True)
    want = np.zeros(1 << n, dtype=np.int64)
    for c in range(C):
        s0 = oracle.initial_state(seed, 3 + c, n)
        props, acc = oracle.run_chain(J, h, masks, w, s0, H, 1.1, (0.25, 0.6), 0.8, seed, chain_id=3 + c)
        assert np.array_equal(b.proposals[c], props) and np.array_equal(b.accepted[c], acc), c
        mc = npo.markov_chain_from(s0, props, acc)
        assert int(b.final_state[c]) == int(mc[-1]) and int(b.acc_count[c]) == int(acc.sum())
        ham = np.bincount([bin(int(x) ^ int(y)).count("1") for x, y in zip(mc[:-1], props)], minlength=n + 1)
        assert np.array_equal(b.hamming_hist[c], ham)
        want += np.bincount(mc.astype(np.int64), minlength=1 << n)
    assert np.array_equal(b.visit_hist, want)
    h1 = H // 2
    b1 = q.quantum_enhanced_mcmc(h1, m, mx, (0.25, 0.6), temperature=1.1, n_chains=C, seed=seed, dtype="fp64", chain_id0=3)
    b2 = q.quantum_enhanced_mcmc(H - h1, m, mx, (0.25, 0.6), initial_state=[int(x) for x in b1.final_state], temperature=1.1,
                                 n_chains=C, seed=seed, dtype="fp64", chain_id0=3, hop0=h1)
    assert np.array_equal(np.concatenate([b1.proposals, b2.proposals], axis=1), b.proposals)
    monkeypatch.setenv("QUMCMC_NO_WARP_KERNEL", "1")
    old = q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.1, n_chains=C, seed=seed, dtype="fp64", chain_id0=3)
    assert np.array_equal(old.proposals, b.proposals) and np.array_equal(old.accepted, b.accepted)


def test_warp_kernel_fp32_chains_follow_the_oracle_sampler():
    """complex64: sampled indices equal the oracle's sequential inverse CDF except within fp32 rounding of a CDF boundary;
    bookkeeping follows exactly from the device's proposals."""
    q = _api()
    n, C, H, seed = 9, 48, 30, 5
    J, h, _ = models.bas3("both")
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.OneBodyMixer(n)
    masks, w = models.one_body(n)
    b = q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), initial_state="111000111", temperature=1 / 1.5, n_chains=C, seed=seed,
                                dtype="fp32")
    al = oracle.alpha(J, h)
    near = 0
    for c in range(C):
        cur = int("111000111", 2)
        e_cur = oracle.energy(J, h, cur)
        for hop in range(H):
            ug, t, _, um, ua = oracle.hop_draws(seed, c, hop)
            gamma = 0.25 + (0.6 - 0.25) * ug
            st = oracle.evolve(J, h, al, gamma, 0.8, oracle.trotter_steps(t, 0.8), masks, w, cur, gatewise=False)
            want = oracle.sample_index(st, um)
            sp = int(b.proposals[c, hop])
            if sp != want:
                cdf = np.cumsum(np.abs(st) ** 2)
                lo = cdf[sp - 1] if sp > 0 else 0.0
                assert lo - 5e-6 <= um <= cdf[sp] + 5e-6, (c, hop, sp, want)
                near += 1
            e_sp = oracle.energy(J, h, sp)
            acc = oracle.test_accept(e_cur, e_sp, 1 / 1.5, ua)
            assert bool(b.accepted[c, hop]) == acc
            if acc:
                cur, e_cur = sp, e_sp
        assert int(b.final_state[c]) == cur
    assert near <= C * H // 50


# ---------------------------------------------------------------------------------------------------------------
# Hadamard-basis mode of the CTA-per-chain kernel: groups with >= 2 n + 4 terms run as H . exp(-i gamma dt E_x) . H
# ---------------------------------------------------------------------------------------------------------------
def _kbody_mixers(q, n):
    return {"two": q.GenericMixer.TwoBodyMixer(n), "three": q.GenericMixer.ThreeBodyMixer(n),
            "coh": q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.GenericMixer.TwoBodyMixer(n)], [0.6, 0.4]),
            "inc": q.IncoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.GenericMixer.ThreeBodyMixer(n)], [0.5, 0.5])}


@pytest.mark.parametrize("n", [6, 9, 10, 12, 13])
@pytest.mark.parametrize("dtype,tol_p,tol_a", [("fp64", 1e-10, 1e-9), ("fp32", 1e-6, 2e-5)])
def test_hadamard_basis_mode_transition_probabilities(n, dtype, tol_p, tol_a):
    """Two-body, three-body, coherent and incoherent sums (one group below the threshold: term loop; one above) against
    the oracle; odd and even n (the last butterfly pass of an odd n handles one qubit), r = 1 and r > 1."""
    q = _api()
    if n == 13 and dtype == "fp64":
        pytest.skip("complex128 statevectors live in shared memory up to n = 12")
    J, h = models.packaged_random_model(n, 11)
    m = q.IsingEnergyFunction(J, h)
    al = oracle.alpha(J, h)
    rng = np.random.default_rng(100 + n)
    for name, mx in _kbody_mixers(q, n).items():
        offs, mk, wk, _ = mx.lower()
        for group in range(len(offs) - 1):
            masks, w = [int(x) for x in mk[offs[group]:offs[group + 1]]], [float(x) for x in wk[offs[group]:offs[group + 1]]]
            for r in (1, 2, 5):
                s = int(rng.integers(0, 1 << n))
                gamma = float(rng.uniform(0.1, 0.8))
                p, a = q.transition_probabilities(m, mx, s, gamma, r, dtype=dtype, return_amplitudes=True, group=group)
                ref = oracle.evolve(J, h, al, gamma, 0.8, r, masks, w, s, gatewise=False)
                assert np.abs(p - np.abs(ref) ** 2).max() < tol_p, (n, name, group, r, dtype)
                assert np.linalg.norm(a - ref) < tol_a, (n, name, group, r, np.linalg.norm(a - ref))


def test_hadamard_basis_mode_equals_the_term_loop(monkeypatch):
    """Same proposals from both formulations of the mixer (complex128, BAS 3x3 with a two-body mixer)."""
    q = _api()
    J, h, _ = models.bas3("both")
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.TwoBodyMixer(9)
    kw = dict(initial_state="111000111", temperature=1 / 1.5, n_chains=64, seed=5, dtype="fp64")
    got = q.quantum_enhanced_mcmc(40, m, mx, (0.25, 0.6), **kw)
    monkeypatch.setenv("QUMCMC_SMALL_HX", "0")
    ref = q.quantum_enhanced_mcmc(40, m, mx, (0.25, 0.6), **kw)
    assert np.array_equal(ref.proposals, got.proposals)
    assert np.array_equal(ref.accepted, got.accepted)


def test_hadamard_basis_mode_chains_equal_the_oracle_chain():
    q = _api()
    n = 8
    J, h = models.packaged_random_model(n, 4)
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.ThreeBodyMixer(n)
    _, mk, wk, _ = mx.lower()
    b = q.quantum_enhanced_mcmc(12, m, mx, (0.25, 0.6), temperature=1.0, n_chains=5, seed=3, dtype="fp64")
    for c in range(5):
        s0 = oracle.initial_state(3, c, n)
        props, acc = oracle.run_chain(J, h, mk, wk, s0, 12, 1.0, (0.25, 0.6), 0.8, 3, chain_id=c)
        assert np.array_equal(np.asarray(props, dtype=np.int64), b.proposals[c].astype(np.int64)), c
        assert np.array_equal(np.asarray(acc, dtype=np.uint8), b.accepted[c].astype(np.uint8)), c
