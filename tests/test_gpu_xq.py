"""Quadratic X mixers on the sweep kernels (sweep.cu, XQ mode): one coherent group of one- and two-qubit X strings
(GenericMixer.TwoBodyMixer, pair CustomMixer, CoherentMixerSum of them; qumcmc/mixers.py:21-130) at 14 <= n <= 20 in
complex64.  U = H Dx H (P H Dx H)^(r-1) with the mixer a quadratic diagonal in the x basis: 2 r sweeps per proposal.
Against the oracle's fused complex128 simulator (<= 1e-6 on probabilities) and against the Hadamard-basis general path."""
import numpy as np
import pytest

import oracle
from tests import models

pytestmark = pytest.mark.gpu


def _mixer(q, n, kind):
    if kind == "two":
        return q.GenericMixer.TwoBodyMixer(n)
    if kind == "pairs":   # pairs across and inside the qubit groups + a few singles, repeated qubits
        return q.CustomMixer(n, [[0, n - 1], [3, 7], [1], [n - 3, n - 2], [5, n - 2], [11, 12], [n - 1], [2, 9], [0, 1], [6]])
    if kind == "coh":
        return q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.GenericMixer.TwoBodyMixer(n)], [0.7, 0.3])
    raise ValueError(kind)


def _check(q, n, mixer, rs, seed=3, dtype="fp32", group=0):
    tol_p, tol_a = (1e-6, 5e-5) if dtype == "fp32" else (1e-10, 1e-9)
    J, h = models.packaged_random_model(n, seed)
    m = q.IsingEnergyFunction(J, h)
    offs, masks, w, _ = mixer.lower()
    masks, w = masks[offs[group]:offs[group + 1]], w[offs[group]:offs[group + 1]]
    al = oracle.alpha(J, h)
    rng = np.random.default_rng(7 * n + len(masks))
    for r in rs:
        s = int(rng.integers(0, 1 << n))
        gamma = float(rng.uniform(0.25, 0.6))
        p, a = q.transition_probabilities(m, mixer, s, gamma, r, dtype=dtype, return_amplitudes=True, group=group)
        ref = oracle.evolve(J, h, al, gamma, 0.8, r, masks, w, s, gatewise=False)
        assert abs(p.sum() - 1.0) < 1e-5
        assert np.abs(p - np.abs(ref) ** 2).max() < tol_p, (n, r, np.abs(p - np.abs(ref) ** 2).max())
        assert np.linalg.norm(a - ref) < tol_a, (n, r, np.linalg.norm(a - ref))


@pytest.mark.parametrize("n", [14, 15, 16, 17, 18, 19, 20])
def test_xq_two_body_vs_oracle(n):
    """Every HI geometry (W = n - 12 = 2..8, odd and even: the transform normalisation 2^(-W/2) is irrational for odd W),
    r = 1 (HI FIRST + LO FINAL only), even and odd r."""
    import qumcmc_b200 as q
    _check(q, n, _mixer(q, n, "two"), (1, 2, 3, 6) if n < 19 else (1, 4))


@pytest.mark.parametrize("n,kind", [(14, "pairs"), (17, "pairs"), (20, "pairs"), (15, "coh"), (18, "coh")])
def test_xq_custom_pairs_and_coherent_sums_vs_oracle(n, kind):
    import qumcmc_b200 as q
    _check(q, n, _mixer(q, n, kind), (1, 2, 5))


@pytest.mark.parametrize("n,kind", [(13, "two"), (14, "two"), (15, "coh"), (16, "two"), (17, "pairs"), (18, "two"), (19, "coh"), (20, "two")])
def test_xq_complex128_vs_oracle(n, kind):
    """The same in the reference's own precision (sweep128.cu, XQ mode): every HI geometry of the complex128 tiles
    (LC = 22 - n = 9..2), tolerance 1e-10 on the probabilities."""
    import qumcmc_b200 as q
    _check(q, n, _mixer(q, n, kind), (1, 2, 5) if n < 19 else (1, 3), dtype="fp64")


def test_xq_complex128_chains_equal_the_oracle_chain():
    """complex128 chains with a two-body mixer at n = 13: proposals, accept flags and final states equal oracle.run_chain
    hop for hop."""
    import qumcmc_b200 as q
    n, C, H, seed = 13, 6, 6, 41
    J, h = models.packaged_random_model(n, 7)
    m = q.IsingEnergyFunction(J, h)
    mixer = q.GenericMixer.TwoBodyMixer(n)
    offs, masks, w, _ = mixer.lower()
    b = q.quantum_enhanced_mcmc(H, m, mixer, (0.25, 0.6), temperature=1.0, n_chains=C, seed=seed, dtype="fp64")
    for c in range(C):
        s0 = oracle.initial_state(seed, c, n)
        props, acc = oracle.run_chain(J, h, masks, w, s0, H, 1.0, (0.25, 0.6), 0.8, seed, chain_id=c)
        assert np.array_equal(np.asarray(props, dtype=np.int64), b.proposals[c].astype(np.int64)), c
        assert np.array_equal(np.asarray(acc, dtype=np.uint8), b.accepted[c].astype(np.uint8)), c


@pytest.mark.parametrize("n,dtype", [(14, "fp32"), (17, "fp32"), (20, "fp32"), (13, "fp64"), (18, "fp64")])
def test_xq_incoherent_sum_uses_the_tables_of_the_chosen_group(n, dtype):
    """IncoherentMixerSum (qumcmc/mixers.py:133-151) picks one sub-mixer per proposal: every group -- one-body, two-body,
    pairs -- has its own mixer-diagonal tables, selected per chain inside the HI sweeps."""
    import qumcmc_b200 as q
    inc = q.IncoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.GenericMixer.TwoBodyMixer(n), _mixer(q, n, "pairs")], [0.3, 0.4, 0.3])
    for group in range(3):
        _check(q, n, inc, (1, 3) if n < 20 else (2,), dtype=dtype, group=group)


def test_xq_incoherent_chains_equal_the_general_path(monkeypatch):
    """Chains whose proposals draw different sub-mixers inside one launch (complex128, n = 13): proposals and accept flags
    equal the general path's, which evolves every group with its own diagonal."""
    import qumcmc_b200 as q
    n = 13
    J, h = models.packaged_random_model(n, 9)
    m = q.IsingEnergyFunction(J, h)
    inc = q.IncoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.GenericMixer.TwoBodyMixer(n)], [0.5, 0.5])
    kw = dict(temperature=1.0, n_chains=24, seed=77, dtype="fp64")
    got = q.quantum_enhanced_mcmc(4, m, inc, (0.25, 0.6), **kw)
    monkeypatch.setenv("QUMCMC_NO_XQ", "1")
    ref = q.quantum_enhanced_mcmc(4, m, inc, (0.25, 0.6), **kw)
    assert np.array_equal(ref.proposals, got.proposals)
    assert np.array_equal(ref.accepted, got.accepted)


def test_xq_gamma_extremes_and_absent_phase_layer():
    """gamma = 1 (no phase layer: c = 0), gamma = 0 (identity mixer) and a model below the skip threshold of
    fn_qckt_problem_half (quantum_mcmc_routines.py:64-67)."""
    import qumcmc_b200 as q
    n = 16
    J, h = models.packaged_random_model(n, 5)
    mixer = _mixer(q, n, "two")
    offs, masks, w, _ = mixer.lower()
    for Jm, hm in ((J, h), (J * 1e-4, h * 1e-4)):
        m = q.IsingEnergyFunction(Jm, hm)
        al = oracle.alpha(Jm, hm)
        for gamma in (1.0, 0.0, 0.37):
            p = q.transition_probabilities(m, mixer, 12345, gamma, 3, dtype="fp32")
            ref = oracle.evolve(Jm, hm, al, gamma, 0.8, 3, masks, w, 12345, gatewise=False)   # skip rule inside the oracle
            assert np.abs(p - np.abs(ref) ** 2).max() < 1e-6, gamma


@pytest.mark.parametrize("n", [14, 20])
def test_xq_matches_the_general_path(monkeypatch, n):
    import qumcmc_b200 as q
    J, h = models.packaged_random_model(n, 2)
    m = q.IsingEnergyFunction(J, h)
    mixer = _mixer(q, n, "coh")
    p_xq = q.transition_probabilities(m, mixer, 4321, 0.41, 4, dtype="fp32")
    monkeypatch.setenv("QUMCMC_NO_XQ", "1")
    p_gen = q.transition_probabilities(m, mixer, 4321, 0.41, 4, dtype="fp32")
    assert np.abs(p_xq - p_gen).max() < 2e-6


def test_xq_chains_follow_the_oracle():
    """quantum_enhanced_mcmc with a two-body mixer at n = 14: proposals equal the oracle's sequential inverse CDF (up to
    CDF-boundary ties in fp32); accept flags / counts / final states / Hamming histograms follow bit-exactly."""
    import qumcmc_b200 as q
    n, C, H, seed = 14, 8, 5, 19
    J, h = models.packaged_random_model(n, 8)
    m = q.IsingEnergyFunction(J, h)
    mixer = q.GenericMixer.TwoBodyMixer(n)
    offs, masks, w, _ = mixer.lower()
    b = q.quantum_enhanced_mcmc(H, m, mixer, (0.25, 0.6), temperature=1.0, n_chains=C, seed=seed, dtype="fp32")
    al = oracle.alpha(J, h)
    mismatched = 0
    for c in range(C):
        cur = oracle.initial_state(seed, c, n)
        assert int(b.initial[c]) == cur
        e_cur = oracle.energy(J, h, cur)
        hist = np.zeros(n + 1, dtype=np.int64)
        nacc = 0
        for hop in range(H):
            ug, t, _, um, ua = oracle.hop_draws(seed, c, hop)
            gamma = 0.25 + (0.6 - 0.25) * ug
            r = oracle.trotter_steps(t, 0.8)
            st = oracle.evolve(J, h, al, gamma, 0.8, r, masks, w, cur, gatewise=False)
            want = oracle.sample_index(st, um)
            sp = int(b.proposals[c, hop])
            if sp != want:
                cdf = np.cumsum(np.abs(st) ** 2)
                lo = cdf[sp - 1] if sp > 0 else 0.0
                assert lo - 5e-6 <= um <= cdf[sp] + 5e-6, (c, hop, sp, want)
                mismatched += 1
            e_sp = oracle.energy(J, h, sp)
            acc = oracle.test_accept(e_cur, e_sp, 1.0, ua)
            assert bool(b.accepted[c, hop]) == acc
            hist[bin(cur ^ sp).count("1")] += 1
            if acc:
                cur, e_cur, nacc = sp, e_sp, nacc + 1
        assert int(b.final_state[c]) == cur and int(b.acc_count[c]) == nacc
        assert np.array_equal(b.hamming_hist[c], hist)
    assert mismatched <= 2


@pytest.mark.parametrize("n,C,mib", [(14, 150, 1), (18, 30, 4), (20, 11, 24)])
def test_xq_group_schedule_is_bit_identical(monkeypatch, n, C, mib):
    """The L2-resident group schedule with 2 r sweeps per proposal (all chains in the even class: HI first, LO last)."""
    import qumcmc_b200 as q
    J, h = models.packaged_random_model(n, 6)
    m = q.IsingEnergyFunction(J, h)
    mixer = _mixer(q, n, "two")
    kw = dict(temperature=1.0, n_chains=C, seed=31, dtype="fp32")
    monkeypatch.setenv("QUMCMC_L2_GROUP_MIB", "0")
    ref = q.quantum_enhanced_mcmc(2, m, mixer, (0.25, 0.6), **kw)
    monkeypatch.setenv("QUMCMC_L2_GROUP_MIB", str(mib))
    got = q.quantum_enhanced_mcmc(2, m, mixer, (0.25, 0.6), **kw)
    for f in ("proposals", "accepted", "final_state", "acc_count", "hamming_hist"):
        assert np.array_equal(getattr(ref, f), getattr(got, f)), f


def test_mixers_outside_xq_keep_their_paths(monkeypatch):
    """Groups with a three-body term are not quadratic: with the tabulated mode switched off they run on the general path and
    agree with the oracle."""
    monkeypatch.setenv("QUMCMC_NO_XT", "1")
    import qumcmc_b200 as q
    n = 14
    J, h = models.packaged_random_model(n, 3)
    m = q.IsingEnergyFunction(J, h)
    al = oracle.alpha(J, h)
    inc = q.IncoherentMixerSum([q.GenericMixer.TwoBodyMixer(n), q.CustomMixer(n, [[0, 1, 2], [3, 4], [5]])], [0.5, 0.5])
    for mixer, dtype, group in ((q.CustomMixer(n, [[0, 1, 2], [3, 4], [5]]), "fp32", 0), (inc, "fp32", 0), (inc, "fp64", 1)):
        offs, masks, w, _ = mixer.lower()
        gm, gw = masks[offs[group]:offs[group + 1]], w[offs[group]:offs[group + 1]]
        p = q.transition_probabilities(m, mixer, 77, 0.4, 3, dtype=dtype, group=group)
        ref = oracle.evolve(J, h, al, 0.4, 0.8, 3, gm, gw, 77, gatewise=False)
        assert np.abs(p - np.abs(ref) ** 2).max() < (1e-6 if dtype == "fp32" else 1e-10)


# ---------------------------------------------------------------------------------------------------------------
# XT mode: every other X-string mixer at 14 <= n <= 20 in complex64 -- the mixer diagonal from the tabulated exponent
# ---------------------------------------------------------------------------------------------------------------
def _xt_mixer(q, n, kind):
    if kind == "three":
        return q.GenericMixer.ThreeBodyMixer(n)
    if kind == "custom":   # weights 1..4, terms inside and across the qubit groups
        return q.CustomMixer(n, [[0, 5, n - 1], [1, 2, 3, 4], [n - 2], [12, 13], [6, 11, 12], [n - 3, n - 2, n - 1], [0, 1]])
    if kind == "coh":
        return q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.GenericMixer.ThreeBodyMixer(n)], [0.8, 0.2])
    raise ValueError(kind)


@pytest.mark.parametrize("n,kind", [(14, "three"), (15, "custom"), (16, "coh"), (17, "three"), (18, "custom"), (19, "coh"), (20, "three")])
def test_xt_mixers_vs_oracle(n, kind):
    import qumcmc_b200 as q
    _check(q, n, _xt_mixer(q, n, kind), (1, 2, 5) if n < 19 else (1, 3))


def test_xt_incoherent_sum_and_general_path_agreement(monkeypatch):
    """Incoherent sum with a three-body member (every group reads its own table) per group against the oracle, and the
    whole-path agreement with general.cu."""
    import qumcmc_b200 as q
    n = 15
    inc = q.IncoherentMixerSum([q.GenericMixer.TwoBodyMixer(n), q.GenericMixer.ThreeBodyMixer(n)], [0.5, 0.5])
    for group in (0, 1):
        _check(q, n, inc, (2, 3), group=group)
    J, h = models.packaged_random_model(n, 2)
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.ThreeBodyMixer(n)
    p_xt = q.transition_probabilities(m, mx, 999, 0.33, 4, dtype="fp32")
    monkeypatch.setenv("QUMCMC_NO_XT", "1")
    p_gen = q.transition_probabilities(m, mx, 999, 0.33, 4, dtype="fp32")
    assert np.abs(p_xt - p_gen).max() < 2e-6


def test_xt_chains_and_group_schedule(monkeypatch):
    """Three-body chains at n = 14: bit-identical between the schedules, accept bookkeeping follows the proposals."""
    import qumcmc_b200 as q
    n, C, H, seed = 14, 120, 3, 11
    J, h = models.packaged_random_model(n, 6)
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.ThreeBodyMixer(n)
    kw = dict(temperature=1.0, n_chains=C, seed=seed, dtype="fp32")
    monkeypatch.setenv("QUMCMC_L2_GROUP_MIB", "0")
    ref = q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), **kw)
    monkeypatch.setenv("QUMCMC_L2_GROUP_MIB", "1")
    got = q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), **kw)
    for f in ("proposals", "accepted", "final_state", "acc_count", "hamming_hist"):
        assert np.array_equal(getattr(ref, f), getattr(got, f)), f
    for c in range(0, C, 17):
        cur = oracle.initial_state(seed, c, n)
        e_cur = oracle.energy(J, h, cur)
        for hop in range(H):
            _, _, _, _, ua = oracle.hop_draws(seed, c, hop)
            sp = int(got.proposals[c, hop])
            e_sp = oracle.energy(J, h, sp)
            acc = oracle.test_accept(e_cur, e_sp, 1.0, ua)
            assert bool(got.accepted[c, hop]) == acc
            if acc:
                cur, e_cur = sp, e_sp
        assert int(got.final_state[c]) == cur


@pytest.mark.parametrize("n,kind", [(13, "three"), (14, "custom"), (16, "coh"), (17, "three"), (20, "custom")])
def test_xt_complex128_vs_oracle(n, kind):
    """XT mode in the reference's precision (sweep128.cu): tabulated fp64 exponent, sincospi per amplitude; 1e-10."""
    import qumcmc_b200 as q
    _check(q, n, _xt_mixer(q, n, kind), (1, 2, 4) if n < 19 else (1, 3), dtype="fp64")


def test_xt_complex128_chains_equal_the_oracle_chain():
    import qumcmc_b200 as q
    n, C, H, seed = 13, 5, 5, 23
    J, h = models.packaged_random_model(n, 5)
    m = q.IsingEnergyFunction(J, h)
    mixer = q.GenericMixer.ThreeBodyMixer(n)
    offs, masks, w, _ = mixer.lower()
    b = q.quantum_enhanced_mcmc(H, m, mixer, (0.25, 0.6), temperature=1.0, n_chains=C, seed=seed, dtype="fp64")
    for c in range(C):
        s0 = oracle.initial_state(seed, c, n)
        props, acc = oracle.run_chain(J, h, masks, w, s0, H, 1.0, (0.25, 0.6), 0.8, seed, chain_id=c)
        assert np.array_equal(np.asarray(props, dtype=np.int64), b.proposals[c].astype(np.int64)), c
        assert np.array_equal(np.asarray(acc, dtype=np.uint8), b.accepted[c].astype(np.uint8)), c
