"""GPU parity for the HBM sweep path (14 <= n <= 20, complex64): transition probabilities,
measurement indices and chain bookkeeping vs the CPU oracle; full-size properties at n = 20."""
import numpy as np
import pytest

import oracle
from oracle import np_oracle as npo
from tests import models

pytestmark = pytest.mark.gpu


def _api():
    import qumcmc_b200 as q
    return q


def _setup(n, seed=3):
    q = _api()
    J, h = models.packaged_random_model(n, seed)
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.OneBodyMixer(n)
    masks, w = models.one_body(n)
    return q, J, h, m, mx, masks, w


@pytest.mark.parametrize("n", [14, 15, 16, 17, 18, 19, 20])
def test_sweep_transition_probabilities(n):
    """Every tile geometry (W = n-12 = 2..8), LO-first (odd r) and HI-first (even r) schedules."""
    q, J, h, m, mx, masks, w = _setup(n)
    al = oracle.alpha(J, h)
    rng = np.random.default_rng(n)
    rs = (1, 2, 3, 6, 13) if n <= 18 else (2, 3, 8)
    for r in rs:
        s = int(rng.integers(0, 1 << n))
        gamma = float(rng.uniform(0.25, 0.6))
        p, a = q.transition_probabilities(m, mx, s, gamma, r, dtype="fp32", return_amplitudes=True)
        ref = oracle.evolve(J, h, al, gamma, 0.8, r, masks, w, s, gatewise=False)
        pref = np.abs(ref) ** 2
        assert abs(p.sum() - 1.0) < 1e-5
        assert np.abs(p - pref).max() < 1e-6, (n, r)                      # north-star tolerance (fp32)
        assert np.linalg.norm(a - ref) < 2e-5, (n, r, np.linalg.norm(a - ref))  # relative L2 (|ref| = 1)


def test_sweep_gamma_extremes_and_skip_rule():
    q, J, h, m, mx, masks, w = _setup(15, seed=8)
    al = oracle.alpha(J, h)
    for gamma in (0.01, 0.95):
        p = q.transition_probabilities(m, mx, 77, gamma, 7, dtype="fp32")
        ref = oracle.transition_probs(J, h, al, gamma, 0.8, 7, masks, w, 77, gatewise=False)
        assert np.abs(p - ref).max() < 1e-6
    n = 14
    Jz = np.full((n, n), 4e-4); np.fill_diagonal(Jz, 0.0)
    hz = np.full(n, 4e-4)
    mz = q.IsingEnergyFunction(Jz, hz)
    mk, wk = models.one_body(n)
    p = q.transition_probabilities(mz, q.GenericMixer.OneBodyMixer(n), 5, 0.3, 6, dtype="fp32")
    ref = oracle.transition_probs(Jz, hz, oracle.alpha(Jz, hz), 0.3, 0.8, 6, mk, wk, 5, gatewise=False)
    assert np.abs(p - ref).max() < 1e-6


@pytest.mark.parametrize("n,C", [(14, 96), (16, 64), (20, 24), (22, 8), (25, 3)])   # n > 20: two-level (coarse) scan
def test_sweep_propose_indices(n, C):
    """qmc_propose with explicit (s, gamma, r, u): the sampled index equals the oracle's sequential
    inverse CDF except where u sits within fp32 rounding of a CDF boundary."""
    import torch
    from qumcmc_b200 import _lib
    q, J, h, m, mx, masks, w = _setup(n, seed=5)
    dm = m.device_model(0)
    dm.set_mixer(mx)
    rng = np.random.default_rng(n)
    s_in = rng.integers(0, 1 << n, size=C).astype(np.int64)
    gam = rng.uniform(0.25, 0.6, size=C)
    rr = rng.choice([1, 2, 3, 5, 6, 7, 8, 10, 11, 12, 13], size=C).astype(np.int32)
    uu = rng.uniform(0, 1, size=C)
    dev = torch.device("cuda", 0)
    t = lambda x: torch.from_numpy(x).to(dev)
    d_s, d_g, d_r, d_u = t(s_in), t(gam), t(rr), t(uu)
    out = torch.empty(C, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().qmc_propose(dm.handle, C, d_s.data_ptr(), d_g.data_ptr(), d_r.data_ptr(), d_u.data_ptr(),
                                      None, 0.8, _lib.DTYPE_F32, out.data_ptr(), None))
    got = out.cpu().numpy()
    al = oracle.alpha(J, h)
    near = 0
    for c in range(C):
        st = oracle.evolve(J, h, al, gam[c], 0.8, int(rr[c]), masks, w, int(s_in[c]), gatewise=False)
        exp = oracle.sample_index(st, uu[c])
        if exp != got[c]:
            cdf = np.cumsum(np.abs(st) ** 2)
            # the fp32 CDF may be off by ~1e-6: the device index must be a neighbour in CDF terms
            lo = cdf[got[c] - 1] if got[c] > 0 else 0.0
            assert lo - 5e-6 <= uu[c] <= cdf[got[c]] + 5e-6, (c, exp, int(got[c]))
            near += 1
    assert near <= max(2, C // 3)


def test_sweep_chain_bookkeeping_exact():
    """K4 is independent of the statevector numerics: given the device's proposals, the accept
    flags, acceptance counts, final states and Hamming histograms must equal the oracle's
    Metropolis rule with the same Philox uniforms, bit for bit."""
    q, J, h, m, mx, masks, w = _setup(14, seed=9)
    C, H, seed, beta = 48, 40, 99, 1.0
    b = q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1 / beta, n_chains=C, seed=seed, dtype="fp32",
                                chain_id0=5)
    for c in range(C):
        cur = oracle.initial_state(seed, 5 + c, 14)
        assert int(b.initial[c]) == cur
        e_cur = oracle.energy(J, h, cur)
        nacc = 0
        hist = np.zeros(15, dtype=np.int64)
        for hop in range(H):
            _, _, _, _, ua = oracle.hop_draws(seed, 5 + c, hop)
            sp = int(b.proposals[c, hop])
            e_sp = oracle.energy(J, h, sp)
            acc = oracle.test_accept(e_cur, e_sp, 1 / beta, ua)
            assert bool(b.accepted[c, hop]) == acc
            hist[bin(cur ^ sp).count("1")] += 1
            if acc:
                cur, e_cur = sp, e_sp
                nacc += 1
        assert int(b.final_state[c]) == cur and int(b.acc_count[c]) == nacc
        assert np.array_equal(b.hamming_hist[c], hist)


def test_sweep_chain_statistics_vs_oracle_chains():
    """Acceptance rate and proposal-Hamming distribution at n = 14 vs oracle chains (same model,
    different random streams are fine: compare distributions)."""
    q, J, h, m, mx, masks, w = _setup(14, seed=4)
    beta = 1.0
    b = q.quantum_enhanced_mcmc(60, m, mx, (0.25, 0.6), temperature=1 / beta, n_chains=256, seed=7, dtype="fp32")
    acc_gpu = b.acceptance_rate()
    ham_gpu = b.hamming_hist.sum(axis=0) / b.hamming_hist.sum()
    accs, hams = [], np.zeros(15)
    for c in range(24):
        s0 = oracle.initial_state(7, c, 14)
        props, acc = oracle.run_chain(J, h, masks, w, s0, 60, 1 / beta, (0.25, 0.6), 0.8, 7, chain_id=c)
        accs.append(acc.mean())
        mc = npo.markov_chain_from(s0, props, acc)
        hams += np.bincount([bin(int(x) ^ int(y)).count("1") for x, y in zip(mc[:-1], props)], minlength=15)
    assert abs(acc_gpu - np.mean(accs)) < 0.03
    assert np.abs(ham_gpu - hams / hams.sum()).max() < 0.03


def test_sweep_full_size_properties_n20():
    """BASELINE config C3 shape (n = 20): norm conservation through 13 layers, symmetry
    Q(s->s') = Q(s'->s) (U^T = U), and chain-sharding invariance of the Philox streams."""
    q, J, h, m, mx, masks, w = _setup(20, seed=1)
    s, sp = 123456, 654321
    p1 = q.transition_probabilities(m, mx, s, 0.4, 13, dtype="fp32")
    p2 = q.transition_probabilities(m, mx, sp, 0.4, 13, dtype="fp32")
    assert abs(p1.sum() - 1.0) < 1e-5
    assert abs(p1[sp] - p2[s]) < 1e-9 + 1e-4 * p1[sp]
    full = q.quantum_enhanced_mcmc(3, m, mx, (0.25, 0.6), temperature=1.0, n_chains=8, seed=5, dtype="fp32")
    half = q.quantum_enhanced_mcmc(3, m, mx, (0.25, 0.6), temperature=1.0, n_chains=4, seed=5, dtype="fp32", chain_id0=4)
    assert np.array_equal(full.proposals[4:], half.proposals)
    assert np.array_equal(full.accepted[4:], half.accepted)


def test_unsupported_configurations_fail_loudly():
    """Above the general path's limit (n = 30) k-body mixers and complex128 have no kernel -- and no CPU fallback."""
    q = _api()
    from qumcmc_b200 import _lib
    J, h = models.packaged_random_model(31, 2)
    m = q.IsingEnergyFunction(J, h)
    with pytest.raises(_lib.QmcUnsupported):
        q.transition_probabilities(m, q.GenericMixer.TwoBodyMixer(31), 0, 0.3, 2, dtype="fp32")
    with pytest.raises(_lib.QmcUnsupported):
        q.transition_probabilities(m, q.GenericMixer.OneBodyMixer(31), 0, 0.3, 2, dtype="fp64")


@pytest.mark.parametrize("n,rs", [(21, (1, 2, 3, 6)), (22, (2, 5)), (23, (3, 4)), (24, (2, 7)), (25, (3,)), (26, (4,))])
def test_generic_path_transition_probabilities(n, rs):
    """n >= 21: LO group + an HI8 group + a partial HI6 group with 1..6 active qubits (n = 21..26), vs the
    oracle's fused simulator.  (HI8 windows with 7 active qubits: tests/test_gpu_shard.py, 19 local qubits.)"""
    q, J, h, m, mx, masks, w = _setup(n, seed=2)
    al = oracle.alpha(J, h)
    rng = np.random.default_rng(n)
    for r in rs:
        s = int(rng.integers(0, 1 << n))
        gamma = float(rng.uniform(0.25, 0.6))
        p, a = q.transition_probabilities(m, mx, s, gamma, r, dtype="fp32", return_amplitudes=True)
        ref = oracle.evolve(J, h, al, gamma, 0.8, r, masks, w, s, gatewise=False)
        assert abs(p.sum() - 1.0) < 1e-5
        assert np.abs(p - np.abs(ref) ** 2).max() < 1e-6, (n, r)
        assert np.linalg.norm(a - ref) < 3e-5, (n, r, np.linalg.norm(a - ref))


def test_generic_path_chain_bookkeeping_n22():
    q, J, h, m, mx, masks, w = _setup(22, seed=6)
    C, H, seed = 6, 5, 31
    b = q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.0, n_chains=C, seed=seed, dtype="fp32")
    for c in range(C):
        cur = oracle.initial_state(seed, c, 22)
        e_cur = oracle.energy(J, h, cur)
        for hop in range(H):
            _, _, _, _, ua = oracle.hop_draws(seed, c, hop)
            sp = int(b.proposals[c, hop])
            e_sp = oracle.energy(J, h, sp)
            acc = oracle.test_accept(e_cur, e_sp, 1.0, ua)
            assert bool(b.accepted[c, hop]) == acc
            if acc:
                cur, e_cur = sp, e_sp
        assert int(b.final_state[c]) == cur


# ---------------------------------------------------------------------------------------------------------------
# kernel geometries and variants that ran without an oracle comparison
# ---------------------------------------------------------------------------------------------------------------
def _fresh_model(n, seed):
    """Model with its own device handle (the handle caches the group geometry / phase tables it was built with)."""
    from qumcmc_b200 import _device
    _device._cache.clear()
    return _setup(n, seed)


def _check_vs_oracle(q, J, h, m, mx, masks, w, n, r, s, gamma, tol_p=1e-6, tol_a=3e-5):
    al = oracle.alpha(J, h)
    p, a = q.transition_probabilities(m, mx, s, gamma, r, dtype="fp32", return_amplitudes=True)
    ref = oracle.evolve(J, h, al, gamma, 0.8, r, masks, w, s, gatewise=False)
    assert abs(p.sum() - 1.0) < 1e-5
    err_p = np.abs(p - np.abs(ref) ** 2).max()
    err_a = np.linalg.norm(a - ref)
    assert err_p < tol_p and err_a < tol_a, (n, r, err_p, err_a)


@pytest.mark.parametrize("n,groups,rs", [
    (22, "6:4,6:3,6:3", (2, 3, 4, 7)),        # K = 3 register groups: the (carried + 1) % K rotation of run_generic_schedule
    (24, "6:3,6:3,6:3,6:3", (2, 5, 6)),       # K = 4
    (23, "6:6,6:5", (3, 4)),                  # full HI6 group on top of a partial one
    (24, "8:8,6:2,6:2", (2, 3, 5)),           # HI8 on top of two HI6 groups (the shape of C5's local shard)
    (26, "8:7,8:7", (2, 3)),                  # two stacked HI8 windows, 7 active qubits each
])
def test_generic_schedules_with_forced_group_geometry(monkeypatch, n, groups, rs):
    """QUMCMC_HI_GROUPS forces the register-group cover, so the schedules that BASELINE's large configurations use
    (C4 n = 28: two HI8 windows; n >= 29 and C5's 33 local qubits: K = 3) run at sizes the oracle finishes in seconds."""
    monkeypatch.setenv("QUMCMC_HI_GROUPS", groups)
    q, J, h, m, mx, masks, w = _fresh_model(n, seed=12)
    rng = np.random.default_rng(n + len(groups))
    try:
        for r in rs:
            _check_vs_oracle(q, J, h, m, mx, masks, w, n, r, int(rng.integers(0, 1 << n)), float(rng.uniform(0.25, 0.6)))
    finally:
        from qumcmc_b200 import _device
        _device._cache.clear()


def test_forced_group_geometry_is_validated(monkeypatch):
    from qumcmc_b200 import _device
    for bad in ("6:4,6:3", "8:6,6:4", "6:4;6:6", "6:7,6:3"):
        monkeypatch.setenv("QUMCMC_HI_GROUPS", bad)
        q, J, h, m, mx, masks, w = _fresh_model(22, seed=12)
        with pytest.raises(ValueError):
            q.transition_probabilities(m, mx, 1, 0.3, 2, dtype="fp32")
    _device._cache.clear()


@pytest.mark.slow
@pytest.mark.parametrize("n,rs", [(27, (2, 3)), (28, (2, 3, 6)), (29, (3,))])
def test_large_n_geometries_vs_oracle(n, rs):
    """The default geometries of the large single-GPU configurations against the oracle's fused complex128 simulator:
    n = 27 (HI8 + HI8 window with 7 active qubits), n = 28 = BASELINE C4 (two HI8 windows at p = 20 and p = 12),
    n = 29 (K = 3: HI8, HI8, HI6 with one active qubit).  Needs 2^n x 16 B of host memory for the oracle (8 GiB at n = 29)."""
    q, J, h, m, mx, masks, w = _fresh_model(n, seed=1)
    rng = np.random.default_rng(n)
    for r in rs:
        _check_vs_oracle(q, J, h, m, mx, masks, w, n, r, int(rng.integers(0, 1 << n)), float(rng.uniform(0.25, 0.6)),
                         tol_a=5e-5)
    from qumcmc_b200 import _device
    _device._cache.clear()   # frees the 2^n workspace


def _nonfast_mixers(q, n):
    """One-body mixers for which the per-chain constant-bank multiplier does not apply (sweep kernels with FAST = false)."""
    subset = [[b] for b in range(n) if b % 3 != 1]                     # CustomMixer of single-qubit terms on a subset
    custom = q.CustomMixer(n, subset)
    mixed = q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.CustomMixer(n, [[0], [5], [n - 1]])], [0.7, 0.3])
    return {"subset": custom, "sum": mixed}


@pytest.mark.parametrize("n", [14, 17, 20])
@pytest.mark.parametrize("which", ["subset", "sum", "env"])
def test_non_fast_sweep_kernels(monkeypatch, n, which):
    """lo_sweep_kernel<.., FAST=false> / hi_sweep_kernel<.., FAST=false> (per-qubit multipliers from the parameter record,
    normalisation inside the phase factor): reached by one-body mixers with unequal weights -- a CustomMixer of single-qubit
    terms on a subset of the qubits, a CoherentMixerSum of one-body mixers (weights on a shared qubit add up) -- and,
    for the uniform mixer, with QUMCMC_SWEEP_FAST=0.  r odd and even (LO-first and HI-first schedules), r = 1."""
    q, J, h, m, mx, masks, w = _setup(n, seed=13)
    if which == "env":
        monkeypatch.setenv("QUMCMC_SWEEP_FAST", "0")
    else:
        mx = _nonfast_mixers(q, n)[which]
        _, mk, wk, _ = mx.lower()
        masks, w = [int(x) for x in mk], [float(x) for x in wk]
    rng = np.random.default_rng(n)
    for r in ((1, 2, 3, 6, 13) if n < 20 else (2, 5)):
        _check_vs_oracle(q, J, h, m, mx, masks, w, n, r, int(rng.integers(0, 1 << n)), float(rng.uniform(0.25, 0.6)), tol_a=2e-5)


def test_non_fast_fold_underflow_case_n20():
    """n = 20, gamma ~ 0.95, r = 13: (prod cos)^r leaves the fp32 range, so the fold is refused and the unfolded
    kernels run even for the uniform one-body mixer."""
    q, J, h, m, mx, masks, w = _setup(20, seed=3)
    _check_vs_oracle(q, J, h, m, mx, masks, w, 20, 13, 4242, 0.95, tol_a=5e-5)


def test_non_fast_chains_match_the_oracle_sampler():
    """Chain bookkeeping through the non-FAST kernels: proposals of a CoherentMixerSum of one-body mixers at n = 14 equal
    the oracle's sampler up to fp32 CDF rounding, accept flags follow exactly."""
    q, J, h, m, _, _, _ = _setup(14, seed=21)
    mx = _nonfast_mixers(q, 14)["sum"]
    _, mk, wk, _ = mx.lower()
    masks, w = [int(x) for x in mk], [float(x) for x in wk]
    C, H, seed = 16, 6, 77
    b = q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.0, n_chains=C, seed=seed, dtype="fp32")
    al = oracle.alpha(J, h)
    near = 0
    for c in range(C):
        cur = oracle.initial_state(seed, c, 14)
        e_cur = oracle.energy(J, h, cur)
        for hop in range(H):
            ug, t, _, um, ua = oracle.hop_draws(seed, c, hop)
            gamma = 0.25 + (0.6 - 0.25) * ug
            r = max(1, int(np.floor(t / 0.8)))
            st = oracle.evolve(J, h, al, gamma, 0.8, r, masks, w, cur, gatewise=False)
            exp = oracle.sample_index(st, um)
            sp = int(b.proposals[c, hop])
            if sp != exp:
                cdf = np.cumsum(np.abs(st) ** 2)
                lo = cdf[sp - 1] if sp > 0 else 0.0
                assert lo - 5e-6 <= um <= cdf[sp] + 5e-6, (c, hop, exp, sp)
                near += 1
            e_sp = oracle.energy(J, h, sp)
            acc = oracle.test_accept(e_cur, e_sp, 1.0, ua)
            assert bool(b.accepted[c, hop]) == acc
            if acc:
                cur, e_cur = sp, e_sp
        assert int(b.final_state[c]) == cur
    assert near <= C * H // 10


@pytest.mark.parametrize("n,C,mib,streams", [(14, 200, 1, 3), (16, 96, 2, 2), (18, 40, 6, 4), (20, 20, 24, 2), (20, 14, 48, 1)])
def test_l2_resident_group_schedule_is_bit_identical(monkeypatch, n, C, mib, streams):
    """The L2-resident schedule (groups of chains run all their sweeps back to back on side streams) launches the same
    kernels on the same data as the step-major schedule: proposals, accept flags, counts and histograms must be equal
    bit for bit, for groups that straddle r values, ragged last groups and both r parities."""
    q, J, h, m, mx, masks, w = _setup(n, seed=6)
    kw = dict(temperature=1.0, n_chains=C, seed=31, dtype="fp32", chain_id0=2)
    monkeypatch.setenv("QUMCMC_L2_GROUP_MIB", "0")
    ref = q.quantum_enhanced_mcmc(3, m, mx, (0.25, 0.6), **kw)
    monkeypatch.setenv("QUMCMC_L2_GROUP_MIB", str(mib))
    monkeypatch.setenv("QUMCMC_L2_STREAMS", str(streams))
    got = q.quantum_enhanced_mcmc(3, m, mx, (0.25, 0.6), **kw)
    assert np.array_equal(ref.proposals, got.proposals)
    assert np.array_equal(ref.accepted, got.accepted)
    assert np.array_equal(ref.final_state, got.final_state)
    assert np.array_equal(ref.acc_count, got.acc_count)
    assert np.array_equal(ref.hamming_hist, got.hamming_hist)


def test_l2_resident_group_schedule_vs_oracle_sampler(monkeypatch):
    """Grouped schedule against the oracle directly: every proposal of a 64-chain hop at n = 14 equals the index the
    oracle samples from its own complex128 evolution with the same Philox draws (groups of 8 chains, 3 streams)."""
    monkeypatch.setenv("QUMCMC_L2_GROUP_MIB", "1")
    monkeypatch.setenv("QUMCMC_L2_STREAMS", "3")
    q, J, h, m, mx, masks, w = _setup(14, seed=2)
    C, seed = 64, 17
    b = q.quantum_enhanced_mcmc(1, m, mx, (0.25, 0.6), temperature=1.0, n_chains=C, seed=seed, dtype="fp32")
    bad = 0
    for c in range(C):
        s0 = oracle.initial_state(seed, c, 14)
        props, _ = oracle.run_chain(J, h, masks, w, s0, 1, 1.0, (0.25, 0.6), 0.8, seed, chain_id=c)
        bad += int(props[0]) != int(b.proposals[c, 0])
    assert bad <= 1, bad  # fp32 vs fp64 CDF: a draw within 1e-6 of a boundary may differ


@pytest.mark.parametrize("n,C", [(21, 9), (22, 5)])
def test_l2_resident_schedule_on_the_generic_path_is_bit_identical(monkeypatch, n, C):
    """n = 21, 22: a few statevectors still fit the L2, every chain runs all its passes back to back on a side stream
    (4 / 2 chains in flight).  Same kernels, same data: chains equal the r-major schedule bit for bit."""
    q, J, h, m, mx, masks, w = _setup(n, seed=5)
    kw = dict(temperature=1.0, n_chains=C, seed=23, dtype="fp32")
    monkeypatch.setenv("QUMCMC_L2_GENERIC", "0")
    ref = q.quantum_enhanced_mcmc(2, m, mx, (0.25, 0.6), **kw)
    monkeypatch.delenv("QUMCMC_L2_GENERIC")
    got = q.quantum_enhanced_mcmc(2, m, mx, (0.25, 0.6), **kw)
    for f in ("proposals", "accepted", "final_state", "acc_count", "hamming_hist"):
        assert np.array_equal(getattr(ref, f), getattr(got, f)), f
