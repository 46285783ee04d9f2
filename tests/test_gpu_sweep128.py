"""GPU parity for the complex128 sweep kernels (sweep128.cu: one-body mixers, 13 <= n <= 20) -- the reference's own
precision (qulacs evolves complex128, qumcmc/quantum_mcmc_routines.py:150-152).  Tolerance: 1e-10 on the transition
probabilities (the north star's fp64 figure); sampled indices and whole chains equal the oracle's."""
import numpy as np
import pytest

import oracle
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


def _check(q, J, h, m, mx, masks, w, n, r, s, gamma):
    al = oracle.alpha(J, h)
    p, a = q.transition_probabilities(m, mx, s, gamma, r, dtype="fp64", return_amplitudes=True)
    ref = oracle.evolve(J, h, al, gamma, 0.8, r, masks, w, s, gatewise=False)
    assert p.dtype == np.float64 and abs(p.sum() - 1.0) < 1e-12
    err_p, err_a = np.abs(p - np.abs(ref) ** 2).max(), np.linalg.norm(a - ref)
    assert err_p < 1e-10 and err_a < 1e-9, (n, r, err_p, err_a)      # north-star tolerance (fp64)
    return err_p, err_a


@pytest.mark.parametrize("n", [13, 14, 15, 16, 17, 18, 19, 20])
def test_sweep128_transition_probabilities(n):
    """Every HI tile geometry (LC = 22 - n = 9..2: single-layout tiles for n <= 15, two layouts above), LO-first (odd r)
    and HI-first (even r) schedules, r = 1 (analytic product state)."""
    q, J, h, m, mx, masks, w = _setup(n)
    dm = m.device_model(0)
    launches0 = dm.launch_count()
    rng = np.random.default_rng(n)
    for r in ((1, 2, 3, 6, 13) if n <= 17 else (2, 3, 8)):
        _check(q, J, h, m, mx, masks, w, n, r, int(rng.integers(0, 1 << n)), float(rng.uniform(0.25, 0.6)))
    assert m.device_model(0).launch_count() > launches0


def test_sweep128_is_the_path_taken_and_beats_the_general_path_accuracy_contract(monkeypatch):
    """Same inputs through the fp64 sweeps and (QUMCMC_NO_SWEEP128) the general tiled path: both within 1e-10 of the
    oracle, and within 1e-12 of each other."""
    q, J, h, m, mx, masks, w = _setup(16, seed=5)
    p1 = q.transition_probabilities(m, mx, 4321, 0.37, 7, dtype="fp64")
    monkeypatch.setenv("QUMCMC_NO_SWEEP128", "1")
    p2 = q.transition_probabilities(m, mx, 4321, 0.37, 7, dtype="fp64")
    ref = oracle.transition_probs(J, h, oracle.alpha(J, h), 0.37, 0.8, 7, masks, w, 4321, gatewise=False)
    assert np.abs(p1 - ref).max() < 1e-10 and np.abs(p2 - ref).max() < 1e-10 and np.abs(p1 - p2).max() < 1e-12


@pytest.mark.parametrize("n", [14, 18])
def test_sweep128_unequal_weights_gamma_extremes_and_skip_rule(n):
    q, J, h, m, mx, masks, w = _setup(n, seed=8)
    rng = np.random.default_rng(n)
    sub = q.CustomMixer(n, [[b] for b in range(n) if b % 3 != 1])
    mixed = q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.CustomMixer(n, [[0], [5], [n - 1]])], [0.7, 0.3])
    for mixer in (sub, mixed):
        _, mk, wk, _ = mixer.lower()
        for r in (2, 5):
            _check(q, J, h, m, mixer, [int(x) for x in mk], [float(x) for x in wk], n, r, int(rng.integers(0, 1 << n)),
                   float(rng.uniform(0.25, 0.6)))
    for gamma in (0.01, 0.95):
        _check(q, J, h, m, mx, masks, w, n, 7, 77, gamma)
    Jz = np.full((n, n), 4e-4); np.fill_diagonal(Jz, 0.0)          # skip rule: mean |J|, mean |h| <= 1e-3 => no phase layer
    hz = np.full(n, 4e-4)
    mz = q.IsingEnergyFunction(Jz, hz)
    _check(q, Jz, hz, mz, mx, masks, w, n, 6, 5, 0.3)


@pytest.mark.parametrize("n,C", [(13, 40), (16, 24), (20, 6)])
def test_sweep128_propose_indices_equal_the_oracle_sampler(n, C):
    """qmc_propose with explicit (s, gamma, r, u): in complex128 the sampled index equals the oracle's sequential inverse
    CDF (a mismatch would need u within ~1e-13 of a CDF boundary)."""
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
                                      None, 0.8, _lib.DTYPE_F64, out.data_ptr(), None))
    got = out.cpu().numpy()
    al = oracle.alpha(J, h)
    for c in range(C):
        st = oracle.evolve(J, h, al, gam[c], 0.8, int(rr[c]), masks, w, int(s_in[c]), gatewise=False)
        assert oracle.sample_index(st, uu[c]) == got[c], (c, int(rr[c]))


@pytest.mark.parametrize("n,C,H", [(14, 12, 12), (13, 6, 8)])
def test_sweep128_chains_equal_oracle_run_chain(n, C, H):
    """fp64 chains at n = 14 equal oracle.run_chain hop for hop (proposals and accept flags), plus
    counts, final states and Hamming histograms."""
    q, J, h, m, mx, masks, w = _setup(n, seed=9)
    seed = 99
    b = q.quantum_enhanced_mcmc(H, m, mx, (0.25, 0.6), temperature=1.0, n_chains=C, seed=seed, dtype="fp64", chain_id0=5)
    for c in range(C):
        s0 = oracle.initial_state(seed, 5 + c, n)
        assert int(b.initial[c]) == s0
        props, acc = oracle.run_chain(J, h, masks, w, s0, H, 1.0, (0.25, 0.6), 0.8, seed, chain_id=5 + c)
        assert np.array_equal(b.proposals[c], props), (c, b.proposals[c], props)
        assert np.array_equal(b.accepted[c], acc)
        assert int(b.acc_count[c]) == int(acc.sum())
        cur, hist = s0, np.zeros(n + 1, dtype=np.int64)
        for p_, a_ in zip(props, acc):
            hist[bin(cur ^ int(p_)).count("1")] += 1
            if a_:
                cur = int(p_)
        assert int(b.final_state[c]) == cur and np.array_equal(b.hamming_hist[c], hist)


def test_sweep128_full_size_properties_n20():
    """n = 20 in complex128 (16 MiB per chain): norm conservation through 13 layers to 1e-12, symmetry
    Q(s -> s') = Q(s' -> s), chain-sharding invariance, incoherent sums of one-body mixers."""
    q, J, h, m, mx, masks, w = _setup(20, seed=1)
    s, sp = 123456, 654321
    p1 = q.transition_probabilities(m, mx, s, 0.4, 13, dtype="fp64")
    p2 = q.transition_probabilities(m, mx, sp, 0.4, 13, dtype="fp64")
    assert abs(p1.sum() - 1.0) < 1e-12
    assert abs(p1[sp] - p2[s]) < 1e-14 + 1e-9 * p1[sp]
    full = q.quantum_enhanced_mcmc(3, m, mx, (0.25, 0.6), temperature=1.0, n_chains=8, seed=5, dtype="fp64")
    half = q.quantum_enhanced_mcmc(3, m, mx, (0.25, 0.6), temperature=1.0, n_chains=4, seed=5, dtype="fp64", chain_id0=4)
    assert np.array_equal(full.proposals[4:], half.proposals) and np.array_equal(full.accepted[4:], half.accepted)
    inc = q.IncoherentMixerSum([q.GenericMixer.OneBodyMixer(20), q.CustomMixer(20, [[b] for b in range(0, 20, 2)])], [0.5, 0.5])
    b = q.quantum_enhanced_mcmc(2, m, inc, (0.25, 0.6), temperature=1.0, n_chains=6, seed=2, dtype="fp64")
    assert b.proposals.shape == (6, 2)


@pytest.mark.parametrize("n,C,mib,streams", [(13, 150, 1, 3), (16, 48, 3, 2), (20, 9, 16, 4), (20, 7, 40, 1)])
def test_sweep128_l2_resident_group_schedule_is_bit_identical(monkeypatch, n, C, mib, streams):
    """complex128 chains under the L2-resident group schedule (group_sched.cuh) equal the step-major schedule bit for
    bit: same kernels on the same data, only the launch order and the streams differ."""
    q, J, h, m, mx, masks, w = _setup(n, seed=8)
    kw = dict(temperature=1.0, n_chains=C, seed=13, dtype="fp64", chain_id0=1)
    monkeypatch.setenv("QUMCMC_L2_GROUP_MIB", "0")
    ref = q.quantum_enhanced_mcmc(3, m, mx, (0.25, 0.6), **kw)
    monkeypatch.setenv("QUMCMC_L2_GROUP_MIB", str(mib))
    monkeypatch.setenv("QUMCMC_L2_STREAMS", str(streams))
    got = q.quantum_enhanced_mcmc(3, m, mx, (0.25, 0.6), **kw)
    for f in ("proposals", "accepted", "final_state", "acc_count", "hamming_hist"):
        assert np.array_equal(getattr(ref, f), getattr(got, f)), f
