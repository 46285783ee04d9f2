"""GPU parity tests: the CUDA path (through the C-ABI) vs the CPU oracle on the same seeded inputs,
vs the golden fixtures, and through size-independent properties at larger n."""
import numpy as np
import pytest

import oracle
from oracle import np_oracle as npo
from tests import models

pytestmark = pytest.mark.gpu

CS = models.golden("chain_stats.json")
KA = models.golden("known_answers.json")


def _api():
    import qumcmc_b200 as q
    return q


def _model(J, h, **kw):
    return _api().IsingEnergyFunction(np.array(J, dtype=float), np.array(h, dtype=float), **kw)


# ------------------------------------------------------------------------------- energies (a2)
@pytest.mark.parametrize("n,seed", [(1, 1), (5, 23564), (9, 3), (10, 4), (16, 5), (20, 6), (28, 7), (34, 8)])
def test_energies_bit_exact(n, seed):
    J, h = models.packaged_random_model(n, seed)
    m = _model(J, h)
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 1 << n, size=4096, dtype=np.uint64)
    idx[:2] = [0, (1 << n) - 1]
    got = m.get_energies(idx)
    exp = oracle.energies(J, h, idx)
    assert np.array_equal(got, exp)  # bit-exact (fixed fp64 summation order, no FMA contraction)


def test_energy_api_forms_and_golden():
    J, h, data = models.bas3("both")
    m = _model(J, h)
    assert m.get_energy("111000111") == KA["bas3_bars_stripes"]["energy"]
    s = np.array([1, 1, 1, -1, -1, -1, 1, 1, 1])
    assert m.get_energy(s) == -48.0
    assert m.alpha == KA["bas3_bars_stripes"]["alpha"]
    assert m.get_energies([]).shape == (0,)
    with pytest.raises(ValueError):
        m.get_energy("1010")


# ------------------------------------------------------------------------------- exact sampler (a13, a14)
@pytest.mark.parametrize("which,beta,key", [("both", 1.0, "entropy_beta_1.0"), ("both", 1.5, "entropy_beta_1.5")])
def test_exact_sampling_bas_golden(which, beta, key):
    q = _api()
    J, h, _ = models.bas3(which)
    ex = q.Exact_Sampling(_model(J, h), beta)
    ka = KA["bas3_bars_stripes"]
    assert ex.get_entropy() == pytest.approx(ka[key], abs=1e-12)
    logZ, E, P = oracle.exact_boltzmann(J, h, beta)
    assert np.array_equal(ex.energies, E)
    assert ex.logZ == pytest.approx(logZ, rel=1e-13)
    assert np.abs(ex.probs - P).max() < 1e-13
    pd = ex.boltzmann_pd
    assert len(pd) == 512 and list(pd.values()) == sorted(pd.values(), reverse=True)
    assert sum(1 for v in pd.values() if v > 0.01) == ka["n_states_gt_0.01_beta1"]


def test_exact_sampling_random_models_and_quirks():
    q = _api()
    kr = KA["random_ising_model_n5"]
    for seed, cnt in zip(kr["seeds"], kr["n_states_gt_0.01"]):
        ex = q.Exact_Sampling(q.random_ising_model(5, seed), kr["beta"])
        assert int((ex.probs > 0.01).sum()) == cnt
        assert ex.get_entropy() is None  # quirk Q7
    J, h = models.recipe_b(10, 4)
    ex = q.Exact_Sampling(_model(J, h), 1.0)
    assert ex.get_entropy() == pytest.approx(KA["recipe_b_n10_seed4"]["entropy_beta_1.0"], abs=1e-12)
    unif = {k: 1.0 / 1024 for k in ex.boltzmann_pd}
    kl = ex.get_kldiv(unif)
    exp = float(np.sum(ex.probs * np.log2(ex.probs * 1024)))
    assert kl == pytest.approx(exp, rel=1e-10)


@pytest.mark.parametrize("n", [16, 20, 24])
def test_exact_sampling_large_properties(n):
    """Full-size properties: sum p = 1, logZ vs the oracle (n <= 20), overflow robustness (Q13)."""
    q = _api()
    J, h = models.packaged_random_model(n, 11)
    ex = q.Exact_Sampling(_model(J, h), 1.0)
    assert abs(ex.probs.sum() - 1.0) < 1e-10
    if n <= 20:
        logZ, E, _ = oracle.exact_boltzmann(J, h, 1.0, want_probs=False)
        assert np.array_equal(ex.energies, E)
        assert ex.logZ == pytest.approx(logZ, rel=1e-13)
    big = q.Exact_Sampling(_model(J, h), 400.0)  # exp(-beta E) overflows in the reference
    assert np.isfinite(big.logZ) and abs(big.probs.sum() - 1.0) < 1e-10


# ------------------------------------------------------------------------------- proposal probabilities (a4-a8)
def _mixers(n):
    q = _api()
    out = {"one": q.GenericMixer.OneBodyMixer(n)}
    if n >= 3:
        out["two"] = q.GenericMixer.TwoBodyMixer(n)
    if n == 9:
        bars = q.CustomMixer(9, [[0, 1, 2], [3, 4, 5], [6, 7, 8]])
        out["custom_bars"] = bars
        out["coherent"] = q.CoherentMixerSum([out["one"], bars], weights=[0.75, 0.25])
    return out


@pytest.mark.parametrize("n", [1, 2, 5, 9, 10, 12])
@pytest.mark.parametrize("dtype,tol", [("fp64", 1e-10), ("fp32", 1e-6)])
def test_transition_probabilities_small_n(n, dtype, tol):
    q = _api()
    J, h = models.recipe_b(n, 2) if n > 1 else (np.zeros((1, 1)), np.array([0.3]))
    m = _model(J, h)
    al = oracle.alpha(J, h)
    rng = np.random.default_rng(n)
    for name, mx in _mixers(n).items():
        offs, masks, w, _ = mx.lower()
        for r in (1, 2, 7, 13):
            s = int(rng.integers(0, 1 << n))
            gamma = float(rng.uniform(0.05, 0.95))
            p, a = q.transition_probabilities(m, mx, s, gamma, r, dtype=dtype, return_amplitudes=True)
            ref = oracle.evolve(J, h, al, gamma, 0.8, r, masks.tolist(), w.tolist(), s, gatewise=True)
            assert np.abs(p - np.abs(ref) ** 2).max() < tol, (name, r)
            assert np.abs(a - ref).max() < (1e-9 if dtype == "fp64" else 2e-5)


def test_phase_layer_skip_rule_and_decoy_hamiltonian():
    """mean|h|,mean|J| <= 1e-3 drops the problem layer (quantum_mcmc_routines.py:64-67);
    circuit_J replaces J in the circuit only (energy_models.py:35)."""
    q = _api()
    n = 6
    J = np.full((n, n), 5e-4); np.fill_diagonal(J, 0.0)
    h = np.full(n, 5e-4)
    mx = q.GenericMixer.OneBodyMixer(n)
    masks, w = models.one_body(n)
    p = q.transition_probabilities(_model(J, h), mx, 5, 0.3, 6, dtype="fp64")
    ref = oracle.transition_probs(J, h, oracle.alpha(J, h), 0.3, 0.8, 6, masks, w, 5)
    assert np.abs(p - ref).max() < 1e-12
    J2, h2 = models.recipe_b(n, 9)
    Jd, _ = models.recipe_b(n, 10)
    m = _model(J2, h2, circuit_J=Jd)
    p = q.transition_probabilities(m, mx, 3, 0.4, 5, dtype="fp64")
    ref = oracle.transition_probs(Jd, h2, oracle.alpha(J2, h2), 0.4, 0.8, 5, masks, w, 3)
    assert np.abs(p - ref).max() < 1e-10


# ------------------------------------------------------------------------------- measurement (a9) + explicit draws
def test_propose_matches_oracle_inverse_cdf():
    import torch
    from qumcmc_b200 import _device, _lib
    q = _api()
    n = 9
    J, h, _ = models.bas3("both")
    m = _model(J, h)
    mx = q.GenericMixer.OneBodyMixer(n)
    masks, w = models.one_body(n)
    dm = m.device_model(0)
    dm.set_mixer(mx)
    rng = np.random.default_rng(5)
    C = 256
    s_in = rng.integers(0, 512, size=C).astype(np.int64)
    gam = rng.uniform(0.25, 0.6, size=C)
    rr = rng.choice([2, 3, 5, 6, 7, 8, 10, 11, 12, 13], size=C).astype(np.int32)
    uu = rng.uniform(0, 1, size=C)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(a).to(dev)
    d_s, d_g, d_r, d_u = t(s_in), t(gam), t(rr), t(uu)
    out = torch.empty(C, dtype=torch.int64, device=dev)
    for dtype in (_lib.DTYPE_F64, _lib.DTYPE_F32):
        _lib.check(_lib.lib().qmc_propose(dm.handle, C, d_s.data_ptr(), d_g.data_ptr(), d_r.data_ptr(), d_u.data_ptr(),
                                          None, 0.8, dtype, out.data_ptr(), None))
        got = out.cpu().numpy()
        al = oracle.alpha(J, h)
        bad = 0
        for c in range(C):
            st = oracle.evolve(J, h, al, gam[c], 0.8, int(rr[c]), masks, w, int(s_in[c]), gatewise=False)
            exp = oracle.sample_index(st, uu[c])
            if exp != got[c]:
                cdf = np.cumsum(np.abs(st) ** 2)
                margin = np.abs(cdf - uu[c]).min()
                assert margin < (1e-12 if dtype == _lib.DTYPE_F64 else 2e-6), (c, exp, got[c], margin)
                bad += 1
        assert bad <= (0 if dtype == _lib.DTYPE_F64 else 2)


# ------------------------------------------------------------------------------- chains (a10-a12): bit-exact vs oracle
@pytest.mark.parametrize("case", ["bas_one", "bas_coherent", "bas_incoherent", "readme10"])
def test_chain_bit_exact_vs_oracle_fp64(case):
    """Same Philox streams on both sides: proposals and accept decisions must be identical."""
    q = _api()
    if case == "readme10":
        J, h = models.readme_model()
        n, beta, gamma = 10, 1.100209, (0.25, 0.6)
        mx = q.GenericMixer.OneBodyMixer(n)
    else:
        J, h, _ = models.bas3("both")
        n, beta, gamma = 9, 1.5, 0.5
        one = q.GenericMixer.OneBodyMixer(9)
        bars = q.CustomMixer(9, [[0, 1, 2], [3, 4, 5], [6, 7, 8]])
        mx = {"bas_one": one, "bas_coherent": q.CoherentMixerSum([one, bars], [0.75, 0.25]),
              "bas_incoherent": q.IncoherentMixerSum([one, bars], [0.75, 0.25])}[case]
    offs, masks, w, probs = mx.lower()
    m = _model(J, h)
    H, C, seed = 300, 4, 0x5EED
    batch = q.quantum_enhanced_mcmc(H, m, mx, gamma, temperature=1 / beta, n_chains=C, seed=seed, dtype="fp64",
                                    chain_id0=10)
    assert batch.proposals.shape == (C, H)
    for c in range(C):
        s0 = oracle.initial_state(seed, 10 + c, n)
        assert int(batch.initial[c]) == s0
        props, acc = oracle.run_chain(J, h, masks, w, s0, H, 1 / beta, gamma, 0.8, seed, chain_id=10 + c,
                                      group_off=offs, group_prob=probs)
        assert np.array_equal(batch.proposals[c], props)
        assert np.array_equal(batch.accepted[c], acc)
        assert int(batch.acc_count[c]) == int(acc.sum())
        mc = npo.markov_chain_from(s0, props, acc)
        assert int(batch.final_state[c]) == int(mc[-1])
        ham = np.bincount([bin(int(a) ^ int(b)).count("1") for a, b in zip(mc[:-1], props)], minlength=n + 1)
        assert np.array_equal(batch.hamming_hist[c], ham)
        ch = batch[c]
        assert ch.markov_chain == [f"{int(k):0{n}b}" for k in mc]


def test_chain_resume_and_sharding_invariance():
    """chain_id0/hop0 make results independent of how chains and hops are split across calls/GPUs."""
    q = _api()
    J, h, _ = models.bas3("both")
    m = _model(J, h)
    mx = q.GenericMixer.OneBodyMixer(9)
    full = q.quantum_enhanced_mcmc(64, m, mx, (0.25, 0.6), temperature=1 / 1.5, n_chains=8, seed=77)
    lo = q.quantum_enhanced_mcmc(64, m, mx, (0.25, 0.6), temperature=1 / 1.5, n_chains=4, seed=77, chain_id0=0)
    hi = q.quantum_enhanced_mcmc(64, m, mx, (0.25, 0.6), temperature=1 / 1.5, n_chains=4, seed=77, chain_id0=4)
    assert np.array_equal(full.proposals, np.concatenate([lo.proposals, hi.proposals]))
    first = q.quantum_enhanced_mcmc(40, m, mx, (0.25, 0.6), temperature=1 / 1.5, n_chains=8, seed=77)
    states = [f"{int(s):09b}" for s in first.final_state]
    rest = q.quantum_enhanced_mcmc(24, m, mx, (0.25, 0.6), initial_state=states, temperature=1 / 1.5, n_chains=8,
                                   seed=77, hop0=40)
    assert np.array_equal(full.proposals[:, 40:], rest.proposals)
    assert np.array_equal(full.accepted[:, 40:], rest.accepted)


# ------------------------------------------------------------------------------- chain statistics vs recorded chains
@pytest.mark.parametrize("dtype", ["fp64", "fp32"])
def test_chain_statistics_vs_reference_fixture(dtype):
    """Config C2 shape: BAS 3x3, beta 1.5, init 111000111, gamma 0.5, many chains. Acceptance and the
    proposal-Hamming distribution vs the reference's recorded chains (5 x 10 000 hops)."""
    q = _api()
    J, h, _ = models.bas3("both")
    m = _model(J, h)
    mx = q.GenericMixer.OneBodyMixer(9)
    C, H = 1024, 2000
    b = q.quantum_enhanced_mcmc(H, m, mx, 0.5, initial_state="111000111", temperature=1 / 1.5, n_chains=C,
                                seed=1234, dtype=dtype)
    fx = CS["bas"]["BAS_new_code_final_plots_wt1.pkl"]
    acc = b.acceptance_rate()
    # fixture: mean of 5 chains, std 0.0041 -> sem 0.0018; ours: 2e6 proposals
    assert abs(acc - fx["acceptance_mean"]) < 3 * 0.0041 / np.sqrt(5) + 0.003
    ham = b.hamming_hist.sum(axis=0) / b.hamming_hist.sum()
    assert np.abs(ham - np.array(fx["hamming_mean"])).max() < 0.006
    # KL(Boltzmann || visits), the reference's running-KL end point (ln, EPS floor)
    ex = q.Exact_Sampling(m, 1.5)
    visits = np.zeros(512)
    for c in range(0, C, 64):
        visits += b[c].visit_histogram()
    kl = npo.vectorised_kl(ex.probs, visits / visits.sum())
    assert kl < 0.02


def test_single_chain_api_and_readme_call_shape():
    q = _api()
    J, h = models.readme_model()
    m = _model(J, h, name="my_model")
    np.random.seed(3)
    ch = q.quantum_enhanced_mcmc(n_hops=500, model=m, temperature=1 / 1.100209)  # README call shape
    assert isinstance(ch, q.MCMCChain) and len(ch.states) == 501 and ch.states[0].accepted
    assert len(ch.markov_chain) == 501 and sum(ch.get_accepted_dict().values()) == 501
    assert "GenericMixer" in ch.name
    cl = q.classical_mcmc(n_hops=300, model=m, temperature=1 / 1.100209)
    assert len(cl.states) == 301
    s = q.run_qmcmc_quantum_ckt("0101010101", m, q.GenericMixer.OneBodyMixer(10), lambda: 0.4, m.alpha, 10, 0.8)
    assert len(s) == 10 and set(s) <= {"0", "1"}
    sampler = q.QuantumMCMCSampler(50, m, q.GenericMixer.OneBodyMixer(10), 0.5, initial_state="0000011111")
    assert len(sampler.run().states) == 51


def test_reference_mixer_matrix_runs():
    """testing.py:50-83: every quantum mixer of the reference's smoke script, 10 hops each."""
    q = _api()
    J, h, _ = models.bas3("both")
    m = _model(J, h)
    n = 9
    generic = q.GenericMixer.OneBodyMixer(n)
    stripes = q.CustomMixer(n, [[0, 3, 6], [1, 4, 7], [2, 5, 8]])
    bars = q.CustomMixer(n, [[0, 1, 2], [3, 4, 5], [6, 7, 8]])
    bas = q.CustomMixer(n, [[0, 1, 2], [3, 4, 5], [6, 7, 8], [0, 3, 6], [1, 4, 7], [2, 5, 8]])
    mixers = [generic, bars, bas, q.CoherentMixerSum([generic, bars], weights=[0.75, 0.25]),
              q.CoherentMixerSum([generic, bars, stripes], weights=[0.7, 0.15, 0.15]),
              q.IncoherentMixerSum([generic, bars], probabilities=[0.75, 0.25]),
              q.IncoherentMixerSum([generic, bars, stripes], probabilities=[0.7, 0.15, 0.15]),
              q.GenericMixer.ThreeBodyMixer(n)]
    for mx in mixers:
        ch = q.QuantumMCMCSampler(10, m, mx, 0.5, initial_state="111000111").run()
        assert len(ch.states) == 11


def test_exact_partial_ranges_combine_to_full_logz():
    """Sharded enumeration: ranges of the index space combine to the single-GPU log Z."""
    import torch
    from qumcmc_b200 import _lib
    from qumcmc_b200.distributed import combine_logsumexp, shard_range
    q = _api()
    n, beta = 18, 0.9
    J, h = models.packaged_random_model(n, 21)
    m = _model(J, h)
    ex = q.Exact_Sampling(m, beta)
    dm = m.device_model(0)
    dev = torch.device("cuda", 0)
    pairs = []
    for r in range(5):
        a, b = shard_range(1 << n, r, 5)
        pair = torch.empty(2, dtype=torch.float64, device=dev)
        en = torch.empty(b - a, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().qmc_exact_partial(dm.handle, beta, a, b - a, pair.data_ptr(), en.data_ptr(), None))
        pr = torch.empty(b - a, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().qmc_exact_probs_range(dm.handle, beta, ex.logZ, a, b - a, en.data_ptr(), pr.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(en.cpu().numpy(), ex.energies[a:b])
        assert np.abs(pr.cpu().numpy() - ex.probs[a:b]).max() < 1e-15
        pairs.append(pair.cpu().numpy())
    assert combine_logsumexp(np.stack(pairs)) == pytest.approx(ex.logZ, rel=1e-14)


@pytest.mark.parametrize("n", [16, 19])
def test_exact_fast_enumerator_matches_ordered_path(n):
    """log Z only / probabilities only take the hi/lo-split enumerator (O(1) energy per state); it must agree with
    the ordered evaluation (bit-exact energies) to fp64 rounding, also for a non-symmetric J with a diagonal."""
    import torch
    from qumcmc_b200 import _lib
    from qumcmc_b200.distributed import combine_logsumexp
    rng = np.random.default_rng(n)
    J = rng.uniform(-1, 1, size=(n, n))            # neither symmetric nor zero-diagonal: get_energy takes any J
    h = rng.normal(0, 0.5, size=n)
    m = _model(J, h)
    beta = 1.3
    dm = m.device_model(0)
    dev = torch.device("cuda", 0)
    N = 1 << n
    lz_o = torch.zeros(1, dtype=torch.float64, device=dev)
    en = torch.zeros(N, dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().qmc_exact_sampling(dm.handle, beta, lz_o.data_ptr(), en.data_ptr(), None, None))
    lz_f = torch.zeros(1, dtype=torch.float64, device=dev)
    pr = torch.zeros(N, dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().qmc_exact_sampling(dm.handle, beta, lz_f.data_ptr(), None, pr.data_ptr(), None))
    torch.cuda.synchronize()
    E = en.cpu().numpy()
    assert np.array_equal(E[:64], np.array([oracle.energy(J, h, i) for i in range(64)]))
    assert abs(lz_f.item() - lz_o.item()) < 1e-12
    ref_p = np.exp(-beta * E - lz_o.item())
    assert np.abs(pr.cpu().numpy() - ref_p).max() < 1e-13 * max(1.0, ref_p.max())
    # fast-eligible ranges (multiples of 2^16) through the sharded entry points
    pairs = []
    for a in range(0, N, N // 2):
        pair = torch.empty(2, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().qmc_exact_partial(dm.handle, beta, a, N // 2, pair.data_ptr(), None, None))
        p2 = torch.empty(N // 2, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().qmc_exact_probs_range(dm.handle, beta, lz_o.item(), a, N // 2, None, p2.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.abs(p2.cpu().numpy() - ref_p[a:a + N // 2]).max() < 1e-13 * max(1.0, ref_p.max())
        pairs.append(pair.cpu().numpy())
    assert combine_logsumexp(np.stack(pairs)) == pytest.approx(lz_o.item(), abs=1e-12)
