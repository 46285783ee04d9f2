"""classical_mcmc on the device: oracle properties on CPU, bit-exact parity on the GPU."""
import numpy as np
import pytest

import oracle
from tests import models


def _spec():
    # CombineProposals([Uniform, FixedWeight(2), FixedWeight(5), Custom]) flattened: (kinds, params, probs)
    return [0, 1, 1, 2], [0, 2, 5, 0b1000000101], [0.2, 0.4, 0.3, 0.1]


def test_oracle_classical_rules_flip_what_the_reference_flips():
    """FixedWeightProposals xor-s exactly k ones (random_bstr, basic_utils.py:383-388), CustomProposals a fixed
    pattern (classical_mixers.py:72-78), UniformProposals stays inside [0, 2^n) (:38-40)."""
    n = 10
    J, h = models.packaged_random_model(n, 3)
    for kind, param, weight in [(1, 1, 1), (1, 3, 3), (1, 10, 10), (2, 0b0100100001, 3)]:
        props, acc = oracle.classical_chain(J, h, [kind], [param], None, 5, 400, 1.0, seed=9)
        cur = 5
        for p_, a_ in zip(props, acc):
            assert bin(int(p_) ^ cur).count("1") == weight
            if kind == 2:
                assert int(p_) ^ cur == param
            if a_:
                cur = int(p_)
    props, _ = oracle.classical_chain(J, h, [0], [0], None, 5, 4000, 1.0, seed=9)
    assert props.max() < (1 << n) and len(np.unique(props)) > 900          # ~ uniform over 1024 states


def test_oracle_fixed_weight_positions_are_uniform():
    n, k, H = 8, 3, 56000
    J, h = np.zeros((n, n)), np.zeros(n)                                   # flat energy: every proposal accepted
    props, acc = oracle.classical_chain(J, h, [1], [k], None, 0, H, 1.0, seed=4)
    assert acc.all()
    flips = np.concatenate([np.zeros(1, dtype=np.uint64), props[:-1]]) ^ props
    counts = np.bincount(flips.astype(np.int64), minlength=1 << n)
    used = counts[counts > 0]
    assert len(used) == 56                                                 # C(8,3) subsets, all reached
    assert abs(used / H - 1 / 56).max() < 0.25 / 56


def test_oracle_classical_chain_samples_the_boltzmann_distribution():
    """Uniform proposals + Metropolis at n=5: visit frequencies of a long chain match exp(-E/T)/Z."""
    n = 5
    J, h = models.packaged_random_model(n, 1)
    H = 200000
    props, acc = oracle.classical_chain(J, h, [0, 1], [0, 1], [0.5, 0.5], 3, H, 1.0, seed=21)
    cur, visits = 3, np.zeros(1 << n)
    for p_, a_ in zip(props, acc):
        if a_:
            cur = int(p_)
        visits[cur] += 1
    _, _, P = oracle.exact_boltzmann(J, h, 1.0)
    assert np.abs(visits / H - P).max() < 0.01


def test_lowering_of_the_reference_proposal_classes():
    import qumcmc_b200 as q
    from qumcmc_b200.classical_mcmc_routines import lower_proposals
    from qumcmc_b200.classical_mixers import (ClassicalMixer, CombineProposals, CustomProposals, FixedWeightProposals,
                                              UniformProposals)
    n = 6
    mix = CombineProposals([UniformProposals(n), CombineProposals([FixedWeightProposals(n, 2), CustomProposals(n, [0, 5])],
                                                                  [1, 3])], [0.5, 0.5])
    kinds, params, probs = lower_proposals(mix)
    assert kinds.tolist() == [0, 1, 2] and params.tolist() == [0, 2, 0b100001]
    assert np.allclose(probs, [0.5, 0.125, 0.375])

    class Mine(ClassicalMixer):
        def propose_transition(self, current_state):
            return current_state[::-1]
    assert lower_proposals(Mine(n)) is None and lower_proposals(CombineProposals([Mine(n)], [1.0])) is None


@pytest.mark.gpu
@pytest.mark.parametrize("n", [9, 20, 36])
def test_device_classical_chains_equal_the_oracle(n):
    import qumcmc_b200 as q
    from qumcmc_b200.classical_mcmc_routines import run_classical_chains
    from qumcmc_b200.classical_mixers import CombineProposals, CustomProposals, FixedWeightProposals, UniformProposals
    J, h = models.packaged_random_model(n, 5)
    m = q.IsingEnergyFunction(J, h)
    pattern = [0, 3, n - 1]
    mix = CombineProposals([UniformProposals(n), FixedWeightProposals(n, 2), FixedWeightProposals(n, min(n, 7)),
                            CustomProposals(n, pattern)], [0.2, 0.4, 0.3, 0.1])
    C, H, seed = 40, 60, 77
    b = run_classical_chains(m, mix, H, C, None, 0.9, seed, chain_id0=3)
    kinds, params, probs = [0, 1, 1, 2], [0, 2, min(n, 7), sum(1 << (n - 1 - i) for i in pattern)], [0.2, 0.4, 0.3, 0.1]
    for c in range(C):
        s0 = oracle.initial_state(seed, 3 + c, n)
        assert int(b.initial[c]) == s0
        props, acc = oracle.classical_chain(J, h, kinds, params, probs, s0, H, 0.9, seed, chain_id=3 + c)
        assert np.array_equal(b.proposals[c], props) and np.array_equal(b.accepted[c], acc)
        assert int(b.acc_count[c]) == int(acc.sum())
        cur, hist = s0, np.zeros(n + 1, dtype=np.int64)
        for p_, a_ in zip(props, acc):
            hist[bin(int(p_) ^ cur).count("1")] += 1
            if a_:
                cur = int(p_)
        assert int(b.final_state[c]) == cur and np.array_equal(b.hamming_hist[c], hist)


@pytest.mark.gpu
def test_classical_mcmc_api_shapes_and_resume():
    import qumcmc_b200 as q
    from qumcmc_b200.classical_mixers import ClassicalMixer, FixedWeightProposals
    J, h = models.packaged_random_model(8, 2)
    m = q.IsingEnergyFunction(J, h)
    np.random.seed(3)
    ch = q.classical_mcmc(n_hops=200, model=m, temperature=1.0)                       # README call shape
    assert isinstance(ch, q.MCMCChain) and len(ch.states) == 201
    whole = q.classical_mcmc(100, m, FixedWeightProposals(8, 1), temperature=1.0, n_chains=6, seed=5)
    first = q.classical_mcmc(40, m, FixedWeightProposals(8, 1), temperature=1.0, n_chains=6, seed=5)
    rest = q.classical_mcmc(60, m, FixedWeightProposals(8, 1), initial_state=[int(s) for s in first.final_state],
                            temperature=1.0, n_chains=6, seed=5, hop0=40)
    assert np.array_equal(whole.proposals[:, :40], first.proposals) and np.array_equal(whole.proposals[:, 40:], rest.proposals)
    part = q.classical_mcmc(100, m, FixedWeightProposals(8, 1), temperature=1.0, n_chains=2, seed=5, chain_id0=4)
    assert np.array_equal(part.proposals, whole.proposals[4:6])                       # independent of the sharding

    class Reverse(ClassicalMixer):                                                    # host-only rule: reference loop
        def propose_transition(self, current_state):
            return current_state[::-1]
    ch = q.classical_mcmc(20, m, Reverse(8), initial_state="11010000", temperature=1.0)
    assert all(s.bitstring in ("11010000", "00001011") for s in ch.states)
    with pytest.raises(q._lib.QmcUnsupported):
        q.classical_mcmc(20, m, Reverse(8), temperature=1.0, n_chains=4)


@pytest.mark.gpu
def test_classical_and_trajectory_edge_cases():
    """Zero chains / zero hops are no-ops; malformed proposal specs and missing inputs fail loudly (no fallback)."""
    import torch
    import qumcmc_b200 as q
    from qumcmc_b200 import _lib
    from qumcmc_b200.classical_mcmc_routines import run_classical_chains
    from qumcmc_b200.classical_mixers import CustomProposals, FixedWeightProposals, UniformProposals
    J, h = models.packaged_random_model(7, 2)
    m = q.IsingEnergyFunction(J, h)
    b = run_classical_chains(m, UniformProposals(7), 0, 3, np.array([1, 2, 3], dtype=np.uint64), 1.0, 5)
    assert b.proposals.shape == (3, 0) and b.final_state.tolist() == [1, 2, 3] and b.acc_count.tolist() == [0, 0, 0]
    dm = m.device_model(0)
    L = _lib.lib()
    fin = torch.zeros(4, dtype=torch.int64, device="cuda")
    kinds = np.array([7], dtype=np.int32)
    params = np.array([0], dtype=np.uint64)
    with pytest.raises((ValueError, _lib.QmcError)):                                        # unknown proposal kind
        _lib.check(L.qmc_classical_chains(dm.handle, 4, 2, None, 0, 0, 1.0, 1, kinds.ctypes.data, params.ctypes.data, None,
                                          1, None, None, fin.data_ptr(), None, None, None))
    kinds[0], params[0] = _lib.CLASSICAL_FIXED_WEIGHT, 9
    with pytest.raises((ValueError, _lib.QmcError)):                                        # bodyness > num_spins
        _lib.check(L.qmc_classical_chains(dm.handle, 4, 2, None, 0, 0, 1.0, 1, kinds.ctypes.data, params.ctypes.data, None,
                                          1, None, None, fin.data_ptr(), None, None, None))
    kinds[0], params[0] = _lib.CLASSICAL_CUSTOM, 1 << 9
    with pytest.raises((ValueError, _lib.QmcError)):                                        # flip mask outside the model
        _lib.check(L.qmc_classical_chains(dm.handle, 4, 2, None, 0, 0, 1.0, 1, kinds.ctypes.data, params.ctypes.data, None,
                                          1, None, None, fin.data_ptr(), None, None, None))
    with pytest.raises((ValueError, _lib.QmcError)):                                        # non-positive temperature
        run_classical_chains(m, FixedWeightProposals(7, 1), 5, 2, None, 0.0, 1)
    with pytest.raises(ValueError):                                           # running KL without the target distribution
        q.batch_trajectory_statistics(m, np.zeros(1, np.uint64), np.zeros((1, 3), np.uint64), np.zeros((1, 3), np.uint8),
                                      None, to_observe=("kldiv",))
    res = q.batch_trajectory_statistics(m, np.array([5], np.uint64), np.zeros((1, 0), np.uint64), np.zeros((1, 0), np.uint8),
                                        np.full(128, 1 / 128), to_observe=("markov_chain", "kldiv"))
    assert res["markov_chain"].tolist() == [[5]] and res["kldiv"].shape == (1, 1)
    assert abs(res["kldiv"][0, 0] - (np.log(1 / 128) - (1 / 128) * np.log(1.0) - (127 / 128) * np.log(1e-8))) < 1e-9
    ch = q.classical_mcmc(30, m, CustomProposals(7, [0, 6]), initial_state="0000000", temperature=1e9, seed=1)
    assert set(ch.markov_chain) <= {"0000000", "1000001"}                     # one fixed flip pattern, always accepted
