"""CPU-only tests: host-side mirror of the reference API, the C-ABI surface, the chain-sharding
logic (gloo, world_size 2) and the no-CPU-fallback guarantee."""
import os
import pickle
import re
import sys

import numpy as np
import pytest

import oracle
from tests import models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from qumcmc_b200 import _lib
    header = open(os.path.join(ROOT, "include", "qumcmc_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(qmc_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)          # loads without a GPU
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().qmc_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import qumcmc_b200 as q
    from qumcmc_b200._lib import QmcError
    m = q.random_ising_model(5, 1)
    with pytest.raises(QmcError):
        m.get_energy("01010")
    with pytest.raises(QmcError):
        q.quantum_enhanced_mcmc(10, m, q.GenericMixer.OneBodyMixer(5), 0.5)
    with pytest.raises(QmcError):
        q.Exact_Sampling(m, 1.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "qumcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "qmc_oracle" not in src, f


def test_model_construction_alpha_and_quirks():
    import qumcmc_b200 as q
    J, h, _ = models.bas3("both")
    m = q.IsingEnergyFunction(J, h, name="bas")
    assert m.alpha == 0.125 and m.num_spins == 9 and m.circuit_J is J and m.circuit_h is h
    assert q.IsingEnergyFunction(np.zeros((3, 3)), np.ones(3)).alpha == 1.0      # np.allclose(J, 0) branch
    Jr, hr = models.packaged_random_model(7, 3)
    rm = q.random_ising_model(7, 3)
    assert np.array_equal(rm.J, Jr) and np.array_equal(rm.h, hr)
    assert rm.alpha == pytest.approx(oracle.alpha(Jr, hr), rel=1e-15)
    # quirk Q4: circuit_h given => circuit_J is installed in its place (energy_models.py:36)
    dec = q.IsingEnergyFunction(Jr, hr, circuit_J=2 * Jr, circuit_h=3 * hr)
    assert dec.circuit_h is dec.circuit_J
    with pytest.raises(ValueError):
        dec._circuit_arrays()


def test_mixer_lowering_matches_reference_semantics():
    import qumcmc_b200 as q
    one = q.GenericMixer.OneBodyMixer(4)
    offs, masks, w, probs = one.lower()
    assert offs.tolist() == [0, 4] and masks.tolist() == [1, 2, 4, 8] and probs is None   # qubit index = bit (Q2)
    assert one.possible_qubit_combinations == [[0], [1], [2], [3]]
    two = q.GenericMixer.TwoBodyMixer(4)
    assert len(two.possible_qubit_combinations) == 6 and two.lower()[1].tolist() == [3, 5, 9, 6, 10, 12]
    assert len(q.GenericMixer.ThreeBodyMixer(9).lower()[1]) == 84
    bars = q.CustomMixer(9, [[0, 1, 2], [3, 4, 5], [6, 7, 8]])
    assert bars.lower()[1].tolist() == [7, 56, 448]
    coh = q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(9), bars], weights=[3, 1])
    offs, masks, w, probs = coh.lower()
    assert offs.tolist() == [0, 12] and probs is None
    assert w.tolist() == [0.75] * 9 + [0.25] * 3                                         # gamma * w_m (mixers.py:126-129)
    inc = q.IncoherentMixerSum([q.GenericMixer.OneBodyMixer(9), bars], probabilities=[3, 1])
    offs, masks, w, probs = inc.lower()
    assert offs.tolist() == [0, 9, 12] and probs.tolist() == [0.75, 0.25] and set(w.tolist()) == {1.0}
    with pytest.raises(NotImplementedError):
        one.get_mixer_circuit(0.5, 0.8)
    with pytest.raises(ValueError):
        q.CustomMixer(3, [[0, 3]]).lower()
    with pytest.raises(AssertionError):
        q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(3), q.GenericMixer.OneBodyMixer(4)], [1, 1])


def test_mcmc_chain_semantics_and_pickle():
    import qumcmc_b200 as q
    S = q.MCMCState
    ch = q.MCMCChain([S("000", True)], name="t")
    for bs, acc in [("001", False), ("011", True), ("111", False), ("100", True), ("101", False)]:
        ch.add_state(S(bs, acc))
    assert ch.markov_chain == ["000", "000", "011", "011", "100", "100"]          # basic_utils.py:84-115
    assert ch.current_state == S("100", True)
    assert ch.accepted_states == ["000", "011", "100"]
    assert ch.get_accepted_dict() == {"000": 2, "011": 2, "100": 2}
    assert ch.get_accepted_dict(normalize=True, until_index=4) == {"000": 0.5, "011": 0.5}
    assert [s.bitstring for s in ch.states] == ["000", "001", "011", "111", "100", "101"]
    arr = q.MCMCChain.from_arrays(3, 0, np.array([1, 3, 7, 4, 5], dtype=np.uint64), np.array([0, 1, 0, 1, 0]), name="t")
    assert arr.markov_chain == ch.markov_chain and arr.states == ch.states
    assert arr.acceptance_rate() == pytest.approx(0.4)
    clone = pickle.loads(pickle.dumps(arr))
    assert clone.markov_chain == ch.markov_chain and clone.name == "t"
    # legacy layout written by the reference (attrs name,_states,_current_state,_states_accepted,markov_chain)
    legacy = q.MCMCChain.__new__(q.MCMCChain)
    legacy.__setstate__({"name": "old", "_states": ch.states, "_current_state": S("100", True),
                         "_states_accepted": [], "markov_chain": ch.markov_chain})
    assert legacy.markov_chain == ch.markov_chain and legacy.name == "old"
    empty = q.MCMCChain(None)                                                     # quirk Q11: reference crashes
    assert len(empty) == 0 and empty.current_state is None


def test_qumcmc_alias_package_resolves_reference_module_names():
    import qumcmc
    import qumcmc.basic_utils as bu
    from qumcmc.energy_models import IsingEnergyFunction, Exact_Sampling  # noqa: F401
    from qumcmc.mixers import GenericMixer  # noqa: F401
    from qumcmc.quantum_mcmc_routines import quantum_enhanced_mcmc  # noqa: F401
    from qumcmc.classical_mcmc_routines import classical_mcmc  # noqa: F401
    import qumcmc_b200.basic_utils
    assert bu is qumcmc_b200.basic_utils and sys.modules["qumcmc.prob_dist"] is qumcmc_b200.prob_dist
    assert qumcmc.MCMCChain is bu.MCMCChain


def test_datasets_and_prob_dist_helpers():
    import qumcmc_b200 as q
    bas = q.bas_dataset(3)
    assert sorted(bas.bas_dict["bars"]) == sorted(["000000111", "000111000", "111000000", "000111111", "111000111", "111111000"])
    assert len(bas.dataset) == 12
    assert q.get_cardinality_dataset(4, 2) == ["0011", "0101", "0110", "1001", "1010", "1100"]
    W = q.hebbing_learning(["101", "110"])
    assert np.array_equal(W, np.array([[0, 0, 0], [0, 0, -2], [0, -2, 0]], dtype=float))
    d = q.DiscreteProbabilityDistribution({"00": 2.0, "01": 1.0, "10": 1.0, "11": 0.0})
    assert d["00"] == 0.5 and list(d.value_sorted_dict(reverse=True))[0] == "00"
    assert q.kl_divergence({"0": 0.5, "1": 0.5}, {"0": 0.5, "1": 0.5}) == 0
    assert q.kl_divergence({"0": 1.0}, {}) == pytest.approx(-np.log(1e-8))
    t = np.array([0.5, 0.5, 0.0]); m = np.array([0.25, 0.75, 0.0])
    assert q.vectoried_KL(t, m) == pytest.approx(0.5 * np.log(0.5 / 0.25) + 0.5 * np.log(0.5 / 0.75))
    assert q.js_divergence({"0": 0.5, "1": 0.5}, {"0": 0.5, "1": 0.5}) == pytest.approx(0.0, abs=1e-12)
    assert d.get_entropy() is None or d.get_entropy() >= 0


def test_host_philox_replica_matches_oracle():
    from qumcmc_b200.rng import initial_states_philox, philox4x32
    w = philox4x32(0x5EED1234ABCD, np.arange(3, 8, dtype=np.uint64), 17, 1)
    for k, c in enumerate(range(3, 8)):
        assert w[k].tolist() == oracle.philox(0x5EED1234ABCD, c, 17, 1).tolist()
    init = initial_states_philox(99, 5, 4, 20)
    assert init.tolist() == [oracle.initial_state(99, 5 + c, 20) for c in range(4)]


def test_shard_range_partitions_chains():
    from qumcmc_b200.distributed import shard_range
    for C in (0, 1, 7, 4096, 4099):
        for W in (1, 2, 3, 8):
            spans = [shard_range(C, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == C
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _gloo_worker(rank, world, port, C, width, q):
    import torch.distributed as dist
    from qumcmc_b200.distributed import run_chains_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def run_local(chain_id0, count):  # stand-in for the CUDA call: statistics are functions of the global id
        ids = np.arange(chain_id0, chain_id0 + count)
        return ids * 3, np.stack([ids * 10 + k for k in range(width)], axis=1).reshape(count, width), (chain_id0, count)

    acc, hist, payload = run_chains_sharded(run_local, C, None)
    # visit histogram of this rank's chains (here: chain id mod 8) -> one all-reduce, as numpy and as a tensor
    from qumcmc_b200.distributed import gather_statistics, reduce_visit_histogram, shard_range
    a, b = shard_range(C, rank, world)
    mine = np.bincount(np.arange(a, b) % 8, minlength=8).astype(np.int64)
    import torch
    total_t = reduce_visit_histogram(torch.from_numpy(mine.copy()), None)
    _, _, total_np = gather_statistics(np.arange(a, b) * 3, np.zeros((b - a, width), dtype=np.int64), C, None, visit_hist=mine)
    q.put((rank, acc.tolist(), hist.tolist(), payload, total_np.tolist(), total_t.tolist()))
    dist.destroy_process_group()


def test_chain_sharding_gather_gloo_world2():
    import torch.multiprocessing as mp
    C, width, world = 11, 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, C, width, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = np.arange(C)
    for rank, acc, hist, payload, visits_np, visits_t in outs:
        assert acc == (ids * 3).tolist()
        assert hist == np.stack([ids * 10 + k for k in range(width)], axis=1).tolist()
        assert visits_np == visits_t == np.bincount(ids % 8, minlength=8).tolist()   # K7: all-reduced visit histogram
    assert sorted(p[3] for p in outs) == [(0, 6), (6, 5)]


def _gloo_logz_worker(rank, world, port, n, beta, q):
    import torch.distributed as dist
    from qumcmc_b200.distributed import exact_logz_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    J, h = models.packaged_random_model(n, 5)

    class M:  # only num_spins is needed when partial_fn replaces the CUDA call
        num_spins = n

    def partial(begin, count):  # oracle energies stand in for qmc_exact_partial on CPU
        x = -beta * oracle.energies(J, h, np.arange(begin, begin + count, dtype=np.uint64))
        m = x.max()
        return float(m), float(np.exp(x - m).sum())

    q.put((rank, exact_logz_sharded(M(), beta, partial_fn=partial)))
    dist.destroy_process_group()


def test_exact_logz_sharding_gloo_world2():
    """The 'allreduce for Z' of the sharded exact sampler: per-rank (max, sumexp) pairs, all-gather, combine."""
    import torch.multiprocessing as mp
    n, beta, world = 10, 1.3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_gloo_logz_worker, args=(r, world, port, n, beta, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    J, h = models.packaged_random_model(n, 5)
    logZ, _, _ = oracle.exact_boltzmann(J, h, beta, want_probs=False)
    for _, val in outs:
        assert val == pytest.approx(logZ, rel=1e-13)


def test_combine_logsumexp_edge_cases():
    from qumcmc_b200.distributed import combine_logsumexp
    assert combine_logsumexp([[0.0, 1.0], [0.0, 1.0]]) == pytest.approx(np.log(2.0))
    assert combine_logsumexp([[1000.0, 1.0], [-1000.0, 5.0]]) == pytest.approx(1000.0)
    assert combine_logsumexp([[float("-inf"), 0.0], [2.0, 3.0]]) == pytest.approx(2.0 + np.log(3.0))


def test_initial_states_are_validated_on_the_host():
    """the reference asserts len(bitstring) == num_spins; an out-of-range index must not reach the device."""
    from qumcmc_b200.quantum_mcmc_routines import initial_state_indices
    assert initial_state_indices("0101", 4, 3).tolist() == [5, 5, 5]
    assert initial_state_indices(["0001", 7], 4, 2).tolist() == [1, 7]
    for bad in ("01010", "012", 16, -1, ["0000", "00000"]):
        with pytest.raises(ValueError):
            initial_state_indices(bad, 4, 2 if isinstance(bad, list) else 1)
    with pytest.raises(ValueError):
        initial_state_indices(["0000"], 4, 2)   # one state per chain


def test_chain_batch_without_a_record():
    """record=False batches keep acceptance counts / final states; views that need the record say so."""
    from qumcmc_b200.quantum_mcmc_routines import ChainBatch
    b = ChainBatch(4, np.array([1, 2], dtype=np.uint64), None, None, np.array([3, 4], dtype=np.uint64),
                   np.array([5, 7], dtype=np.int64), np.zeros((2, 5), dtype=np.int64), "x", n_hops=10)
    assert not b.recorded and b.acceptance_rate() == pytest.approx(12 / 20)
    with pytest.raises(ValueError, match="record=False"):
        b[0]
    from qumcmc_b200 import io as qio
    with pytest.raises(ValueError, match="record=False"):
        qio.save_chains("/tmp/never_written.npz", b)
    full = ChainBatch(4, np.array([1], dtype=np.uint64), np.array([[2, 3]], dtype=np.uint64),
                      np.array([[1, 0]], dtype=np.uint8), np.array([2], dtype=np.uint64), np.array([1], dtype=np.int64),
                      np.zeros((1, 5), dtype=np.int64), "y")
    assert full.recorded and full.n_hops == 2 and full.acceptance_rate() == 0.5 and full[0].markov_chain[-1] == "0010"


def test_overridden_builtin_proposal_rules_run_on_the_host():
    """a user subclass of a built-in rule that overrides propose_transition must not be lowered to the
    built-in device rule (the reference always calls propose_transition)."""
    from qumcmc_b200.classical_mixers import CombineProposals, CustomProposals, FixedWeightProposals, UniformProposals

    class Sticky(UniformProposals):
        def propose_transition(self, current_state=None):
            return current_state

    class Plain(FixedWeightProposals):
        pass

    assert UniformProposals(5).device_lowering() is not None
    assert Plain(5, 2).device_lowering() == FixedWeightProposals(5, 2).device_lowering()
    assert Sticky(5).device_lowering() is None
    assert CombineProposals([Sticky(5), CustomProposals(5, [0])], [0.5, 0.5]).device_lowering() is None
    assert CombineProposals([UniformProposals(5), CustomProposals(5, [0])], [0.5, 0.5]).device_lowering() is not None
