"""On-disk compatibility: reference-written pickles load, chains written here keep the
reference's attribute layout, and the compact .npz format round-trips."""
import os
import pickle
import pickletools

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LEGACY = os.path.join(HERE, "golden", "legacy_chains.pkl")


def test_fixture_is_in_the_reference_byte_format():
    """The pickle names qumcmc.basic_utils.MCMCChain / MCMCState (what the reference's files contain), not this package."""
    ops = [(op.name, arg) for op, arg, _ in pickletools.genops(open(LEGACY, "rb").read())]
    strings = {arg for name, arg in ops if isinstance(arg, str)}
    assert "qumcmc.basic_utils" in strings and "MCMCChain" in strings and "MCMCState" in strings
    assert not any("qumcmc_b200" in s for s in strings)


def test_legacy_pickle_loads_with_reference_semantics():
    from qumcmc_b200.io import load_legacy_pickle
    from qumcmc_b200.basic_utils import MCMCChain
    d = load_legacy_pickle(LEGACY)
    assert len(d["chains"]) == 2
    for ch in d["chains"]:
        assert isinstance(ch, MCMCChain) and len(ch.states) == 300 and ch.num_spins == 9
        assert ch.states[0].bitstring == "111000111" and ch.states[0].accepted          # BAS runs start from 111000111
        mc = [ch.states[0].bitstring]                                                   # basic_utils.py:105-115
        for s in ch.states[1:]:
            mc.append(s.bitstring if s.accepted else mc[-1])
        assert ch.markov_chain == mc
        assert ch.accepted_states == [s.bitstring for s in ch.states if s.accepted]
        assert ch.current_state.bitstring == mc[-1]
        assert sum(ch.get_accepted_dict().values()) == 300


def test_chains_written_here_keep_the_reference_attribute_layout():
    from qumcmc_b200.basic_utils import MCMCChain
    ch = MCMCChain.from_arrays(5, 3, np.array([7, 1, 30], dtype=np.uint64), np.array([True, False, True]), name="x")
    raw = pickle.dumps(ch)
    state = pickle.loads(raw).__getstate__()
    assert set(state) == {"name", "_states", "_current_state", "_states_accepted", "markov_chain"}   # basic_utils.py:60-89
    assert state["markov_chain"] == ["00011", "00111", "00111", "11110"]


def test_compact_format_round_trip(tmp_path):
    from qumcmc_b200.io import load_chains, load_legacy_pickle, save_chains
    chains = load_legacy_pickle(LEGACY)["chains"]
    path = str(tmp_path / "chains.npz")
    save_chains(path, chains)
    assert os.path.getsize(path) < 0.5 * os.path.getsize(LEGACY)
    back = load_chains(path)
    assert len(back) == 2
    for a, b in zip(chains, back):
        assert np.array_equal(a.proposal_indices, b.proposal_indices) and np.array_equal(a.accepted_mask, b.accepted_mask)
        assert a.name == b.name and a.markov_chain == b.markov_chain
    n, initial, proposals, accepted, names = load_chains(path, as_batch=True)
    assert n == 9 and proposals.shape == (2, 299) and accepted.shape == (2, 299) and initial.dtype == np.uint64


REF_PICKLE = "/root/reference/gamma_variation/num_spins_8_Gamma_var_num_spins_8_beta_1.5_seed_1.pkl"


@pytest.mark.skipif(not os.path.exists(REF_PICKLE), reason="reference checkout not present (GPU box)")
def test_reference_written_pickle_loads():
    """A pickle the reference itself wrote (gamma-variation experiment): a dict gamma-range -> MCMCChain of 10 001 states.
    The acceptance rates must equal the ones tests/golden/make_golden.py extracted from the same file."""
    import json
    from qumcmc_b200.basic_utils import MCMCChain
    from qumcmc_b200.io import load_chains, load_legacy_pickle, save_chains
    d = load_legacy_pickle(REF_PICKLE)
    with open(os.path.join(HERE, "golden", "chain_stats.json")) as f:
        fx = json.load(f)["gamma_variation"]["n8_seed1"]["ranges"]
    assert set(d) == set(fx)
    for key, ch in d.items():
        assert isinstance(ch, MCMCChain) and ch.num_spins == 8 and len(ch) == fx[key]["n_hops"] + 1
        assert abs(ch.acceptance_rate() - fx[key]["acceptance"]) < 1e-12
        assert int(ch.proposal_indices[0]) == fx[key]["initial"]
        top = sorted(zip(*np.unique(ch.markov_chain_indices(), return_counts=True)), key=lambda kc: -kc[1])[:4]
        assert [int(c) for _, c in top] == [c for _, c in fx[key]["visit_top"][:4]]
