"""Chain/state containers and small helpers of the reference's ``qumcmc.basic_utils``.

``MCMCState`` / ``MCMCChain`` keep the reference's attribute and method names
(/root/reference/qumcmc/basic_utils.py:54-131) but are backed by integer arrays coming straight from
the device (``proposals[H]``, ``accepted[H]``), so a 4096 x 10 000 run never materialises 4e7 Python
objects unless ``.states`` is asked for.  Pickles written by the reference load into these classes
(legacy attributes ``name, _states, _current_state, _states_accepted, markov_chain``), and pickles
written here carry the same five attributes.
"""
from __future__ import annotations

import itertools
import random
from collections import Counter
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np


# --------------------------------------------------------------------------------------------
# bitstring helpers
# --------------------------------------------------------------------------------------------
def int_to_binary(state_obtained: int, n_spins: int) -> str:
    return f"{int(state_obtained):0{n_spins}b}"


int_to_str = int_to_binary


def binary_to_bipolar(string: str) -> float:
    return 2.0 * float(string) - 1.0


def states(num_spins: int) -> List[str]:
    """All 2^n bitstrings in index order (basic_utils.py:160-172)."""
    return [int_to_binary(k, num_spins) for k in range(1 << num_spins)]


def xor_strings(s1: str, s2: str) -> str:
    if len(s1) != len(s2):
        raise AssertionError("Strings must be of the same length")
    return int_to_binary(int(s1, 2) ^ int(s2, 2), len(s1))


def hamming_dist(str1: str, str2: str) -> int:
    return bin(int(str1, 2) ^ int(str2, 2)).count("1")


def magnetization_of_state(bitstring: str) -> float:
    ones = bitstring.count("1")
    return ones - (len(bitstring) - ones)


def dict_magnetization_of_all_states(list_all_possible_states: Sequence[str]) -> Dict[str, float]:
    return {s: magnetization_of_state(s) for s in list_all_possible_states}


def value_sorted_dict(dict_in: dict, reverse: bool = False) -> dict:
    return dict(sorted(dict_in.items(), key=lambda kv: kv[1], reverse=reverse))


def sort_dict_by_keys(dict_in: dict) -> dict:
    return dict(sorted(dict_in.items()))


def merge_2_dict(dict1: dict, dict2: dict) -> dict:
    out = dict(dict1)
    out.update(dict2)
    return out


def uncommon_els_2_lists(list_1, list_2):
    return list(set(list_1) ^ set(list_2))


def get_distn(list_of_samples: Sequence[str]) -> Dict[str, float]:
    total = len(list_of_samples)
    return {s: c * (1.0 / total) for s, c in Counter(list_of_samples).items()}


def random_bstr(n: int, k: int = 1) -> str:
    """Random bitstring with exactly k ones (basic_utils.py:383-388; uses ``random``)."""
    assert n >= k
    chars = ["1"] * k + ["0"] * (n - k)
    random.shuffle(chars)
    return "".join(chars)


def get_random_state(num_spins: int) -> str:
    """Uniform bitstring via the global numpy RNG (basic_utils.py:395-404)."""
    return int_to_binary(np.random.randint(0, 2 ** num_spins), num_spins)


# --------------------------------------------------------------------------------------------
# chain containers
# --------------------------------------------------------------------------------------------
@dataclass
class MCMCState:
    bitstring: str
    accepted: bool


class MCMCChain:
    """Record of one chain: the initial state followed by one ``MCMCState`` per hop.

    Array-backed: ``from_arrays`` wraps device results without creating Python objects; the
    list views (``states``, ``markov_chain``, ``accepted_states``) are built on first use.
    """

    def __init__(self, states: Optional[List[MCMCState]] = None, name: Optional[str] = "MCMC"):
        self.name = name
        self._n: Optional[int] = None
        self._idx = np.zeros(0, dtype=np.uint64)      # proposal indices incl. the initial state
        self._acc = np.zeros(0, dtype=bool)
        self._states_cache: Optional[List[MCMCState]] = None
        self._mc_cache: Optional[List[str]] = None
        if states:  # the reference crashes on None (its `len(states) is None` test, Q11); accept it
            self._n = len(states[0].bitstring)
            self._idx = np.array([int(s.bitstring, 2) for s in states], dtype=np.uint64)
            self._acc = np.array([bool(s.accepted) for s in states], dtype=bool)
            self._states_cache = list(states)

    # ---- construction from device results
    @classmethod
    def from_arrays(cls, num_spins: int, initial_index: int, proposals: np.ndarray, accepted: np.ndarray,
                    name: str = "MCMC") -> "MCMCChain":
        ch = cls(None, name=name)
        ch._n = int(num_spins)
        ch._idx = np.concatenate([np.array([initial_index], dtype=np.uint64), np.asarray(proposals, dtype=np.uint64)])
        ch._acc = np.concatenate([np.array([True]), np.asarray(accepted).astype(bool)])
        return ch

    # ---- array views (no Python object blow-up)
    @property
    def num_spins(self) -> Optional[int]:
        return self._n

    @property
    def proposal_indices(self) -> np.ndarray:
        return self._idx

    @property
    def accepted_mask(self) -> np.ndarray:
        return self._acc

    def markov_chain_indices(self) -> np.ndarray:
        """State after every hop (index form of ``markov_chain``)."""
        if len(self._idx) == 0:
            return self._idx
        pos = np.where(self._acc, np.arange(len(self._idx)), 0)
        pos = np.maximum.accumulate(pos)
        return self._idx[pos]

    def acceptance_rate(self) -> float:
        return float(self._acc[1:].mean()) if len(self._acc) > 1 else float("nan")

    # ---- reference API
    def add_state(self, state: MCMCState):
        if self._n is None:
            self._n = len(state.bitstring)
        self._idx = np.append(self._idx, np.uint64(int(state.bitstring, 2)))
        self._acc = np.append(self._acc, bool(state.accepted))
        if self._states_cache is not None:
            self._states_cache.append(state)
        self._mc_cache = None

    @property
    def states(self) -> List[MCMCState]:
        if self._states_cache is None:
            n = self._n
            self._states_cache = [MCMCState(int_to_binary(i, n), bool(a)) for i, a in zip(self._idx.tolist(), self._acc.tolist())]
        return self._states_cache

    _states = states  # legacy attribute name

    @property
    def current_state(self) -> Optional[MCMCState]:
        hit = np.flatnonzero(self._acc)
        if len(hit) == 0:
            return None
        k = int(hit[-1])
        return MCMCState(int_to_binary(int(self._idx[k]), self._n), True)

    _current_state = current_state

    @property
    def _states_accepted(self) -> List[MCMCState]:
        return [MCMCState(int_to_binary(int(i), self._n), True) for i in self._idx[self._acc]]

    @property
    def accepted_states(self) -> List[str]:
        return [int_to_binary(int(i), self._n) for i in self._idx[self._acc]]

    def get_list_markov_chain(self) -> List[str]:
        n = self._n
        self._mc_cache = [int_to_binary(int(i), n) for i in self.markov_chain_indices()]
        return self._mc_cache

    @property
    def markov_chain(self) -> List[str]:
        if self._mc_cache is None:
            self.get_list_markov_chain()
        return self._mc_cache

    def get_accepted_dict(self, normalize: bool = False, until_index: int = -1):
        mc = self.markov_chain_indices()
        if until_index != -1:
            mc = mc[:until_index]
        vals, counts = np.unique(mc, return_counts=True)
        n = self._n
        if normalize:
            length = len(mc)
            return Counter({int_to_binary(int(v), n): c / length for v, c in zip(vals, counts)})
        return Counter({int_to_binary(int(v), n): int(c) for v, c in zip(vals, counts)})

    def visit_histogram(self) -> np.ndarray:
        """counts[2^n] of the markov chain (dense; use only for small n)."""
        return np.bincount(self.markov_chain_indices().astype(np.int64), minlength=1 << self._n)

    def __len__(self) -> int:
        return len(self._idx)

    def __repr__(self) -> str:
        return f"MCMCChain(name={self.name!r}, states={len(self._idx)}, num_spins={self._n})"

    # ---- pickle compatibility with the reference's layout
    def __getstate__(self):
        return {
            "name": self.name,
            "_states": self.states,
            "_current_state": self.current_state,
            "_states_accepted": self._states_accepted,
            "markov_chain": self.markov_chain,
        }

    def __setstate__(self, state):
        sts = state.get("_states") or []
        self.__init__(list(sts), name=state.get("name", "MCMC"))


# --------------------------------------------------------------------------------------------
# data sets used by the reference's benchmarks (basic_utils.py:581-660)
# --------------------------------------------------------------------------------------------
class bas_dataset:
    """Bars-and-stripes patterns on a grid_size x grid_size grid (all-0 / all-1 excluded)."""

    def __init__(self, grid_size: int):
        self.grid_size = grid_size
        combos = ["".join(p) for p in itertools.product("01", repeat=grid_size)]
        combos.sort(key=lambda s: s.count("1"))
        self._combos = combos[1:-1]
        self.bas_dict = self.bars_and_stripes_dataset()
        self.dataset = self.bas_dict["stripes"] + self.bas_dict["bars"] + self.bas_dict.get("both", [])

    def vertical_stripes(self) -> List[str]:
        return [row * self.grid_size for row in self._combos]

    def horizontal_bars(self) -> List[str]:
        return ["".join(ch * self.grid_size for ch in pattern) for pattern in self._combos]

    def bars_and_stripes_dataset(self) -> Dict[str, List[str]]:
        return {"stripes": self.vertical_stripes(), "bars": self.horizontal_bars()}

    def bit_string_to_2d_matrix(self, bitstring: str, array_shape: int) -> np.ndarray:
        return np.array([int(c) for c in bitstring]).reshape(array_shape, array_shape)


def hebbing_learning(list_bas_state: Iterable[str]) -> np.ndarray:
    """Hebbian weights sum_s s s^T - |data| * I with s in {-1,+1}^n (basic_utils.py:633-643)."""
    data = list(list_bas_state)
    size = len(data[0])
    S = np.array([[1 if c == "1" else -1 for c in s] for s in data], dtype=float)
    return S.T @ S - len(data) * np.identity(size)


def get_cardinality_dataset(n_qubits: int, card: int = 2) -> List[str]:
    return [s for s in states(n_qubits) if s.count("1") == card]
