"""Quantum mixers of the reference's ``qumcmc.mixers`` as data.

In the reference a mixer *returns a qulacs circuit* (``get_mixer_circuit(gamma, delta_time)``,
/root/reference/qumcmc/mixers.py:8-18).  Here a mixer lowers to the term list the CUDA kernels
consume: groups of ``(mask, weight)`` where the term evolves as
``exp(-i * gamma * weight * delta_time * X_mask)``.  ``mask`` has bit ``q`` set for every *qulacs
qubit index* ``q`` in the X-string (LSB-based, no ``n-1-j`` reversal: mixers.py:40-50, quirk Q2).

    GenericMixer / CustomMixer  -> one group, weight 1            (mixers.py:21-105)
    CoherentMixerSum            -> one group, weight w_m / sum(w) (mixers.py:108-130)
    IncoherentMixerSum          -> one group per sub-mixer, drawn per proposal (mixers.py:133-151)
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from itertools import combinations
from typing import List, Optional, Sequence, Tuple

import numpy as np

Term = Tuple[int, float]  # (mask, weight)


def _mask_of(qubits: Sequence[int], n_qubits: int) -> int:
    mask = 0
    for q in qubits:
        q = int(q)
        if not 0 <= q < n_qubits:
            raise ValueError(f"qubit index {q} outside [0, {n_qubits})")
        if mask >> q & 1:
            raise ValueError(f"qubit {q} repeated in X-string {list(qubits)}")
        mask |= 1 << q
    if mask == 0:
        raise ValueError("empty X-string")
    return mask


class Mixer(ABC):
    def __init__(self, n_qubits: int) -> None:
        self.n_qubits = n_qubits
        self._precompute_properties()

    def get_mixer_circuit(self, gamma: float, delta_time: float):
        raise NotImplementedError(
            "qumcmc_b200 mixers do not build qulacs circuits; they lower to (mask, weight) term lists "
            "consumed by libqumcmc_b200 (see Mixer.lower()).")

    def _precompute_properties(self):
        pass

    @abstractmethod
    def term_groups(self) -> Tuple[List[List[Term]], Optional[List[float]]]:
        """(groups, probabilities): probabilities is None for a single coherent group."""

    def lower(self):
        """(group_offsets int32[G+1], masks uint64[T], weights float64[T], group_prob float64[G] | None)."""
        groups, probs = self.term_groups()
        offs = [0]
        masks: List[int] = []
        weights: List[float] = []
        for g in groups:
            for mask, w in g:
                masks.append(mask)
                weights.append(w)
            offs.append(len(masks))
        return (np.asarray(offs, dtype=np.int32), np.asarray(masks, dtype=np.uint64),
                np.asarray(weights, dtype=np.float64),
                None if probs is None else np.asarray(probs, dtype=np.float64))


class GenericMixer(Mixer):
    """All C(n, bodyness) X-strings of a given weight."""

    def __init__(self, n_qubits: int, bodyness: int) -> None:
        self.bodyness = bodyness
        super().__init__(n_qubits)

    def __repr__(self):
        return f" GenericMixer : \n ------------- \n num-qubits = {self.n_qubits } | bodyness = {self.bodyness}"

    def _precompute_properties(self):
        self.possible_qubit_combinations = [list(c) for c in combinations(range(self.n_qubits), self.bodyness)]

    def term_groups(self):
        return [[(_mask_of(c, self.n_qubits), 1.0) for c in self.possible_qubit_combinations]], None

    @classmethod
    def OneBodyMixer(cls, n_qubits: int):
        return cls(n_qubits, 1)

    @classmethod
    def TwoBodyMixer(cls, n_qubits: int):
        return cls(n_qubits, 2)

    @classmethod
    def ThreeBodyMixer(cls, n_qubits: int):
        return cls(n_qubits, 3)


class CustomMixer(Mixer):
    """User-supplied X-strings (lists of qubit indices)."""

    def __init__(self, n_qubits: int, possible_qubit_combinations: List[List[int]]) -> None:
        super().__init__(n_qubits)
        self.possible_qubit_combinations = possible_qubit_combinations

    def __repr__(self):
        return (f" GenericMixer : \n ------------- \n num-qubits = {self.n_qubits } \n -------- \n "
                f"possible-qubit-combintations = {self.possible_qubit_combinations}")

    def term_groups(self):
        return [[(_mask_of(c, self.n_qubits), 1.0) for c in self.possible_qubit_combinations]], None


def _check_same_width(mixers: Sequence[Mixer]) -> int:
    n = mixers[0].n_qubits
    assert all(n == m.n_qubits for m in mixers), "Mixers don't have the same number of qubits"
    return n


class CoherentMixerSum(Mixer):
    """All sub-mixers applied in every layer, sub-mixer m with gamma * w_m (w normalised)."""

    def __init__(self, mixers: List[Mixer], weights: List[float]) -> None:
        self.n_qubits = _check_same_width(mixers)
        assert len(weights) == len(mixers), "Length of list of mixers and weights is not equal"
        self.mixers = mixers
        self.weights = np.array(weights) / sum(weights)

    def __repr__(self):
        return f"CoherentMixerSum -> num_mixers : {len(self.mixers)} |  weights : {self.weights}"

    def term_groups(self):
        merged: List[Term] = []
        for w, mx in zip(self.weights, self.mixers):
            groups, probs = mx.term_groups()
            if probs is not None:
                raise NotImplementedError("an IncoherentMixerSum inside a CoherentMixerSum is not supported")
            merged.extend((mask, float(w) * tw) for mask, tw in groups[0])
        return [merged], None


class IncoherentMixerSum(Mixer):
    """One sub-mixer per proposal, drawn with the given probabilities."""

    def __init__(self, mixers: List[Mixer], probabilities: List[float]) -> None:
        self.n_qubits = _check_same_width(mixers)
        assert len(probabilities) == len(mixers), "Length of list of mixers and probabilities is not equal"
        self.mixers = mixers
        self.probabilities = np.array(probabilities) / sum(probabilities)

    def __repr__(self):
        return f"IncoherentMixerSum -> num_mixers : {len(self.mixers)} |  probabilities : {self.probabilities}"

    def term_groups(self):
        out_groups: List[List[Term]] = []
        out_probs: List[float] = []
        for p, mx in zip(self.probabilities, self.mixers):
            groups, probs = mx.term_groups()
            if probs is None:
                out_groups.append(groups[0])
                out_probs.append(float(p))
            else:  # nested incoherent sum: probabilities multiply
                for g, q in zip(groups, probs):
                    out_groups.append(g)
                    out_probs.append(float(p) * float(q))
        return out_groups, out_probs
