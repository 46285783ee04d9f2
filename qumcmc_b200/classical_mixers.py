"""Classical proposal rules of the reference's ``qumcmc.classical_mixers``
(/root/reference/qumcmc/classical_mixers.py).  ``propose_transition`` keeps the reference's host-side string
semantics (user subclasses plug in there); the four built-in rules also ``lower()`` to the
``(kinds, params, probabilities)`` mixture that ``qmc_classical_chains`` runs on the device."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List

import numpy as np

from . import _lib
from .basic_utils import get_random_state, random_bstr, xor_strings


class ClassicalMixer(ABC):
    def __init__(self, num_spins: int) -> None:
        self.num_spins = num_spins
        self._precompute_properties()

    @abstractmethod
    def propose_transition(self, current_state: str) -> str: ...

    def _precompute_properties(self):
        pass

    def _check(self, current_state: str) -> None:
        assert len(current_state) == self.num_spins, "Wrong 'current_state' "

    def lower(self):
        """``[(kind, param, probability), ...]`` for the device kernel, or ``None`` when the rule only exists as
        host code (a user subclass)."""
        return None


class UniformProposals(ClassicalMixer):
    def __repr__(self):
        return f"UniformProposals({self.num_spins})"

    def propose_transition(self, current_state=None):
        self._check(current_state)
        return get_random_state(self.num_spins)

    def lower(self):
        return [(_lib.CLASSICAL_UNIFORM, 0, 1.0)]


class FixedWeightProposals(ClassicalMixer):
    def __init__(self, num_spins: int, bodyness: int) -> None:
        assert bodyness <= num_spins, "Incorrect"
        self.bodyness = bodyness
        super().__init__(num_spins)

    def __repr__(self):
        return f"FixedWeightProposals({self.num_spins}, ({self.bodyness}))"

    def propose_transition(self, current_state: str):
        self._check(current_state)
        return xor_strings(current_state, random_bstr(self.num_spins, self.bodyness))

    def lower(self):
        return [(_lib.CLASSICAL_FIXED_WEIGHT, int(self.bodyness), 1.0)]


class CustomProposals(ClassicalMixer):
    def __init__(self, num_spins: int, flip_pattern: List[int]) -> None:
        self.flip_pattern = flip_pattern
        super().__init__(num_spins)

    def __repr__(self):
        return f"CustomProposals({self.num_spins}, pattern: {self.flip_pattern})"

    def propose_transition(self, current_state: str):
        self._check(current_state)
        flips = "".join("1" if i in self.flip_pattern else "0" for i in range(self.num_spins))
        return xor_strings(current_state, flips)

    def lower(self):
        mask = 0
        for i in set(self.flip_pattern):  # string position i (MSB first) <-> index bit n-1-i
            if 0 <= i < self.num_spins:
                mask |= 1 << (self.num_spins - 1 - i)
        return [(_lib.CLASSICAL_CUSTOM, mask, 1.0)]


class CombineProposals(ClassicalMixer):
    def __init__(self, proposal_methods: List[ClassicalMixer], probabilities: List[float]) -> None:
        self.num_spins = proposal_methods[0].num_spins
        assert all(self.num_spins == pm.num_spins for pm in proposal_methods), \
            "Mixers don't have the same number of qubits"
        assert len(probabilities) == len(proposal_methods), \
            "Length of list of mixers and probabilities is not equal"
        self.probabilities = np.array(probabilities) / sum(probabilities)
        self.proposal_methods = proposal_methods

    def __repr__(self):
        return f"CombinedProposals( num-proposal-methods = {len(self.proposal_methods)} , p = {self.probabilities})"

    def propose_transition(self, current_state: str):
        self._check(current_state)
        method = np.random.choice(self.proposal_methods, p=self.probabilities)
        return method.propose_transition(current_state)

    def lower(self):
        out = []
        for method, p in zip(self.proposal_methods, self.probabilities):
            sub = method.lower()
            if sub is None:
                return None
            out.extend((kind, param, float(p) * q) for kind, param, q in sub)
        return out
