"""Classical proposal rules with the surface of the reference's ``qumcmc.classical_mixers``
(/root/reference/qumcmc/classical_mixers.py:27-101).

Every built-in rule is described by what it does to the current state -- replace it (uniform) or xor it with a flip
pattern -- and ``lower()`` turns that description into the ``(kind, param, probability)`` mixture that
``qmc_classical_chains`` runs on the device.  ``propose_transition`` is the host-side
entry point the reference's callers use; user subclasses only need to override it (they then run in the reference's
host loop, see ``classical_mcmc``)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .basic_utils import get_random_state, random_bstr, xor_strings

Lowered = List[Tuple[int, int, float]]


class ClassicalMixer(ABC):
    """A rule ``current bitstring -> proposed bitstring`` over ``num_spins`` spins."""

    def __init__(self, num_spins: int) -> None:
        self.num_spins = num_spins
        self._precompute_properties()

    def _precompute_properties(self):
        """Hook kept for subclasses written against the reference."""

    def _require_width(self, state: str) -> None:
        assert len(state) == self.num_spins, "Wrong 'current_state' "

    @abstractmethod
    def propose_transition(self, current_state: str) -> str: ...

    def lower(self) -> Optional[Lowered]:
        """Device form ``[(kind, param, probability), ...]``; ``None`` = host-only rule (the default for subclasses)."""
        return None

    def device_lowering(self) -> Optional[Lowered]:
        """``lower()`` -- unless a user subclass of a built-in rule overrides ``propose_transition`` (or the flip
        pattern behind it): the device rule would silently ignore the override, whereas the reference always calls
        ``propose_transition`` (classical_mcmc_routines.py:44-99).  Such rules run in the host loop."""
        for attr in ("propose_transition", "_pattern", "_positions"):
            mine = getattr(type(self), attr, None)
            owner = next((k for k in type(self).__mro__ if attr in k.__dict__), None)
            if mine is not None and owner is not None and owner.__module__ != __name__:
                return None
        return self.lower()


class _XorRule(ClassicalMixer):
    """Built-in rules of the form ``state xor pattern``; subclasses supply the pattern."""

    def _pattern(self) -> str:
        raise NotImplementedError

    def propose_transition(self, current_state: str) -> str:
        self._require_width(current_state)
        return xor_strings(current_state, self._pattern())


class UniformProposals(ClassicalMixer):
    """Independent uniform draw over all 2^n states (reference :27-40)."""

    def __repr__(self):
        return f"UniformProposals({self.num_spins})"

    def propose_transition(self, current_state=None) -> str:
        self._require_width(current_state)
        return get_random_state(self.num_spins)

    def lower(self) -> Lowered:
        return [(_lib.CLASSICAL_UNIFORM, 0, 1.0)]


class FixedWeightProposals(_XorRule):
    """Flip ``bodyness`` spins chosen uniformly at random (reference :43-59, ``random_bstr``)."""

    def __init__(self, num_spins: int, bodyness: int) -> None:
        assert bodyness <= num_spins, "Incorrect"
        self.bodyness = bodyness
        super().__init__(num_spins)

    def __repr__(self):
        return f"FixedWeightProposals({self.num_spins}, ({self.bodyness}))"

    def _pattern(self) -> str:
        return random_bstr(self.num_spins, self.bodyness)

    def lower(self) -> Lowered:
        return [(_lib.CLASSICAL_FIXED_WEIGHT, int(self.bodyness), 1.0)]


class CustomProposals(_XorRule):
    """Flip a fixed set of string positions (reference :62-78)."""

    def __init__(self, num_spins: int, flip_pattern: List[int]) -> None:
        self.flip_pattern = flip_pattern
        super().__init__(num_spins)

    def __repr__(self):
        return f"CustomProposals({self.num_spins}, pattern: {self.flip_pattern})"

    def _positions(self) -> Sequence[int]:
        return sorted(i for i in set(self.flip_pattern) if 0 <= i < self.num_spins)

    def _pattern(self) -> str:
        marks = ["0"] * self.num_spins
        for i in self._positions():
            marks[i] = "1"
        return "".join(marks)

    def lower(self) -> Lowered:
        n = self.num_spins  # string position i (MSB first) <-> index bit n-1-i
        return [(_lib.CLASSICAL_CUSTOM, sum(1 << (n - 1 - i) for i in self._positions()), 1.0)]


class CombineProposals(ClassicalMixer):
    """Mixture: pick one of ``proposal_methods`` with the given probabilities, then apply it (reference :81-101)."""

    def __init__(self, proposal_methods: List[ClassicalMixer], probabilities: List[float]) -> None:
        widths = {pm.num_spins for pm in proposal_methods}
        assert len(widths) == 1, "Mixers don't have the same number of qubits"
        assert len(probabilities) == len(proposal_methods), "Length of list of mixers and probabilities is not equal"
        self.num_spins = widths.pop()
        self.proposal_methods = proposal_methods
        self.probabilities = np.array(probabilities) / sum(probabilities)

    def __repr__(self):
        return f"CombinedProposals( num-proposal-methods = {len(self.proposal_methods)} , p = {self.probabilities})"

    def propose_transition(self, current_state: str) -> str:
        self._require_width(current_state)
        chosen = np.random.choice(len(self.proposal_methods), p=self.probabilities)
        return self.proposal_methods[int(chosen)].propose_transition(current_state)

    def lower(self) -> Optional[Lowered]:
        flat: Lowered = []
        for weight, method in zip(self.probabilities, self.proposal_methods):
            leaves = method.device_lowering()
            if leaves is None:
                return None
            flat += [(kind, param, float(weight) * q) for kind, param, q in leaves]
        return flat
