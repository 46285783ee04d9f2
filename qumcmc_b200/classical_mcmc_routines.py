"""``classical_mcmc`` and ``test_accept`` (/root/reference/qumcmc/classical_mcmc_routines.py).

Classical proposals are string flips on the host (not the quantum hot path); the
energies still come from the CUDA energy kernel (one batched table for n <= 20)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .basic_utils import MCMCChain, MCMCState, get_random_state
from .classical_mixers import ClassicalMixer, UniformProposals
from .energy_models import IsingEnergyFunction


def test_accept(energy_s: float, energy_sprime: float, temperature: float = 1.0) -> bool:
    """Metropolis rule: min(1, exp(-(E(s')-E(s))/T)) > U[0,1) (classical_mcmc_routines.py:28-41)."""
    with np.errstate(over="ignore"):
        factor = np.exp(-(energy_sprime - energy_s) / temperature)
    return min(1, factor) > np.random.rand()


test_accept.__test__ = False  # not a pytest test


def classical_mcmc(n_hops: int, model: IsingEnergyFunction, proposition_method: Optional[ClassicalMixer] = None,
                   initial_state: Optional[str] = None, temperature: float = 1.0, verbose: bool = False,
                   name="cl-mcmc", *, device: int = 0):
    """Metropolis chain with classical proposals.  ``proposition_method`` defaults to
    ``UniformProposals`` (the README call shape passes none)."""
    n = model.num_spins
    if proposition_method is None:
        proposition_method = UniformProposals(n)
    start = get_random_state(n) if initial_state is None else initial_state
    table = model.get_energies(np.arange(1 << n, dtype=np.uint64), device=device) if n <= 20 else None
    energy_of = (lambda s: float(table[int(s, 2)])) if table is not None else model.get_energy
    current = start
    energy_s = energy_of(current)
    if verbose:
        print("starting with: ", current, "with energy:", energy_s)
    proposals = np.empty(n_hops, dtype=np.uint64)
    accepted = np.empty(n_hops, dtype=bool)
    for k in range(n_hops):
        s_prime = proposition_method.propose_transition(current)
        energy_sprime = energy_of(s_prime)
        ok = test_accept(energy_s, energy_sprime, temperature=temperature)
        proposals[k] = int(s_prime, 2)
        accepted[k] = ok
        if ok:
            current, energy_s = s_prime, energy_sprime
    return MCMCChain.from_arrays(n, int(start, 2), proposals, accepted, name=name)
