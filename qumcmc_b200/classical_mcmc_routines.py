"""``classical_mcmc`` and ``test_accept`` (/root/reference/qumcmc/classical_mcmc_routines.py).

The built-in proposal rules run on the device, all chains and hops in one launch (``qmc_classical_chains``, SURVEY
section 8 row f2): Philox draws, fixed-order fp64 energies and the Metropolis rule are shared with the quantum path."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _device, _lib
from .basic_utils import MCMCChain, MCMCState, get_random_state
from .classical_mixers import ClassicalMixer, UniformProposals
from .energy_models import IsingEnergyFunction
from .quantum_mcmc_routines import ChainBatch, initial_state_indices


def test_accept(energy_s: float, energy_sprime: float, temperature: float = 1.0) -> bool:
    """Metropolis rule: min(1, exp(-(E(s')-E(s))/T)) > U[0,1) (classical_mcmc_routines.py:28-41)."""
    with np.errstate(over="ignore"):
        factor = np.exp(-(energy_sprime - energy_s) / temperature)
    return min(1, factor) > np.random.rand()


test_accept.__test__ = False  # not a pytest test


def lower_proposals(proposition_method: ClassicalMixer):
    """(kinds int32[M], params uint64[M], probs float64[M]) of a built-in proposal rule, or None for host-only rules."""
    spec = proposition_method.device_lowering()
    if spec is None:
        return None
    if len(spec) > _lib.MAX_CLASSICAL_METHODS:
        raise ValueError(f"CombineProposals with {len(spec)} leaf rules (device limit {_lib.MAX_CLASSICAL_METHODS})")
    kinds = np.array([k for k, _, _ in spec], dtype=np.int32)
    params = np.array([p for _, p, _ in spec], dtype=np.uint64)
    probs = np.array([q for _, _, q in spec], dtype=np.float64)
    return kinds, params, probs


def run_classical_chains(model: IsingEnergyFunction, proposition_method: ClassicalMixer, n_hops: int, n_chains: int,
                         initial_states: Optional[np.ndarray], temperature: float, seed: int, device: int = 0,
                         chain_id0: int = 0, hop0: int = 0, name: str = "cl-mcmc", record: bool = True,
                         visit_histogram: bool = False) -> ChainBatch:
    """Batched driver over qmc_classical_chains (device buffers are torch tensors): every chain, every hop, one launch."""
    spec = lower_proposals(proposition_method)
    if spec is None:
        raise _lib.QmcUnsupported(f"{proposition_method!r} has no device lowering (host-only proposal rule)")
    kinds, params, probs = spec
    torch = _device._torch()
    dm = model.device_model(device)
    n = model.num_spins
    dev = _device.require_cuda(device)
    C_, H = int(n_chains), int(n_hops)
    s0 = None
    if initial_states is not None:
        s0 = torch.from_numpy(np.ascontiguousarray(initial_states, dtype=np.uint64).view(np.int64)).to(dev)
    props = torch.empty((C_, H), dtype=torch.int64, device=dev) if record else None
    acc = torch.empty((C_, H), dtype=torch.uint8, device=dev) if record else None
    fin = torch.empty(C_, dtype=torch.int64, device=dev)
    cnt = torch.empty(C_, dtype=torch.int64, device=dev)
    hist = torch.empty((C_, n + 1), dtype=torch.int64, device=dev)
    ptr = lambda t: None if t is None else t.data_ptr()
    visits = torch.zeros(1 << n, dtype=torch.int64, device=dev) if visit_histogram else None
    if visits is not None:
        dm.attach_visit_histogram(visits)
    try:
        _lib.check(_lib.lib().qmc_classical_chains(
            dm.handle, C_, H, ptr(s0), int(chain_id0), int(hop0) & 0xFFFFFFFF, float(temperature), len(kinds),
            kinds.ctypes.data, params.ctypes.data, probs.ctypes.data, int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(props),
            ptr(acc), ptr(fin), ptr(cnt), ptr(hist), _device.current_stream_ptr(device)))
        torch.cuda.synchronize(device)
    finally:
        if visits is not None:
            dm.attach_visit_histogram(None)
    to_u64 = lambda t: t.cpu().numpy().view(np.uint64)
    if initial_states is None:
        from .rng import initial_states_philox
        init = initial_states_philox(int(seed), int(chain_id0), C_, n)
    else:
        init = np.ascontiguousarray(initial_states, dtype=np.uint64)
    return ChainBatch(n, init, to_u64(props) if record else None, acc.cpu().numpy() if record else None,
                      to_u64(fin), cnt.cpu().numpy(), hist.cpu().numpy(), name, n_hops=H,
                      visit_hist=None if visits is None else visits.cpu().numpy())


def classical_mcmc(n_hops: int, model: IsingEnergyFunction, proposition_method: Optional[ClassicalMixer] = None,
                   initial_state: Optional[str] = None, temperature: float = 1.0, verbose: bool = False,
                   name="cl-mcmc", *, n_chains: Optional[int] = None, seed: Optional[int] = None, device: int = 0,
                   chain_id0: int = 0, hop0: int = 0, record: bool = True, visit_histogram: bool = False):
    """Metropolis chain with classical proposals.  ``proposition_method`` defaults to
    ``UniformProposals`` (the README call shape passes none).  The built-in proposal rules run on the device
    (``qmc_classical_chains``: returns an ``MCMCChain``, or a ``ChainBatch`` when ``n_chains`` is given); a user
    subclass of ``ClassicalMixer`` keeps the reference's host loop around its ``propose_transition`` with energies
    from the CUDA energy kernel."""
    n = model.num_spins
    if proposition_method is None:
        proposition_method = UniformProposals(n)
    if proposition_method.device_lowering() is not None:
        if seed is None:  # stays under the control of np.random.seed, like the reference's global RNG use
            seed = int(np.random.randint(0, 2 ** 62))
        single = n_chains is None
        C_ = 1 if single else int(n_chains)
        if initial_state is None:
            init = np.array([int(get_random_state(n), 2)], dtype=np.uint64) if single else None
        else:
            init = initial_state_indices(initial_state, n, C_)
        if verbose and init is not None:
            print("starting with: ", f"{int(init[0]):0{n}b}", "with energy:", model.get_energy(f"{int(init[0]):0{n}b}"))
        batch = run_classical_chains(model, proposition_method, n_hops, C_, init, temperature, seed, device=device,
                                     chain_id0=chain_id0, hop0=hop0, name=name, record=record or single,
                                     visit_histogram=visit_histogram)
        return batch[0] if single else batch
    if n_chains is not None:
        raise _lib.QmcUnsupported(f"{proposition_method!r} has no device lowering: n_chains needs a built-in proposal rule")
    start = get_random_state(n) if initial_state is None else initial_state
    if not isinstance(start, str):
        start = f"{int(start):0{n}b}"
    initial_state_indices(start, n, 1)  # width / alphabet check
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
