"""``quantum_enhanced_mcmc`` / ``run_qmcmc_quantum_ckt`` with the reference's signatures
(/root/reference/qumcmc/quantum_mcmc_routines.py:110-227), executed by libqumcmc_b200.

The whole hop loop -- draw (gamma, t), build U = M (P M)^{r-1}, measure once, get_energy,
test_accept, record -- runs on the device for all chains of the call; Python crosses the C-ABI once
per call, not once per hop.  Additive keyword arguments only: ``n_chains``, ``seed``, ``dtype``,
``device``, ``chain_id0``, ``hop0``.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _device, _lib
from .basic_utils import MCMCChain, MCMCState, get_random_state, int_to_binary
from .energy_models import IsingEnergyFunction
from .mixers import GenericMixer, Mixer

GammaType = Union[float, Tuple[float, float]]


def initial_state_indices(initial_state, num_spins: int, n_chains: int) -> np.ndarray:
    """uint64[n_chains] basis indices of the caller's initial state(s).  The reference asserts
    ``len(bitstring) == num_spins`` inside get_energy (qumcmc/energy_models.py:124-144); here an out-of-range index
    would reach the device unchecked, so it is rejected on the host."""
    n = int(num_spins)

    def one(s) -> int:
        if isinstance(s, str):
            if len(s) != n or any(ch not in "01" for ch in s):
                raise ValueError(f"initial state {s!r} is not a bitstring of {n} spins")
            return int(s, 2)
        v = int(s)
        if v < 0 or (n < 64 and v >= (1 << n)):
            raise ValueError(f"initial state index {v} outside [0, 2^{n})")
        return v

    if isinstance(initial_state, str) or np.isscalar(initial_state):
        return np.full(n_chains, one(initial_state), dtype=np.uint64)
    init = np.array([one(s) for s in initial_state], dtype=np.uint64)
    if len(init) != n_chains:
        raise ValueError("one initial state per chain expected")
    return init

# README-era call shape quantum_enhanced_mcmc(n_hops, model, temperature) (README.md:62-66):
# single-qubit X mixer and gamma ~ U(0.25, 0.6) (fossil in qumcmc/restricted_samplng.py:124-127).
DEFAULT_GAMMA = (0.25, 0.6)


def _resolve_dtype(dtype, n: int) -> int:
    if dtype is None:
        return _lib.DTYPE_F64 if n <= _lib.SMEM_MAX_SPINS_F64 else _lib.DTYPE_F32
    if dtype in (_lib.DTYPE_F32, "fp32", "float32", "complex64", np.float32, np.complex64):
        return _lib.DTYPE_F32
    if dtype in (_lib.DTYPE_F64, "fp64", "float64", "complex128", np.float64, np.complex128):
        return _lib.DTYPE_F64
    raise ValueError(f"unknown dtype {dtype!r}")


def _gamma_range(gamma: GammaType) -> Tuple[float, float]:
    if isinstance(gamma, (float, int, np.floating, np.integer)):
        return float(gamma), float(gamma)
    lo, hi = gamma
    return float(lo), float(hi)


class ChainBatch(Sequence):
    """Result of a multi-chain call: arrays plus a lazy ``MCMCChain`` view per chain."""

    def __init__(self, num_spins, initial, proposals, accepted, final_state, acc_count, hamming_hist, name,
                 n_hops=None, visit_hist=None):
        self.num_spins = num_spins
        self.initial = initial              # uint64[C]
        self.proposals = proposals          # uint64[C, H]; None when the run was not recorded (record=False)
        self.accepted = accepted            # uint8[C, H];  None when the run was not recorded
        self.final_state = final_state      # uint64[C]
        self.acc_count = acc_count          # int64[C]
        self.hamming_hist = hamming_hist    # int64[C, n+1]
        self.name = name
        self.n_hops = int(n_hops) if n_hops is not None else (0 if proposals is None else int(proposals.shape[1]))
        self.visit_hist = visit_hist        # int64[2^n] visits of the Markov chains (device histogram) or None

    @property
    def recorded(self) -> bool:
        return self.proposals is not None and self.accepted is not None

    def require_recorded(self, what: str) -> None:
        if not self.recorded:
            raise ValueError(f"{what} needs the per-hop record, but this batch was run with record=False "
                             "(only final states, acceptance counts and histograms were kept)")

    def __len__(self):
        return len(self.initial)

    def __getitem__(self, c) -> MCMCChain:
        if isinstance(c, slice):
            return [self[i] for i in range(*c.indices(len(self)))]
        self.require_recorded("ChainBatch[c] (an MCMCChain view)")
        return MCMCChain.from_arrays(self.num_spins, int(self.initial[c]), self.proposals[c], self.accepted[c],
                                     name=self.name)

    def acceptance_rate(self) -> float:
        total = len(self.initial) * self.n_hops
        return float(self.acc_count.sum()) / float(total) if total else float("nan")


def run_chains(model: IsingEnergyFunction, mixer: Mixer, gamma: GammaType, n_hops: int, n_chains: int,
               initial_states: Optional[np.ndarray], temperature: float, delta_time: float, seed: int,
               dtype=None, device: int = 0, chain_id0: int = 0, hop0: int = 0, name: str = "Q-MCMC",
               record: bool = True, visit_histogram: bool = False) -> ChainBatch:
    """Batched driver over qmc_run_chains (device buffers are torch tensors).  ``visit_histogram=True`` also returns
    the dense visit counts of the Markov chains (``ChainBatch.visit_hist``: int64[2^n], summed over the chains of the
    call, accumulated on the device -- MCMCChain.get_accepted_dict without materialising the chains)."""
    torch = _device._torch()
    dm = model.device_model(device)
    dm.set_mixer(mixer)
    n = model.num_spins
    dt = _resolve_dtype(dtype, n)
    g_lo, g_hi = _gamma_range(gamma)
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
        _lib.check(_lib.lib().qmc_run_chains(
            dm.handle, C_, H, ptr(s0), int(chain_id0), int(hop0) & 0xFFFFFFFF, float(temperature), g_lo, g_hi,
            float(delta_time), int(seed) & 0xFFFFFFFFFFFFFFFF, dt, ptr(props), ptr(acc), ptr(fin), ptr(cnt), ptr(hist),
            _device.current_stream_ptr(device)))
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


def quantum_enhanced_mcmc(
    n_hops: int,
    model: IsingEnergyFunction,
    mixer: Optional[Mixer] = None,
    gamma: Optional[GammaType] = None,
    initial_state: Optional[str] = None,
    temperature: float = 1,
    delta_time=0.8,
    verbose: bool = False,
    name: str = "Q-MCMC",
    *,
    n_chains: Optional[int] = None,
    seed: Optional[int] = None,
    dtype=None,
    device: int = 0,
    chain_id0: int = 0,
    hop0: int = 0,
    record: bool = True,
    visit_histogram: bool = False,
):
    """Quantum-enhanced MCMC.  Returns an ``MCMCChain`` (or a ``ChainBatch`` when ``n_chains`` is
    given).  ``mixer``/``gamma`` default to the README-era sampler when omitted."""
    n = model.num_spins
    if mixer is None:
        mixer = GenericMixer.OneBodyMixer(n)
    if gamma is None:
        gamma = DEFAULT_GAMMA
    if seed is None:  # stays under the control of np.random.seed, like the reference's global RNG use
        seed = int(np.random.randint(0, 2 ** 62))
    chain_name = name + ": " + str(mixer)
    single = n_chains is None
    C_ = 1 if single else int(n_chains)
    if initial_state is None:
        init = np.array([int(get_random_state(n), 2)], dtype=np.uint64) if single else None
    else:
        init = initial_state_indices(initial_state, n, C_)
    if verbose and init is not None:
        e0 = model.get_energy(int_to_binary(int(init[0]), n))
        print("starting with: ", int_to_binary(int(init[0]), n), "with energy:", e0)
    batch = run_chains(model, mixer, gamma, n_hops, C_, init, temperature, delta_time, seed, dtype=dtype,
                       device=device, chain_id0=chain_id0, hop0=hop0, name=chain_name,
                       record=record or single, visit_histogram=visit_histogram)
    return batch[0] if single else batch


def run_qmcmc_quantum_ckt(state_s: str, model: IsingEnergyFunction, mixer: Mixer, gammafunc: callable,
                          alpha: float, num_spins: int, delta_time: float, *, dtype=None, device: int = 0) -> str:
    """One proposal s -> s' (quantum_mcmc_routines.py:110-156).  gamma and the evolution time come
    from the caller's gammafunc / the global numpy RNG exactly as in the reference; the single
    measurement shot uses a numpy uniform in place of qulacs' internal RNG."""
    torch = _device._torch()
    gamma = float(gammafunc())
    time = np.random.choice(list(range(2, 12)))
    r = int(np.floor(time / delta_time))
    u = float(np.random.random_sample())
    if float(alpha) != float(model.alpha):
        model = _AlphaOverride(model, alpha)
    dm = model.device_model(device)
    dm.set_mixer(mixer)
    groups, probs = mixer.term_groups()
    grp = 0 if probs is None else int(np.random.choice(len(groups), p=np.asarray(probs) / np.sum(probs)))
    dev = _device.require_cuda(device)
    s_in = torch.tensor([int(state_s, 2)], dtype=torch.int64, device=dev)
    g = torch.tensor([gamma], dtype=torch.float64, device=dev)
    rr = torch.tensor([r], dtype=torch.int32, device=dev)
    uu = torch.tensor([u], dtype=torch.float64, device=dev)
    gg = torch.tensor([grp], dtype=torch.int32, device=dev)
    out = torch.empty(1, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().qmc_propose(dm.handle, 1, s_in.data_ptr(), g.data_ptr(), rr.data_ptr(), uu.data_ptr(),
                                      gg.data_ptr(), float(delta_time), _resolve_dtype(dtype, num_spins),
                                      out.data_ptr(), _device.current_stream_ptr(device)))
    return int_to_binary(int(out.cpu().numpy().view(np.uint64)[0]), num_spins)


class _AlphaOverride(IsingEnergyFunction):
    def __init__(self, model: IsingEnergyFunction, alpha: float):
        self.__dict__.update(model.__dict__)
        self.alpha = float(alpha)


def transition_probabilities(model: IsingEnergyFunction, mixer: Mixer, state: Union[str, int], gamma: float, r: int,
                             delta_time: float = 0.8, group: int = 0, dtype=None, device: int = 0,
                             return_amplitudes: bool = False):
    """|<s'|U|s>|^2 for all s' (test hook over qmc_transition_probs)."""
    torch = _device._torch()
    n = model.num_spins
    dm = model.device_model(device)
    dm.set_mixer(mixer)
    dt = _resolve_dtype(dtype, n)
    dev = _device.require_cuda(device)
    real = torch.float64 if dt == _lib.DTYPE_F64 else torch.float32
    probs = torch.empty(1 << n, dtype=real, device=dev)
    amps = torch.empty((1 << n, 2), dtype=real, device=dev) if return_amplitudes else None
    s = int(state, 2) if isinstance(state, str) else int(state)
    _lib.check(_lib.lib().qmc_transition_probs(dm.handle, s, float(gamma), int(r), float(delta_time), int(group), dt,
                                               probs.data_ptr(), None if amps is None else amps.data_ptr(),
                                               _device.current_stream_ptr(device)))
    torch.cuda.synchronize(device)
    if return_amplitudes:
        a = amps.cpu().numpy()
        return probs.cpu().numpy(), a[:, 0] + 1j * a[:, 1]
    return probs.cpu().numpy()
