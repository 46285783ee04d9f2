"""Trajectory statistics of MCMC chains (the reference's ``qumcmc.trajectory_processing`` functions that consume an
``MCMCChain``: /root/reference/qumcmc/trajectory_processing.py:114-153,183-210,213-321), computed on the device by
``qmc_trajectory_stats`` -- one pass over the recorded proposals per chain, every chain of a batch in one launch
(SURVEY section 8 row f3).  The running KL divergence needs no 2^n histogram per chain and no O(2^n) work per step.

``batch_trajectory_statistics`` is the batched entry point (a ``ChainBatch`` or arrays); the reference-named functions
take one ``MCMCChain`` and return what the reference returns."""
from __future__ import annotations

from typing import Dict, Optional, Union

import numpy as np

from . import _device, _lib
from .basic_utils import MCMCChain
from .energy_models import Exact_Sampling, IsingEnergyFunction

_ALL = ("markov_chain", "energy", "magnetisation", "hamming", "kldiv")


def _probs_vector(actual_boltz_distn, n: int) -> np.ndarray:
    """Dense probabilities in index order from a ``{bitstring: p}`` mapping (missing states: 0) or an array."""
    if isinstance(actual_boltz_distn, np.ndarray):
        p = np.ascontiguousarray(actual_boltz_distn, dtype=np.float64)
        if p.shape != (1 << n,):
            raise ValueError(f"expected {1 << n} probabilities")
        return p
    p = np.zeros(1 << n, dtype=np.float64)
    for k, v in actual_boltz_distn.items():
        p[int(k, 2)] = v
    return p


def batch_trajectory_statistics(model: IsingEnergyFunction, initial: np.ndarray, proposals: np.ndarray,
                                accepted: np.ndarray, boltzmann_probs: Optional[np.ndarray] = None,
                                to_observe=_ALL, device: int = 0) -> Dict[str, np.ndarray]:
    """Statistics of ``C`` recorded chains (``initial[C]``, ``proposals[C, H]``, ``accepted[C, H]``):
    ``markov_chain[C, H+1]``, ``energy[C, H]`` (E(proposal) - E(current)), ``magnetisation[C, H]`` (running mean),
    ``hamming[C, n+1, 2]`` (accepted, rejected) and ``kldiv[C, H+1]`` (needs ``boltzmann_probs[2^n]``)."""
    torch = _device._torch()
    dev = _device.require_cuda(device)
    dm = model.device_model(device)
    n = model.num_spins
    proposals = np.ascontiguousarray(proposals, dtype=np.uint64)
    accepted = np.ascontiguousarray(accepted, dtype=np.uint8)
    C_, H = proposals.shape
    want = set(to_observe)
    unknown = want - set(_ALL)
    if unknown:
        raise ValueError(f"unknown statistics {sorted(unknown)}")
    if "kldiv" in want and boltzmann_probs is None:
        raise ValueError("kldiv needs the Boltzmann probabilities")
    up = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
    d_s0, d_p, d_a = up(np.ascontiguousarray(initial, dtype=np.uint64)), up(proposals), up(accepted)
    d_probs = up(np.ascontiguousarray(boltzmann_probs, dtype=np.float64)) if "kldiv" in want else None
    out = {
        "markov_chain": torch.empty((C_, H + 1), dtype=torch.int64, device=dev) if "markov_chain" in want else None,
        "energy": torch.empty((C_, H), dtype=torch.float64, device=dev) if "energy" in want else None,
        "magnetisation": torch.empty((C_, H), dtype=torch.float64, device=dev) if "magnetisation" in want else None,
        "hamming": torch.empty((C_, n + 1, 2), dtype=torch.int64, device=dev) if "hamming" in want else None,
        "kldiv": torch.empty((C_, H + 1), dtype=torch.float64, device=dev) if "kldiv" in want else None,
    }
    ptr = lambda t: None if t is None else t.data_ptr()
    _lib.check(_lib.lib().qmc_trajectory_stats(dm.handle, C_, H, ptr(d_s0), ptr(d_p), ptr(d_a), ptr(d_probs),
                                               ptr(out["markov_chain"]), ptr(out["energy"]), ptr(out["magnetisation"]),
                                               ptr(out["hamming"]), ptr(out["kldiv"]),
                                               _device.current_stream_ptr(device)))
    torch.cuda.synchronize(device)
    res = {k: v.cpu().numpy() for k, v in out.items() if v is not None}
    if "markov_chain" in res:
        res["markov_chain"] = res["markov_chain"].view(np.uint64)
    return res


def _one(chain: MCMCChain, model, probs, to_observe, device=0):
    idx, acc = chain.proposal_indices, chain.accepted_mask        # entry 0 = the initial state (accepted)
    return batch_trajectory_statistics(model, idx[:1].astype(np.uint64), idx[None, 1:], acc[None, 1:].astype(np.uint8),
                                       probs, to_observe, device)


class _WidthOnly(IsingEnergyFunction):
    """Stand-in model for statistics that need no couplings (KL, magnetisation): n zero-coupled spins."""

    def __init__(self, n: int):
        super().__init__(np.zeros((n, n)), np.zeros(n))


def calculate_running_kl_divergence(actual_boltz_distn, mcmc_chain: MCMCChain, skip_steps: int = 1, verbose=False,
                                    *, device: int = 0) -> list:
    """KL(boltzmann || empirical distribution of the Markov chain) after every entry of ``mcmc_chain.markov_chain``
    (trajectory_processing.py:114-153; ``skip_steps`` is ignored there too)."""
    if skip_steps > 1:
        print("skip_steps currently not available and defaults to 1")
    n = mcmc_chain.num_spins
    res = _one(mcmc_chain, _WidthOnly(n), _probs_vector(actual_boltz_distn, n), ("kldiv",), device)
    return list(res["kldiv"][0])


def calculate_runnning_magnetisation(mcmc_chain: MCMCChain, skip_steps: int = 1, verbose: bool = False, *,
                                     device: int = 0) -> list:
    """Running mean magnetisation of the Markov chain (trajectory_processing.py:183-210), entries ``1, 1+skip, ...``."""
    n = mcmc_chain.num_spins
    res = _one(mcmc_chain, _WidthOnly(n), None, ("magnetisation",), device)
    return list(res["magnetisation"][0][::max(1, int(skip_steps))])


def get_trajectory_statistics(mcmc_chain: MCMCChain, model: Union[IsingEnergyFunction, Exact_Sampling],
                              to_observe: set = frozenset({"acceptance_prob", "kldiv", "hamming", "energy",
                                                           "magnetisation"}),
                              verbose: bool = False, *, device: int = 0) -> dict:
    """``{'acceptance_prob', 'energy', 'hamming', 'kldiv'}`` of one chain, shaped like the reference's
    (trajectory_processing.py:213-321)."""
    want = []
    if "energy" in to_observe:
        want.append("energy")
    if "hamming" in to_observe:
        want.append("hamming")
    probs = None
    if "kldiv" in to_observe:
        want.append("kldiv")
        probs = _probs_vector(model.boltzmann_pd, model.num_spins)
    res = _one(mcmc_chain, model, probs, tuple(want), device) if want else {}
    out = {}
    if "acceptance_prob" in to_observe:  # 1 / dwell time, host side: variable length (:266-277)
        idx, acc = mcmc_chain.proposal_indices, mcmc_chain.accepted_mask.astype(bool)
        cur, counter, stat = int(idx[0]), 1, []
        for st_idx, ok in zip(idx.tolist(), acc.tolist()):
            if ok and st_idx != cur:
                stat.append(1 / counter)
                cur, counter = st_idx, 1
            else:
                counter += 1
        stat.append(1 / counter)
        out["acceptance_prob"] = np.array(stat)
    if "energy" in to_observe:
        out["energy"] = res["energy"][0]
    if "hamming" in to_observe:
        ham = res["hamming"][0]
        out["hamming"] = {d: {"accepted": int(ham[d, 0]), "rejected": int(ham[d, 1]), "total": int(ham[d].sum())}
                          for d in range(model.num_spins + 1)}
    if "kldiv" in to_observe:
        out["kldiv"] = res["kldiv"][0]
    return out
