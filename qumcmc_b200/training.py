"""Contrastive-divergence training of an Ising model against a data distribution -- the reference's principal caller of
the proposal path (/root/reference/qumcmc/training.py:104-364).

Per epoch the reference runs one chain from a data sample and then walks every (i, j) in Python, re-scanning the
chain's accepted states for each pair (O(n^2 H) string operations).  Here the chain(s) come from the device samplers
and the model expectations <s_i>, <s_i s_j> are one ``qmc_state_moments`` launch over the recorded states (exact integer
counts), so the whole gradient is two small matrices; the data expectations are cached once.

Semantics kept: the model expectation runs over ``mcmc_chain.accepted_states`` (initial state + accepted proposals, each
counted once: training.py:44-56), the update is ``theta -= lr * (<.>_data - <.>_model)`` (:236-262), the 'all' and
'random' update strategies (:187-262), the two KL bookkeeping modes and the training_history layout (:285-364)."""
from __future__ import annotations

import pickle as pkl
import random
from copy import deepcopy
from typing import List, Optional, Tuple, Union

import numpy as np

from . import _device, _lib
from .basic_utils import MCMCChain
from .energy_models import Exact_Sampling, IsingEnergyFunction
from .mcmc_sampler_base import ClassicalMCMCSampler, QuantumMCMCSampler
from .prob_dist import DiscreteProbabilityDistribution, kl_divergence
from .quantum_mcmc_routines import ChainBatch


def correlation_spins(state: str, indices: Union[tuple, List]) -> float:
    """Product of the +-1 spins of ``state`` at ``indices`` (training.py:59-67)."""
    assert len(indices) <= len(state)
    prod = 1.0
    for j in indices:
        prod *= 2.0 * float(state[j]) - 1.0
    return prod


def get_observable_expectation(observable: callable, mcmc_chain: MCMCChain, skip_init: int = 100) -> float:
    """Mean of ``observable`` over ``mcmc_chain.accepted_states`` (training.py:44-56; ``skip_init`` is unused there too)."""
    return float(np.mean([observable(s) for s in mcmc_chain.accepted_states], dtype=float))


def spin_moments(model: IsingEnergyFunction, states: np.ndarray, mask: Optional[np.ndarray] = None,
                 device: int = 0) -> Tuple[int, np.ndarray, np.ndarray]:
    """``(N, <s_i>[n], <s_i s_j>[n, n])`` over ``states`` (uint64 indices; ``mask`` selects entries) on the device."""
    torch = _device._torch()
    dev = _device.require_cuda(device)
    dm = model.device_model(device)
    n = model.num_spins
    st = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1)
    d_st = torch.from_numpy(st.view(np.int64)).to(dev)
    d_mask = None
    if mask is not None:
        d_mask = torch.from_numpy(np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1)).to(dev)
        if d_mask.numel() != d_st.numel():
            raise ValueError("mask and states differ in length")
    used = torch.zeros(1, dtype=torch.int64, device=dev)
    ones = torch.zeros(n, dtype=torch.int64, device=dev)
    agree = torch.zeros((n, n), dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().qmc_state_moments(dm.handle, int(st.size), d_st.data_ptr(),
                                            None if d_mask is None else d_mask.data_ptr(), used.data_ptr(),
                                            ones.data_ptr(), agree.data_ptr(), _device.current_stream_ptr(device)))
    torch.cuda.synchronize(device)
    N = int(used.item())
    if N == 0:
        raise ValueError("no states selected")
    return N, (2.0 * ones.cpu().numpy() - N) / N, (2.0 * agree.cpu().numpy() - N) / N


def chain_moments(model: IsingEnergyFunction, chains: Union[MCMCChain, ChainBatch], device: int = 0):
    """Moments over the accepted states of one chain or of every chain of a batch (pooled)."""
    if isinstance(chains, ChainBatch):
        chains.require_recorded("chain_moments")
        states = np.concatenate([chains.initial.reshape(-1), chains.proposals.reshape(-1)])
        mask = np.concatenate([np.ones(len(chains.initial), dtype=np.uint8), chains.accepted.reshape(-1)])
    else:
        states, mask = chains.proposal_indices, chains.accepted_mask     # entry 0 = the initial state (accepted)
    return spin_moments(model, states, mask, device)


def data_moments(data_dist: DiscreteProbabilityDistribution) -> Tuple[np.ndarray, np.ndarray]:
    """``(<s_i>, <s_i s_j>)`` under the data distribution (DiscreteProbabilityDistribution.get_observable_expectation)."""
    keys = list(data_dist.keys())
    p = np.array([data_dist[k] for k in keys], dtype=np.float64)
    S = np.array([[2.0 * float(c) - 1.0 for c in k] for k in keys], dtype=np.float64)
    return p @ S, (S * p[:, None]).T @ S


def cd_J(index, data_distribution: DiscreteProbabilityDistribution, mcmc_chain: MCMCChain) -> float:
    """<s_i s_j>_data - <s_i s_j>_chain (training.py:70-82)."""
    assert len(index) == 2
    obs = lambda s: correlation_spins(s, [index[0], index[1]])
    return data_distribution.get_observable_expectation(obs) - get_observable_expectation(obs, mcmc_chain)


def cd_h(index: int, data_distribution: DiscreteProbabilityDistribution, mcmc_chain: MCMCChain) -> float:
    """<s_i>_data - <s_i>_chain (training.py:85-99)."""
    assert isinstance(index, int)
    obs = lambda s: correlation_spins(s, [index])
    return data_distribution.get_observable_expectation(obs) - get_observable_expectation(obs, mcmc_chain)


class CDTraining:
    """model: initial (J, h) at inverse temperature ``beta``; data_dist: the distribution to learn (training.py:104-136)."""

    def __init__(self, model: IsingEnergyFunction, beta: float, data_dist: DiscreteProbabilityDistribution,
                 name: str = "train", pickle_loc: Union[None, str] = None, *, device: int = 0) -> None:
        self.model = deepcopy(model)
        self.model_beta = beta
        self.data_distribution = data_dist
        self.training_history = {"specifications": [], "kl_div": [], "max-min-gradient": []}
        self.kl_div = []
        n = self.model.num_spins
        self.list_pair_of_indices = [[i, j] for i in range(1, n) for j in range(i, n) if j != i]
        self.name = name
        self.pickle_fileloc = pickle_loc
        self.device = device
        self._data_moments = None
        self.mcmc_chain = None

    # the per-parameter gradients of the reference, kept for callers that use them directly (:138-156)
    def cd_J(self, index, mcmc_chain: MCMCChain) -> float:
        return cd_J(index, self.data_distribution, mcmc_chain)

    def cd_h(self, index: int, mcmc_chain: MCMCChain) -> float:
        return cd_h(index, self.data_distribution, mcmc_chain)

    def gradients(self, chains: Union[MCMCChain, ChainBatch]) -> Tuple[np.ndarray, np.ndarray]:
        """``(grad_h[n], grad_J[n, n])``: data minus model expectations for every parameter at once."""
        if self._data_moments is None:
            self._data_moments = data_moments(self.data_distribution)
        _, m1, m2 = chain_moments(self.model, chains, self.device)
        return self._data_moments[0] - m1, self._data_moments[1] - m2

    def _train_on_mcmc_chain(self, mcmc_sampler: Union[ClassicalMCMCSampler, QuantumMCMCSampler], lr: float = 0.01,
                             mcmc_steps: int = 1000, update_strategy=("all", []), save_gradient_info=False):
        initial_state = self.data_distribution.get_sample(1)[0]  # a state drawn from the data (:171-173)
        mcmc_sampler.n_hops = mcmc_steps
        mcmc_sampler.initial_state = initial_state
        mcmc_sampler.model = self.model
        self.mcmc_chain = mcmc_sampler.run()
        grad_h, grad_J = self.gradients(self.mcmc_chain)
        n = self.model.num_spins
        if update_strategy[0] == "random":
            opts = update_strategy[1]
            assert opts["num_random_bias"] <= n, f"update_strategy[1]['num_random_bias'] should be <= num_spins (= {n})"
            assert opts["num_random_interactions"] <= len(self.list_pair_of_indices), \
                "update_strategy[1]['num_random_interactions'] should be <= len(self.list_pair_of_indices)"
            picked = random.sample(range(0, n), opts["num_random_bias"])
            pairs = [[i, random.choice(list(range(0, i)) + list(range(i + 1, n)))] for i in picked]
            gj = [grad_J[i, j] for i, j in pairs]
            gh = [grad_h[i] for i in picked]
            for (i, j), g in zip(pairs, gj):
                self.model._update_J(self.model.J[i, j] - lr * g, [i, j])
            for i, g in zip(picked, gh):
                self.model._update_h(self.model.h[i] - lr * g, i)
        elif update_strategy[0] == "all":
            iu = np.tril_indices(n, -1)
            gj, gh = grad_J[iu], grad_h
            J = np.array(self.model.J, dtype=np.float64, copy=True)
            J[iu] -= lr * gj
            J[(iu[1], iu[0])] = J[iu]                       # _update_J writes both triangles (energy_models.py:146-152)
            for i, j, v in zip(iu[0], iu[1], J[iu]):
                self.model._update_J(v, [int(i), int(j)])
            for i in range(n):
                self.model._update_h(self.model.h[i] - lr * gh[i], i)
        else:
            raise ValueError(f"unknown update strategy {update_strategy[0]!r}")
        if save_gradient_info:
            aj, ah = np.abs(np.asarray(gj)), np.abs(np.asarray(gh))
            self.training_history["max-min-gradient"].append(
                [(float(aj.max()) if aj.size else 0, float(aj.min()) if aj.size else 0),
                 (float(ah.max()) if ah.size else 0, float(ah.min()) if ah.size else 0)])

    def train(self, mcmc_sampler: Union[ClassicalMCMCSampler, QuantumMCMCSampler], mcmc_steps: int = 500,
              lr: float = 0.01, epochs: int = 10, update_strategy=("all", []),
              save_training_data={"kl_div": [True, "exact-sampling"], "max-min-gradient": True}, verbose=True,
              save_picle: bool = False):
        """``update_strategy``: ``['all', []]`` updates every parameter; ``['random', {'num_random_bias': b,
        'num_random_interactions': k}]`` updates b random biases and b random couplings (training.py:285-364)."""
        self.training_history["specifications"].append(
            {"lr": lr, "sampler": mcmc_sampler, "epochs": epochs, "update_strategy": update_strategy})
        for epoch in range(epochs):
            self._train_on_mcmc_chain(mcmc_sampler=mcmc_sampler, lr=lr, mcmc_steps=mcmc_steps,
                                      update_strategy=update_strategy,
                                      save_gradient_info=save_training_data["max-min-gradient"])
            if save_training_data["kl_div"][0]:
                how = save_training_data["kl_div"][1]
                if how == "last-mcmc-chain":
                    chain = self.mcmc_chain[0] if isinstance(self.mcmc_chain, ChainBatch) else self.mcmc_chain
                    self.training_history["kl_div"].append(
                        kl_divergence(self.data_distribution, chain.get_accepted_dict(normalize=True)))
                if how == "exact-sampling":
                    exact = Exact_Sampling(self.model, self.model_beta)
                    self.training_history["kl_div"].append(kl_divergence(self.data_distribution, exact.boltzmann_pd))
                if verbose:
                    print(f"epoch {epoch + 1}/{epochs}  sampler {mcmc_sampler.name}  kl div {self.training_history['kl_div'][-1]:.6f}")
        if save_picle:
            if not self.pickle_fileloc:
                raise ValueError(f"{self.pickle_fileloc} must be .pkl file")
            with open(self.pickle_fileloc, "wb") as f:
                pkl.dump(self, f)
