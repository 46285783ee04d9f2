"""``IsingEnergyFunction``, ``Exact_Sampling`` and ``random_ising_model`` with the reference's
names, argument order and defaults (/root/reference/qumcmc/energy_models.py), computing through
libqumcmc_b200: batched energies (qmc_energies) and the 2^n enumeration + grid-wide logsumexp
(qmc_exact_sampling).  No CPU fallback: without the CUDA extension these raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Union

import numpy as np

from . import _device, _lib
from .basic_utils import int_to_binary
from .prob_dist import DiscreteProbabilityDistribution, value_sorted_dict  # noqa: F401


def _state_to_index(state, n: int) -> int:
    if isinstance(state, str):
        if len(state) != n:
            raise ValueError(f"bitstring of length {len(state)} for a {n}-spin model")
        return int(state, 2)
    arr = np.asarray(state)
    if arr.shape != (n,) or not np.all(np.abs(arr) == 1):
        raise ValueError("array states must be bipolar (+1/-1) vectors of length num_spins")
    idx = 0
    for v in arr:
        idx = (idx << 1) | (1 if v > 0 else 0)
    return idx


class IsingEnergyFunction:
    """E(s) = 1/2 s^T J s + h.s over s in {-1,+1}^n; bitstring '1' <-> +1, leftmost char = spin 0."""

    def __init__(self, J: np.array, h: np.array, name: str = None, circuit_J=None, circuit_h=None) -> None:
        self.J = J
        self.h = h
        self.circuit_J = circuit_J if circuit_J is not None else J
        # The reference installs circuit_J when circuit_h is given (energy_models.py:36, quirk Q4).
        # Reproduced: a drop-in must evolve under the same decoy Hamiltonian the reference would.
        self.circuit_h = circuit_J if circuit_h is not None else h
        self.num_spins = len(h)
        if np.allclose(J, 0):
            self.alpha = 1.0
        else:
            n = self.num_spins
            self.alpha = np.sqrt(n) / np.sqrt(
                sum([J[i][j] ** 2 for i in range(n) for j in range(i)]) + sum([h[j] ** 2 for j in range(n)]))
        self.name = "JK_random" if name is None else name

    def __repr__(self):
        return f"Ising Model:{self.name}| num-spins: {self.num_spins}"

    def __eq__(self, other: object) -> bool:
        return (self.J == other.J).all() and (self.h == other.h).all()

    __hash__ = None

    @property
    def get_J(self):
        return self.J

    @property
    def get_h(self):
        return self.h

    # ---- device handle
    def _circuit_arrays(self):
        cJ = np.asarray(self.circuit_J, dtype=np.float64)
        ch = np.asarray(self.circuit_h, dtype=np.float64)
        if ch.shape != (self.num_spins,):
            raise ValueError(
                "circuit_h has shape %s: the reference assigns circuit_J to circuit_h when circuit_h is passed "
                "(energy_models.py:36) and then fails inside fn_qckt_problem_half; pass circuit_J only" % (ch.shape,))
        return cJ, ch

    def device_model(self, device: int = 0) -> "_device.DeviceModel":
        cJ, ch = self._circuit_arrays()
        return _device.device_model(np.asarray(self.J, dtype=np.float64), np.asarray(self.h, dtype=np.float64),
                                    cJ, ch, float(self.alpha), device)

    def model_summary(self, plot=False):
        print("=============================================")
        print("            MODEL : " + str(self.name))
        print("=============================================")
        print("Non-zero Interactions (J) : " + str(int(np.count_nonzero(self.J) / 2)) + " / "
              + str(int(0.5 * self.num_spins * (self.num_spins - 1))))
        print("Non-zero Bias (h) : " + str(int(np.count_nonzero(self.h))) + " / " + str(self.num_spins))
        print("---------------------------------------------")
        print("Average Interaction Strength <|J|> : ", np.mean(np.abs(self.J)))
        print("Average Bias Strength <|h|>: ", np.mean(np.abs(self.h)))
        print("alpha : ", self.alpha)
        print("---------------------------------------------")
        # plot=True in the reference draws a seaborn heat-map and, as a side effect, writes h onto
        # the diagonal of J (quirk Q6). Plotting is out of scope here and J is left untouched.

    def get_hamiltonian(self):
        raise NotImplementedError("qulacs Observable export is not part of the B200 path")

    def get_energies(self, states, device: int = 0) -> np.ndarray:
        """Batched get_energy: iterable of bitstrings / +-1 arrays, or an integer index array."""
        if isinstance(states, np.ndarray) and states.dtype.kind in "iu" and states.ndim == 1:
            idx = np.ascontiguousarray(states, dtype=np.uint64)
        else:
            idx = np.array([_state_to_index(s, self.num_spins) for s in states], dtype=np.uint64)
        out = np.empty(len(idx), dtype=np.float64)
        dm = self.device_model(device)
        _lib.check(_lib.lib().qmc_energies_host(dm.handle, idx.ctypes.data, len(idx), out.ctypes.data))
        return out

    def get_energy(self, state: Union[str, np.array]) -> float:
        """Energy of one configuration (energy_models.py:124-144)."""
        return float(self.get_energies([state])[0])

    def get_boltzmann_factor(self, state: Union[str, np.array], beta: float = 1.0) -> float:
        return np.exp(-1 * beta * self.get_energy(state))

    def _update_J(self, new_param: float, index: Union[tuple, List]):
        assert len(index) == 2
        self.J[index[0], index[1]] = new_param
        self.J[index[1], index[0]] = new_param

    def _update_h(self, new_param: float, index: int):
        self.h[index] = new_param


class Exact_Sampling(IsingEnergyFunction):
    """Exhaustive Boltzmann distribution of a model at inverse temperature beta."""

    # above this many spins the {bitstring: p} dict is not materialised (use .probs / .logZ)
    DICT_MAX_SPINS = 22

    def __init__(self, model: IsingEnergyFunction, beta: float = 1.0, verbose=False, device: int = 0) -> None:
        super().__init__(model.get_J, model.get_h, model.name)  # circuit_* dropped like the reference (Q5)
        self.beta = beta
        self.exact_sampling_status = False
        self._device = device
        self._pd: Optional[DiscreteProbabilityDistribution] = None
        self.probs: Optional[np.ndarray] = None       # p(idx), index order
        self.energies: Optional[np.ndarray] = None    # E(idx), index order
        self.logZ: Optional[float] = None
        self.run_exact_sampling(self.beta, verbose=verbose)

    # ---- device enumeration
    def _enumerate(self, beta: float, want_arrays: bool = True):
        dm = self.device_model(self._device)
        N = 1 << self.num_spins
        logZ = np.empty(1, dtype=np.float64)
        E = np.empty(N, dtype=np.float64) if want_arrays else None
        P = np.empty(N, dtype=np.float64) if want_arrays else None
        _lib.check(_lib.lib().qmc_exact_sampling_host(dm.handle, float(beta), logZ.ctypes.data,
                                                      None if E is None else E.ctypes.data,
                                                      None if P is None else P.ctypes.data))
        return float(logZ[0]), E, P

    def _dict_from_probs(self, probs: np.ndarray, sort_desc: bool) -> dict:
        n = self.num_spins
        if n > self.DICT_MAX_SPINS:
            raise MemoryError(f"a {{bitstring: p}} dict over 2^{n} states is not materialised; use .probs/.logZ")
        order = np.argsort(-probs, kind="stable") if sort_desc else np.arange(len(probs))
        vals = probs[order].tolist()
        return {int_to_binary(int(k), n): v for k, v in zip(order.tolist(), vals)}

    @property
    def boltzmann_pd(self) -> DiscreteProbabilityDistribution:
        if not self.exact_sampling_status:
            raise RuntimeError("Please Run Exact Sampling at any specified temperature first")
        if self._pd is None:
            self._pd = DiscreteProbabilityDistribution(self._dict_from_probs(self.probs, sort_desc=True))
        return self._pd

    @boltzmann_pd.setter
    def boltzmann_pd(self, value):
        self._pd = value

    def get_boltzmann_distribution(self, beta: float = 1.0, sorted: bool = False, save_distribution: bool = False,
                                   return_dist: bool = True, plot_dist: bool = False, verbose: bool = False) -> dict:
        logZ, E, P = self._enumerate(beta, want_arrays=True)
        if save_distribution:
            self.probs, self.energies, self.logZ = P, E, logZ
            self._pd = None
        if return_dist:
            return self._dict_from_probs(P, sort_desc=sorted)

    def run_exact_sampling(self, beta: float, verbose: bool = False) -> None:
        self.exact_sampling_status = True
        self.beta = beta
        if verbose:
            print("Running Exact Sampling | beta : ", beta)
        if self.num_spins > 26:  # 2^n doubles x2 on the host stops being sensible; keep log Z only
            self.logZ, self.energies, self.probs = self._enumerate(beta, want_arrays=False)
        else:
            self.get_boltzmann_distribution(beta=beta, save_distribution=True, return_dist=False, verbose=verbose)

    def sampling_summary(self, plot_dist: bool = False, show_threshold=0.01, ylabel="Prob"):
        if not self.exact_sampling_status:
            raise RuntimeError("Please Run Exact Sampling at any specified temperature first")
        print("=============================================")
        print("     MODEL : " + str(self.name) + " |  beta : " + str(self.beta))
        print("=============================================")
        print("Num Most Probable States : " + str(int(np.count_nonzero(self.probs > show_threshold))))
        print("Entropy : " + str(self.get_entropy()))
        print("---------------------------------------------")

    def get_observable_expectation(self, observable) -> float:
        n = self.num_spins
        return sum(p * observable(int_to_binary(k, n)) for k, p in enumerate(self.probs.tolist()))

    def get_entropy(self):
        """log2 entropy over p > 1e-5 in descending order; None when every p > 1e-5
        (energy_models.py:319-326, quirk Q7)."""
        entropy = 0
        for val in np.sort(self.probs)[::-1]:
            if val > 0.00001:
                entropy += -1 * val * np.log2(val)
            else:
                return entropy

    def _check_beta(self, beta):
        if beta is None:
            return self.beta
        if isinstance(beta, float) and beta != self.beta:
            raise ValueError("Current beta is different from model beta. Please 'run_exact_sampling' with "
                             "appropriate beta value ")
        return beta

    def get_kldiv(self, q: dict, beta: Union[float, None] = None) -> float:
        """D_kl(boltzmann || q) in log2 (energy_models.py:328-375)."""
        self._check_beta(beta)
        n = self.num_spins
        assert np.sum(list(q.values())) == 1, " given distribution is not normalised "
        if len(q) != (1 << n) or any(len(k) != n for k in q):
            raise ValueError(" given distribution is not defined over all possible configurations ")
        qd = DiscreteProbabilityDistribution(q)
        qd.normalise()
        pv = list(DiscreteProbabilityDistribution(self.boltzmann_pd).index_sorted_dict().values())
        qv = list(qd.index_sorted_dict().values())
        return sum(pv[i] * np.log2(pv[i] / qv[i]) for i in range(len(pv)) if pv[i] != 0)

    def get_jsdiv(self, q, beta: Union[float, None] = None) -> float:
        self._check_beta(beta)
        bltz = self.boltzmann_pd
        assert np.sum(list(q.values())) == 1, " given distribution is not normalised "
        m = {key: 0.5 * (bltz[key] + q[key]) for key in bltz.keys()}
        return 0.5 * self.get_kldiv(bltz, m) + 0.5 * self.get_kldiv(q, m)


def random_ising_model(n_spins: int, seed: int, print_model: bool = False):
    """J ~ U(-1,1) symmetrised, zero diagonal, 3 decimals; h = round(0.4 randn, 2)
    (energy_models.py:419-446); numpy's legacy global stream so the models match bit for bit."""
    np.random.seed(seed)
    J = np.random.uniform(low=-1, high=1, size=(n_spins, n_spins))
    J = 0.5 * (J + J.transpose())
    J = np.round(J - np.diag(np.diag(J)), decimals=3)
    h = np.round(0.4 * np.random.randn(n_spins), decimals=2)
    model = IsingEnergyFunction(J, h, name="param_model")
    if print_model:
        model.model_summary()
    return model
