"""numpy/scipy layer of the oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* dense proposal matrix Q(s'|s) = |<s'|M (P M)^{r-1}|s>|^2 from the C gate-wise simulator (n <= 10)
* closed-form stationary chain statistics (acceptance rate, proposal Hamming distribution) from Q,
  which is how the reference's recorded chains are compared without running long chains
* restatement of the exact-unitary path (quantum_mcmc_routines_exact.py) with its quirks as flags
* the reference's distribution metrics (prob_dist.py, energy_models.py:319-326)
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from . import alpha as _alpha
from . import energies as _energies
from . import evolve, trotter_steps

EPS = 1e-8  # prob_dist.py:8


# --------------------------------------------------------------------------- proposal matrices
def dense_unitary(circJ, circh, alpha_, gamma, delta_time, r, masks, weights, gatewise=True) -> np.ndarray:
    """U[s', s] for the proposal circuit (columns = evolve(|s>))."""
    n = len(circh)
    U = np.empty((1 << n, 1 << n), dtype=np.complex128)
    for s in range(1 << n):
        U[:, s] = evolve(circJ, circh, alpha_, gamma, delta_time, r, masks, weights, s, gatewise)
    return U


def proposal_matrix(J, h, masks, weights, gamma, delta_time=0.8, times=range(2, 12), alpha_=None,
                    circJ=None, circh=None, n_gamma_nodes: int = 12) -> np.ndarray:
    """Q[s', s] averaged over the evolution time t ~ U{2..11} (quantum_mcmc_routines.py:125) and,
    if gamma is a (lo, hi) tuple, over gamma ~ U(lo, hi) (:186) by Gauss-Legendre quadrature."""
    circJ = J if circJ is None else circJ
    circh = h if circh is None else circh
    a = _alpha(J, h) if alpha_ is None else alpha_
    if np.isscalar(gamma):
        nodes, wts = [float(gamma)], [1.0]
    else:
        x, w = np.polynomial.legendre.leggauss(n_gamma_nodes)
        lo, hi = gamma
        nodes = 0.5 * (hi - lo) * x + 0.5 * (hi + lo)
        wts = 0.5 * w
    n = len(h)
    Q = np.zeros((1 << n, 1 << n))
    rs = [trotter_steps(t, delta_time) for t in times]
    for g, wg in zip(nodes, wts):
        cache = {}
        for r in rs:
            if r not in cache:
                U = dense_unitary(circJ, circh, a, g, delta_time, r, masks, weights, gatewise=False)
                cache[r] = np.abs(U) ** 2
            Q += wg * cache[r] / len(rs)
    return Q


def stationary_chain_stats(J, h, Q: np.ndarray, beta: float) -> Tuple[float, np.ndarray]:
    """Stationary acceptance rate and proposal Hamming-distance distribution of the
    Metropolis chain with symmetric proposal Q at inverse temperature beta
    (accept rule: classical_mcmc_routines.py:28-41)."""
    n = len(h)
    N = 1 << n
    E = _energies(J, h, np.arange(N, dtype=np.uint64))
    w = np.exp(-beta * (E - E.min()))
    pi = w / w.sum()
    dE = E[:, None] - E[None, :]  # E(s') - E(s), rows = s'
    A = np.minimum(1.0, np.exp(-beta * dE))
    acc = float(np.sum(pi[None, :] * Q * A))
    idx = np.arange(N)
    ham = np.array([[bin(a ^ b).count("1") for b in idx] for a in idx]) if N <= 1024 else None
    if ham is None:
        x = idx[:, None] ^ idx[None, :]
        ham = np.zeros_like(x)
        for b in range(n):
            ham += (x >> b) & 1
    hist = np.zeros(n + 1)
    flow = pi[None, :] * Q
    for d in range(n + 1):
        hist[d] = flow[ham == d].sum()
    return acc, hist


# --------------------------------------------------------------------------- exact-unitary path
def _pauli_string_matrix(n: int, ops: dict) -> np.ndarray:
    """Kronecker product with qubit 0 = least-significant bit (qulacs get_matrix convention)."""
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
    I = np.eye(2, dtype=np.complex128)
    m = np.array([[1.0 + 0j]])
    for q in range(n - 1, -1, -1):  # most-significant first
        g = ops.get(q, "I")
        m = np.kron(m, {"I": I, "X": X, "Z": Z}[g])
    return m


def exact_path_unitary(J, h, alpha_, gamma, t, mixer_terms: Sequence[Sequence[int]],
                       reverse_labels: bool = True, round_decimals: Optional[int] = 6) -> np.ndarray:
    """U = round(expm(-i H t), 6), H = (1-gamma) alpha (-1) H_prob + gamma H_mix
    (quantum_mcmc_routines_exact.py:56,141-143).  H_prob = sum h_i Z_i + sum_{i>j} J_ij Z_i Z_j on
    qubit i = spin i (energy_models.py:105-122) when reverse_labels (quirk Q10: un-reversed w.r.t.
    the bitstring convention); reverse_labels=False uses qubit n-1-i, the Trotter path's labelling.
    H_mix = sum over X-strings (create_X_mixer_hamiltonian :66-79)."""
    from scipy.linalg import expm

    J = np.asarray(J, dtype=float)
    h = np.asarray(h, dtype=float)
    n = len(h)
    q_of = (lambda i: i) if reverse_labels else (lambda i: n - 1 - i)
    Hp = np.zeros((1 << n, 1 << n), dtype=np.complex128)
    for i in range(n):
        Hp += h[i] * _pauli_string_matrix(n, {q_of(i): "Z"})
        for j in range(i):
            Hp += J[i, j] * _pauli_string_matrix(n, {q_of(i): "Z", q_of(j): "Z"})
    Hm = np.zeros_like(Hp)
    for term in mixer_terms:
        Hm += _pauli_string_matrix(n, {q: "X" for q in term})
    H = (1 - gamma) * alpha_ * (-1) * Hp + gamma * Hm
    U = expm(-1j * H * t)
    return np.round(U, decimals=round_decimals) if round_decimals is not None else U


def contiguous_x_mixer_terms(n: int, weight: int):
    """create_X_mixer_hamiltonian (quantum_mcmc_routines_exact.py:66-79): windows X_i..X_{i+w-1}."""
    return [list(range(i, i + weight)) for i in range(0, n - weight + 1)]


# --------------------------------------------------------------------------- distribution metrics
def entropy_ref(probs: np.ndarray):
    """Exact_Sampling.get_entropy (energy_models.py:319-326): log2, only p > 1e-5, returns None if
    no p <= 1e-5 is ever reached (quirk Q7)."""
    tmp = sorted(np.asarray(probs, dtype=float), reverse=True)
    ent = 0.0
    for v in tmp:
        if v > 0.00001:
            ent += -1 * v * np.log2(v)
        else:
            return ent
    return None


def vectorised_kl(target: np.ndarray, model: np.ndarray) -> float:
    """prob_dist.py:154-161 (vectoried_KL): ln, target > 1e-10, model floored at EPS."""
    inds = target > 1e-10
    t = target[inds]
    m = model[inds]
    m = np.where(m < EPS, EPS, m)
    return float(np.sum(t * np.log(t) - t * np.log(m)))


def running_kl(target_probs: np.ndarray, markov_chain_idx: np.ndarray) -> np.ndarray:
    """calculate_running_kl_divergence (trajectory_processing.py:114-153)."""
    mod = np.zeros_like(target_probs)
    out = np.empty(len(markov_chain_idx))
    for ii, s in enumerate(markov_chain_idx, start=1):
        mod[int(s)] += 1
        out[ii - 1] = vectorised_kl(target_probs, mod / ii)
    return out


def markov_chain_from(s0: int, proposals: np.ndarray, accepted: np.ndarray) -> np.ndarray:
    """MCMCChain.get_list_markov_chain (basic_utils.py:105-115): state after every hop,
    starting with the initial state."""
    out = np.empty(len(proposals) + 1, dtype=np.uint64)
    out[0] = s0
    cur = s0
    for k, (p, a) in enumerate(zip(proposals, accepted)):
        if a:
            cur = p
        out[k + 1] = cur
    return out


def running_magnetisation(markov_chain_idx: np.ndarray, n: int) -> np.ndarray:
    """calculate_runnning_magnetisation (trajectory_processing.py:183-210): for k = 1..H the expectation of
    magnetization_of_state (#'1' - #'0', basic_utils.py:174-186) over get_accepted_dict(normalize=True, until_index=k),
    i.e. over markov_chain[0:k]."""
    mags = np.array([2 * bin(int(s)).count("1") - n for s in markov_chain_idx], dtype=np.float64)
    H = len(markov_chain_idx) - 1
    return np.array([mags[:k].sum() / k for k in range(1, H + 1)])


def trajectory_statistics(J, h, s0: int, proposals: np.ndarray, accepted: np.ndarray):
    """get_trajectory_statistics (trajectory_processing.py:213-321): per hop E(proposal) - E(current) (:235-263) and
    the Hamming-distance table of accepted / rejected proposals w.r.t. the current state (:279-301)."""
    n = len(h)
    ediff = np.empty(len(proposals))
    ham = np.zeros((n + 1, 2), dtype=np.int64)
    cur = int(s0)
    e_cur = float(_energies(J, h, np.array([cur], dtype=np.uint64))[0])
    for k, (p, a) in enumerate(zip(proposals, accepted)):
        e_p = float(_energies(J, h, np.array([int(p)], dtype=np.uint64))[0])
        ediff[k] = e_p - e_cur
        ham[bin(int(p) ^ cur).count("1"), 0 if a else 1] += 1
        if a:
            cur, e_cur = int(p), e_p
    return ediff, ham
