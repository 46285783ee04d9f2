"""CPU oracle for the quMCMC quantum-proposal hot path -- TEST INFRASTRUCTURE ONLY.

ctypes bindings over ``oracle/liboracle.so`` (built from ``qmc_oracle.c`` by ``make -C oracle``)
plus numpy-level helpers in :mod:`oracle.np_oracle`.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; the
product (``qumcmc_b200``) never does.

Parity status: qulacs (the reference's un-pinned simulator dependency,
/root/reference/setup.py:37) is not installable here; the oracle restates its published gate
semantics and is pinned against the reference's notebook known-answers and recorded chains
(tests/golden/*.json, tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib: Optional[C.CDLL] = None


def build(force: bool = False) -> str:
    """Compile liboracle.so with the system gcc (seconds)."""
    src = os.path.join(_HERE, "qmc_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        u64p = C.POINTER(C.c_uint64)
        L.orc_philox4x32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_philox4x32.restype = None
        L.orc_hop_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, dp, C.POINTER(C.c_int), dp, dp, dp]
        L.orc_hop_draws.restype = None
        L.orc_initial_state.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
        L.orc_initial_state.restype = C.c_uint64
        L.orc_energy.argtypes = [C.c_int, dp, dp, C.c_uint64]
        L.orc_energy.restype = C.c_double
        L.orc_energies.argtypes = [C.c_int, dp, dp, u64p, C.c_int64, dp]
        L.orc_energies.restype = None
        L.orc_alpha.argtypes = [C.c_int, dp, dp]
        L.orc_alpha.restype = C.c_double
        L.orc_phase_D.argtypes = [C.c_int, dp, dp, C.c_uint64]
        L.orc_phase_D.restype = C.c_double
        L.orc_problem_layer_active.argtypes = [C.c_int, dp, dp]
        L.orc_problem_layer_active.restype = C.c_int
        L.orc_trotter_steps.argtypes = [C.c_int, C.c_double]
        L.orc_trotter_steps.restype = C.c_int
        ev = [C.c_int, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, u64p, dp, C.c_uint64, dp]
        L.orc_evolve_gatewise.argtypes = ev
        L.orc_evolve_gatewise.restype = None
        L.orc_evolve_fused.argtypes = ev
        L.orc_evolve_fused.restype = None
        L.orc_sample_index.argtypes = [C.c_int, dp, C.c_double]
        L.orc_sample_index.restype = C.c_uint64
        L.orc_test_accept.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_test_accept.restype = C.c_int
        L.orc_run_chain.argtypes = [
            C.c_int, dp, dp, dp, dp, C.c_double, C.c_int, C.POINTER(C.c_int), u64p, dp, dp,
            C.c_uint64, C.c_int64, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_double,
            C.c_uint64, C.c_uint64, C.c_int, u64p, C.POINTER(C.c_uint8)]
        L.orc_run_chain.restype = C.c_int64
        L.orc_classical_chain.argtypes = [C.c_int, dp, dp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_uint64), dp,
                                          C.c_uint64, C.c_int64, C.c_uint32, C.c_double, C.c_uint64, C.c_uint64,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)]
        L.orc_classical_chain.restype = C.c_int64
        L.orc_exact_boltzmann.argtypes = [C.c_int, dp, dp, C.c_double, dp, dp]
        L.orc_exact_boltzmann.restype = C.c_double
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_set_num_threads.restype = None
        L.orc_num_threads.argtypes = []
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _u64p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def bitstring_to_index(s: str) -> int:
    return int(s, 2)


def index_to_bitstring(idx: int, n: int) -> str:
    return f"{int(idx):0{n}b}"


def philox(seed: int, chain: int, hop: int, slot: int) -> np.ndarray:
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32(seed, chain, hop & 0xFFFFFFFF, slot, out)
    return np.array(list(out), dtype=np.uint32)


def hop_draws(seed: int, chain: int, hop: int) -> Tuple[float, int, float, float, float]:
    ug, ugrp, um, ua = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    t = C.c_int()
    lib().orc_hop_draws(seed, chain, hop, C.byref(ug), C.byref(t), C.byref(ugrp), C.byref(um), C.byref(ua))
    return ug.value, t.value, ugrp.value, um.value, ua.value


def initial_state(seed: int, chain: int, n: int) -> int:
    return int(lib().orc_initial_state(seed, chain, n))


def energy(J, h, idx: int) -> float:
    J, h = _d(J), _d(h)
    return float(lib().orc_energy(len(h), _dp(J), _dp(h), int(idx)))


def energies(J, h, idx) -> np.ndarray:
    J, h = _d(J), _d(h)
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    out = np.empty(idx.shape[0], dtype=np.float64)
    lib().orc_energies(len(h), _dp(J), _dp(h), _u64p(idx), idx.shape[0], _dp(out))
    return out


def alpha(J, h) -> float:
    J, h = _d(J), _d(h)
    return float(lib().orc_alpha(len(h), _dp(J), _dp(h)))


def phase_D(J, h, idx: int) -> float:
    J, h = _d(J), _d(h)
    return float(lib().orc_phase_D(len(h), _dp(J), _dp(h), int(idx)))


def phase_D_table(J, h) -> np.ndarray:
    n = len(h)
    return np.array([phase_D(J, h, i) for i in range(1 << n)])


def problem_layer_active(J, h) -> bool:
    J, h = _d(J), _d(h)
    return bool(lib().orc_problem_layer_active(len(h), _dp(J), _dp(h)))


def trotter_steps(t: int, delta_time: float) -> int:
    return int(lib().orc_trotter_steps(int(t), float(delta_time)))


def evolve(circJ, circh, alpha_: float, gamma: float, delta_time: float, r: int,
           masks: Sequence[int], weights: Sequence[float], s_idx: int, gatewise: bool = True) -> np.ndarray:
    """State after the proposal circuit, complex128[2^n] (run_qmcmc_quantum_ckt :110-152)."""
    circJ, circh = _d(circJ), _d(circh)
    n = len(circh)
    m = np.ascontiguousarray(masks, dtype=np.uint64)
    w = _d(weights)
    st = np.empty(2 << n, dtype=np.float64)
    fn = lib().orc_evolve_gatewise if gatewise else lib().orc_evolve_fused
    fn(n, _dp(circJ), _dp(circh), float(alpha_), float(gamma), float(delta_time), int(r), len(m),
       _u64p(m), _dp(w), int(s_idx), _dp(st))
    return st.view(np.complex128)


def transition_probs(circJ, circh, alpha_, gamma, delta_time, r, masks, weights, s_idx,
                     gatewise: bool = True) -> np.ndarray:
    a = evolve(circJ, circh, alpha_, gamma, delta_time, r, masks, weights, s_idx, gatewise)
    return a.real * a.real + a.imag * a.imag


def sample_index(state: np.ndarray, u: float) -> int:
    st = np.ascontiguousarray(state, dtype=np.complex128).view(np.float64)
    n = int(np.log2(state.shape[0]))
    return int(lib().orc_sample_index(n, _dp(st), float(u)))


def test_accept(e_s: float, e_sp: float, temperature: float, u: float) -> bool:
    return bool(lib().orc_test_accept(e_s, e_sp, temperature, u))


test_accept.__test__ = False  # not a pytest test


def run_chain(J, h, masks, weights, s0: int, n_hops: int, temperature: float, gamma, delta_time: float,
              seed: int, chain_id: int = 0, *, circJ=None, circh=None, alpha_: Optional[float] = None,
              group_off=None, group_prob=None, hop0: int = 0, gatewise: bool = False):
    """One chain under the Philox stream contract. Returns (proposals u64[H], accepted u8[H])."""
    J, h = _d(J), _d(h)
    circJ = J if circJ is None else _d(circJ)
    circh = h if circh is None else _d(circh)
    n = len(h)
    a = alpha(J, h) if alpha_ is None else alpha_
    m = np.ascontiguousarray(masks, dtype=np.uint64)
    w = _d(weights)
    if group_off is None:
        group_off = [0, len(m)]
    go = np.ascontiguousarray(group_off, dtype=np.int32)
    gp = None if group_prob is None else _d(group_prob)
    g_lo, g_hi = (gamma, gamma) if np.isscalar(gamma) else gamma
    props = np.empty(n_hops, dtype=np.uint64)
    acc = np.empty(n_hops, dtype=np.uint8)
    lib().orc_run_chain(n, _dp(J), _dp(h), _dp(circJ), _dp(circh), float(a), len(go) - 1,
                        go.ctypes.data_as(C.POINTER(C.c_int)), _u64p(m), _dp(w),
                        None if gp is None else _dp(gp), int(s0), int(n_hops), int(hop0),
                        float(temperature), float(g_lo), float(g_hi), float(delta_time), int(seed),
                        int(chain_id), int(bool(gatewise)), _u64p(props),
                        acc.ctypes.data_as(C.POINTER(C.c_uint8)))
    return props, acc


def classical_chain(J, h, kinds, params, probs, s0: int, n_hops: int, temperature: float, seed: int,
                    chain_id: int = 0, hop0: int = 0):
    """One classical_mcmc chain (classical_mcmc_routines.py:44-99) under the classical draw contract.
    kinds: 0 uniform / 1 fixed weight (param k) / 2 custom (param mask); returns (proposals u64[H], accepted u8[H])."""
    J, h = _d(J), _d(h)
    n = len(h)
    k = np.ascontiguousarray(kinds, dtype=np.int32)
    p = np.ascontiguousarray(params, dtype=np.uint64)
    pr = np.full(len(k), 1.0 / len(k)) if probs is None else np.asarray(probs, dtype=np.float64)
    cum = np.cumsum(pr / pr.sum())
    cum[-1] = 1.0
    cum = _d(cum)
    props = np.empty(n_hops, dtype=np.uint64)
    acc = np.empty(n_hops, dtype=np.uint8)
    lib().orc_classical_chain(n, _dp(J), _dp(h), len(k), k.ctypes.data_as(C.POINTER(C.c_int)), _u64p(p), _dp(cum),
                              int(s0), int(n_hops), int(hop0), float(temperature), int(seed), int(chain_id),
                              _u64p(props), acc.ctypes.data_as(C.POINTER(C.c_uint8)))
    return props, acc


def exact_boltzmann(J, h, beta: float, want_probs: bool = True):
    """(logZ, energies[2^n], probs[2^n] or None) -- energy_models.py:218-275."""
    J, h = _d(J), _d(h)
    n = len(h)
    E = np.empty(1 << n, dtype=np.float64)
    P = np.empty(1 << n, dtype=np.float64) if want_probs else None
    logZ = lib().orc_exact_boltzmann(n, _dp(J), _dp(h), float(beta), _dp(E), None if P is None else _dp(P))
    return float(logZ), E, P


def set_num_threads(t: int) -> None:
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm may use all host cores."""
    lib().orc_set_num_threads(int(t))


def num_threads() -> int:
    return int(lib().orc_num_threads())
