"""CPU checks of the identities behind the Hadamard-basis mixer modes (DESIGN sections 4.1, 4.2b, 4.4), against the oracle's
term-by-term simulator (oracle.evolve follows qumcmc/mixers.py:34-52 + quantum_mcmc_routines.py:95-106 gate by gate):
  (1) E_x(x) = sum_S w_S (-1)^popcount(x & S) is the Walsh-Hadamard transform of the coefficient vector w[S]       (XT tables)
  (2) for terms of weight <= 2 it is the quadratic form sum_q h_q z_q + sum_{q<q'} J_qq' z_q z_q'                   (XQ tables)
  (3) M = prod_S exp(-i gamma w_S dt X_S) = H exp(-i gamma dt E_x) H, and U = M (P M)^(r-1) built from it equals the oracle
  (4) with B = [[1, 1], [-1, 1]] per qubit (the butterfly with multiplier +1) H D H = 2^-n B^T D B, and B^(x)n |s> has
      components (-1)^popcount(x & ~s)                                                                  (the kernels' convention)
"""
import numpy as np
import pytest

import oracle
from tests import models


def _wht(v):
    v = np.array(v, dtype=np.complex128)
    n = int(np.log2(v.size))
    for b in range(n):
        v = v.reshape(-1, 2, 1 << b)
        v = np.stack([v[:, 0] + v[:, 1], v[:, 0] - v[:, 1]], axis=1).reshape(-1)
    return v


def _ex_table(n, masks, w):
    x = np.arange(1 << n, dtype=np.uint64)
    e = np.zeros(1 << n)
    for mk, wk in zip(masks, w):
        par = np.array([bin(int(v)).count("1") & 1 for v in (x & np.uint64(mk))])
        e += np.where(par == 1, -wk, wk)
    return e


def _mixers(n):
    import qumcmc_b200 as q
    return {"one": q.GenericMixer.OneBodyMixer(n), "two": q.GenericMixer.TwoBodyMixer(n), "three": q.GenericMixer.ThreeBodyMixer(n),
            "custom": q.CustomMixer(n, [[0, n - 1], [1, 2, 3], [2], [0, 1, 2, 3]])}


@pytest.mark.parametrize("n", [5, 8])
def test_mixer_exponent_is_the_walsh_transform_of_the_coefficients(n):
    for name, mx in _mixers(n).items():
        _, masks, w, _ = mx.lower()
        coef = np.zeros(1 << n)
        for mk, wk in zip(masks, w):
            coef[int(mk)] += float(wk)
        assert np.allclose(_wht(coef).real, _ex_table(n, masks, w), atol=1e-12), name


def test_quadratic_form_of_weight_two_mixers():
    n = 7
    import qumcmc_b200 as q
    mx = q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.GenericMixer.TwoBodyMixer(n)], [0.7, 0.3])
    _, masks, w, _ = mx.lower()
    xh, xJ = np.zeros(n), np.zeros((n, n))
    for mk, wk in zip(masks, w):
        qs = [b for b in range(n) if (int(mk) >> b) & 1]
        if len(qs) == 1:
            xh[qs[0]] += wk
        else:
            xJ[qs[0], qs[1]] += wk
    x = np.arange(1 << n)
    z = 1.0 - 2.0 * ((x[:, None] >> np.arange(n)[None, :]) & 1)
    quad = z @ xh + np.einsum("xi,ij,xj->x", z, xJ, z)
    assert np.allclose(quad, _ex_table(n, masks, w), atol=1e-12)


@pytest.mark.parametrize("n,r", [(5, 1), (6, 3), (8, 4)])
def test_proposal_unitary_through_the_hadamard_basis_equals_the_oracle(n, r):
    J, h = models.packaged_random_model(n, 3)
    al = oracle.alpha(J, h)
    gamma, dt, s = 0.37, 0.8, 11 % (1 << n)
    idx = np.arange(1 << n)
    # phase layer exp(+i c D(idx)), D with z = +1 / -1 for bit clear / set, qubit a = spin n - 1 - a
    z = 1.0 - 2.0 * ((idx[:, None] >> np.arange(n)[None, :]) & 1)
    Jq, hq = J[::-1, ::-1], h[::-1]
    D = z @ hq + 0.5 * np.einsum("xi,ij,xj->x", z, Jq - np.diag(np.diag(Jq)), z)
    P = np.exp(1j * (1.0 - gamma) * al * dt * D)
    for name, mx in _mixers(n).items():
        _, masks, w, _ = mx.lower()
        ex = _ex_table(n, masks, w)
        Dx = np.exp(-1j * gamma * dt * ex)
        N = 1 << n

        def M(v):
            return _wht(Dx * _wht(v)) / N

        st = np.zeros(N, dtype=np.complex128)
        st[s] = 1.0
        st = M(st)
        for _ in range(r - 1):
            st = M(P * st)
        ref = oracle.evolve(J, h, al, gamma, dt, r, masks, w, s, gatewise=False)
        assert np.abs(st - ref).max() < 1e-12, name


def test_butterfly_convention_of_the_kernels():
    """B = [[1, 1], [-1, 1]] (multiplier +1), B^T (multiplier -1): H D H = 2^-n B^T D B for any diagonal D, and the first
    sweep's analytic start state B^(x)n |s> = (-1)^popcount(x & ~s)."""
    n = 6
    N = 1 << n
    rng = np.random.default_rng(1)
    D = np.exp(1j * rng.uniform(0, 6.28, N))

    def apply(v, t):  # lo' = lo + t hi, hi' = hi - t lo on every qubit
        v = np.array(v, dtype=np.complex128)
        for b in range(n):
            v = v.reshape(-1, 2, 1 << b)
            lo, hi = v[:, 0].copy(), v[:, 1].copy()
            v = np.stack([lo + t * hi, hi - t * lo], axis=1).reshape(-1)
        return v

    v = rng.normal(size=N) + 1j * rng.normal(size=N)
    assert np.abs(apply(D * apply(v, +1.0), -1.0) / N - _wht(D * _wht(v)) / N).max() < 1e-12
    s = 0b101101
    e = np.zeros(N)
    e[s] = 1.0
    x = np.arange(N)
    sign = np.array([(-1.0) ** bin(int(v) & ~s & (N - 1)).count("1") for v in x])
    assert np.abs(apply(e, +1.0) - sign).max() < 1e-12
