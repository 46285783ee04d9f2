"""numpy executor of the sharded-statevector op list (test infrastructure, CPU only).

Runs ``qumcmc_b200.sharded.sharded_plan`` on complex128 shards in physical coordinates with the same semantics
as qmc_shard_op, so the schedule / layout bookkeeping can be checked against the oracle without a GPU (the exchange
is a real ``all_to_all_single`` over gloo in the multi-process test, a reshape in the single-process one)."""
import numpy as np

from qumcmc_b200.sharded import EXCHANGE, FIRST, LO_FINAL, LO_SINGLE, MID, SINGLE, shard_geometry


class ShardSim:
    def __init__(self, J, h, alpha, rank, g):
        self.n = n = len(h)
        self.g, self.rank, self.nl = g, rank, n - g
        # logical couplings in qubit order: qubit a = spin n-1-a
        self.Jl = np.asarray(J, dtype=np.float64)[::-1, ::-1].copy()
        self.hl = np.asarray(h, dtype=np.float64)[::-1].copy()
        self.alpha = alpha
        self.perm = [np.arange(n), np.arange(n)]
        for j in range(g):
            a, b = self.nl - g + j, self.nl + j
            self.perm[1][a], self.perm[1][b] = self.perm[1][b], self.perm[1][a]
        self.groups = shard_geometry(self.nl)
        self.state = None

    def begin(self, s, gamma, r, dt, weights=None):
        self.s, self.gamma, self.r, self.dt = s, gamma, r, dt
        self.w = np.ones(self.n) if weights is None else np.asarray(weights, dtype=np.float64)

    def _theta(self, layout):
        return self.gamma * self.w[self.perm[layout]] * self.dt

    def _bits(self, a):
        """bit of position a for every local amplitude (array) or for the rank (scalar)"""
        if a < self.nl:
            return (np.arange(1 << self.nl, dtype=np.int64) >> a) & 1
        return (self.rank >> (a - self.nl)) & 1

    def _mix(self, positions, layout):
        th = self._theta(layout)
        st = self.state
        for a in positions:
            assert a < self.nl
            c, s_ = np.cos(th[a]), np.sin(th[a])
            v = st.reshape(-1, 2, 1 << a)
            lo, hi = v[:, 0, :].copy(), v[:, 1, :].copy()
            v[:, 0, :] = c * lo - 1j * s_ * hi
            v[:, 1, :] = c * hi - 1j * s_ * lo
        self.state = st

    def _phase(self, layout):
        if self.r <= 1:
            return
        n, perm = self.n, self.perm[layout]
        Jp = self.Jl[np.ix_(perm, perm)]
        hp = self.hl[perm]
        z = [1.0 - 2.0 * self._bits(a) for a in range(n)]
        D = np.zeros(1 << self.nl)
        for a in range(n):
            D = D + hp[a] * z[a]
            for b in range(a + 1, n):
                if Jp[a, b] != 0.0:
                    D = D + Jp[a, b] * (z[a] * z[b])
        c = (1.0 - self.gamma) * self.alpha * self.dt
        self.state = self.state * np.exp(1j * c * D)

    def _product_state(self, layout):
        th = self._theta(layout)
        perm = self.perm[layout]
        amp = np.ones(1 << self.nl, dtype=np.complex128)
        for a in range(self.n):
            sbit = (self.s >> int(perm[a])) & 1
            differs = self._bits(a) ^ sbit
            amp = amp * np.where(differs == 1, -1j * np.sin(th[a]), np.cos(th[a]))
        self.state = amp

    def op(self, kind, k, pre, layout):
        lo_pos = range(0, 12)
        top = self.groups[0]
        swapped = range(self.nl - self.g, self.nl)
        if kind == FIRST:
            self._product_state(layout)
            self._phase(layout)
            self._mix(range(*top), layout)
        elif kind == MID:
            self._mix(swapped, layout)
            self._phase(layout)
            self._mix(range(*top), layout)
        elif kind == SINGLE:
            self._mix(swapped if (k == 0 and pre > 0) else range(*self.groups[k]), layout)
        elif kind == LO_SINGLE:
            self._mix(lo_pos, layout)
        elif kind == LO_FINAL:
            if self.r <= 1:
                self._product_state(layout)
            else:
                self._mix(lo_pos, layout)
        else:
            raise ValueError(kind)

    def probabilities(self):
        return np.abs(self.state) ** 2


def logical_indices(rank, nl, g, layout):
    """logical basis index of every local amplitude of ``rank`` in ``layout``"""
    L = np.arange(1 << nl, dtype=np.int64)
    if layout == 0:
        return L | (rank << nl)
    low = L & ((1 << (nl - g)) - 1)
    top = L >> (nl - g)
    return low | (rank << (nl - g)) | (top << nl)
