"""Host replica of the Philox4x32-10 stream contract (include/qumcmc_b200.h) in numpy integer
arithmetic -- used for bookkeeping only (e.g. reporting the initial states the device drew);
the device draws its own numbers."""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_LO = np.uint64(0xFFFFFFFF)


def philox4x32(seed: int, chain: np.ndarray, hop: int, slot: int) -> np.ndarray:
    """words[len(chain), 4] as uint32 for counter (chain lo, chain hi, hop, slot)."""
    chain = np.asarray(chain, dtype=np.uint64)
    c0 = chain & _LO
    c1 = chain >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(hop & 0xFFFFFFFF))
    c3 = np.full_like(c0, np.uint64(slot))
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & _LO
        n1 = p1 & _LO
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & _LO
        n3 = p0 & _LO
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def initial_states_philox(seed: int, chain_id0: int, n_chains: int, n: int) -> np.ndarray:
    w = philox4x32(seed, np.arange(chain_id0, chain_id0 + n_chains, dtype=np.uint64), 0xFFFFFFFF, 0).astype(np.uint64)
    v = (w[:, 1] << np.uint64(32)) | w[:, 0]
    return v if n >= 64 else v & np.uint64((1 << n) - 1)
