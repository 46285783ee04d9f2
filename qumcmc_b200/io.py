"""On-disk formats.

* **Legacy pickles**: the reference stores experiment results as pickled ``MCMCChain`` objects / lists / dicts of them
  (``gamma_variation/*.pkl``, ``final_results_notebooks/**/*.pkl``; class layout /root/reference/qumcmc/basic_utils.py:54-131).
  ``load_legacy_pickle`` reads them through the ``qumcmc`` alias package (the pickles name ``qumcmc.basic_utils.MCMCChain``)
  and also accepts files written under numpy 1.x (``numpy.core`` module paths).  Chains pickled by this package carry the
  reference's five attributes, so the reference's own analysis code can read them back.
* **Compact arrays** (``.npz``): ``save_chains`` / ``load_chains`` store a batch of chains as the device path produces
  them -- ``initial[C]``, ``proposals[C, H]`` (uint64 indices), ``accepted[C, H]`` bit-packed -- 8 B + 1 bit per hop instead
  of a Python object per state."""
from __future__ import annotations

import pickle
from typing import Sequence, Union

import numpy as np

from .basic_utils import MCMCChain
from .quantum_mcmc_routines import ChainBatch

FORMAT = "qumcmc_b200.chains/1"


class _LegacyUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        import qumcmc  # noqa: F401  registers qumcmc.basic_utils & co. as aliases of this package's modules

        if module.startswith("numpy.core") and np.lib.NumpyVersion(np.__version__) >= "2.0.0":
            module = module.replace("numpy.core", "numpy._core", 1)
        return super().find_class(module, name)


def load_legacy_pickle(path: str):
    """Object tree of a reference-written pickle; every ``MCMCChain`` inside comes back array-backed."""
    with open(path, "rb") as f:
        return _LegacyUnpickler(f).load()


def _as_batch(chains: Union[ChainBatch, MCMCChain, Sequence[MCMCChain]]):
    if isinstance(chains, ChainBatch):
        chains.require_recorded("save_chains")
        return chains.num_spins, chains.initial, chains.proposals, chains.accepted, [chains.name] * len(chains.initial)
    if isinstance(chains, MCMCChain):
        chains = [chains]
    chains = list(chains)
    if not chains:
        raise ValueError("no chains to save")
    n = chains[0].num_spins
    H = len(chains[0]) - 1
    if any(c.num_spins != n or len(c) - 1 != H for c in chains):
        raise ValueError("chains differ in width or length")
    idx = np.stack([c.proposal_indices for c in chains]).astype(np.uint64)
    acc = np.stack([c.accepted_mask for c in chains]).astype(np.uint8)
    return n, idx[:, 0], idx[:, 1:], acc[:, 1:], [c.name for c in chains]


def save_chains(path: str, chains: Union[ChainBatch, MCMCChain, Sequence[MCMCChain]]) -> None:
    n, initial, proposals, accepted, names = _as_batch(chains)
    C_, H = proposals.shape
    np.savez_compressed(path, format=np.array(FORMAT), num_spins=np.int64(n), n_hops=np.int64(H),
                        initial=np.ascontiguousarray(initial, dtype=np.uint64),
                        proposals=np.ascontiguousarray(proposals, dtype=np.uint64),
                        accepted_bits=np.packbits(np.ascontiguousarray(accepted, dtype=np.uint8) != 0, axis=1),
                        names=np.array(names))


def load_chains(path: str, as_batch: bool = False):
    """List of ``MCMCChain`` (or ``(num_spins, initial, proposals, accepted, names)`` arrays with ``as_batch``)."""
    with np.load(path, allow_pickle=False) as z:
        if str(z["format"]) != FORMAT:
            raise ValueError(f"{path}: not a {FORMAT} file")
        n, H = int(z["num_spins"]), int(z["n_hops"])
        initial, proposals = z["initial"], z["proposals"]
        accepted = np.unpackbits(z["accepted_bits"], axis=1, count=H).astype(np.uint8) if H else np.zeros((len(initial), 0), np.uint8)
        names = [str(s) for s in z["names"]]
    if as_batch:
        return n, initial, proposals, accepted, names
    return [MCMCChain.from_arrays(n, int(initial[c]), proposals[c], accepted[c].astype(bool), name=names[c])
            for c in range(len(initial))]
