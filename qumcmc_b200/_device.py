"""Host glue between the qumcmc Python API and the C-ABI: handle cache, torch tensors as device
buffers, stream passing.  PyTorch is plumbing here (device memory, streams, torch.distributed)."""
from __future__ import annotations

import ctypes as C
import hashlib
import weakref
from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib


def _torch():
    import torch

    return torch


def require_cuda(device: int = 0):
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.QmcError("qumcmc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", device)


def current_stream_ptr(device: int) -> int:
    torch = _torch()
    return int(torch.cuda.current_stream(device).cuda_stream)


def _hp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class DeviceModel:
    """Owns one qmc_model handle (one device)."""

    def __init__(self, J: np.ndarray, h: np.ndarray, circJ: np.ndarray, circh: np.ndarray, alpha: float, device: int):
        self.n = len(h)
        self.device = device
        require_cuda(device)
        L = _lib.lib()
        self._J = np.ascontiguousarray(J, dtype=np.float64)
        self._h = np.ascontiguousarray(h, dtype=np.float64)
        self._cJ = np.ascontiguousarray(circJ, dtype=np.float64)
        self._ch = np.ascontiguousarray(circh, dtype=np.float64)
        handle = C.c_void_p()
        _lib.check(L.qmc_model_create(C.byref(handle), self.n, _hp(self._J), _hp(self._h), _hp(self._cJ),
                                      _hp(self._ch), float(alpha), int(device)))
        self.handle = handle
        self._mixer_key = None
        self._finalizer = weakref.finalize(self, L.qmc_model_destroy, handle)

    def set_mixer(self, mixer) -> None:
        offs, masks, weights, probs = mixer.lower()
        key = (offs.tobytes(), masks.tobytes(), weights.tobytes(), None if probs is None else probs.tobytes())
        if key == self._mixer_key:
            return
        _lib.check(_lib.lib().qmc_mixer_set(self.handle, len(offs) - 1, _hp(offs), _hp(masks), _hp(weights), _hp(probs)))
        self._mixer_key = key

    def attach_visit_histogram(self, hist) -> None:
        """Attach (torch int64[2^n] CUDA tensor, zeroed by the caller) or detach (None) the visit-histogram accumulator:
        while attached every chain call on this handle counts the states of its Markov chains into it."""
        if hist is not None and (hist.numel() != (1 << self.n) or str(hist.dtype) != "torch.int64" or not hist.is_cuda):
            raise ValueError("visit histogram must be a CUDA int64 tensor of 2^n entries")
        _lib.check(_lib.lib().qmc_set_visit_histogram(self.handle, None if hist is None else hist.data_ptr()))
        self._visit = hist  # keeps the tensor alive while attached

    def launch_count(self) -> int:
        return int(_lib.lib().qmc_model_launch_count(self.handle))


_cache: Dict[Tuple, DeviceModel] = {}


def device_model(J, h, circJ, circh, alpha: float, device: int = 0) -> DeviceModel:
    """Handle for (model parameters, device); rebuilt when the parameters change (CD training
    mutates J/h in place through _update_J/_update_h)."""
    hs = hashlib.blake2b(digest_size=16)
    for a in (J, h, circJ, circh):
        hs.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    hs.update(np.float64(alpha).tobytes())
    key = (hs.hexdigest(), int(device))
    dm = _cache.get(key)
    if dm is None:
        if len(_cache) > 8:
            _cache.pop(next(iter(_cache)))
        dm = DeviceModel(J, h, circJ, circh, alpha, device)
        _cache[key] = dm
    return dm
