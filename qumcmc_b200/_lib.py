"""ctypes binding of ``csrc/libqumcmc_b200.so`` (C-ABI in ``include/qumcmc_b200.h``).

There is no CPU fallback: if the shared library is missing or cannot be loaded, every entry point of
the hot path raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C qumcmc_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libqumcmc_b200.so")

QMC_OK = 0
QMC_ERR_BADARG = -1
QMC_ERR_CUDA = -2
QMC_ERR_NCCL = -3
QMC_ERR_UNSUPPORTED = -4
DTYPE_F32 = 0
DTYPE_F64 = 1
SMEM_MAX_SPINS_F32 = 13
SMEM_MAX_SPINS_F64 = 12
MAX_SPINS = 40          # model limit; one GPU holds a statevector up to MAX_LOCAL_SPINS qubits
MAX_LOCAL_SPINS = 34
IPC_HANDLE_BYTES = 64
VISIT_HIST_MAX_SPINS = 30
CLASSICAL_UNIFORM, CLASSICAL_FIXED_WEIGHT, CLASSICAL_CUSTOM = range(3)
MAX_CLASSICAL_METHODS = 16
SHARD_OP_FIRST, SHARD_OP_MID, SHARD_OP_SINGLE, SHARD_OP_LO_SINGLE, SHARD_OP_LO_FINAL = range(5)

# every symbol include/qumcmc_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "qmc_version", "qmc_last_error", "qmc_model_create", "qmc_model_destroy", "qmc_mixer_set",
    "qmc_energies", "qmc_energies_host", "qmc_exact_sampling", "qmc_exact_sampling_host",
    "qmc_exact_partial", "qmc_exact_probs_range",
    "qmc_transition_probs", "qmc_propose", "qmc_run_chains", "qmc_run_chains_host",
    "qmc_model_num_spins", "qmc_model_launch_count", "qmc_profile_enable", "qmc_profile_read",
    "qmc_shard_create", "qmc_shard_destroy", "qmc_shard_num_groups", "qmc_shard_begin", "qmc_shard_op",
    "qmc_shard_norm", "qmc_shard_select", "qmc_shard_final_probs", "qmc_shard_set_peers", "qmc_shard_op_exchange",
    "qmc_enable_peer_access", "qmc_ipc_export", "qmc_ipc_open", "qmc_ipc_close_all",
    "qmc_classical_chains", "qmc_classical_chains_host", "qmc_trajectory_stats", "qmc_state_moments", "qmc_shard_op_chunk",
    "qmc_set_visit_histogram", "qmc_profile_read_moved", "qmc_shard_set_exchange_ctas", "qmc_shard_exchange_can_persist", "qmc_shard_copy_chunk",
    "qmc_debug_smem_slot", "qmc_debug_layout128", "qmc_debug_general_slot", "qmc_debug_group_plan", "qmc_debug_mixer_class",
]


class QmcError(RuntimeError):
    """Non-zero return code from libqumcmc_b200."""


class QmcUnsupported(QmcError, NotImplementedError):
    """QMC_ERR_UNSUPPORTED: the configuration has no CUDA path (and there is no CPU fallback)."""


_lib: Optional[C.CDLL] = None


STAMP_PATH = os.path.join(CSRC, ".build_stamp")


def source_digest() -> str:
    """sha256 over every input of the build (sources, headers, Makefile): staleness is decided by content, not by
    mtimes (a snapshot copied to another box does not keep them)."""
    import hashlib
    h = hashlib.sha256()
    inc = os.path.join(os.path.dirname(_HERE), "include")
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh")) or f == "Makefile"]
    files += [os.path.join(inc, f) for f in sorted(os.listdir(inc)) if f.endswith(".h")]
    for path in files:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != source_digest()


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile the CUDA extension for sm_100a in-tree (nvcc cross-compiles without a GPU).  A no-op when the library
    was built from exactly the present sources (content digest in csrc/.build_stamp)."""
    if not force and not is_stale():
        return LIB_PATH
    digest = source_digest()
    cmd = ["make", "-C", CSRC, "-j8"]
    if os.environ.get("GRAFT_REPO_ROOT"):  # a snapshot on another box: file times say nothing, rebuild everything
        cmd.append("-B")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libqumcmc_b200.so failed")
    with open(STAMP_PATH, "w") as f:
        f.write(digest + "\n")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QmcError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run __graft_entry__.build() or `make -C qumcmc_b200/csrc`). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, dp, u64p, i64p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses (host or device)
    L.qmc_version.restype = C.c_int
    L.qmc_last_error.restype = C.c_char_p
    L.qmc_model_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, dp, dp, dp, dp, C.c_double, C.c_int]
    L.qmc_model_destroy.argtypes = [C.c_void_p]
    L.qmc_mixer_set.argtypes = [C.c_void_p, C.c_int, vp, vp, vp, vp]
    L.qmc_energies.argtypes = [C.c_void_p, u64p, C.c_int64, dp, vp]
    L.qmc_energies_host.argtypes = [C.c_void_p, u64p, C.c_int64, dp]
    L.qmc_exact_sampling.argtypes = [C.c_void_p, C.c_double, dp, dp, dp, vp]
    L.qmc_exact_sampling_host.argtypes = [C.c_void_p, C.c_double, dp, dp, dp]
    L.qmc_exact_partial.argtypes = [C.c_void_p, C.c_double, C.c_uint64, C.c_int64, dp, dp, vp]
    L.qmc_exact_probs_range.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint64, C.c_int64, dp, dp, vp]
    L.qmc_transition_probs.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.c_double, C.c_int,
                                       C.c_int, vp, vp, vp]
    L.qmc_propose.argtypes = [C.c_void_p, C.c_int64, u64p, dp, vp, dp, vp, C.c_double, C.c_int, u64p, vp]
    L.qmc_run_chains.argtypes = [C.c_void_p, C.c_int64, C.c_int64, u64p, C.c_uint64, C.c_uint32, C.c_double,
                                 C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int, u64p, vp, u64p,
                                 i64p, i64p, vp]
    L.qmc_run_chains_host.argtypes = [C.c_void_p, C.c_int64, C.c_int64, u64p, C.c_uint64, C.c_uint32,
                                      C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int,
                                      u64p, vp, u64p, i64p, i64p]
    L.qmc_model_num_spins.argtypes = [C.c_void_p]
    L.qmc_model_launch_count.argtypes = [C.c_void_p]
    L.qmc_model_launch_count.restype = C.c_int64
    L.qmc_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.qmc_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.qmc_shard_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int]
    L.qmc_shard_destroy.argtypes = [C.c_void_p]
    L.qmc_shard_num_groups.argtypes = [C.c_void_p]
    L.qmc_shard_begin.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.c_double, C.c_int]
    L.qmc_shard_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.qmc_shard_norm.argtypes = [C.c_void_p, dp, vp]
    L.qmc_shard_select.argtypes = [C.c_void_p, vp, C.c_double, C.c_int, u64p, vp]
    L.qmc_shard_final_probs.argtypes = [C.c_void_p, C.c_int, vp, vp, vp]
    L.qmc_shard_set_peers.argtypes = [C.c_void_p, C.c_int, vp]
    L.qmc_enable_peer_access.argtypes = [C.c_int, C.c_int]
    L.qmc_ipc_export.argtypes = [C.c_int, vp, vp, u64p]
    L.qmc_ipc_open.argtypes = [C.c_int, vp, C.c_uint64, C.POINTER(C.c_void_p)]
    L.qmc_ipc_close_all.argtypes = []
    L.qmc_classical_chains.argtypes = [C.c_void_p, C.c_int64, C.c_int64, u64p, C.c_uint64, C.c_uint32, C.c_double, C.c_int,
                                       vp, vp, vp, C.c_uint64, u64p, vp, u64p, i64p, i64p, vp]
    L.qmc_trajectory_stats.argtypes = [C.c_void_p, C.c_int64, C.c_int64, u64p, u64p, vp, dp, u64p, dp, dp, i64p, dp, vp]
    L.qmc_state_moments.argtypes = [C.c_void_p, C.c_int64, u64p, vp, i64p, i64p, i64p, vp]
    L.qmc_classical_chains_host.argtypes = [C.c_void_p, C.c_int64, C.c_int64, u64p, C.c_uint64, C.c_uint32, C.c_double,
                                            C.c_int, vp, vp, vp, C.c_uint64, u64p, vp, u64p, i64p, i64p]
    L.qmc_shard_op_chunk.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int64, vp]
    L.qmc_set_visit_histogram.argtypes = [C.c_void_p, i64p]
    L.qmc_profile_read_moved.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.qmc_debug_smem_slot.argtypes = [C.c_int, C.c_int, C.c_int]
    L.qmc_debug_mixer_class.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.qmc_debug_group_plan.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    L.qmc_debug_general_slot.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.qmc_debug_layout128.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    L.qmc_shard_copy_chunk.argtypes = [C.c_void_p, vp, C.c_int, C.c_int, C.c_int64, vp]
    L.qmc_shard_set_exchange_ctas.argtypes = [C.c_void_p, C.c_int]
    L.qmc_shard_exchange_can_persist.argtypes = [C.c_void_p]
    L.qmc_shard_op_exchange.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("qmc_last_error", "qmc_model_launch_count"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc == QMC_OK:
        return
    msg = lib().qmc_last_error()
    text = msg.decode(errors="replace") if msg else ""
    if rc == QMC_ERR_UNSUPPORTED:
        raise QmcUnsupported(f"libqumcmc_b200: unsupported ({text})")
    if rc == QMC_ERR_BADARG:
        raise ValueError(f"libqumcmc_b200: bad argument ({text})")
    raise QmcError(f"libqumcmc_b200: error {rc} ({text})")
