"""Statevectors sharded over GPUs by global qubits (BASELINE config C5: n >= 35 over 8 B200).

Rank ``R`` of ``W = 2^g`` holds the ``2^(n-g)`` amplitudes whose top ``g`` index bits equal ``R``.  One proposal
``U = M (P M)^(r-1)`` of ``run_qmcmc_quantum_ckt`` (/root/reference/qumcmc/quantum_mcmc_routines.py:110-156) becomes a
list of local passes (``qmc_shard_op``: one CUDA kernel over the local shard each) and all-to-all exchanges on equal
chunks -- an ``all_to_all_single`` between two state buffers swaps the top ``g`` local index bits with the rank
bits, so the qubits that were global become local and get mixed by the next pass.  The layout is never swapped
back: the library keeps the logical-qubit-at-position map for both layout parities.

``sharded_plan`` is the schedule; ``ShardHandle`` wraps the C-ABI; ``ShardedProposer`` runs it over
``torch.distributed`` (NCCL on GPUs).  Sampling: every rank reduces its local sum of |a|^2, the W sums are
all-gathered, the rank that owns ``u * total`` resolves the index and broadcasts it.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _device, _lib

FIRST, MID, SINGLE, LO_SINGLE, LO_FINAL, EXCHANGE = "first", "mid", "single", "lo_single", "lo_final", "exchange"
_OPCODE = {FIRST: _lib.SHARD_OP_FIRST, MID: _lib.SHARD_OP_MID, SINGLE: _lib.SHARD_OP_SINGLE,
           LO_SINGLE: _lib.SHARD_OP_LO_SINGLE, LO_FINAL: _lib.SHARD_OP_LO_FINAL}


def sharded_plan(r: int, n_groups: int, log2_world: int) -> List[Tuple[str, int, int]]:
    """Op list ``(kind, group, pre)`` of one proposal with ``r`` mixer layers.

    The proposal starts in layout ``initial_layout(r)`` and ends in the identity layout.
    Positions 0..11 are the LO group; ``n_groups`` register groups of up to 8 cover the remaining local positions, group 0
    on top (it contains the ``g`` positions an exchange swaps with the rank bits).  Layer 1 is analytic (FIRST builds
    M|s> over all n qubits).  Every later layer j: the top group was mixed by the preceding FIRST/MID pass (after
    its phase), the other groups by SINGLE passes, the g global qubits after the exchange brought them in -- as the
    restricted first half of the next MID (which also carries the phase layer and layer j+1 of the top group), or,
    in the last layer, a restricted SINGLE followed by the read-only LO_FINAL that emits the |a|^2 segment sums.
    """
    if r <= 1:
        return [(LO_FINAL, 0, 0)]
    ops: List[Tuple[str, int, int]] = [(FIRST, 0, 0)]
    for layer in range(2, r + 1):
        last = layer == r
        if not last:
            ops.append((LO_SINGLE, 0, 0))
        ops.extend((SINGLE, k, 0) for k in range(1, n_groups))
        ops.append((EXCHANGE, 0, 0))
        if last:
            ops.append((SINGLE, 0, log2_world))
            ops.append((LO_FINAL, 0, 0))
        else:
            ops.append((MID, 0, log2_world))
    return ops


def initial_layout(r: int) -> int:
    """Layout parity a proposal with ``r`` mixer layers starts in.  There is one exchange per layer after the first and
    the layout is never swapped back, so starting in layout ``(r - 1) & 1`` (FIRST builds the analytic product state
    directly in it) ends every proposal in the identity layout: the measurement CDF runs in index order, like the
    reference's sequential sampler, without a trailing exchange."""
    return (r - 1) & 1 if r >= 2 else 0


def shard_geometry(n_local: int) -> List[Tuple[int, int]]:
    """Active position range [lo, hi) of every register group, top group first (mirrors hi_group_geometry in
    csrc/sweep.cu): positions 0..11 belong to the LO passes, full groups of 8 (HI8 passes) stack down from the top,
    the remaining 1..7 positions form the lowest group."""
    import os
    override = os.environ.get("QUMCMC_HI_GROUPS", "")
    if override:  # test hook, see hi_group_geometry: "kind:act,..." from the top; a group owns the top `act` of its window
        out, top = [], n_local
        for item in override.split(","):
            _, act = (int(x) for x in item.split(":"))
            out.append((top - act, top))
            top -= act
        if top != 12:
            raise ValueError(f"QUMCMC_HI_GROUPS={override} does not cover the {n_local - 12} upper local qubits")
        return out
    hi_qubits = n_local - 12
    out = [(n_local - 8 * (k + 1), n_local - 8 * k) for k in range(hi_qubits // 8)]
    if hi_qubits % 8:
        out.append((12, 12 + hi_qubits % 8))
    return out


def max_chunk_bits(n_local: int) -> int:
    """Largest chunk_bits for which every pass below the top group still fits inside a chunk (mirrors the check in
    qmc_shard_op_chunk): the chunk keeps the LO tile (13 positions) and the windows of the groups k >= 1."""
    geo = shard_geometry(n_local)
    top = max([13] + [hi for _, hi in geo[1:]])
    return n_local - top


def chunk_order(chunk_bits: int, log2_world: int, rank: int) -> List[int]:
    """Chunks of a pipelined exchange pass in issue order: the destination rank (top ``log2_world`` bits of the chunk
    index) cycles fastest and starts at ``rank + 1``, so at every step all ranks store to different peers."""
    W = 1 << log2_world
    per = 1 << (chunk_bits - log2_world)
    return [(((d + rank + 1) % W) << (chunk_bits - log2_world)) | j for j in range(per) for d in range(W)]


def plan_traffic_bytes(ops: Sequence[Tuple[str, int, int]], n_local: int, world: int) -> Tuple[float, float]:
    """(HBM bytes moved by the local passes, bytes sent over NVLink) per rank for a plan."""
    amp = 8.0 * (1 << n_local)
    hbm = nv = 0.0
    for kind, _, _ in ops:
        if kind == EXCHANGE:
            nv += amp * (world - 1) / world
            hbm += 2.0 * amp
        elif kind == FIRST:
            hbm += amp
        elif kind == LO_FINAL:
            hbm += amp
        else:
            hbm += 2.0 * amp
    return hbm, nv


class ShardHandle:
    """``qmc_shard`` of one rank: local passes, norm and inverse-CDF selection on that rank's buffers."""

    def __init__(self, model, mixer, rank: int, log2_world: int, device: int = 0):
        self.model, self.rank, self.g, self.device = model, int(rank), int(log2_world), int(device)
        self.n = model.num_spins
        self.n_local = self.n - self.g
        self.dm = model.device_model(device)
        self.dm.set_mixer(mixer)
        h = C.c_void_p()
        _lib.check(_lib.lib().qmc_shard_create(C.byref(h), self.dm.handle, self.rank, self.g))
        self.handle = h
        self.n_groups = int(_lib.lib().qmc_shard_num_groups(h))

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().qmc_shard_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def begin(self, s: int, gamma: float, r: int, delta_time: float, group: int = 0):
        _lib.check(_lib.lib().qmc_shard_begin(self.handle, int(s), float(gamma), int(r), float(delta_time), int(group)))

    def op(self, kind: str, k: int, pre: int, layout: int, state) -> None:
        _lib.check(_lib.lib().qmc_shard_op(self.handle, _OPCODE[kind], int(k), int(pre), int(layout), state.data_ptr(),
                                           _device.current_stream_ptr(self.device)))

    def set_peers(self, which: int, pointers: Sequence[int]) -> None:
        """Device-accessible addresses of buffer ``which`` of every rank (own buffer included)."""
        arr = (C.c_void_p * len(pointers))(*[C.c_void_p(int(p_)) for p_ in pointers])
        _lib.check(_lib.lib().qmc_shard_set_peers(self.handle, int(which), arr))

    def op_exchange(self, kind: str, k: int, pre: int, layout: int, state, dst_which: int) -> None:
        """The pass ``kind`` fused with the qubit-swap exchange: its stores land in buffer ``dst_which`` of the
        ranks that own the amplitudes after the swap."""
        _lib.check(_lib.lib().qmc_shard_op_exchange(self.handle, _OPCODE[kind], int(k), int(pre), int(layout),
                                                    state.data_ptr(), int(dst_which),
                                                    _device.current_stream_ptr(self.device)))

    def copy_chunk(self, state, dst_which: int, chunk_bits: int, chunk: int) -> None:
        """Copy-engine exchange of one chunk into buffer ``dst_which`` of the owning rank (no SM involved)."""
        _lib.check(_lib.lib().qmc_shard_copy_chunk(self.handle, state.data_ptr(), int(dst_which), int(chunk_bits), int(chunk),
                                                   _device.current_stream_ptr(self.device)))

    def op_chunk(self, kind: str, k: int, pre: int, layout: int, state, dst_which: int, chunk_bits: int, chunk: int) -> None:
        """One chunk of a SINGLE / LO_SINGLE pass; ``dst_which`` -1 = local, 0 / 1 = fused with the exchange."""
        _lib.check(_lib.lib().qmc_shard_op_chunk(self.handle, _OPCODE[kind], int(k), int(pre), int(layout),
                                                 state.data_ptr(), int(dst_which), int(chunk_bits), int(chunk),
                                                 _device.current_stream_ptr(self.device)))

    def set_exchange_ctas(self, ctas: int) -> None:
        """Overlap mode: exchange passes of register-only (HI6) groups as a persistent kernel on ``ctas`` CTAs (0 = off)."""
        _lib.check(_lib.lib().qmc_shard_set_exchange_ctas(self.handle, int(ctas)))

    def exchange_can_persist(self) -> bool:
        return bool(_lib.lib().qmc_shard_exchange_can_persist(self.handle))

    def norm(self, total) -> None:
        _lib.check(_lib.lib().qmc_shard_norm(self.handle, total.data_ptr(), _device.current_stream_ptr(self.device)))

    def final_probabilities(self, state, layout: int, probs) -> None:
        """Test hook: |a|^2 of the local shard after the last mixer half-layer (float32, local index order)."""
        _lib.check(_lib.lib().qmc_shard_final_probs(self.handle, int(layout), state.data_ptr(), probs.data_ptr(),
                                                    _device.current_stream_ptr(self.device)))

    def select(self, state, target: float, layout: int, index) -> None:
        _lib.check(_lib.lib().qmc_shard_select(self.handle, state.data_ptr(), float(target), int(layout),
                                               index.data_ptr(), _device.current_stream_ptr(self.device)))


def pick_owner(totals: np.ndarray, u: float) -> Tuple[int, float]:
    """Rank whose slice of the global running sum of |a|^2 contains ``u * total`` and the target inside it.
    Ranks are in index order (rank = top bits), matching the sequential cumulative sum of
    QuantumState.sampling (call at qumcmc/quantum_mcmc_routines.py:153)."""
    totals = np.asarray(totals, dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(totals)])
    target = float(u) * cum[-1]
    owner = int(np.searchsorted(cum[1:], target, side="right"))
    owner = min(owner, len(totals) - 1)
    return owner, target - cum[owner]


# Schedule defaults of ShardedProposer, measured on 8 B200 at n = 36 (profiles/r2j_shard_sweep_8gpu.jsonl), 5-layer proposal:
# fused 0.645 s, overlap (8-16 chunks, 222 persistent exchange CTAs) 0.552 s, dma (32 chunks) 0.513 s; 2-layer proposal:
# 0.168 / 0.156 / 0.148 s.  On 2 B200 (n = 34) the three are within 4 % (0.513 / 0.492 / 0.508 s).
DEFAULT_MODE = "dma"
DEFAULT_XCHG_CTAS = 222
DEFAULT_OVERLAP_CHUNK_BITS = 5


class ShardedProposer:
    """One process per GPU (torchrun); every rank calls :meth:`propose` with the same arguments and gets the same
    measured bitstring back."""

    def __init__(self, model, mixer, device: int = 0, group=None, fused: bool = True, chunk_bits: Optional[int] = None,
                 mode: Optional[str] = None, xchg_ctas: Optional[int] = None):
        """``fused``: the pass before every exchange stores straight into the peers' buffers over NVLink (CUDA IPC
        mappings of the two state buffers of every rank) and the exchange shrinks to a barrier; ``False`` keeps the
        pass local and calls NCCL ``all_to_all_single``.

        ``mode``: ``"fused"`` -- one kernel per pass over the whole shard, the exchange pass serial after the local ones;
        ``"overlap"`` -- the passes before each exchange run chunk by chunk (``chunk_bits``), the local ones on one
        stream and the NVLink-bound exchange pass as a persistent kernel on ``xchg_ctas`` CTAs on a second (higher
        priority) stream, so that HBM-bound and NVLink-bound work proceed side by side; ``"dma"`` -- every pass stays local
        and each finished chunk is sent by the copy engines (``cudaMemcpyAsync`` into the peer mapping) while the following
        chunks are computed: no SM is spent on the exchange; ``None`` = the measured default
        (environment override QUMCMC_SHARD_MODE / QUMCMC_SHARD_CTAS / QUMCMC_SHARD_CHUNK_BITS for A/B runs)."""
        import os
        if mode is None:
            mode = os.environ.get("QUMCMC_SHARD_MODE") or DEFAULT_MODE
        if mode not in ("fused", "overlap", "dma"):
            raise ValueError(f"unknown sharded mode {mode!r}")
        if xchg_ctas is None:
            xchg_ctas = int(os.environ.get("QUMCMC_SHARD_CTAS", DEFAULT_XCHG_CTAS))
        if chunk_bits is None and os.environ.get("QUMCMC_SHARD_CHUNK_BITS"):
            chunk_bits = int(os.environ["QUMCMC_SHARD_CHUNK_BITS"])
        import torch
        import torch.distributed as dist

        self.dist, self.group, self._torch = dist, group, torch
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        g = self.world.bit_length() - 1
        if self.world < 2 or (1 << g) != self.world:
            raise ValueError(f"world size must be a power of two >= 2 (got {self.world})")
        self.handle = ShardHandle(model, mixer, self.rank, g, device)
        dev = _device.require_cuda(device)
        self.dev = dev
        amps = 1 << self.handle.n_local
        self.buf = [torch.empty(amps, dtype=torch.complex64, device=dev) for _ in range(2)]
        self.cur = 0
        self.layout = 0
        self.total = torch.zeros(1, dtype=torch.float64, device=dev)
        self.index = torch.zeros(1, dtype=torch.int64, device=dev)
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.exchanges = 0
        self.timeline = None        # set to a list to collect (label, start_event, end_event) per op
        self.fused = bool(fused)
        # pipelined exchange: the passes of a layer that precede the exchange run chunk by chunk, the last of them storing
        # into the peers' buffers on a second stream while the next chunk's local passes run (chunk_bits=0 disables)
        cmax = max_chunk_bits(self.handle.n_local)
        # (round 1's chunked pipeline with a FULL-grid exchange kernel gained nothing: 0.653 s vs 0.647 s unchunked -- the
        # NVLink-bound CTAs took every slot; the persistent-kernel and copy-engine schedules below are what overlaps)
        if mode == "overlap" and (not self.fused or not self.handle.exchange_can_persist() or cmax < g):
            mode = "fused"   # no persistent variant of this shard's exchange pass (or nothing to chunk): serial schedule
        if mode == "dma" and (not self.fused or cmax < g):
            mode = "fused"
        if mode in ("overlap", "dma") and chunk_bits is None:
            chunk_bits = max(g, min(DEFAULT_OVERLAP_CHUNK_BITS, cmax))
        self.mode = mode
        want = 0 if chunk_bits is None else min(int(chunk_bits), cmax)
        self.chunk_bits = want if (self.fused and want >= g) else 0
        if mode == "overlap":
            self.handle.set_exchange_ctas(int(xchg_ctas))
        self.xchg_ctas = int(xchg_ctas) if mode == "overlap" else 0
        self._make_streams()
        if self.fused:
            self._map_peer_buffers(device)

    def _make_streams(self) -> None:
        """Side streams of the chunked schedules.  The exchange stream has the HIGHER priority: its kernels are small
        (overlap mode: a bounded number of persistent CTAs) and must get their CTA slots as soon as a chunk's local
        passes are done; the local passes (thousands of CTAs, queued chunk after chunk) fill every other slot.  With the
        priorities the other way round the block scheduler serves the queued local-pass CTAs first and the exchange only
        starts when all of them have drained -- a serial schedule (measured: 2 GPUs, n = 34, no gain)."""
        torch = self._torch
        if self.chunk_bits and getattr(self, "_xstream", None) is None:
            self._xstream = torch.cuda.Stream(device=self.dev, priority=-1)   # exchange passes
            self._cstream = torch.cuda.Stream(device=self.dev)                # local passes
        elif not self.chunk_bits and not hasattr(self, "_xstream"):
            self._xstream = self._cstream = None

    def reconfigure(self, mode: str, chunk_bits: Optional[int] = None, xchg_ctas: Optional[int] = None) -> None:
        """Switch the exchange schedule of an existing proposer (same buffers and peer mappings): A/B runs."""
        torch, g = self._torch, self.handle.g
        cmax = max_chunk_bits(self.handle.n_local)
        if mode == "overlap" and (not self.fused or not self.handle.exchange_can_persist() or cmax < g):
            mode = "fused"
        if mode == "dma" and (not self.fused or cmax < g):
            mode = "fused"
        if mode in ("overlap", "dma") and chunk_bits is None:
            chunk_bits = max(g, min(DEFAULT_OVERLAP_CHUNK_BITS, cmax))
        want = 0 if chunk_bits is None else min(int(chunk_bits), cmax)
        self.mode = mode
        self.chunk_bits = want if (self.fused and want >= g) else 0
        self.xchg_ctas = int(xchg_ctas if xchg_ctas is not None else DEFAULT_XCHG_CTAS) if mode == "overlap" else 0
        self.handle.set_exchange_ctas(self.xchg_ctas)
        self._make_streams()

    def _map_peer_buffers(self, device: int) -> None:
        """all-gather CUDA IPC handles of both state buffers, map the peers' buffers into this device's context
        (qmc_ipc_export / qmc_ipc_open) and hand the address tables to the library."""
        dist, L = self.dist, _lib.lib()
        mine = []
        for b in self.buf:
            hb = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
            off = C.c_uint64(0)
            _lib.check(L.qmc_ipc_export(int(device), C.c_void_p(b.data_ptr()), hb, C.byref(off)))
            mine.append((hb.raw, int(off.value)))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (int(device), mine), group=self.group)
        for which in (0, 1):
            ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(self.buf[which].data_ptr())
                    continue
                peer_dev, handles = gathered[r]
                raw, off = handles[which]
                out = C.c_void_p()
                _lib.check(L.qmc_ipc_open(int(device), raw, C.c_uint64(off), C.byref(out)))
                ptrs.append(int(out.value))
            self.handle.set_peers(which, ptrs)
        dist.barrier(group=self.group)

    def close(self) -> None:
        """Unmap the peers' buffers (every rank, before any rank frees its state buffers) and free the handle."""
        if self.fused:
            self._torch.cuda.synchronize(self.dev)
            self.dist.barrier(group=self.group)
            _lib.lib().qmc_ipc_close_all()
            self.dist.barrier(group=self.group)
            self.fused = False
        self.handle.close()

    def _timed(self, label: str, fn) -> None:
        """Run one op; with ``self.timeline`` set to a list, bracket it with CUDA events (label, start, end)."""
        if self.timeline is None:
            fn()
            return
        torch = _device._torch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        self.timeline.append((label, e0, e1))

    def _pipelined(self, run) -> None:
        """``run`` = the LO_SINGLE / SINGLE passes before an exchange: per chunk, all but the last on the current
        stream, the last (fused with the exchange) on the side stream once the chunk's local passes are done."""
        torch, h = self._torch, self.handle
        main = torch.cuda.current_stream(self.dev)
        src, dst = self.buf[self.cur], 1 - self.cur
        self._xstream.wait_stream(main)
        self._cstream.wait_stream(main)
        dma = self.mode == "dma"
        for c in chunk_order(self.chunk_bits, h.g, self.rank):
            ev = None
            local = run if dma else run[:-1]
            if local:
                with torch.cuda.stream(self._cstream):
                    for kind, k, pre in local:
                        h.op_chunk(kind, k, pre, self.layout, src, -1, self.chunk_bits, c)
                    ev = torch.cuda.Event()
                    ev.record(self._cstream)
            kind, k, pre = run[-1]
            with torch.cuda.stream(self._xstream):
                if ev is not None:
                    self._xstream.wait_event(ev)
                if dma:   # copy engines: the chunk's amplitudes go to the owning rank's other buffer
                    h.copy_chunk(src, dst, self.chunk_bits, c)
                else:
                    h.op_chunk(kind, k, pre, self.layout, src, dst, self.chunk_bits, c)
        main.wait_stream(self._cstream)
        main.wait_stream(self._xstream)
        self.dist.all_reduce(self._flag, group=self.group)   # barrier in stream order: every rank's stores have landed

    def propose(self, s: int, gamma: float, r: int, u: float, delta_time: float = 0.8, group_id: int = 0) -> int:
        torch, dist, h = _device._torch(), self.dist, self.handle
        h.begin(s, gamma, r, delta_time, group_id)
        ops = sharded_plan(r, h.n_groups, h.g)
        self.layout = initial_layout(r)
        i = 0
        while i < len(ops):
            kind, k, pre = ops[i]
            if self.chunk_bits and kind in (LO_SINGLE, SINGLE) and not (kind == SINGLE and k == 0):
                # run of chunkable passes that ends in an exchange?
                j = i
                while j < len(ops) and (ops[j][0] == LO_SINGLE or (ops[j][0] == SINGLE and ops[j][1] > 0)):
                    j += 1
                if j < len(ops) and ops[j][0] == EXCHANGE:
                    self._timed("pipelined passes+exchange", lambda lo=i, hi=j: self._pipelined(ops[lo:hi]))
                    self.cur ^= 1
                    self.layout ^= 1
                    self.exchanges += 1
                    i = j + 1
                    continue
            if self.fused and i + 1 < len(ops) and ops[i + 1][0] == EXCHANGE:
                def fused_step(kind=kind, k=k, pre=pre):
                    h.op_exchange(kind, k, pre, self.layout, self.buf[self.cur], 1 - self.cur)
                    dist.all_reduce(self._flag, group=self.group)   # barrier in stream order: every rank's stores have landed
                self._timed("pass+exchange", fused_step)
                self.cur ^= 1
                self.layout ^= 1
                self.exchanges += 1
                i += 2
                continue
            if kind == EXCHANGE:
                self._timed("exchange", lambda: dist.all_to_all_single(self.buf[1 - self.cur], self.buf[self.cur],
                                                                       group=self.group))
                self.cur ^= 1
                self.layout ^= 1
                self.exchanges += 1
            else:
                self._timed("pass", lambda: h.op(kind, k, pre, self.layout, self.buf[self.cur]))
            i += 1
        h.norm(self.total)
        totals = [torch.empty_like(self.total) for _ in range(self.world)]
        dist.all_gather(totals, self.total, group=self.group)
        owner, target = pick_owner(np.array([float(t.item()) for t in totals]), u)
        if self.rank == owner:
            h.select(self.buf[self.cur], target, self.layout, self.index)
        dist.broadcast(self.index, src=dist.get_global_rank(self.group, owner) if self.group is not None else owner,
                       group=self.group)
        return int(self.index.item()) & ((1 << h.n) - 1) if h.n < 63 else int(self.index.item())

