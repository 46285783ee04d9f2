#!/usr/bin/env python
"""Config C5: statevector sharded over the GPUs of one node by global qubits (NCCL all-to-all qubit swaps) plus
the sharded exact enumeration (all-gather of (max, sumexp) pairs = the allreduce for Z).

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 benchmarks/sharded_bench.py --spins 36 --layers 2 5

One JSON line per layer count on rank 0: seconds per proposal (device events, max over ranks), HBM GB/s of the local
passes, NVLink GB/s per rank of the exchanges, total norm, measured index (identical on every rank)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def measure_proposals(prop, s, layers, reps, world, rank):
    """Rows (one per layer count) of the sharded proposal: seconds per proposal (device events, max over ranks), the
    HBM rate of the local passes, the NVLink rate per rank and direction of the exchanges, total norm, measured index."""
    import torch
    import torch.distributed as dist
    from qumcmc_b200.sharded import EXCHANGE, plan_traffic_bytes, sharded_plan
    h = prop.handle
    rows = []
    for r in layers:
        prop.propose(s, 0.45, r, 0.5)  # warm-up (allocations, NCCL channels)
        best = None
        for rep in range(reps):
            prop.timeline = []
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            idx = prop.propose(s, 0.45, r, 0.37 + 0.1 * rep)
            e1.record()
            torch.cuda.synchronize()
            per = {}
            for label, a, b in prop.timeline:
                per[label] = per.get(label, 0.0) + a.elapsed_time(b)
            t = torch.tensor([e0.elapsed_time(e1), per.get("pass", 0.0), per.get("exchange", 0.0),
                              per.get("pass+exchange", 0.0) + per.get("pipelined passes+exchange", 0.0) + per.get("overlapped layer", 0.0)],
                             dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            row = [float(x) for x in t.tolist()] + [idx]
            if best is None or row[0] < best[0]:
                best = row
        prop.timeline = None
        ms, pass_ms, ex_ms, fx_ms, idx = best
        ops = sharded_plan(r, h.n_groups, h.g)
        amp = 8.0 * (1 << h.n_local)
        n_ex = sum(1 for o in ops if o[0] == EXCHANGE)
        hbm, nv = plan_traffic_bytes(ops, h.n_local, world)
        hbm_local = hbm - 2.0 * amp * n_ex  # the in-place passes; without the copies NCCL makes
        tot = [torch.empty_like(prop.total) for _ in range(world)]
        dist.all_gather(tot, prop.total)
        line = {"config": "C5 sharded statevector", "n": h.n, "world": world, "n_local": h.n_local, "layers": r,
                "fused_exchange": bool(prop.fused), "mode": getattr(prop, "mode", None), "chunk_bits": prop.chunk_bits,
                "seconds_per_proposal": ms * 1e-3, "exchanges": n_ex,
                "local_passes": len(ops) - n_ex, "total_norm": float(sum(float(t.item()) for t in tot)),
                "index": idx, "state_GiB_per_rank": amp / 2 ** 30,
                "hbm_GB_per_rank": hbm_local / 1e9, "nvlink_GB_per_rank_per_direction": nv / 1e9,
                "whole_proposal_hbm_GBps": hbm_local / 1e9 / max(1e-9, ms * 1e-3),
                "whole_proposal_nvlink_GBps_per_direction": nv / 1e9 / max(1e-9, ms * 1e-3)}
        if prop.fused and getattr(prop, "mode", "") == "overlap":
            line.update({"pass_seconds": pass_ms * 1e-3, "overlapped_layer_seconds": fx_ms * 1e-3})
        elif prop.fused and prop.chunk_bits:   # passes before each exchange pipelined against the NVLink stores, chunk by chunk
            line.update({"pass_seconds": pass_ms * 1e-3, "pipelined_passes_exchange_seconds": fx_ms * 1e-3,
                         "nvlink_GBps_per_rank": (nv / 1e9) / max(1e-9, fx_ms * 1e-3)})
        elif prop.fused:   # the pass before each exchange stores over NVLink: its time is pass + exchange + barrier
            line.update({"pass_seconds": pass_ms * 1e-3, "pass_exchange_seconds": fx_ms * 1e-3,
                         "local_pass_hbm_GBps": (hbm_local - 2.0 * amp * n_ex) / 1e9 / max(1e-9, pass_ms * 1e-3),
                         "pass_exchange_GBps_per_rank": (amp * n_ex / 1e9) / max(1e-9, fx_ms * 1e-3),
                         "nvlink_GBps_per_rank": (nv / 1e9) / max(1e-9, fx_ms * 1e-3)})
        else:
            line.update({"pass_seconds": pass_ms * 1e-3, "exchange_seconds": ex_ms * 1e-3,
                         "local_pass_hbm_GBps": hbm_local / 1e9 / max(1e-9, pass_ms * 1e-3),
                         "nvlink_GBps_per_rank": (nv / 1e9) / max(1e-9, ex_ms * 1e-3)})
        rows.append(line)
    return rows


def single_gpu_reference(n, r, reps=2, device=0):
    """One proposal with r mixer layers of an n-qubit statevector held by ONE GPU (qmc_propose): with n = n_local of a
    sharded run this is the weak-scaling reference (same bytes per GPU, no exchange)."""
    import torch
    import qumcmc_b200 as q
    from qumcmc_b200 import _device, _lib
    m = q.random_ising_model(n, 1)
    dm = m.device_model(device)
    dm.set_mixer(q.GenericMixer.OneBodyMixer(n))
    dev = torch.device("cuda", device)
    t = lambda x: torch.from_numpy(x).to(dev)
    s_in = t(np.array([0x5EED5EED5 & ((1 << n) - 1)], dtype=np.int64))
    gam, rr, uu = t(np.array([0.45])), t(np.array([r], dtype=np.int32)), t(np.array([0.37]))
    out = torch.empty(1, dtype=torch.int64, device=dev)
    call = lambda: _lib.check(_lib.lib().qmc_propose(dm.handle, 1, s_in.data_ptr(), gam.data_ptr(), rr.data_ptr(),
                                                     uu.data_ptr(), None, 0.8, _lib.DTYPE_F32, out.data_ptr(),
                                                     _device.current_stream_ptr(device)))
    call()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    del dm
    _device._cache.clear()      # frees the 2^n workspace
    torch.cuda.empty_cache()
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spins", dest="n", type=int, default=36)
    ap.add_argument("--layers", type=int, nargs="+", default=[2, 5])
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--logz", action="store_true")
    ap.add_argument("--chunk-bits", type=int, default=-1, help="pipelined exchange: 2^k chunks (-1: default, 0: off)")
    ap.add_argument("--fused", type=int, default=1, help="1: passes store into the peers' buffers (NVLink P2P); 0: NCCL all_to_all_single")
    ap.add_argument("--mode", default=None, help="exchange schedule of the fused path (ShardedProposer mode=...)")
    ap.add_argument("--reference", action="store_true", help="also time n_local qubits on one GPU (rank 0) for the weak-scaling ratio")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import qumcmc_b200 as q
    from qumcmc_b200.distributed import exact_logz_sharded
    from qumcmc_b200.sharded import ShardedProposer

    model = q.random_ising_model(args.n, 1)
    kw = {} if args.mode is None else {"mode": args.mode}
    prop = ShardedProposer(model, q.GenericMixer.OneBodyMixer(args.n), device=local, fused=bool(args.fused),
                           chunk_bits=None if args.chunk_bits < 0 else args.chunk_bits, **kw)
    s = 0x5EED5EED5 & ((1 << args.n) - 1)
    rows = measure_proposals(prop, s, args.layers, args.reps, world, rank)
    if rank == 0:
        for row in rows:
            print(json.dumps(row), flush=True)
    if args.logz:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lz = exact_logz_sharded(model, 1.0, device=local)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            print(json.dumps({"config": "C5 sharded exact log Z", "n": args.n, "world": world, "logZ": lz, "seconds": dt,
                              "Gstates_per_s": (1 << args.n) / dt / 1e9}), flush=True)
    n_local = prop.handle.n_local
    prop.close()
    del prop
    torch.cuda.empty_cache()
    if args.reference and rank == 0:
        for r in args.layers:
            print(json.dumps({"config": "one GPU, one proposal (weak-scaling reference)", "n": n_local, "layers": r,
                              "seconds_per_proposal": single_gpu_reference(n_local, r, args.reps, local)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
