#!/usr/bin/env python
"""BASELINE config C3 at full length: random_ising_model(20, seed=1), 4096 independent chains x 10 000 hops,
chain-sharded over the GPUs of one node (torchrun), one gather of the statistics at the end.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 benchmarks/c3_full.py [--hops 10000]

One JSON line on rank 0: wall seconds (device events, max over ranks), proposals/s, acceptance rate, mean proposal
Hamming distance, and a checksum of the final states (identical for any number of GPUs: Philox keyed by global chain id)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=4096)
    ap.add_argument("--hops", type=int, default=10000)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import qumcmc_b200 as q
    from qumcmc_b200.distributed import gather_statistics, shard_range
    from qumcmc_b200.quantum_mcmc_routines import run_chains

    n = 20
    model = q.random_ising_model(n, 1)
    mixer = q.GenericMixer.OneBodyMixer(n)
    lo, hi = shard_range(args.chains, rank, world)
    run_chains(model, mixer, (0.25, 0.6), 2, hi - lo, None, 1.0, 0.8, 0x5EED, device=local, chain_id0=lo, record=False)  # warm
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    b = run_chains(model, mixer, (0.25, 0.6), args.hops, hi - lo, None, 1.0, 0.8, 0x5EED, device=local, chain_id0=lo,
                   record=False)
    acc, hist = gather_statistics(b.acc_count, b.hamming_hist, args.chains) if world > 1 else (b.acc_count, b.hamming_hist)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
    chk = torch.tensor([int(np.bitwise_xor.reduce(b.final_state.astype(np.uint64)) & np.uint64(0x7FFFFFFFFFFFFFFF))],
                       dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parts = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(parts, chk)
        x = 0
        for p_ in parts:
            x ^= int(p_.item())
    else:
        x = int(chk.item())
    if rank == 0:
        total = args.chains * args.hops
        hs = np.asarray(hist).sum(axis=0)
        print(json.dumps({"config": "C3 full length: n=20, %d chains x %d hops" % (args.chains, args.hops), "n_gpus": world,
                          "seconds": float(t.item()), "proposals_per_s": total / float(t.item()),
                          "acceptance_rate": float(np.asarray(acc).sum()) / total,
                          "mean_proposal_hamming": float((hs * np.arange(n + 1)).sum() / hs.sum()),
                          "final_state_xor": x}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
