#!/usr/bin/env python
"""A/B of the exchange schedules of the sharded statevector on the GPUs of one node (one proposer, same buffers and peer
mappings, ShardedProposer.reconfigure between runs):

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 benchmarks/shard_sweep.py --local-spins 33

One JSON line per (schedule, layer count) on rank 0; the first line is the fused baseline.  --reference adds the same
shard on one GPU (weak-scaling reference)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--local-spins", type=int, default=33)
    ap.add_argument("--layers", type=int, nargs="+", default=[5])
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--chunk-bits", type=int, nargs="+", default=[3, 5])
    ap.add_argument("--ctas", type=int, nargs="+", default=[32, 64, 128])
    ap.add_argument("--dma-chunk-bits", type=int, nargs="*", default=[3, 5], help="copy-engine exchange: chunk counts 2^k")
    ap.add_argument("--reference", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import sharded_bench as sb
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import qumcmc_b200 as q
    from qumcmc_b200.sharded import ShardedProposer
    g = world.bit_length() - 1
    n = args.local_spins + g
    model = q.random_ising_model(n, 1)
    prop = ShardedProposer(model, q.GenericMixer.OneBodyMixer(n), device=local, fused=True, mode="fused")
    s = 0x5EED5EED5 & ((1 << n) - 1)
    configs = ([("fused", None, None)] + [("overlap", cb, c) for cb in args.chunk_bits for c in args.ctas]
               + [("dma", cb, None) for cb in args.dma_chunk_bits])
    base_idx = {}
    for mode, cb, ctas in configs:
        prop.reconfigure(mode, cb, ctas)
        rows = sb.measure_proposals(prop, s, args.layers, args.reps, world, rank)
        for row in rows:
            row["xchg_ctas"] = prop.xchg_ctas
            base_idx.setdefault(row["layers"], row["index"])
            row["same_index_as_fused"] = row["index"] == base_idx[row["layers"]]
            if rank == 0:
                print(json.dumps(row), flush=True)
    n_local = prop.handle.n_local
    prop.close()
    del prop
    torch.cuda.empty_cache()
    if args.reference and rank == 0:
        for r in args.layers:
            print(json.dumps({"config": "one GPU, one proposal (weak-scaling reference)", "n": n_local, "layers": r,
                              "seconds_per_proposal": sb.single_gpu_reference(n_local, r, args.reps, local)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
