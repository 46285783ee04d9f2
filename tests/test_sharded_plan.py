"""Sharded statevector (config C5) host logic on CPU: the op schedule, the layout bookkeeping and the exchange
are executed with the numpy shard simulator (tests/shard_sim.py) and compared with the oracle's U|s>."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle
from qumcmc_b200.sharded import EXCHANGE, initial_layout, pick_owner, plan_traffic_bytes, shard_geometry, sharded_plan
from tests import models
from tests.shard_sim import ShardSim, logical_indices


def _run_single_process(J, h, s, gamma, r, g, dt=0.8):
    """all ranks in one process; the exchange is the chunk transpose an all_to_all_single performs"""
    n = len(h)
    W, nl = 1 << g, n - g
    al = oracle.alpha(J, h)
    sims = [ShardSim(J, h, al, R, g) for R in range(W)]
    for sm in sims:
        sm.begin(s, gamma, r, dt)
    layout = initial_layout(r)
    for kind, k, pre in sharded_plan(r, len(shard_geometry(nl)), g):
        if kind == EXCHANGE:
            chunks = np.stack([sm.state.reshape(W, -1) for sm in sims])          # [rank][chunk][...]
            for R, sm in enumerate(sims):
                sm.state = np.ascontiguousarray(chunks[:, R, :]).reshape(-1)      # out_R[c] = in_c[R]
            layout ^= 1
        else:
            for sm in sims:
                sm.op(kind, k, pre, layout)
    p = np.zeros(1 << n)
    for R, sm in enumerate(sims):
        p[logical_indices(R, nl, g, layout)] = sm.probabilities()
    return p, layout


def test_plan_shape_and_final_layout():
    for g in (1, 2, 3):
        for K in (1, 2, 4):
            assert sharded_plan(1, K, g) == [("lo_final", 0, 0)]
            for r in range(2, 9):
                ops = sharded_plan(r, K, g)
                kinds = [o[0] for o in ops]
                assert kinds[0] == "first" and kinds[-1] == "lo_final"
                assert kinds.count("exchange") == r - 1                     # one exchange per layer after the first ...
                assert (initial_layout(r) + kinds.count("exchange")) % 2 == 0  # ... and the identity layout at the end
                assert kinds.count("first") + kinds.count("mid") == r - 1   # one phase layer per pass with P
                assert kinds.count("lo_single") == r - 2
    hbm, nv = plan_traffic_bytes(sharded_plan(5, 4, 3), 33, 8)
    assert nv == 4 * 8.0 * 2 ** 33 * 7 / 8 and hbm > nv


def test_geometry_covers_every_local_position_once():
    for nl in range(18, 35):
        covered = list(range(12))
        for lo, hi in shard_geometry(nl):
            covered += list(range(lo, hi))
        assert sorted(covered) == list(range(nl))
        top = shard_geometry(nl)[0]
        assert top[1] == nl and top[1] - top[0] >= 6                     # the top group holds the exchanged positions


def test_pick_owner_is_the_sequential_cdf():
    totals = np.array([0.1, 0.0, 0.6, 0.3])
    assert pick_owner(totals, 0.05)[0] == 0
    o, t = pick_owner(totals, 0.1 + 1e-12)
    assert o == 2 and abs(t - 1e-12) < 1e-15
    assert pick_owner(totals, 0.999999)[0] == 3
    assert pick_owner(totals, 1.0)[0] == 3


@pytest.mark.parametrize("n,g,r", [(19, 1, 1), (19, 1, 2), (19, 1, 3), (20, 2, 4), (21, 1, 5), (21, 3, 2), (20, 2, 6)])
def test_sharded_schedule_reproduces_oracle(n, g, r):
    J, h = models.packaged_random_model(n, 11)
    masks, w = models.one_body(n)
    rng = np.random.default_rng(100 * n + 10 * g + r)
    s = int(rng.integers(0, 1 << n))
    gamma = float(rng.uniform(0.25, 0.6))
    p, layout = _run_single_process(J, h, s, gamma, r, g)
    assert layout == 0
    ref = oracle.transition_probs(J, h, oracle.alpha(J, h), gamma, 0.8, r, masks, w, s, gatewise=False)
    assert abs(p.sum() - 1.0) < 1e-10
    assert np.abs(p - ref).max() < 1e-12


def _gloo_worker(rank, world, port, n, r, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = world.bit_length() - 1
        nl = n - g
        J, h = models.packaged_random_model(n, 11)
        s, gamma = 123457 % (1 << n), 0.41
        sim = ShardSim(J, h, oracle.alpha(J, h), rank, g)
        sim.begin(s, gamma, r, 0.8)
        layout = initial_layout(r)
        for kind, k, pre in sharded_plan(r, len(shard_geometry(nl)), g):
            if kind == EXCHANGE:
                src = torch.from_numpy(np.ascontiguousarray(sim.state).view(np.float64))
                dst = torch.empty_like(src)
                dist.all_to_all_single(dst, src)
                sim.state = dst.numpy().view(np.complex128).copy()
                layout ^= 1
            else:
                sim.op(kind, k, pre, layout)
        total = torch.tensor([sim.probabilities().sum()], dtype=torch.float64)
        tot = [torch.empty_like(total) for _ in range(world)]
        dist.all_gather(tot, total)
        q.put((rank, layout, sim.probabilities(), [float(t.item()) for t in tot]))
    finally:
        dist.destroy_process_group()


def test_sharded_schedule_gloo_world2():
    """two processes, real all_to_all_single over gloo: same probabilities as the oracle, totals all-gathered"""
    import torch.multiprocessing as mp
    n, r, world = 19, 3, 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(rk, world, port, n, r, q)) for rk in range(world)]
    for p_ in procs:
        p_.start()
    res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda t: t[0])
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    J, h = models.packaged_random_model(n, 11)
    masks, w = models.one_body(n)
    ref = oracle.transition_probs(J, h, oracle.alpha(J, h), 0.41, 0.8, r, masks, w, 123457 % (1 << n), gatewise=False)
    p = np.zeros(1 << n)
    for rank, layout, probs, totals in res:
        assert layout == 0
        p[logical_indices(rank, n - 1, 1, layout)] = probs
        assert abs(sum(totals) - 1.0) < 1e-10
    assert np.abs(p - ref).max() < 1e-12


def test_chunk_order_rotates_destinations_and_covers_every_chunk():
    from qumcmc_b200.sharded import chunk_order, max_chunk_bits
    for g in (1, 2, 3):
        W = 1 << g
        for cb in range(g, g + 3):
            orders = [chunk_order(cb, g, rank) for rank in range(W)]
            for rank, order in enumerate(orders):
                assert sorted(order) == list(range(1 << cb))                    # every chunk exactly once
            for step in range(1 << cb):
                dests = [orders[rank][step] >> (cb - g) for rank in range(W)]  # destination = top g bits of the chunk
                assert sorted(dests) == list(range(W))                          # all ranks store to different peers
    # a chunk must keep the LO tile (13 positions) and the windows of the groups below the top one
    for nl in range(18, 35):
        geo = shard_geometry(nl)
        cb = max_chunk_bits(nl)
        assert nl - cb >= 13 and all(hi <= nl - cb for _, hi in geo[1:])
        assert cb >= 0
