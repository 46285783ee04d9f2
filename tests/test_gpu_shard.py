"""GPU parity for the sharded statevector ops (qmc_shard_*, config C5).

Single-GPU tests emulate W ranks in one process: W handles, W pairs of buffers on cuda:0, the exchange done with
the same chunk transpose an all_to_all_single performs.  The NCCL test needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest

import oracle
from tests import models
from tests.shard_sim import logical_indices

pytestmark = pytest.mark.gpu


def _emulated_proposal(n, g, r, s, gamma, seed=11, us=(), fused=False, chunk_bits=0, on_device=False, xchg_ctas=0, dma=False):
    import torch
    import qumcmc_b200 as q
    from qumcmc_b200.sharded import EXCHANGE, ShardHandle, initial_layout, pick_owner, sharded_plan
    J, h = models.packaged_random_model(n, seed)
    model = q.IsingEnergyFunction(J, h)
    mixer = q.GenericMixer.OneBodyMixer(n)
    W, nl = 1 << g, n - g
    dev = torch.device("cuda", 0)
    hs = [ShardHandle(model, mixer, R, g, 0) for R in range(W)]
    bufs = [torch.zeros(1 << nl, dtype=torch.complex64, device=dev) for _ in range(W)]
    for hd in hs:
        hd.begin(s, gamma, r, 0.8)
        if xchg_ctas:
            assert hd.exchange_can_persist()
            hd.set_exchange_ctas(xchg_ctas)
    layout = initial_layout(r)
    ops = sharded_plan(r, hs[0].n_groups, g)
    if fused:   # the pass before every exchange stores straight into the owning "rank"'s other buffer (same device here)
        both = [bufs, [torch.zeros(1 << nl, dtype=torch.complex64, device=dev) for _ in range(W)]]
        for hd in hs:
            for which in (0, 1):
                hd.set_peers(which, [b.data_ptr() for b in both[which]])
        from qumcmc_b200.sharded import LO_SINGLE, SINGLE, chunk_order
        cur, i = 0, 0
        while i < len(ops):
            kind, k, pre = ops[i]
            j = i
            while chunk_bits and j < len(ops) and (ops[j][0] == LO_SINGLE or (ops[j][0] == SINGLE and ops[j][1] > 0)):
                j += 1
            if chunk_bits and j > i and j < len(ops) and ops[j][0] == EXCHANGE:   # pipelined run, chunk by chunk
                for R, hd in enumerate(hs):
                    for c in chunk_order(chunk_bits, g, R):
                        for kind2, k2, pre2 in ops[i:j - (0 if dma else 1)]:
                            hd.op_chunk(kind2, k2, pre2, layout, both[cur][R], -1, chunk_bits, c)
                        if dma:   # copy-engine exchange of the finished chunk
                            hd.copy_chunk(both[cur][R], 1 - cur, chunk_bits, c)
                        else:
                            kind2, k2, pre2 = ops[j - 1]
                            hd.op_chunk(kind2, k2, pre2, layout, both[cur][R], 1 - cur, chunk_bits, c)
                cur ^= 1
                layout ^= 1
                i = j + 1
                continue
            if i + 1 < len(ops) and ops[i + 1][0] == EXCHANGE:
                for R, hd in enumerate(hs):
                    hd.op_exchange(kind, k, pre, layout, both[cur][R], 1 - cur)
                cur ^= 1
                layout ^= 1
                i += 2
                continue
            assert kind != EXCHANGE
            for R, hd in enumerate(hs):
                hd.op(kind, k, pre, layout, both[cur][R])
            i += 1
        bufs, ops = both[cur], []
    for kind, k, pre in ops:
        if kind == EXCHANGE:
            stacked = torch.stack([b.view(W, -1) for b in bufs])                  # [rank][chunk][...]
            bufs = [stacked[:, R, :].contiguous().view(-1) for R in range(W)]
            layout ^= 1
        else:
            for R, hd in enumerate(hs):
                hd.op(kind, k, pre, layout, bufs[R])
    totals = []
    for R, hd in enumerate(hs):
        t = torch.zeros(1, dtype=torch.float64, device=dev)
        hd.norm(t)
        totals.append(float(t.item()))
    if on_device:  # large n: per-rank probabilities stay on the device (the caller compares them there)
        p = []
        for R, hd in enumerate(hs):
            pr = torch.empty(1 << nl, dtype=torch.float32, device=dev)
            hd.final_probabilities(bufs[R], layout, pr)
            p.append(pr)
    else:
        p = np.zeros(1 << n)
    for R, hd in enumerate(hs):
        if on_device:
            break
        pr = torch.empty(1 << nl, dtype=torch.float32, device=dev)
        hd.final_probabilities(bufs[R], layout, pr)
        p[logical_indices(R, nl, g, layout)] = pr.cpu().numpy().astype(np.float64)
    picks = []
    for u in us:
        owner, target = pick_owner(np.array(totals), u)
        idx = torch.zeros(1, dtype=torch.int64, device=dev)
        hs[owner].select(bufs[owner], target, layout, idx)
        picks.append(int(idx.item()))
    for hd in hs:
        hd.close()
    return J, h, p, totals, layout, picks


@pytest.mark.parametrize("n,g,r", [(19, 1, 1), (19, 1, 2), (19, 1, 3), (20, 2, 4), (21, 1, 5), (22, 3, 6), (25, 1, 3)])
def test_shard_ops_match_oracle(n, g, r):
    rng = np.random.default_rng(1000 * n + 10 * g + r)
    s = int(rng.integers(0, 1 << n))
    gamma = float(rng.uniform(0.25, 0.6))
    us = [0.0, 0.25, 0.5, 0.77, 0.999999]
    J, h, p, totals, layout, picks = _emulated_proposal(n, g, r, s, gamma, us=us)
    assert layout == 0
    masks, w = models.one_body(n)
    ref = oracle.transition_probs(J, h, oracle.alpha(J, h), gamma, 0.8, r, masks, w, s, gatewise=False)
    assert abs(sum(totals) - 1.0) < 1e-5
    assert np.abs(p - ref).max() < 1e-6                                      # north-star tolerance (fp32)
    # measurement: sequential inverse CDF in index order over the probabilities this path produced
    cdf = np.cumsum(p)
    for u, got in zip(us, picks):
        want = int(np.searchsorted(cdf, u * cdf[-1], side="right"))
        if got != want:  # u within fp32 summation rounding of a CDF boundary
            lo, hi = min(got, want), max(got, want)
            assert abs(cdf[hi - 1] - u * cdf[-1]) < 1e-6 * cdf[-1] or abs(cdf[lo] - u * cdf[-1]) < 1e-6 * cdf[-1], (u, got, want)


@pytest.mark.parametrize("n,g,r", [(19, 1, 2), (20, 2, 4), (21, 1, 5), (22, 3, 6), (25, 2, 3)])
def test_shard_fused_exchange_matches_staged(n, g, r):
    """qmc_shard_op_exchange (pass + all-to-all in one kernel, stores into the peers' buffers) gives bit-identical
    amplitudes to pass-then-exchange: same arithmetic, only the store addresses differ."""
    rng = np.random.default_rng(77 * n + g + r)
    s = int(rng.integers(0, 1 << n))
    gamma = float(rng.uniform(0.25, 0.6))
    us = [0.1, 0.6, 0.93]
    _, _, p0, t0, l0, k0 = _emulated_proposal(n, g, r, s, gamma, us=us)
    _, _, p1, t1, l1, k1 = _emulated_proposal(n, g, r, s, gamma, us=us, fused=True)
    assert l0 == l1 == 0
    assert np.array_equal(p0, p1) and k0 == k1
    assert np.allclose(t0, t1, rtol=1e-12, atol=0)       # the norm reduction uses atomics: summation order varies


@pytest.mark.parametrize("n,g,r,cb", [(19, 1, 3, 1), (20, 2, 4, 3), (22, 3, 4, 3), (24, 1, 3, 2), (25, 2, 5, 4), (25, 2, 2, 2)])
def test_shard_chunked_exchange_matches_staged(n, g, r, cb):
    """qmc_shard_op_chunk: the passes before an exchange run chunk by chunk (the last one storing into the peers'
    buffers) -- bit-identical amplitudes to pass-then-exchange on the whole shard."""
    from qumcmc_b200.sharded import max_chunk_bits
    assert g <= cb <= max_chunk_bits(n - g)
    rng = np.random.default_rng(31 * n + g + r)
    s = int(rng.integers(0, 1 << n))
    gamma = float(rng.uniform(0.25, 0.6))
    us = [0.2, 0.71]
    _, _, p0, t0, l0, k0 = _emulated_proposal(n, g, r, s, gamma, us=us)
    _, _, p1, t1, l1, k1 = _emulated_proposal(n, g, r, s, gamma, us=us, fused=True, chunk_bits=cb)
    assert l0 == l1 == 0
    assert np.array_equal(p0, p1) and k0 == k1
    assert np.allclose(t0, t1, rtol=1e-12, atol=0)


@pytest.mark.parametrize("n,g,r,cb,ctas,groups", [(22, 1, 3, 2, 7, ""), (24, 3, 4, 3, 5, ""), (25, 2, 5, 4, 16, ""), (26, 2, 2, 0, 3, ""),
                                                  (24, 3, 3, 3, 6, "8:8,6:1"), (27, 3, 3, 4, 11, "8:8,6:2,6:2")])
def test_shard_persistent_exchange_kernel_matches_staged(monkeypatch, n, g, r, cb, ctas, groups):
    """Overlap mode: the exchange pass as a persistent kernel on a few CTAs (hi6_xchg_kernel), whole-shard (cb = 0) and
    chunk by chunk -- bit-identical amplitudes to pass-then-exchange; also with C5's three-group shard structure."""
    from qumcmc_b200 import _device
    from qumcmc_b200.sharded import max_chunk_bits
    if groups:
        monkeypatch.setenv("QUMCMC_HI_GROUPS", groups)
    _device._cache.clear()
    try:
        assert cb == 0 or g <= cb <= max_chunk_bits(n - g)
        rng = np.random.default_rng(13 * n + g + r)
        s = int(rng.integers(0, 1 << n))
        gamma = float(rng.uniform(0.25, 0.6))
        us = [0.15, 0.8]
        _, _, p0, t0, l0, k0 = _emulated_proposal(n, g, r, s, gamma, us=us)
        _, _, p1, t1, l1, k1 = _emulated_proposal(n, g, r, s, gamma, us=us, fused=True, chunk_bits=cb, xchg_ctas=ctas)
    finally:
        _device._cache.clear()
    assert l0 == l1 == 0
    assert np.array_equal(p0, p1) and k0 == k1
    assert np.allclose(t0, t1, rtol=1e-12, atol=0)


@pytest.mark.parametrize("n,g,r,cb", [(22, 1, 3, 2), (24, 3, 4, 3), (25, 2, 5, 4), (25, 2, 2, 2), (26, 3, 6, 5)])
def test_shard_copy_engine_exchange_matches_staged(n, g, r, cb):
    """"dma" mode: every pass local, each finished chunk sent with qmc_shard_copy_chunk (cudaMemcpyAsync into the owning
    rank's buffer) -- bit-identical amplitudes to pass-then-exchange."""
    from qumcmc_b200.sharded import max_chunk_bits
    assert g <= cb <= max_chunk_bits(n - g)
    rng = np.random.default_rng(17 * n + g + r)
    s = int(rng.integers(0, 1 << n))
    gamma = float(rng.uniform(0.25, 0.6))
    us = [0.33, 0.9]
    _, _, p0, t0, l0, k0 = _emulated_proposal(n, g, r, s, gamma, us=us)
    _, _, p1, t1, l1, k1 = _emulated_proposal(n, g, r, s, gamma, us=us, fused=True, chunk_bits=cb, dma=True)
    assert l0 == l1 == 0
    assert np.array_equal(p0, p1) and k0 == k1
    assert np.allclose(t0, t1, rtol=1e-12, atol=0)


def test_shard_rejects_bad_configurations():
    import qumcmc_b200 as q
    from qumcmc_b200 import _lib
    from qumcmc_b200.sharded import ShardHandle
    J, h = models.packaged_random_model(18, 3)
    with pytest.raises(_lib.QmcUnsupported):                                  # 17 local qubits: below the tile geometry
        ShardHandle(q.IsingEnergyFunction(J, h), q.GenericMixer.OneBodyMixer(18), 0, 1, 0)
    J, h = models.packaged_random_model(19, 3)
    with pytest.raises(ValueError):
        ShardHandle(q.IsingEnergyFunction(J, h), q.GenericMixer.OneBodyMixer(19), 2, 1, 0)   # rank outside the world


def _nccl_worker(rank, world, port, n, cases, q, fused):
    if fused == "overlap":
        os.environ["QUMCMC_SHARD_CTAS"] = "9"
    mode = fused if fused in ("overlap", "dma") else "fused"
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import qumcmc_b200 as qq
        from qumcmc_b200.sharded import ShardedProposer
        J, h = models.packaged_random_model(n, 11)
        prop = ShardedProposer(qq.IsingEnergyFunction(J, h), qq.GenericMixer.OneBodyMixer(n), device=rank, fused=bool(fused),
                               chunk_bits=(2 if fused in ("chunked", "overlap", "dma") else 0), mode=mode)
        assert prop.mode == mode
        out = [prop.propose(s, gamma, r, u) for (s, gamma, r, u) in cases]
        q.put((rank, out, prop.exchanges))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fused", [False, True, "chunked", "overlap", "dma"])
def test_shard_nccl_world2_matches_emulation(fused):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    n, world = (22 if fused == "overlap" else 21), 2   # 21 local qubits: HI8 + HI6 groups (persistent exchange pass)
    rng = np.random.default_rng(5)
    cases = [(int(rng.integers(0, 1 << n)), float(rng.uniform(0.25, 0.6)), int(r), float(rng.uniform(0, 1)))
             for r in (1, 2, 3, 6)]
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(rk, world, port, n, cases, q, fused)) for rk in range(world)]
    for p_ in procs:
        p_.start()
    res = sorted((q.get(timeout=600) for _ in range(world)), key=lambda t: t[0])
    for p_ in procs:
        p_.join(timeout=120)
        assert p_.exitcode == 0
    assert res[0][1] == res[1][1]                                             # every rank reports the same bitstring
    for (s, gamma, r, u), got in zip(cases, res[0][1]):
        _, _, _, _, _, picks = _emulated_proposal(n, 1, r, s, gamma, us=[u])
        assert got == picks[0]


@pytest.mark.parametrize("n,g,groups,r", [(27, 3, "8:8,6:2,6:2", 3), (25, 2, "8:8,6:3", 4), (24, 1, "6:6,6:3,6:2", 5)])
def test_shard_ops_with_three_register_groups(monkeypatch, n, g, groups, r):
    """C5's local shard (33 qubits) has K = 3 register groups; QUMCMC_HI_GROUPS gives an emulated shard
    the same structure (HI8 top group + two lower groups, SINGLE passes on both before every exchange) at a size the
    oracle finishes."""
    from qumcmc_b200 import _device
    monkeypatch.setenv("QUMCMC_HI_GROUPS", groups)
    _device._cache.clear()
    rng = np.random.default_rng(n * 7 + r)
    s = int(rng.integers(0, 1 << n))
    gamma = float(rng.uniform(0.25, 0.6))
    try:
        J, h, p, totals, layout, picks = _emulated_proposal(n, g, r, s, gamma, us=[0.3, 0.9], fused=(g == 3))
    finally:
        _device._cache.clear()
    assert layout == 0
    masks, w = models.one_body(n)
    ref = oracle.transition_probs(J, h, oracle.alpha(J, h), gamma, 0.8, r, masks, w, s, gatewise=False)
    assert abs(sum(totals) - 1.0) < 1e-5
    assert np.abs(p - ref).max() < 1e-6
    assert np.linalg.norm(p - ref) < 1e-4 * np.linalg.norm(ref)     # relative: probabilities are ~2^-n at this size


@pytest.mark.slow
def test_shard_eight_way_split_n30_matches_single_gpu():
    """an (emulated) 8-way split of n = 30 (g = 3, 27 local qubits, fused pass + exchange) against
    the single-GPU n = 30 path -- whose geometry class (K = 3) is checked against the oracle at n = 29.  Everything is
    compared on the device: 2^30 probabilities."""
    import torch
    import qumcmc_b200 as q
    from qumcmc_b200 import _device, _lib
    n, g, r, s, gamma = 30, 3, 3, 0x2F00F00D & ((1 << 30) - 1), 0.43
    nl = n - g
    out = _emulated_proposal(n, g, r, s, gamma, seed=11, us=[0.37], fused=True, on_device=True)
    J, h, parts, totals, layout, picks = out
    assert layout == 0 and abs(sum(totals) - 1.0) < 1e-5
    model = q.IsingEnergyFunction(J, h)
    dm = model.device_model(0)
    dm.set_mixer(q.GenericMixer.OneBodyMixer(n))
    dev = torch.device("cuda", 0)
    single = torch.empty(1 << n, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().qmc_transition_probs(dm.handle, s, gamma, r, 0.8, 0, _lib.DTYPE_F32, single.data_ptr(), None, None))
    torch.cuda.synchronize()
    num = den = 0.0
    worst = 0.0
    for R, pr in enumerate(parts):     # layout 0: logical index = (R << nl) | local index
        ref = single[R << nl:(R + 1) << nl]
        d = (pr.double() - ref.double())
        num += float((d * d).sum())
        den += float((ref.double() ** 2).sum())
        worst = max(worst, float(d.abs().max()))
    assert worst < 1e-6                                   # north-star tolerance (fp32), trivial at 2^-30 per state ...
    assert (num / den) ** 0.5 < 2e-4                      # ... so also relative L2 over all 2^30 probabilities
    # measurement through the sharded selector = sequential inverse CDF over the single-GPU probabilities
    cdf_at = float(single[: picks[0]].double().sum())
    p_at = float(single[picks[0]])
    assert cdf_at - 2e-5 <= 0.37 <= cdf_at + p_at + 2e-5
    del parts, single
    _device._cache.clear()
    torch.cuda.empty_cache()
