#!/usr/bin/env python
"""bench.py -- quantum proposals/sec on BASELINE config C3: n = 20 random Ising
(random_ising_model(20, seed=1), qumcmc/energy_models.py:419-446), one-body X mixer,
gamma ~ U(0.25, 0.6), delta_t = 0.8, t ~ U{2..11}, beta = 1, 4096 independent chains per GPU.

A "step" = one hop of every chain (4096 proposals per GPU): Philox draws, U = M (P M)^{r-1} on a
2^20 complex64 statevector per chain, one measurement, get_energy, Metropolis accept.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # CPU arm (gate-at-a-time)

For N > 1 launch under torchrun (one rank per GPU); chains are sharded, no collective in the hop
loop, one gather of acceptance counts / Hamming histograms at the end (inside the timed region).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_SPINS = 20
MODEL_SEED = 1
BETA = 1.0
GAMMA = (0.25, 0.6)
DELTA_T = 0.8
PHILOX_SEED = 0x5EED
METRIC = "quantum proposals/sec (n=20 Ising, 4096 chains)"
UNIT = "proposals/s"


def workload_config(chains_per_gpu, n_gpus, extra=None):
    cfg = {
        "workload": "C3: random_ising_model(20, seed=1), OneBodyMixer, gamma~U(0.25,0.6), dt=0.8, t~U{2..11}, beta=1.0",
        "n_spins": N_SPINS, "chains_per_gpu": chains_per_gpu, "chains_total": chains_per_gpu * n_gpus,
        "step": "one hop of every chain", "statevector": "complex64, 8 MiB per chain",
        "l2_policy": "inputs larger than the L2: every step rewrites chains x 8 MiB = %.1f GiB of statevectors per GPU "
                     "(>> 126 MB), nothing survives from one timed step to the next; inside a step the L2-resident group "
                     "schedule keeps the statevectors of the 8 chains in flight in the L2 between their own sweeps"
                     % (chains_per_gpu * 8 / 1024),
        "parallelism": "chain-sharded x%d, no collective in the hop loop" % n_gpus,
    }
    if extra:
        cfg.update(extra)
    return cfg


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._stop_evt = threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (one pass over the complex128 state per gate) restated in
# oracle/qmc_oracle.c, OpenMP over the host cores.  qulacs is not installable in this image.
# ---------------------------------------------------------------------------------------------
REF_HOPS_PER_STEP = 4


def cpu_reference_rate(hops, warm=0):
    import oracle
    from qumcmc_b200.energy_models import random_ising_model
    from qumcmc_b200.mixers import GenericMixer

    oracle.set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1
    model = random_ising_model(N_SPINS, MODEL_SEED)
    _, masks, w, _ = GenericMixer.OneBodyMixer(N_SPINS).lower()
    s0 = oracle.initial_state(PHILOX_SEED, 0, N_SPINS)
    if warm:
        oracle.run_chain(model.J, model.h, masks, w, s0, warm, 1 / BETA, GAMMA, DELTA_T, PHILOX_SEED, chain_id=0,
                         gatewise=True)
    t0 = time.perf_counter()
    oracle.run_chain(model.J, model.h, masks, w, s0, hops, 1 / BETA, GAMMA, DELTA_T, PHILOX_SEED, chain_id=0,
                     hop0=warm, gatewise=True)
    dt = time.perf_counter() - t0
    return hops / dt, dt, oracle.num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    # one step = REF_HOPS_PER_STEP hops of one chain (a bounded sample of the workload: the full step is one hop of
    # 4096 chains); several hops per step average over the Trotter depth r, which varies 1..13 from hop to hop
    rate, dt, threads = cpu_reference_rate(args.steps * REF_HOPS_PER_STEP, warm=args.warmup * REF_HOPS_PER_STEP)
    sample = "1 chain x %d hops (+%d warm-up), gate-at-a-time complex128 (n(n-1)/2+n diagonal + n X gates per layer)" % (
        args.steps * REF_HOPS_PER_STEP, args.warmup * REF_HOPS_PER_STEP)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": workload_config(args.chains, args.gpus, {"reference_step": "%d hops of one chain on the host CPU" % REF_HOPS_PER_STEP}),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "oracle/qmc_oracle.c restatement of the qulacs circuit; qulacs itself is not installable here"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def pin_rank_to_cores(local_rank):
    """torchrun leaves every rank on the full CPU mask.  The L2-resident schedule keeps one host thread per rank busy
    launching kernels (about 10 k launches per hop), so each rank gets its own block of logical CPUs: two ranks'
    launch threads never share a core."""
    if os.environ.get("QUMCMC_BENCH_NO_PIN"):   # A/B switch
        return
    try:
        lws = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        cpus = sorted(os.sched_getaffinity(0))
        per = len(cpus) // max(1, lws)
        if lws > 1 and per >= 2:
            os.sched_setaffinity(0, cpus[local_rank * per:(local_rank + 1) * per])
    except (AttributeError, OSError, ValueError):
        pass


def run_ours(args, rank, local_rank, world):
    import torch

    from qumcmc_b200 import _device, _lib
    from qumcmc_b200.distributed import gather_statistics
    from qumcmc_b200.energy_models import random_ising_model
    from qumcmc_b200.mixers import GenericMixer

    torch.cuda.set_device(local_rank)
    pin_rank_to_cores(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    L = _lib.lib()
    model = random_ising_model(N_SPINS, MODEL_SEED)
    if args.model_scale != 1.0:
        from qumcmc_b200.energy_models import IsingEnergyFunction
        model = IsingEnergyFunction(model.J * args.model_scale, model.h * args.model_scale)
    mixer = GenericMixer.OneBodyMixer(N_SPINS)
    dm = model.device_model(local_rank)
    dm.set_mixer(mixer)
    C, K, W = args.chains, args.steps, args.warmup
    chain_id0 = rank * C
    dev = torch.device("cuda", local_rank)
    stream = _device.current_stream_ptr(local_rank)
    n1 = N_SPINS + 1

    Hbuf = max(K, W, 1)
    props_buf = torch.empty(C * Hbuf, dtype=torch.int64, device=dev)
    acc_buf = torch.empty(C * Hbuf, dtype=torch.uint8, device=dev)
    fin = torch.empty(C, dtype=torch.int64, device=dev)
    cnt = torch.empty(C, dtype=torch.int64, device=dev)
    hist = torch.empty((C, n1), dtype=torch.int64, device=dev)

    def run_dev(n_hops, s0, hop0):
        _lib.check(L.qmc_run_chains(dm.handle, C, n_hops, None if s0 is None else s0.data_ptr(), chain_id0, hop0,
                                    1.0 / BETA, GAMMA[0], GAMMA[1], DELTA_T, PHILOX_SEED, _lib.DTYPE_F32,
                                    props_buf.data_ptr(), acc_buf.data_ptr(), fin.data_ptr(), cnt.data_ptr(),
                                    hist.data_ptr(), stream))

    # ---- warm-up (allocates the statevector workspace and phase tables)
    run_dev(max(W, 1), None, 0)
    torch.cuda.synchronize()
    s0 = fin.clone()

    if dist is not None:   # warm-up of the one collective of the path as well: NCCL sets its all-gather channels up on first use
        gather_statistics(cnt.cpu().numpy(), hist.cpu().numpy(), C * world)
        torch.cuda.synchronize()

    # ---- timed region: K steps, inputs resident in HBM
    _lib.check(L.qmc_profile_enable(dm.handle, 1))
    launches0 = dm.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    run_dev(K, s0, W)
    ev_mid = torch.cuda.Event(enable_timing=True)
    ev_mid.record()
    g_acc = g_hist = None
    if dist is not None:  # the one exchange of the chain-parallel path: statistics gather
        g_acc, g_hist = gather_statistics(cnt.cpu().numpy(), hist.cpu().numpy(), C * world)
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    per_rank = None   # diagnostics: every rank's own device time of the K hops and of the gather that follows
    if dist is not None:
        mine = torch.tensor([ev0.elapsed_time(ev_mid), ev_mid.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"hops_ms": [round(float(t[0]), 2) for t in allr], "gather_ms": [round(float(t[1]), 2) for t in allr]}
    clocks = sampler.stop() if sampler else None
    launches = dm.launch_count() - launches0
    sweep_ms, sweep_launches, alg_bytes = (torch.zeros(1, dtype=torch.float64).numpy(), np.zeros(1, dtype=np.int64),
                                           np.zeros(1, dtype=np.float64))
    import ctypes as Ct
    c_ms, c_ln, c_by = Ct.c_double(), Ct.c_int64(), Ct.c_double()
    _lib.check(L.qmc_profile_read(dm.handle, Ct.byref(c_ms), Ct.byref(c_ln), Ct.byref(c_by)))
    c_mv = Ct.c_double()
    _lib.check(L.qmc_profile_read_moved(dm.handle, Ct.byref(c_mv)))
    _lib.check(L.qmc_profile_enable(dm.handle, 0))   # also resets the counters
    value = world * C * K / (ms * 1e-3)
    # acceptance over ALL ranks' chains (the gathered counts when N > 1)
    acc_rate = float(g_acc.sum()) / (C * world * K) if g_acc is not None else float(cnt.sum().item()) / (C * K)

    # ---- end to end through the host-buffer C-ABI: every step copies its inputs H2D and results D2H
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    h_s0 = pin(C, torch.int64)
    h_s0.copy_(s0)
    h_prop, h_acc = pin(C, torch.int64), pin(C, torch.uint8)
    h_fin, h_cnt, h_hist = pin(C, torch.int64), pin(C, torch.int64), pin((C, n1), torch.int64)
    Ke = min(K, args.e2e_steps)

    def step_host(hop0):
        _lib.check(L.qmc_run_chains_host(dm.handle, C, 1, h_s0.data_ptr(), chain_id0, hop0, 1.0 / BETA, GAMMA[0],
                                         GAMMA[1], DELTA_T, PHILOX_SEED, _lib.DTYPE_F32, h_prop.data_ptr(),
                                         h_acc.data_ptr(), h_fin.data_ptr(), h_cnt.data_ptr(), h_hist.data_ptr()))
        h_s0.copy_(h_fin)

    step_host(W)  # warm
    h_s0.copy_(s0)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(Ke):
        step_host(W + k)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * C * Ke / t_e2e
    # same hops through both entry points must give the same chain (device-resident vs host buffers)
    e2e_consistent = bool(torch.equal(h_prop.to(dev), props_buf[: C * K].view(C, K)[:, Ke - 1])) if Ke >= 1 else None
    bi = C * 8
    bo = C * (8 + 1 + 8 + 8 + n1 * 8)

    # ---- secondary measurements (never allowed to lose the headline line: guarded, with a watchdog)
    secondary = {}
    state = {"line": None}
    watchdog = start_watchdog(args.secondary_timeout, rank, state)
    if world > 1 and not args.no_secondary:
        try:   # strong scaling: the metric's 4096 chains IN TOTAL over the N GPUs (same global chain ids as one GPU)
            secondary["strong"] = strong_scaling_block(args, L, dm, rank, world, dev, stream, barrier, max_over_ranks)
        except Exception as e:  # noqa: BLE001
            secondary["strong"] = {"error": repr(e)[:300]}

    if rank == 0:
        state["line"] = build_line(args, world, C, K, W, ms, value, acc_rate, clocks, launches, c_ms.value, c_ln.value,
                                   c_by.value, c_mv.value, e2e_value, Ke, bi, bo, e2e_consistent, g_acc, secondary)
        if per_rank is not None:
            state["line"]["per_rank"] = per_rank
    if world == 1 and not args.no_secondary:
        try:   # mixers beyond one-body (the paper's subject) on the same model sizes: short runs, a few seconds in total
            secondary["mixers"] = mixers_block(args)
        except Exception as e:  # noqa: BLE001
            secondary["mixers"] = {"error": repr(e)[:300]}
    if world >= 2 and (world & (world - 1)) == 0 and not args.no_secondary and not args.no_c5:
        try:   # C5: statevector sharded by global qubits, 33 local qubits per GPU (n = 36 on 8 GPUs)
            del props_buf, acc_buf
            dm = None
            from qumcmc_b200 import _device as _dv
            _dv._cache.clear()           # frees the 32 GiB C3 workspace: C5 needs 2 x 64 GiB per GPU
            torch.cuda.empty_cache()
            secondary["c5"] = c5_block(args, rank, local_rank, world)
        except Exception as e:  # noqa: BLE001
            secondary["c5"] = {"error": repr(e)[:300]}
    if watchdog is not None:
        watchdog.cancel()
    if rank == 0:
        state["line"]["secondary"] = secondary
        print(json.dumps(state["line"]), flush=True)
    if dist is not None:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass


def start_watchdog(seconds, rank, state):
    """If a secondary block hangs (a collective that never completes), rank 0 still prints the headline line."""
    if seconds <= 0:
        return None

    def fire():
        if rank == 0 and state.get("line") is not None:
            state["line"]["secondary"] = {"error": "secondary measurements timed out after %d s" % seconds}
            print(json.dumps(state["line"]), flush=True)
        os._exit(0 if state.get("line") is not None or rank != 0 else 1)

    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


def mixers_block(args):
    """Secondary lines for mixers of higher weight (qumcmc/mixers.py:21-151), through the public Python API: two-body mixers
    and three-body mixers at n = 20 run on the sweep kernels in XQ / XT mode (2 sweeps per Trotter layer), k-body mixers at
    n = 9 (BAS 3x3) on the warp-per-chain kernel in Hadamard-basis mode; one-body lines beside them for the ratio."""
    import torch
    import qumcmc_b200 as q
    from tests import models as tmodels

    out = []
    Jb, hb, _ = tmodels.bas3("both")
    cases = [("n=20 random Ising", q.random_ising_model(N_SPINS, MODEL_SEED), N_SPINS, None, 1.0, 1024, 8, ("two", "three"), ("fp32", "fp64")),
             ("BAS 3x3 (n=9)", q.IsingEnergyFunction(Jb, hb), 9, "111000111", 1 / 1.5, 1024, 400, ("one", "two", "three"), ("fp32", "fp64"))]
    for label, model, n, init, temp, chains, hops, kinds, dtypes in cases:
        for kind in kinds:
            mx = {"one": q.GenericMixer.OneBodyMixer, "two": q.GenericMixer.TwoBodyMixer, "three": q.GenericMixer.ThreeBodyMixer}[kind](n)
            for dtype in dtypes:
                kw = dict(temperature=temp, n_chains=chains, seed=2, dtype=dtype, record=False)
                if init:
                    kw["initial_state"] = init
                q.quantum_enhanced_mcmc(max(1, hops // 20), model, mx, GAMMA, **kw)   # warm: tables, workspace
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                b = q.quantum_enhanced_mcmc(hops, model, mx, GAMMA, **kw)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                out.append({"model": label, "mixer": kind + "-body", "terms": int(len(mx.lower()[1])), "dtype": dtype, "chains": chains,
                            "hops": hops, "proposals_per_s": chains * hops / dt, "acceptance": b.acceptance_rate()})
    return out


def strong_scaling_block(args, L, dm, rank, world, dev, stream, barrier, max_over_ranks):
    """BASELINE's metric reads "4096 chains": the same 4096 global chain ids split over the N GPUs (strong scaling).
    Results do not depend on the split (Philox is keyed by the global chain id)."""
    import torch
    from qumcmc_b200 import _lib
    from qumcmc_b200.distributed import shard_range
    total = args.chains
    a, b = shard_range(total, rank, world)
    Cl, K, W = b - a, args.steps, args.warmup
    props = torch.empty(Cl * max(K, 1), dtype=torch.int64, device=dev)
    accb = torch.empty(Cl * max(K, 1), dtype=torch.uint8, device=dev)
    fin = torch.empty(Cl, dtype=torch.int64, device=dev)
    cnt = torch.empty(Cl, dtype=torch.int64, device=dev)
    hist = torch.empty((Cl, N_SPINS + 1), dtype=torch.int64, device=dev)

    def run(n_hops, s0, hop0):
        _lib.check(L.qmc_run_chains(dm.handle, Cl, n_hops, None if s0 is None else s0.data_ptr(), a, hop0, 1.0 / BETA,
                                    GAMMA[0], GAMMA[1], DELTA_T, PHILOX_SEED, _lib.DTYPE_F32, props.data_ptr(),
                                    accb.data_ptr(), fin.data_ptr(), cnt.data_ptr(), hist.data_ptr(), stream))

    run(max(W, 1), None, 0)
    torch.cuda.synchronize()
    s0 = fin.clone()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(K, s0, max(W, 1))
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    return {"value": total * K / (ms * 1e-3), "unit": UNIT, "chains_total": total, "chains_per_gpu": Cl, "steps": K,
            "ms_per_step": ms / max(1, K), "scaling": "strong"}


def c5_block(args, rank, local_rank, world):
    """BASELINE config C5 scaled to the world size: 33 local qubits per GPU (n = 36 on 8 GPUs, 2 x 64 GiB of state
    buffers each), fused pass + exchange over NVLink peer mappings; plus the same 64 GiB statevector on ONE GPU (n = 33,
    rank 0) as the weak-scaling reference."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import sharded_bench as sb
    import qumcmc_b200 as q
    from qumcmc_b200.sharded import ShardedProposer
    g = world.bit_length() - 1
    n_local = args.c5_local_spins
    n = n_local + g
    model = q.random_ising_model(n, 1)
    prop = ShardedProposer(model, q.GenericMixer.OneBodyMixer(n), device=local_rank, fused=True)
    rows = sb.measure_proposals(prop, 0x5EED5EED5 & ((1 << n) - 1), [2, 5], 2, world, rank)
    prop.close()
    del prop
    torch.cuda.empty_cache()
    out = {"n": n, "n_local": n_local, "world": world, "proposals": rows}
    if rank == 0:
        ref = {}
        for row in rows:
            r = row["layers"]
            ref[r] = sb.single_gpu_reference(n_local, r, 2, local_rank)
            row["single_gpu_same_shard_seconds"] = ref[r]
            row["weak_scaling_efficiency"] = ref[r] / row["seconds_per_proposal"]
    dist.barrier()
    return out


def build_line(args, world, C, K, W, ms, value, acc_rate, clocks, launches, sweep_ms, sweep_launches, alg_bytes, moved_bytes,
               e2e_value, Ke, bi, bo, e2e_consistent, g_acc, secondary):
    peak, peak_src = measured_peak_gbs()
    class _V:  # the three profile counters, named as below
        pass
    c_ms, c_ln, c_by = _V(), _V(), _V()
    c_ms.value, c_ln.value, c_by.value = sweep_ms, sweep_launches, alg_bytes
    achieved = (c_by.value / 1e9) / (c_ms.value * 1e-3) if c_ms.value > 0 else None
    traffic = traffic_alg = dram_ratio = None
    grouped = os.environ.get("QUMCMC_L2_GROUP_MIB", "24") != "0"
    tpath = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tpath) and grouped:   # the committed capture is of the default (L2-resident) schedule
        try:
            tj = json.load(open(tpath))
            traffic, traffic_alg = tj.get("dram_bytes_per_launch"), tj.get("algorithmic_bytes_per_launch")
            dram_ratio = tj.get("all_sweep_kernels", {}).get("dram_over_loadstore")
        except Exception:
            traffic = None
    moved = (moved_bytes / 1e9) / (c_ms.value * 1e-3) if c_ms.value > 0 else None
    if not grouped:
        dram_ratio = 1.0   # step-major schedule: every load / store of a sweep is HBM traffic (ncu: 0.99)
    dram = None if moved is None or dram_ratio is None else moved * dram_ratio
    roofline = {
        "bound": "hbm",
        "kernel": "lo_sweep_kernel<MID> + hiq_sweep_kernel (HI MID) + their FIRST / FINAL roles: all sweep launches of the step "
                  "(fused phase + butterfly sweeps)",
        "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": None if achieved is None else achieved / peak, "traffic": traffic,
        "traffic_algorithmic_bytes_per_launch": traffic_alg,
        "schedule": ("L2-resident groups (group_sched.cuh): the chains of a group run all their sweeps back to back, a sweep "
                     "rewrites the lines its predecessor left in the L2; HBM sees the write-backs only, so `frac` can exceed "
                     "what an HBM-streaming schedule reaches and the kernels are bound by FP32 issue, not by HBM")
                    if grouped else "step-major (QUMCMC_L2_GROUP_MIB=0): every sweep streams all statevectors through HBM",
        "loadstore_achieved": moved, "loadstore_frac": None if moved is None else moved / peak,
        "loadstore_definition": "bytes the sweep launches load + store: 16 B * 2^n * (r - 1) per proposal (first sweep "
                                "write-only, last read-only) / the same device time; equals the DRAM traffic under the "
                                "step-major schedule",
        "dram_achieved": dram, "dram_frac": None if dram is None else dram / peak,
        "dram_definition": "loadstore_achieved x (DRAM bytes / loaded + stored bytes of the sweep kernels in the committed ncu "
                           "capture, profiles/sweep_traffic.json); DRAM counters cannot be read outside a profiler",
        "peak_source": peak_src, "sweep_ms_per_step": c_ms.value / max(1, K),
        "algorithmic_bytes_per_step": c_by.value / max(1, K), "launches_per_step": c_ln.value / max(1, K),
        "share_of_step": c_ms.value / ms if ms > 0 else None,
        "definition": "B_alg = 2 * 2^n * 8 B * r per proposal (SURVEY 8d); achieved = sum B_alg / device time of the sweep launches (CUDA events)",
    }
    cpu = None
    world_is_one = world == 1
    if world_is_one and not args.no_cpu_baseline:
        rate, dt, threads = cpu_reference_rate(args.cpu_hops)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "1 chain x %d hops of the same workload, gate-at-a-time complex128 (reference cost structure), %.1f s"
                         % (args.cpu_hops, dt),
               "published_context": "authors' qulacs runs: ~500 proposals/s at n=10, ~52 at n=12 (BASELINE.md); none at n=20"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / max(1, K), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c64 (fp32 arithmetic; energies/accept fp64)", "data": "synthetic",
        "config": workload_config(C, world, {"acceptance_rate": acc_rate}),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo,
                "steps": Ke, "api": "qmc_run_chains_host (one hop per call, pinned host buffers)",
                "matches_device_resident_run": e2e_consistent},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if g_acc is not None:
        line["config"]["gathered_chains"] = int(len(g_acc))
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=4096, help="chains per GPU")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-hops", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary blocks (strong scaling, C5)")
    ap.add_argument("--no-c5", action="store_true", help="skip the sharded-statevector block")
    ap.add_argument("--c5-local-spins", type=int, default=33, help="local qubits per GPU of the C5 block (33 = 64 GiB per buffer)")
    ap.add_argument("--secondary-timeout", type=int, default=420, help="watchdog for the secondary blocks, seconds")
    ap.add_argument("--model-scale", type=float, default=1.0, help="diagnostics only: scales J,h (1e-6 disables the phase layer)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
