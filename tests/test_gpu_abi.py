"""GPU tests of the reference-facing C-ABI calls: the host-buffer entry points
(qmc_run_chains_host, qmc_classical_chains_host) against the device-pointer variants, the INTEGRATION.md ctypes stub
executed verbatim, and the device visit histogram (MCMCChain.get_accepted_dict, qumcmc/basic_utils.py:117-131)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from oracle import np_oracle as npo
from tests import models

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GAMMA = (0.25, 0.6)


def _api():
    import qumcmc_b200 as q
    return q


def _vp(a):
    return None if a is None else a.ctypes.data


def _host_buffers(C_, H, n):
    return (np.empty((C_, H), dtype=np.uint64), np.empty((C_, H), dtype=np.uint8), np.empty(C_, dtype=np.uint64),
            np.empty(C_, dtype=np.int64), np.empty((C_, n + 1), dtype=np.int64))


def _run_host(dm, C_, H, s0, chain_id0, hop0, T, seed, dtype, n, bufs=None):
    from qumcmc_b200 import _lib
    prop, acc, fin, cnt, hist = bufs if bufs is not None else _host_buffers(C_, H, n)
    _lib.check(_lib.lib().qmc_run_chains_host(dm.handle, C_, H, _vp(s0), chain_id0, hop0, T, GAMMA[0], GAMMA[1], 0.8, seed,
                                              dtype, _vp(prop), _vp(acc), _vp(fin), _vp(cnt), _vp(hist)))
    return prop, acc, fin, cnt, hist


@pytest.mark.parametrize("n,mixer,dtype", [(9, "one", "fp64"), (11, "two", "fp32"), (14, "one", "fp32"), (14, "two", "fp32"),
                                           (13, "one", "fp64")])
def test_run_chains_host_equals_the_device_pointer_variant(n, mixer, dtype):
    """Same (seed, chain ids, hops) through qmc_run_chains_host (numpy buffers, copies inside) and through qmc_run_chains
    (torch device buffers): every output bit-equal, on the shared-memory, sweep and general paths; with and without
    explicit initial states; with a resumed second call (hop0 > 0) that reuses the handle's staging arena."""
    q = _api()
    from qumcmc_b200 import _lib
    J, h = models.packaged_random_model(n, 3)
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.OneBodyMixer(n) if mixer == "one" else q.GenericMixer.TwoBodyMixer(n)
    dt = _lib.DTYPE_F64 if dtype == "fp64" else _lib.DTYPE_F32
    C_, H, seed = 12, 7, 41
    dev = q.quantum_enhanced_mcmc(H, m, mx, GAMMA, temperature=1.3, n_chains=C_, seed=seed, dtype=dtype, chain_id0=3)
    dm = m.device_model(0)
    dm.set_mixer(mx)
    prop, acc, fin, cnt, hist = _run_host(dm, C_, H, None, 3, 0, 1.3, seed, dt, n)
    assert np.array_equal(prop, dev.proposals) and np.array_equal(acc, dev.accepted)
    assert np.array_equal(fin, dev.final_state) and np.array_equal(cnt, dev.acc_count)
    assert np.array_equal(hist, dev.hamming_hist)
    # resumed: hops [H, 2H) from the final states, host vs device; nullable outputs left out on the host side
    dev2 = q.quantum_enhanced_mcmc(H, m, mx, GAMMA, initial_state=[int(x) for x in dev.final_state], temperature=1.3,
                                   n_chains=C_, seed=seed, dtype=dtype, chain_id0=3, hop0=H)
    fin2 = np.empty(C_, dtype=np.uint64)
    _lib.check(_lib.lib().qmc_run_chains_host(dm.handle, C_, H, _vp(fin), 3, H, 1.3, GAMMA[0], GAMMA[1], 0.8, seed, dt,
                                              None, None, _vp(fin2), None, None))
    assert np.array_equal(fin2, dev2.final_state)
    # a smaller call after a larger one (arena reuse) and a zero-chain call
    p3, a3, f3, c3, h3 = _run_host(dm, 5, 3, None, 3, 0, 1.3, seed, dt, n)
    assert np.array_equal(p3, dev.proposals[:5, :3]) and np.array_equal(a3, dev.accepted[:5, :3])
    assert _lib.lib().qmc_run_chains_host(dm.handle, 0, 3, None, 0, 0, 1.0, 0.3, 0.3, 0.8, 1, dt, None, None, _vp(f3),
                                          None, None) == 0
    # error path: final_state is required (BADARG -> ValueError), nothing written
    with pytest.raises(ValueError):
        _lib.check(_lib.lib().qmc_run_chains_host(dm.handle, 2, 2, None, 0, 0, 1.0, 0.3, 0.3, 0.8, 1, dt, None, None, None,
                                                  None, None))


def test_run_chains_host_steady_state_does_not_allocate():
    """after the first call of a shape the host-buffer entry point must not touch the allocator
    (device-free memory unchanged over repeated calls) and must accept pinned buffers."""
    import torch
    q = _api()
    from qumcmc_b200 import _lib
    n, C_ = 14, 64
    J, h = models.packaged_random_model(n, 2)
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.OneBodyMixer(n)
    dm = m.device_model(0)
    dm.set_mixer(mx)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    s0, prop, acc = pin(C_, torch.int64), pin(C_, torch.int64), pin(C_, torch.uint8)
    fin, cnt, hist = pin(C_, torch.int64), pin(C_, torch.int64), pin((C_, n + 1), torch.int64)
    s0.copy_(torch.from_numpy(np.array([oracle.initial_state(5, c, n) for c in range(C_)], dtype=np.uint64).view(np.int64)))

    def step(hop):
        _lib.check(_lib.lib().qmc_run_chains_host(dm.handle, C_, 1, s0.data_ptr(), 0, hop, 1.0, GAMMA[0], GAMMA[1], 0.8, 5,
                                                  _lib.DTYPE_F32, prop.data_ptr(), acc.data_ptr(), fin.data_ptr(),
                                                  cnt.data_ptr(), hist.data_ptr()))
        s0.copy_(fin)

    step(0)
    step(1)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for hop in range(2, 8):
        step(hop)
        assert torch.cuda.mem_get_info()[0] == free0
    # the stepped chain equals one 8-hop call
    ref = q.quantum_enhanced_mcmc(8, m, mx, GAMMA, temperature=1.0, n_chains=C_, seed=5, dtype="fp32")
    assert np.array_equal(fin.numpy().view(np.uint64), ref.final_state)


@pytest.mark.parametrize("n", [9, 20])
def test_classical_chains_host_equals_the_device_pointer_variant(n):
    q = _api()
    from qumcmc_b200 import _lib
    from qumcmc_b200.classical_mcmc_routines import lower_proposals
    from qumcmc_b200.classical_mixers import CombineProposals, CustomProposals, FixedWeightProposals, UniformProposals
    J, h = models.packaged_random_model(n, 4)
    m = q.IsingEnergyFunction(J, h)
    rule = CombineProposals([UniformProposals(n), FixedWeightProposals(n, 2), CustomProposals(n, [0, n - 1])], [0.2, 0.5, 0.3])
    C_, H, seed = 33, 50, 17
    dev = q.classical_mcmc(H, m, rule, temperature=0.9, n_chains=C_, seed=seed, chain_id0=2)
    kinds, params, probs = lower_proposals(rule)
    prop, acc, fin, cnt, hist = _host_buffers(C_, H, n)
    dm = m.device_model(0)
    _lib.check(_lib.lib().qmc_classical_chains_host(dm.handle, C_, H, None, 2, 0, 0.9, len(kinds), kinds.ctypes.data,
                                                    params.ctypes.data, probs.ctypes.data, seed, _vp(prop), _vp(acc),
                                                    _vp(fin), _vp(cnt), _vp(hist)))
    assert np.array_equal(prop, dev.proposals) and np.array_equal(acc, dev.accepted)
    assert np.array_equal(fin, dev.final_state) and np.array_equal(cnt, dev.acc_count) and np.array_equal(hist, dev.hamming_hist)
    # and both equal the oracle's chain
    for c in (0, C_ - 1):
        p_o, a_o = oracle.classical_chain(J, h, list(kinds), [int(x) for x in params], list(probs), int(dev.initial[c]), H,
                                          0.9, seed, chain_id=2 + c)
        assert np.array_equal(prop[c], p_o) and np.array_equal(acc[c], a_o)


def test_integration_md_stub_runs_verbatim():
    """The ctypes stub of INTEGRATION.md section 2 and the quantum_enhanced_mcmc replacement of section 3, executed as
    written (the only things supplied are what the text says the reference supplies: lower(), get_random_state,
    MCMCChain, MCMCState), must reproduce the oracle's chain for the seed it drew."""
    q = _api()
    from qumcmc_b200 import _lib
    _lib.lib()  # loads csrc/libqumcmc_b200.so; its SONAME makes the stub's CDLL("libqumcmc_b200.so") resolve to it
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "class B200Model" in b)
    call = next(b for b in blocks if b.lstrip().startswith("def quantum_enhanced_mcmc("))

    def lower(mixer):
        groups, probs = mixer.term_groups()
        return [[([b for b in range(mixer.n_qubits) if (mask >> b) & 1], w) for mask, w in g] for g in groups], probs

    ns = {"lower": lower, "get_random_state": q.get_random_state, "MCMCChain": q.MCMCChain,
          "MCMCState": q.MCMCState}
    exec(stub, ns)
    exec(call, ns)
    n = 9
    J, h = models.packaged_random_model(n, 7)
    model = q.IsingEnergyFunction(J, h)
    mixer = q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.CustomMixer(n, [[0, 1, 2], [3, 4, 5]])], [0.75, 0.25])
    np.random.seed(123)
    chain = ns["quantum_enhanced_mcmc"](40, model, mixer, (0.3, 0.5), initial_state="101010101", temperature=1.2)
    np.random.seed(123)
    seed = int(np.random.randint(0, 2 ** 62))
    _, masks, w, _ = mixer.lower()
    props, acc = oracle.run_chain(J, h, [int(x) for x in masks], list(w), int("101010101", 2), 40, 1.2, (0.3, 0.5), 0.8, seed,
                                  chain_id=0)
    assert len(chain.states) == 41 and chain.states[0].bitstring == "101010101"
    assert [int(s.bitstring, 2) for s in chain.states[1:]] == [int(x) for x in props]
    assert [s.accepted for s in chain.states[1:]] == [bool(a) for a in acc]
    mc = npo.markov_chain_from(int("101010101", 2), props, acc)
    assert chain.markov_chain == [f"{int(x):09b}" for x in mc]


@pytest.mark.parametrize("n,mixer,dtype,C_,H", [(9, "one", "fp64", 40, 60), (14, "one", "fp32", 24, 12), (14, "two", "fp32", 6, 5)])
def test_device_visit_histogram_quantum(n, mixer, dtype, C_, H):
    """visit_hist[s] = number of (chain, k) with markov_chain_k == s, k = 0..H (entry 0 = the initial state) -- the dense
    form of MCMCChain.get_accepted_dict summed over the chains; a run resumed with hop0 adds up to the same counts."""
    q = _api()
    J, h = models.packaged_random_model(n, 6)
    m = q.IsingEnergyFunction(J, h)
    mx = q.GenericMixer.OneBodyMixer(n) if mixer == "one" else q.GenericMixer.TwoBodyMixer(n)
    b = q.quantum_enhanced_mcmc(H, m, mx, GAMMA, temperature=1.0, n_chains=C_, seed=9, dtype=dtype, visit_histogram=True)
    want = np.zeros(1 << n, dtype=np.int64)
    for c in range(C_):
        want += np.bincount(npo.markov_chain_from(int(b.initial[c]), b.proposals[c], b.accepted[c]).astype(np.int64),
                            minlength=1 << n)
    assert b.visit_hist.sum() == C_ * (H + 1)
    assert np.array_equal(b.visit_hist, want)
    # the same through MCMCChain.get_accepted_dict of one chain
    d = b[0].get_accepted_dict()
    assert sum(d.values()) == H + 1
    # resumed in two calls on one attached accumulator: no double count of the initial state, no record kept
    import torch
    dm = m.device_model(0)
    acc_t = torch.zeros(1 << n, dtype=torch.int64, device="cuda")
    dm.attach_visit_histogram(acc_t)
    try:
        h1 = H // 2
        b1 = q.quantum_enhanced_mcmc(h1, m, mx, GAMMA, temperature=1.0, n_chains=C_, seed=9, dtype=dtype, record=False)
        q.quantum_enhanced_mcmc(H - h1, m, mx, GAMMA, initial_state=[int(x) for x in b1.final_state], temperature=1.0,
                                n_chains=C_, seed=9, dtype=dtype, hop0=h1, record=False)
    finally:
        dm.attach_visit_histogram(None)
    assert np.array_equal(acc_t.cpu().numpy(), want)
    assert not b1.recorded and b1.acceptance_rate() == pytest.approx(b.accepted[:, :h1].mean())


def test_device_visit_histogram_classical_and_detach():
    q = _api()
    n, C_, H = 10, 50, 200
    J, h = models.packaged_random_model(n, 8)
    m = q.IsingEnergyFunction(J, h)
    b = q.classical_mcmc(H, m, temperature=1.0, n_chains=C_, seed=4, visit_histogram=True)
    want = np.zeros(1 << n, dtype=np.int64)
    for c in range(C_):
        want += np.bincount(npo.markov_chain_from(int(b.initial[c]), b.proposals[c], b.accepted[c]).astype(np.int64),
                            minlength=1 << n)
    assert np.array_equal(b.visit_hist, want)
    # empirical distribution of long chains approaches Boltzmann (sanity of the histogram as an evaluation tool)
    ex = q.Exact_Sampling(m, 1.0)
    emp = b.visit_hist / b.visit_hist.sum()
    assert 0.5 * np.abs(emp - ex.probs).sum() < 0.4
    # detached: a later call must not write anywhere
    b2 = q.classical_mcmc(5, m, temperature=1.0, n_chains=4, seed=4)
    assert b2.visit_hist is None
    from qumcmc_b200 import _lib
    big = q.IsingEnergyFunction(*models.packaged_random_model(31, 1))
    with pytest.raises(_lib.QmcUnsupported):
        import torch
        _lib.check(_lib.lib().qmc_set_visit_histogram(big.device_model(0).handle, torch.zeros(8, dtype=torch.int64, device="cuda").data_ptr()))
