"""CPU tests of the L2-resident group schedule (csrc/group_sched.cuh): the launch list of one hop is pure host logic,
exported as qmc_debug_group_plan.  Checked here against the definition of the sweep sequence of a proposal
(qumcmc/quantum_mcmc_routines.py:95-106: U = M (P M)^(r-1), restated in DESIGN section 4.2 as r alternating LO / HI sweeps,
the last one always LO): every chain must see exactly its own r sweeps, in order, on one stream."""
import ctypes as C

import numpy as np
import pytest

from qumcmc_b200 import _lib

LO_FIRST, LO_MID, LO_FINAL, LO_FINAL1, HI_FIRST, HI_MID = range(6)


def _plan(rhist, group, streams):
    L = _lib.lib()
    r_max = len(rhist) - 1
    n = int(sum(rhist))
    rh = (C.c_int * (r_max + 1))(*rhist)
    cap = 16 * (n + 1) * (r_max + 1)
    out = (C.c_int * (4 * cap))()
    cnt = L.qmc_debug_group_plan(n, r_max, rh, group, streams, out, cap)
    assert 0 <= cnt <= cap
    return np.array(out[: 4 * cnt], dtype=np.int64).reshape(cnt, 4)


def _rpos(rhist):
    """r of the chain at every position of order[]: odd r descending, then even r descending (schedule_kernel)."""
    r_max = len(rhist) - 1
    pos = []
    for parity in (1, 0):
        for r in range(r_max, 0, -1):
            if r % 2 == parity:
                pos += [r] * rhist[r]
    return pos


def _expected_sequence(r):
    """Roles of the r sweeps of one proposal: the last sweep is LO, sweeps alternate, the first is a FIRST role."""
    if r == 1:
        return [LO_FINAL1]
    seq = []
    for step in range(1, r + 1):
        lo = (step % 2) == (r % 2)
        if step == r:
            seq.append(LO_FINAL)
        elif step == 1:
            seq.append(LO_FIRST if lo else HI_FIRST)
        else:
            seq.append(LO_MID if lo else HI_MID)
    return seq


@pytest.mark.parametrize("rhist,group,streams", [
    ([0, 3, 0, 5, 2, 0, 4, 1], 2, 3),                 # ragged groups straddling r values, both parities
    ([0, 1], 4, 2),                                   # a single r = 1 chain
    ([0, 0, 7], 3, 4),                                # even r only: HI first
    ([0] + [37, 41, 29, 53, 31, 47, 43, 59, 61, 67, 71, 73, 79], 3, 3),   # C3-like spread, r = 1..13
    ([0, 5, 5, 5, 5], 1, 8),                          # one chain per group
    ([0, 2, 2, 2, 2], 64, 2),                         # one group per parity class
])
def test_every_chain_sees_its_sweeps_in_order_on_one_stream(rhist, group, streams):
    plan = _plan(rhist, group, streams)
    rpos = _rpos(rhist)
    n = len(rpos)
    seen = [[] for _ in range(n)]
    stream_of = [set() for _ in range(n)]
    for kind, off, count, stream in plan:
        assert count >= 1 and 0 <= off and off + count <= n and 0 <= stream < streams
        for p in range(off, off + count):
            seen[p].append(int(kind))
            stream_of[p].add(int(stream))
    for p in range(n):
        assert seen[p] == _expected_sequence(rpos[p]), (p, rpos[p], seen[p])
        assert len(stream_of[p]) == 1          # in-order execution needs one stream per chain
    # a launch never mixes chains of different groups, and groups are handed to the streams round robin
    n_odd = sum(rhist[r] for r in range(1, len(rhist), 2))
    for kind, off, count, stream in plan:
        start = 0 if off < n_odd else n_odd
        assert (off - start) // group == (off + count - 1 - start) // group
    sizes = sorted(set(int(c) for _, _, c, _ in plan))
    assert sizes[-1] <= group


def test_launch_count_matches_the_measured_figure():
    """C3 shape: 4096 chains, t ~ U{2..11}, dt = 0.8 => r in {2, 3, 5, 6, 7, 8, 10, 11, 12, 13}; groups of 3 chains give
    about 10.6 k launches per hop."""
    rs = [int(t / 0.8) for t in range(2, 12)]
    rhist = [0] * 14
    for i in range(4096):
        rhist[rs[i % 10]] += 1
    plan = _plan(rhist, 3, 3)
    assert 9000 < len(plan) < 12500
    sweeps = sum(int(c) for _, _, c, _ in plan)
    assert sweeps == sum(r * c for r, c in enumerate(rhist))


def test_bad_arguments():
    L = _lib.lib()
    rh = (C.c_int * 3)(0, 2, 2)
    out = (C.c_int * 64)()
    assert L.qmc_debug_group_plan(5, 2, rh, 2, 2, out, 16) == -1      # histogram does not sum to n_chains
    assert L.qmc_debug_group_plan(4, 2, rh, 0, 2, out, 16) == -1
    assert L.qmc_debug_group_plan(4, 2, rh, 2, 0, out, 16) == -1


def test_mixer_classification_of_the_reference_mixer_classes():
    """The dispatcher's rule (one-body -> rotation kernels, weight <= 2 -> XQ mode, anything else -> tabulated diagonal) on
    the lowered masks of the reference's mixer classes (qumcmc/mixers.py:21-151)."""
    import qumcmc_b200 as q
    L = _lib.lib()

    def cls(mixer, n):
        _, masks, _, _ = mixer.lower()
        arr = (C.c_uint64 * len(masks))(*[int(x) for x in masks])
        return L.qmc_debug_mixer_class(n, len(masks), arr)

    n = 16
    assert cls(q.GenericMixer.OneBodyMixer(n), n) == 1
    assert cls(q.CustomMixer(n, [[0], [3], [15]]), n) == 1
    assert cls(q.GenericMixer.TwoBodyMixer(n), n) == 2
    assert cls(q.CoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.GenericMixer.TwoBodyMixer(n)], [0.7, 0.3]), n) == 2
    assert cls(q.IncoherentMixerSum([q.GenericMixer.OneBodyMixer(n), q.CustomMixer(n, [[0, 5], [2]])], [0.5, 0.5]), n) == 2
    assert cls(q.GenericMixer.ThreeBodyMixer(n), n) == 3
    assert cls(q.CustomMixer(n, [[0, 1], [2, 3, 4, 5]]), n) == 3
    bad = (C.c_uint64 * 1)(1 << 20)
    assert L.qmc_debug_mixer_class(n, 1, bad) == -1       # qubit >= n
    assert L.qmc_debug_mixer_class(n, 0, bad) == -1
