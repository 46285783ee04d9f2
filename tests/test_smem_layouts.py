"""Static analysis of the shared-memory transpositions of the sweep kernels (CPU; compute-sanitizer is closed on the GPU
pool, see profiles/sanitizer_r2/README.md).  The library exports the slot / offset functions its kernels use
(qmc_debug_smem_slot, qmc_debug_layout128); here every (thread, register) pair is enumerated:

* each layout is a bijection threads x registers -> tile slots: no two threads ever write the same slot, and the steps
  that skip a barrier because "each thread rewrites exactly the slots it read" (sweep.cu: lo_b_to_smem after lo_smem_to_b,
  the HI8 / HIQ B'-layout write-backs) touch thread-private slot sets;
* the two layouts of a tile cover the same slot set (what one layout wrote is what the other reads);
* every warp-wide access is bank-conflict free (16 distinct 8-byte banks per half-warp for 64-bit accesses, 8 distinct
  16-byte bank groups per quarter-warp for 128-bit accesses);
* complex128 tiles: both layouts address the same 4096 amplitudes of the statevector, layout 1 keeps the 32 lanes of a warp
  on consecutive amplitudes, and the two layouts' active register qubits are exactly the qubits of the tile's group."""
import ctypes as C

import numpy as np
import pytest


def _lib():
    from qumcmc_b200 import _lib
    return _lib.lib()


@pytest.mark.parametrize("layout,nreg,width", [(0, 64, 8), (1, 64, 8), (2, 64, 8), (3, 64, 8), (4, 32, 16), (5, 32, 16)])
def test_c64_layouts_are_bijections_without_bank_conflicts(layout, nreg, width):
    L = _lib()
    slots = np.array([[L.qmc_debug_smem_slot(layout, t, j) for j in range(nreg)] for t in range(128)])
    assert sorted(slots.ravel().tolist()) == list(range(128 * nreg))          # bijection: thread-private slot sets
    lanes_per_phase = 128 // width                                            # lanes served per shared-memory wavefront
    for j in range(nreg):
        for w in range(4):
            for ph in range(32 // lanes_per_phase):
                lanes = slots[32 * w + ph * lanes_per_phase: 32 * w + (ph + 1) * lanes_per_phase, j]
                assert len(set((lanes % lanes_per_phase).tolist())) == lanes_per_phase, (layout, j, w, ph)


def test_c64_layout_pairs_cover_the_same_tile():
    L = _lib()
    for a, b, nreg in ((0, 1, 64), (2, 3, 64), (4, 5, 32)):
        sa = {L.qmc_debug_smem_slot(a, t, j) for t in range(128) for j in range(nreg)}
        sb = {L.qmc_debug_smem_slot(b, t, j) for t in range(128) for j in range(nreg)}
        assert sa == sb
    assert L.qmc_debug_smem_slot(9, 0, 0) == -1


@pytest.mark.parametrize("geometry", [0, 2, 3, 4, 5, 6, 7, 8, 9])
def test_c128_tile_geometries(geometry):
    L = _lib()
    out = (C.c_longlong * 4)()
    n = 20 if geometry == 0 else 22 - geometry
    group = set(range(10)) if geometry == 0 else set(range(10, n))
    tab = {}
    for layout in (1, 2):
        rows = np.zeros((128, 32, 4), dtype=np.int64)
        for t in range(128):
            for j in range(32):
                assert L.qmc_debug_layout128(geometry, layout, t, j, out) == 0
                rows[t, j] = list(out)
        tab[layout] = rows
        assert sorted(rows[:, :, 0].ravel().tolist()) == list(range(4096))      # (thread, register) -> local index: bijection
        assert sorted(rows[:, :, 2].ravel().tolist()) == list(range(4096))      # swizzled slots: bijection
        for j in range(32):                                                     # 128-bit accesses: quarter-warps conflict-free
            for q in range(16):
                lanes = rows[8 * q: 8 * q + 8, j, 2]
                assert len(set((lanes % 8).tolist())) == 8, (geometry, layout, j, q)
    # both layouts address the same amplitudes through the same slots
    m1 = dict(zip(tab[1][:, :, 0].ravel().tolist(), zip(tab[1][:, :, 1].ravel().tolist(), tab[1][:, :, 2].ravel().tolist())))
    m2 = dict(zip(tab[2][:, :, 0].ravel().tolist(), zip(tab[2][:, :, 1].ravel().tolist(), tab[2][:, :, 2].ravel().tolist())))
    assert m1 == m2
    offs = np.array(sorted(v[0] for v in m1.values()))
    assert len(set(offs.tolist())) == 4096 and offs.max() < (1 << n)
    # layout 1: the 32 lanes of a warp read consecutive amplitudes in runs of min(32, 2^LC)
    run = 32 if geometry == 0 else min(32, 1 << geometry)
    for w in range(4):
        o = tab[1][32 * w: 32 * w + 32, 0, 1]
        assert all(o[k + 1] - o[k] == 1 for k in range(31) if (k + 1) % run)
    # active register qubits of the two layouts = the tile's qubit group, each exactly once
    seen = []
    for layout in (1, 2):
        act = int(tab[layout][0, 0, 3])
        for b in range(5):
            if (act >> b) & 1:
                d = int(tab[layout][0, 1 << b, 1]) - int(tab[layout][0, 0, 1])   # index offset of register bit b
                assert d > 0 and d & (d - 1) == 0
                seen.append(d.bit_length() - 1)
    assert sorted(seen) == sorted(group), (geometry, seen)


@pytest.mark.parametrize("width", [8, 16])
def test_general_path_register_stages(width):
    """general.cu: the three register stages of the 4096-amplitude tile (256 threads x 16 registers) are bijections onto
    the tile, share one slot set, and are conflict-free for 64-bit (complex64, fp64 residual table) and 128-bit
    (complex128) elements."""
    L = _lib()
    loc = C.c_int()
    lanes_per_phase = 128 // width
    seen = None
    for stage in (0, 1, 2):
        slots = np.zeros((256, 16), dtype=np.int64)
        locs = np.zeros((256, 16), dtype=np.int64)
        for t in range(256):
            for j in range(16):
                slots[t, j] = L.qmc_debug_general_slot(stage, t, j, C.byref(loc))
                locs[t, j] = loc.value
        assert sorted(slots.ravel().tolist()) == list(range(4096)) and sorted(locs.ravel().tolist()) == list(range(4096))
        pairs = set(zip(locs.ravel().tolist(), slots.ravel().tolist()))
        assert seen is None or pairs == seen          # the same local index lives in the same slot in every stage
        seen = pairs
        for j in range(16):
            for w in range(8):
                for ph in range(32 // lanes_per_phase):
                    lanes = slots[32 * w + ph * lanes_per_phase: 32 * w + (ph + 1) * lanes_per_phase, j]
                    assert len(set((lanes % lanes_per_phase).tolist())) == lanes_per_phase, (stage, j, w, ph)
    # stage 2 is the HBM-facing stage: lanes 0..15 of a warp sit on local bits 0..3 (runs of 16 amplitudes)
    got = []
    for t in range(16):
        L.qmc_debug_general_slot(2, t, 0, C.byref(loc))
        got.append(loc.value)
    assert got == list(range(16))
    assert L.qmc_debug_general_slot(3, 0, 0, None) == -1
