"""CPU: the address arithmetic of the unit-compacted slabs, restated in numpy — the pack kernel's layout
(tests/helpers.unit_layout) and the staged SpMM's copy / read offsets, expression by expression as in
csrc/spmm_units.cu (g % 4 == 0) and csrc/spmm_units_even.cu (g % 4 == 2).  It pins the contract the GPU tests
then hold the kernels to: every run the SpMM copies is 16-byte aligned, stays inside the node's dense row, and
every live lane finds its g values inside the copied pieces."""
import numpy as np
import pytest

from helpers import unit_layout


def _pack_row(dense: np.ndarray, live: np.ndarray, g: int):
    """dense [g, h] -> the row as the pack kernel leaves it (float32 [g*h], unused slots = NaN)."""
    h = live.shape[0]
    hdr, slot = unit_layout(live, g)
    row = np.full(g * h, np.nan, dtype=np.float32)
    for u in np.nonzero(live)[0]:
        row[slot[u] * g: slot[u] * g + g] = dense[:, u]
    return hdr, slot, row


@pytest.mark.parametrize("g", [2, 4, 6, 8, 10, 12, 14, 16])
@pytest.mark.parametrize("h", [32, 96, 256, 1024])
@pytest.mark.parametrize("density", [0.0, 0.07, 0.5, 0.93, 1.0])
def test_runs_are_aligned_inside_the_row_and_complete(g, h, density):
    rng = np.random.default_rng(g * 10_000 + h + int(density * 100))
    for trial in range(6):
        live = rng.random(h) < density
        if trial == 0:
            live[:] = density >= 0.5          # all dead / all live
        dense = rng.standard_normal((g, h)).astype(np.float32)
        hdr, slot, row = _pack_row(dense, live, g)
        assert slot.max(initial=-1) < h                              # the compact row fits the dense pitch
        raw = row.view(np.uint8)
        out = np.zeros((g, h), dtype=np.float32)
        for w, (mask, first) in enumerate(hdr):
            k = bin(mask).count("1")
            if g % 4 == 0:                                           # spmm_units_staged_kernel
                G4 = g // 4
                p16 = first * G4                                     # hd.y * G4 float4 from the row start
                n16 = k * G4
                slot_bytes = 512 * G4
            else:                                                    # spmm_units_even_kernel
                G2 = g // 2
                assert first % 2 == 0
                p16 = (first >> 1) * G2
                n16 = (k * G2 + 1) >> 1
                slot_bytes = 256 * G2
            assert 16 * (p16 + n16) <= g * h * 4                      # the copy never leaves the node's row
            assert 16 * n16 <= slot_bytes                             # ... nor its ring slot
            ring = raw[16 * p16: 16 * (p16 + n16)].copy()             # what cp.async brings to shared memory
            sl = 0
            for lane in range(32):
                if not (mask >> lane) & 1:
                    continue
                at = sl * g * 4                                       # 16*sl*G4 resp. 8*sl*G2 bytes
                assert at + g * 4 <= ring.size
                assert at % (16 if g % 4 == 0 else 8) == 0
                out[:, 32 * w + lane] = ring[at: at + g * 4].view(np.float32)
                sl += 1
        assert np.array_equal(out[:, live], dense[:, live])
        assert not out[:, ~live].any()


@pytest.mark.parametrize("G2", [1, 3, 5, 7])
def test_even_group_shared_memory_reads_are_conflict_free(G2):
    """64-bit reads at a slot stride of 2*G2 words (G2 odd): the 16 lanes of a half-warp, holding consecutive
    slots, touch 16 distinct bank pairs whatever the first slot."""
    for first in range(0, 40):
        banks = set()
        for sl in range(first, first + 16):
            word = 2 * G2 * sl
            banks.add(word % 32)
            assert word % 2 == 0
        assert len(banks) == 16
