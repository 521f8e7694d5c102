"""Host logic of the matcher (no GPU): which kernel serves a shape, and that its tiles cover every computed row / column
exactly once within the shared-memory budget of an SM (mrefsr_match_plan)."""
import ctypes

import pytest
from hypothesis import given, settings, strategies as st

from mrefsr_b200 import _lib
from mrefsr_b200 import matcher as MM

SMEM_MAX = 227 * 1024


def plan(c, h_in, w_in, h_ref, w_ref, ps=3, s_in=1, s_ref=1, mode=0):
    meta = (ctypes.c_int * 10)()
    rc = _lib.lib().mrefsr_match_plan(c, h_in, w_in, h_ref, w_ref, ps, s_in, s_ref, mode, meta)
    _lib.check(rc, 'mrefsr_match_plan')
    return list(meta)


def test_kernel_choice():
    # BASELINE config 2 (40 x 40 features, C = 256): diagonal form with the B strip, 120 x 254 outputs per tile
    m = plan(256, 40, 40, 40, 40)
    assert m[0] == 4 and m[3:5] == [120, 254] and m[9] == MM.MATCH_TC_BF16X3
    assert m[5] == m[6] == 37 * 40 + 38 and m[1] == 13 and m[2] == 6 and m[7] >= 2 and m[8] <= SMEM_MAX
    # config 4 / validation grids (128 / 125 wide): the strip does not fit two stages -> plain diagonal form
    assert plan(256, 128, 128, 128, 128)[0] == 3
    assert plan(256, 125, 125, 125, 125)[0] == 3
    assert plan(256, 75, 75, 75, 75)[0] == 4
    # the strip limit is a property of the REFERENCE width only
    assert plan(64, 9, 200, 9, 96)[0] == 4 and plan(64, 9, 20, 9, 97)[0] == 3
    # variants by flag, and the CUDA-core path for other patch sizes / strides / channel counts
    assert plan(256, 40, 40, 40, 40, mode=MM.MATCH_TC_BF16X3 | MM.FLAG_NO_BSTRIP)[0] == 3
    assert plan(256, 40, 40, 40, 40, mode=MM.MATCH_TC_BF16X3 | MM.FLAG_NO_DIAG)[0] == 2
    assert plan(256, 40, 40, 40, 40, mode=MM.MATCH_TC_BF16X3 | MM.FLAG_NO_STRIP)[0] == 1
    assert plan(256, 40, 40, 40, 40, ps=5)[0] == 0 and plan(256, 40, 40, 40, 40, s_ref=2)[0] == 0
    assert plan(48, 40, 40, 40, 40)[0] == 0 and plan(48, 40, 40, 40, 40)[9] == MM.MATCH_FP32
    with pytest.raises(RuntimeError):
        plan(48, 40, 40, 40, 40, mode=MM.MATCH_TC_BF16X3)
    with pytest.raises(RuntimeError):
        plan(64, 2, 40, 40, 40)


@settings(max_examples=200, deadline=None)
@given(st.sampled_from([64, 128, 256]), st.integers(3, 140), st.integers(3, 300), st.integers(3, 140), st.integers(3, 300),
       st.sampled_from([0, 1, 2, 1 | 0x800, 1 | 0x400, 1 | 0x100]))
def test_tiles_cover_the_grids(c, h_in, w_in, h_ref, w_ref, mode):
    m = plan(c, h_in, w_in, h_ref, w_ref, mode=mode)
    kernel, m_tiles, n_tiles, rows_per, cols_per, rows_in, rows_ref, stages, smem = m[:9]
    assert kernel in (1, 2, 3, 4)
    # pixel-linear indices up to the last valid 3 x 3 patch origin
    assert rows_in == (h_in - 3) * w_in + (w_in - 2) and rows_ref == (h_ref - 3) * w_ref + (w_ref - 2)
    # tiles own disjoint row / column ranges of rows_per / cols_per entries: the last one reaches the end, the one before does not
    assert (m_tiles - 1) * rows_per < rows_in <= m_tiles * rows_per
    assert (n_tiles - 1) * cols_per < rows_ref <= n_tiles * cols_per
    assert stages >= 2 and 0 < smem <= SMEM_MAX
    if kernel == 4:     # two stages must fit: up to 96 wide with hi + lo halves (3 passes), 128 with the single bf16 pass
        assert w_ref <= (128 if m[9] == MM.MATCH_TC_BF16 else 96)
    if kernel in (3, 4):
        assert (rows_per, cols_per) == (120, 254)       # 128 x 256 accumulators minus the diagonal look-ahead
