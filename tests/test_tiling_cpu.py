"""The M tiling of dn_gemm (csrc/gemm.cu make_tiling / decode_tile — host-compiled copy of the code the kernel runs, exported as
dn_gemm_tile_rows): with packed rows every (utterance, frame) of a ragged batch is computed by exactly one accumulator row, tiles
are filled across utterance boundaries, and only whole-tile-in-one-utterance tiles take the single-box path.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from diffnorm_b200._lib import lib


def tile_map(B, T, rc, ctas):
    mt = C.c_int32()
    assert lib.dn_gemm_tile_rows(B, T, rc, ctas, 0, 0, C.byref(mt), None, None, None) == 0
    rows = []
    b, t, nb = (C.c_int32 * 128)(), (C.c_int32 * 128)(), C.c_int32()
    for m in range(mt.value):
        for r in range(ctas):
            assert lib.dn_gemm_tile_rows(B, T, rc, ctas, m, r, None, b, t, C.byref(nb)) == 0
            rows.append((np.array(b[:]), np.array(t[:]), nb.value))
    return mt.value, rows


@pytest.mark.parametrize("B,T", [(5, 600), (7, 200), (3, 1000), (9, 136), (1, 77), (64, 1000), (2, 24)])
@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("rc", [0, 8, 16, 32, 64])
def test_every_frame_is_computed_exactly_once(B, T, rc, ctas):
    m_tiles, rows = tile_map(B, T, rc, ctas)
    seen = np.zeros((B, T), dtype=np.int32)
    for b, t, nb in rows:
        ok = (b < B) & (t < T)
        np.add.at(seen, (b[ok], t[ok]), 1)
        if nb == 1:      # one box: 128 consecutive frames of one utterance
            assert (b == b[0]).all() and (np.diff(t) == 1).all()
        else:            # one box per chunk: consecutive frames inside a chunk, chunk origins on chunk multiples
            assert nb == 128 // rc
            for i in range(nb):
                bb, tt = b[i * rc:(i + 1) * rc], t[i * rc:(i + 1) * rc]
                assert (bb == bb[0]).all() and (np.diff(tt) == 1).all() and tt[0] % rc == 0
    assert (seen == 1).all()
    bm = 128 * ctas
    if rc == 0:
        assert m_tiles == B * -(-T // bm)
    else:
        per_t = -(-T // rc)
        assert m_tiles == -(-(B * per_t * rc) // bm)       # only the last chunk of an utterance and the last tile pad
        assert m_tiles * bm - B * T < B * rc + bm


def test_packed_rows_save_the_tile_padding_of_a_ragged_batch():
    # T = 600: 3 pair tiles (768 rows) per utterance without packing, 19 chunks of 32 (608 rows) with it
    plain, _ = tile_map(106, 600, 0, 2)
    packed, rows = tile_map(106, 600, 32, 2)
    assert plain * 256 == 106 * 768 and packed * 256 == -(-106 * 608 // 256) * 256
    straddling = sum(1 for _, _, nb in rows if nb > 1)
    assert 0 < straddling <= 2 * 106         # at most the tiles around each utterance boundary take the per-chunk path


def test_rejects_bad_arguments():
    assert lib.dn_gemm_tile_rows(0, 10, 0, 1, 0, 0, None, None, None, None) == -1
    assert lib.dn_gemm_tile_rows(2, 10, 24, 1, 0, 0, None, None, None, None) == -1
    assert lib.dn_gemm_tile_rows(2, 10, 32, 3, 0, 0, None, None, None, None) == -1
    b = (C.c_int32 * 128)()
    assert lib.dn_gemm_tile_rows(2, 10, 32, 1, 99, 0, None, b, None, None) == -1
