"""The CPU lookahead checker against its own frozen fingerprints (tests/golden/lookahead_golden.json).
Not a reference pin -- libx264 is absent from the reference tree -- but it keeps the target of the
GPU parity tests from moving unnoticed."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_lookahead_golden as mk  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "lookahead_golden.json")))


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_checker_reproduces_its_golden_fingerprints(case):
    fp = mk.fingerprint(case["preset"], case["w"], case["h"], case["frames"], case["over"])
    for k in ("types", "coded_order", "costs", "qp_offset_fnv", "qp_offset_aq_fnv"):
        assert fp[k] == case[k], (case["name"], k)


def test_block_metrics_against_a_matrix_formulation():
    """SAD and SATD of the checker against numpy: SATD(8x8) = sum over the four 4x4 blocks of sum|H D H^T|, halved
    per 8x4 half ([x264] pixel_satd_8x4), H the 4x4 Hadamard matrix -- an independent formulation, not a reference."""
    import numpy as np
    import oracle_lib as ol
    o = ol.oracle()
    H = np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, -1, 1], [1, -1, 1, -1]], dtype=np.int64)
    rng = np.random.default_rng(264)
    for k in range(200):
        a = rng.integers(0, 256, (8, 8), dtype=np.uint8)
        b = rng.integers(0, 256, (8, 8), dtype=np.uint8) if k % 4 else (a if k % 8 else 255 - a)
        if k % 5 == 0:
            a, b = (a > 127).astype(np.uint8) * 255, (b > 127).astype(np.uint8) * 255
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        d = a.astype(np.int64) - b.astype(np.int64)
        sad = int(np.abs(d).sum())
        satd = 0
        for half in range(2):
            s = 0
            for blk in range(2):
                s += int(np.abs(H @ d[4 * half:4 * half + 4, 4 * blk:4 * blk + 4] @ H.T).sum())
            satd += s >> 1
        assert o.orc_test_sad_8x8(a.ctypes.data, b.ctypes.data) == sad
        assert o.orc_test_satd_8x8(a.ctypes.data, b.ctypes.data) == satd
