"""CPU tests: the stage-1 oracle (oracle/csp_oracle.c) is pinned against
 (1) golden hashes generated from the UNMODIFIED reference csp.c (tests/golden/csp_golden.json,
     same values as SURVEY.md appendix A.4), and
 (2) the reference object itself (oracle/_ref/libref_csp.so) on random + adversarial inputs,
     when that object is present (authoring container; it also travels to the GPU box)."""
import itertools
import json
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "csp_golden.json")))["cases"]

I420, YV12, YV16, YV24, NV12, YUYV, UYVY, BGR, BGRA, FLIP = 1, 2, 3, 4, 5, 6, 7, 8, 9, 0x1000
OUTS = [2, 4, 6, 0xc, 0xe, 0xf]


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_oracle_matches_golden(case):
    w, h = case["w"], case["h"]
    n = ol.layout_bytes(ol.src_layout(case["in_csp"], w, h))
    src = ol.lcg_bytes(n, w, h)
    assert ol.fnv(src) == case["src_fnv"]
    dst = ol.oracle_convert(src, case["in_csp"], case["out_csp"], case["colmatrix"], case["fullrange"], w, h)
    if case["ret"] < 0:
        assert dst is None          # convert_fail, csp.c:93-97
    else:
        assert ol.fnv(dst) == case["dst_fnv"]
        assert [int(v) for v in dst[:4]] == case["dst_head"]


def test_rgb_coefficients_match_survey_table():
    import ctypes as C
    want = {
        (2, 0): [269262, 528618, 102662, 17301504, 155423, 305128, 460551, 538968064, 460551, 385654, 74897, 538968064],
        (2, 1): [313524, 615514, 119538, 524288, 176932, 347356, 524288, 538968063, 524288, 439026, 85262, 538968063],
        (1, 0): [191455, 644067, 65019, 17301504, 105533, 355018, 460551, 538968064, 460551, 418321, 42230, 538968064],
        (1, 1): [222927, 749942, 75707, 524288, 120138, 404150, 524288, 538968063, 524288, 476214, 48074, 538968063],
    }
    for (cm, fr), vals in want.items():
        out = (C.c_uint32 * 12)()
        ol.oracle().orc_rgb_coefficients(cm, fr, out)
        assert list(out) == vals


def _inputs(n, seed):
    rng = np.random.default_rng(seed)
    yield rng.integers(0, 256, n, dtype=np.uint8)
    yield np.zeros(n, dtype=np.uint8)
    yield np.full(n, 255, dtype=np.uint8)
    yield (rng.integers(0, 2, n, dtype=np.uint8) * 255).astype(np.uint8)   # saturation mix


@pytest.mark.skipif(not ol.have_ref_csp(), reason="oracle/_ref/libref_csp.so not built (needs /root/reference)")
@pytest.mark.parametrize("size", [(64, 48), (66, 48), (2, 2), (130, 6), (18, 34)])
def test_oracle_matches_reference_object_every_pair(size):
    """Every (input csp, flip, encoder csp, matrix, range) cell of the reference table,
    including the unregistered ones (-1), byte-for-byte."""
    w, h = size
    for in_csp, flip, out_csp in itertools.product(range(1, 10), (0, FLIP), OUTS):
        variants = [(2, 0), (2, 1), (1, 0), (1, 1)] if in_csp in (BGR, BGRA) and out_csp == 2 else [(2, 0)]
        n = ol.layout_bytes(ol.src_layout(in_csp, w, h))
        for cm, fr in variants:
            for src in _inputs(n, 1000 * in_csp + w):
                a = ol.ref_convert(src, in_csp | flip, out_csp, cm, fr, w, h)
                b = ol.oracle_convert(src, in_csp | flip, out_csp, cm, fr, w, h)
                assert (a is None) == (b is None), (in_csp, flip, out_csp)
                if a is not None:
                    assert np.array_equal(a, b), (in_csp, flip, out_csp, cm, fr)


def test_extensions_are_defined_from_reference_results():
    """RGB->NV12 ext == interleave of the reference-defined I420; 4:2:2->I444 ext == I422
    with each chroma sample doubled (DESIGN.md 'Extensions'; no reference path exists)."""
    w, h = 64, 48
    src = ol.lcg_bytes(ol.layout_bytes(ol.src_layout(BGR, w, h)), w, h)
    i420 = ol.oracle_convert(src, BGR | FLIP, 2, 2, 0, w, h)
    nv12 = ol.oracle_convert(src, BGR | FLIP, 4, 2, 0, w, h, ext=1)
    y, u, v = i420[:w * h], i420[w * h:w * h * 5 // 4], i420[w * h * 5 // 4:]
    assert np.array_equal(nv12[:w * h], y)
    assert np.array_equal(nv12[w * h::2], u) and np.array_equal(nv12[w * h + 1::2], v)
    src = ol.lcg_bytes(ol.layout_bytes(ol.src_layout(UYVY, w, h)), w, h)
    i422 = ol.oracle_convert(src, UYVY, 6, 2, 0, w, h)
    i444 = ol.oracle_convert(src, UYVY, 0xc, 2, 0, w, h, ext=2)
    assert np.array_equal(i444[:w * h], i422[:w * h])
    u422 = i422[w * h:w * h * 3 // 2].reshape(h, w // 2)
    assert np.array_equal(i444[w * h:2 * w * h].reshape(h, w), np.repeat(u422, 2, axis=1))
