"""SURVEY 8(f) row 4 -- the decoder-side output conversion (codec.c:2258-2292 -> libswscale).  CPU tests: the checker
(oracle/decode_oracle.c) against fixtures made by libswscale 9.1.100 itself (tests/golden/make_decode_golden.py),
and against the library live wherever this image's opencv wheel can be imported."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import oracle_lib as ol  # noqa: E402
import swsref as sr  # noqa: E402
from make_decode_golden import pixel_bytes  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "decode_golden.json")))


def case_id(c):
    return "%dx%d-%s-csp%x-spc%d-%s" % (c["w"], c["h"], {1: "420", 2: "422", 3: "444"}[c.get("src", 1)], c["csp"], c["spc"], "pc" if c["full"] else "tv")


SMALL = [c for c in GOLDEN["cases"] if c["w"] <= 320]
LARGE = [c for c in GOLDEN["cases"] if c["w"] > 320]


@pytest.mark.parametrize("c", SMALL, ids=case_id)
def test_checker_reproduces_libswscale_fixture(c):
    src = c.get("src", 1)
    y, u, v = ol.decode_source(c["w"], c["h"], seed=c["spc"] + c["full"], pad=24, src_chroma=src)
    dib = ol.oracle_decode_convert(y, u, v, c["csp"], c["spc"], c["full"], src_chroma=src)
    assert dib is not None
    assert ol.fnv(pixel_bytes(dib, c["csp"], c["w"], c["h"])) == c["fnv"]


def test_checker_reproduces_libswscale_fixture_full_sizes():
    for c in LARGE:
        src = c.get("src", 1)
        y, u, v = ol.decode_source(c["w"], c["h"], seed=c["spc"] + c["full"], pad=24, src_chroma=src)
        dib = ol.oracle_decode_convert(y, u, v, c["csp"], c["spc"], c["full"], src_chroma=src)
        assert ol.fnv(pixel_bytes(dib, c["csp"], c["w"], c["h"])) == c["fnv"], case_id(c)


def test_sample_picture_byte_for_byte():
    y, u, v = ol.decode_source(16, 10, seed=2, pad=24)
    dib = ol.oracle_decode_convert(y, u, v, sr.CSP_BGRA, 2, 0)
    assert dib.tolist() == GOLDEN["sample_16x10_bgra"]


@pytest.mark.skipif(not sr.available(), reason="libswscale 9 (opencv wheel) not importable here")
def test_checker_against_the_live_library_on_fresh_inputs():
    """Inputs the fixtures do not hold: other seeds, extreme sample values, flat pictures."""
    rng = np.random.default_rng(2026)
    for w, h in ((24, 18), (88, 50), (200, 120)):
        pics = [tuple(rng.integers(0, 256, s, dtype=np.uint8) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2))),
                tuple(rng.choice(np.array([0, 255], np.uint8), s) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2))),
                (np.full((h, w), 16, np.uint8), np.full((h // 2, w // 2), 128, np.uint8), np.full((h // 2, w // 2), 128, np.uint8))]
        for y, u, v in pics:
            for csp in (sr.CSP_BGRA, sr.CSP_BGR, sr.CSP_YUYV, sr.CSP_UYVY, sr.CSP_NV12, sr.CSP_YV12, sr.CSP_BGRA | sr.CSP_VFLIP):
                for spc, full in ((2, 0), (1, 1)):
                    a = sr.decompress_convert(y, u, v, csp, spc, full)
                    b = ol.oracle_decode_convert(y, u, v, csp, spc, full)
                    assert (pixel_bytes(a, csp, w, h) == pixel_bytes(b, csp, w, h)).all(), (w, h, hex(csp), spc, full)


@pytest.mark.skipif(not sr.available(), reason="libswscale 9 (opencv wheel) not importable here")
def test_checker_against_the_live_library_on_422_and_444_pictures():
    rng = np.random.default_rng(422)
    for w, h in ((24, 18), (88, 50)):
        y, u, v = (rng.integers(0, 256, s, dtype=np.uint8) for s in ((h, w), (h, w // 2), (h, w // 2)))
        for csp in (sr.CSP_BGRA, sr.CSP_BGR, sr.CSP_YUYV, sr.CSP_UYVY, sr.CSP_YV16, sr.CSP_BGRA | sr.CSP_VFLIP):
            for spc, full in ((2, 0), (1, 1), (9, 0)):
                a = sr.decompress_convert(y, u, v, csp, spc, full, src_chroma=2)
                b = ol.oracle_decode_convert(y, u, v, csp, spc, full, src_chroma=2)
                assert (pixel_bytes(a, csp, w, h) == pixel_bytes(b, csp, w, h)).all(), (w, h, hex(csp), spc, full)
    for w, h in ((24, 18), (90, 50)):                                    # High 4:4:4 pictures: RGB and the YV24 copy
        for y, u, v in ((rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (h, w), dtype=np.uint8)),
                        tuple(rng.choice(np.array([0, 255], np.uint8), (h, w)) for _ in range(3))):
            for csp in (sr.CSP_BGRA, sr.CSP_BGR, sr.CSP_YV24, sr.CSP_BGR | sr.CSP_VFLIP):
                for spc, full in ((2, 0), (1, 1), (9, 0), (7, 1)):
                    a = sr.decompress_convert(y, u, v, csp, spc, full, src_chroma=3)
                    b = ol.oracle_decode_convert(y, u, v, csp, spc, full, src_chroma=3)
                    assert (pixel_bytes(a, csp, w, h) == pixel_bytes(b, csp, w, h)).all(), (w, h, hex(csp), spc, full)
    # every YUV output of every picture format: where the chroma resolution changes, libswscale's scaler runs on the chroma planes
    # (4 taps up, 8 taps down), packed 4:2:2 from 4:4:4 adds the single-line packed writers
    for w, h in ((24, 24), (90, 50), (64, 32)):
        for src in (1, 2, 3):
            cw, ch = (w if src == 3 else w // 2), (h if src >= 2 else h // 2)
            for y, u, v in ((rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (ch, cw), dtype=np.uint8), rng.integers(0, 256, (ch, cw), dtype=np.uint8)),
                            (rng.choice(np.array([0, 255], np.uint8), (h, w)), rng.choice(np.array([0, 255], np.uint8), (ch, cw)), rng.choice(np.array([0, 255], np.uint8), (ch, cw)))):
                for csp in (sr.CSP_I420, sr.CSP_YV12, sr.CSP_YV16, sr.CSP_YV24, sr.CSP_NV12, sr.CSP_YUYV, sr.CSP_UYVY):
                    a = sr.decompress_convert(y, u, v, csp, 2, 0, src_chroma=src)
                    b = ol.oracle_decode_convert(y, u, v, csp, 2, 0, src_chroma=src)
                    assert b is not None and (a == b).all(), (w, h, src, csp)
    # too small for libswscale's full tap count: refused, not approximated
    y, u, v = (rng.integers(0, 256, s, dtype=np.uint8) for s in ((10, 16), (10, 8), (10, 8)))
    assert ol.oracle_decode_convert(y, u, v, sr.CSP_I420, src_chroma=2) is None


def test_reference_context_never_gets_full_chroma_interpolation():
    """codec.c:2097 hands `flags` to the context BEFORE codec.c:2110-2111 adds SWS_FULL_CHR_H_INT to the local, so
    RGB output keeps one chroma sample per pixel pair: with flat luma, pixels 2x and 2x+1 of a row are equal."""
    y, u, v = ol.decode_source(64, 32, seed=3)
    y = np.full_like(y, 120)
    dib = ol.oracle_decode_convert(y, u, v, sr.CSP_BGRA, 2, 0).reshape(32, 64, 4)
    assert (dib[:, 0::2] == dib[:, 1::2]).all()


def test_geometry_and_refusals():
    assert ol.decode_picture_size(sr.CSP_BGR, 70, 38) == 212 * 38              # codec.c:489-492: rows padded to 4 bytes
    assert ol.decode_picture_size(sr.CSP_NV12, 64, 32) == 64 * 32 * 3 // 2
    y, u, v = ol.decode_source(64, 32)
    assert ol.oracle_decode_convert(y, u, v, sr.CSP_YUYV | sr.CSP_VFLIP) is None   # only RGB can be flipped (codec.c:510-527)
    assert ol.oracle_decode_convert(y, u, v, 10) is None                         # no such csp
    y, u, v = ol.decode_source(64, 8)
    assert ol.oracle_decode_convert(y, u, v, sr.CSP_BGRA) is None                # fewer than 5 chroma rows: not restated
