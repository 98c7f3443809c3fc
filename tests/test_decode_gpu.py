"""SURVEY 8(f) row 4 on the device: the decoder-side output conversion (codec.c:2258-2292) through the C ABI
(x264vfw_cuda_dec_*), byte for byte against (a) the fixtures libswscale 9.1.100 itself produced
(tests/golden/decode_golden.json, no checker in the loop) and (b) the CPU checker on fresh inputs."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import oracle_lib as ol  # noqa: E402
from make_decode_golden import pixel_bytes  # noqa: E402

pytestmark = pytest.mark.gpu

CSP_I420, CSP_YV12, CSP_NV12, CSP_YUYV, CSP_UYVY, CSP_BGR, CSP_BGRA, VFLIP = 1, 2, 5, 6, 7, 8, 9, 0x1000
ALL = [CSP_I420, CSP_YV12, CSP_NV12, CSP_YUYV, CSP_UYVY, CSP_BGR, CSP_BGRA, CSP_BGR | VFLIP, CSP_BGRA | VFLIP]
GOLDEN = json.load(open(os.path.join(HERE, "golden", "decode_golden.json")))


@pytest.fixture(scope="module")
def dec():
    from x264vfw_b200 import decode
    return decode


def test_every_libswscale_fixture_is_reproduced_on_the_device(dec):
    from x264vfw_b200._lib import Context
    ctx = Context()
    for c in GOLDEN["cases"]:
        if c["h"] < 12:
            continue                   # the device path starts at 12 rows (the checker and libswscale go down to 10)
        src = c.get("src", 1)
        y, u, v = ol.decode_source(c["w"], c["h"], seed=c["spc"] + c["full"], pad=24, src_chroma=src)
        d = dec.Decompressor(c["csp"], c["w"], c["h"], c["spc"], c["full"], ctx=ctx, src_chroma=src)
        dib = d.decompress(y, u, v)
        d.close()
        assert ol.fnv(pixel_bytes(dib, c["csp"], c["w"], c["h"])) == c["fnv"], c


@pytest.mark.parametrize("w,h", [(16, 12), (70, 38), (258, 66), (1920, 1080), (1928, 1088), (4, 12)])
def test_device_matches_checker_on_fresh_inputs(dec, w, h):
    rng = np.random.default_rng(w * 7 + h)
    for kind in range(3):
        if kind == 0:
            y, u, v = (rng.integers(0, 256, s, dtype=np.uint8) for s in ((h, w + 9), (h // 2, w // 2 + 5), (h // 2, w // 2 + 5)))
        elif kind == 1:
            y, u, v = (rng.choice(np.array([0, 255], np.uint8), s) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2)))
        else:
            y, u, v = (rng.integers(0, 256, s, dtype=np.uint8) for s in ((h, w + 64), (h // 2, w // 2 + 32), (h // 2, w // 2 + 32)))
        y, u, v = y[:, :w], u[:, :w // 2], v[:, :w // 2]
        for csp in ALL:
            for spc, full in ((2, 0), (1, 1), (9, 0)):
                want = ol.oracle_decode_convert(y, u, v, csp, spc, full)
                d = dec.Decompressor(csp, w, h, spc, full)
                got = d.decompress(y, u, v)
                d.close()
                assert (got == want).all(), (w, h, hex(csp), spc, full, kind, int((got != want).sum()))


@pytest.mark.parametrize("w,h", [(16, 12), (70, 38), (258, 66), (1920, 1080), (1288, 728)])
def test_device_matches_checker_on_422_pictures(dec, w, h):
    """High 4:2:2 decoder pictures: no vertical chroma filter; YUY2 / UYVY / YV16 are (de)interleaves, RGB uses the single-line writers."""
    rng = np.random.default_rng(w * 3 + h)
    for kind in range(2):
        if kind == 0:
            y, u, v = (rng.integers(0, 256, s, dtype=np.uint8) for s in ((h, w + 9), (h, w // 2 + 5), (h, w // 2 + 5)))
        else:
            y, u, v = (rng.choice(np.array([0, 255], np.uint8), s) for s in ((h, w + 64), (h, w // 2 + 32), (h, w // 2 + 32)))
        y, u, v = y[:, :w], u[:, :w // 2], v[:, :w // 2]
        for csp in (3, CSP_YUYV, CSP_UYVY, CSP_BGR, CSP_BGRA, CSP_BGR | VFLIP, CSP_BGRA | VFLIP):
            for spc, full in ((2, 0), (1, 1), (9, 0)):
                want = ol.oracle_decode_convert(y, u, v, csp, spc, full, src_chroma=2)
                d = dec.Decompressor(csp, w, h, spc, full, src_chroma=2)
                got = d.decompress(y, u, v)
                d.close()
                assert (got == want).all(), (w, h, hex(csp), spc, full, kind, int((got != want).sum()))


@pytest.mark.parametrize("w,h", [(16, 12), (70, 38), (258, 66), (1920, 1080), (1284, 724)])
def test_device_matches_checker_on_444_pictures(dec, w, h):
    """High 4:4:4 decoder pictures: libswscale's per-pixel full-chroma writer (32-bit arithmetic that wraps on saturated colours) and the
    YV24 plane copy."""
    rng = np.random.default_rng(w * 5 + h)
    for kind in range(2):
        if kind == 0:
            y, u, v = (rng.integers(0, 256, (h, w + 12), dtype=np.uint8) for _ in range(3))
        else:
            y, u, v = (rng.choice(np.array([0, 255], np.uint8), (h, w + 64)) for _ in range(3))
        y, u, v = y[:, :w], u[:, :w], v[:, :w]
        for csp in (4, CSP_BGR, CSP_BGRA, CSP_BGR | VFLIP, CSP_BGRA | VFLIP):
            for spc, full in ((2, 0), (1, 1), (9, 0)):
                want = ol.oracle_decode_convert(y, u, v, csp, spc, full, src_chroma=3)
                d = dec.Decompressor(csp, w, h, spc, full, src_chroma=3)
                got = d.decompress(y, u, v)
                d.close()
                assert (got == want).all(), (w, h, hex(csp), spc, full, kind, int((got != want).sum()))


@pytest.mark.parametrize("w,h", [(24, 24), (70, 38), (258, 66), (1920, 1080)])
def test_device_matches_checker_where_the_chroma_resolution_changes(dec, w, h):
    """Every YUV output whose chroma resolution differs from the decoder picture's: libswscale's bicubic chroma scaler (4 taps for 2x up,
    8 taps for 2:1 down; horizontal into 15-bit intermediates, then vertical), luma copied; 4:4:4 -> YUY2 / UYVY through the single-line
    packed writers."""
    rng = np.random.default_rng(w + 9 * h)
    combos = ((1, (3, 4)), (2, (CSP_I420, CSP_YV12, CSP_NV12, 4)), (3, (CSP_I420, CSP_YV12, CSP_NV12, 3, CSP_YUYV, CSP_UYVY)))
    for src, csps in combos:
        cw, ch = (w if src == 3 else w // 2), (h if src >= 2 else h // 2)
        for kind in range(2):
            if kind == 0:
                y, u, v = (rng.integers(0, 256, s, dtype=np.uint8) for s in ((h, w + 9), (ch, cw + 5), (ch, cw + 5)))
            else:
                y, u, v = (rng.choice(np.array([0, 255], np.uint8), s) for s in ((h, w + 64), (ch, cw + 32), (ch, cw + 32)))
            y, u, v = y[:, :w], u[:, :cw], v[:, :cw]
            for csp in csps:
                want = ol.oracle_decode_convert(y, u, v, csp, 2, 0, src_chroma=src)
                assert want is not None
                d = dec.Decompressor(csp, w, h, 2, 0, src_chroma=src)
                got = d.decompress(y, u, v)
                d.close()
                assert (got == want).all(), (w, h, src, csp, kind, int((got != want).sum()))


@pytest.mark.parametrize("csp", ALL)
def test_batch_entry_on_resident_pictures(dec, csp):
    """x264vfw_cuda_dec_convert_batch: N pictures in device memory, one launch; every picture equals the checker's."""
    import torch
    w, h, n = 640, 360, 5
    cw, ch = w // 2, h // 2
    ys, cs = w + 64, cw + 32                                  # decoder-like linesize
    fb = ys * h + 2 * cs * ch
    fb = (fb + 255) & ~255
    host = np.zeros((n, fb), np.uint8)
    pics = []
    for f in range(n):
        y, u, v = ol.decode_source(w, h, seed=f + 1)
        host[f, :ys * h].reshape(h, ys)[:, :w] = y
        host[f, ys * h:ys * h + cs * ch].reshape(ch, cs)[:, :cw] = u
        host[f, ys * h + cs * ch:ys * h + 2 * cs * ch].reshape(ch, cs)[:, :cw] = v
        pics.append((y, u, v))
    src = torch.from_numpy(host).cuda()
    d = dec.Decompressor(csp, w, h, 1, 0)
    dfb = (d.picture_size + 255) & ~255
    dst = torch.zeros((n, dfb), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    base = src.data_ptr()
    d.decompress_batch(dst.data_ptr(), dfb, (base, base + ys * h, base + ys * h + cs * ch), (ys, cs, cs), fb, n)
    d.ctx.sync()
    out = dst.cpu().numpy()
    for f, (y, u, v) in enumerate(pics):
        want = ol.oracle_decode_convert(y, u, v, csp, 1, 0)
        assert (out[f, :d.picture_size] == want).all(), (hex(csp), f)
    d.close()


def test_refusals_and_geometry(dec):
    from x264vfw_b200._lib import CudaError
    assert dec.picture_get_size(CSP_BGR, 70, 38) == 212 * 38
    assert dec.picture_get_size(4, 64, 32) == 64 * 32 * 3 and dec.picture_get_size(10, 64, 32) == -1
    for args in ((CSP_YUYV | VFLIP, 64, 32), (4 | VFLIP, 64, 32), (4, 8, 32), (CSP_BGRA, 64, 8), (CSP_BGRA, 64, 10), (CSP_BGRA, 63, 32), (CSP_BGRA, 64, 0)):
        with pytest.raises(CudaError):
            dec.Decompressor(*args)
    with pytest.raises(CudaError):
        dec.Decompressor(CSP_I420, 16, 10, src_chroma=2)              # 10 chroma rows cannot be halved with libswscale's full 8 taps
    with pytest.raises(CudaError):
        dec.Decompressor(CSP_BGRA, 64, 32, src_chroma=4)
    # codec.c:1930-1980
    from x264vfw_b200.csp import fourcc
    assert dec.decompress_query(64, 32, 0, 32, 64, 32) == dec.ICERR_OK
    assert dec.decompress_query(64, 32, 0, 24, 64, -32) == dec.ICERR_OK
    assert dec.decompress_query(64, 32, fourcc("YUY2"), 16, 64, 32) == dec.ICERR_OK
    assert dec.decompress_query(64, 32, 0, 32, 32, 32) == dec.ICERR_BADFORMAT
    assert dec.decompress_query(64, 32, 0, 16, 64, 32) == dec.ICERR_BADFORMAT
    assert dec.decompress_query(64, 32, 0, 32, 64, 32, out_size_image=100) == dec.ICERR_BADFORMAT


def test_bottom_up_is_the_top_down_picture_with_rows_reversed(dec):
    w, h = 128, 64
    y, u, v = ol.decode_source(w, h, seed=9)
    a = dec.Decompressor(CSP_BGRA, w, h).decompress(y, u, v).reshape(h, w * 4)
    b = dec.Decompressor(CSP_BGRA | VFLIP, w, h).decompress(y, u, v).reshape(h, w * 4)
    assert (a[::-1] == b).all()
