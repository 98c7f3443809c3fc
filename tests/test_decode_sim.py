"""The decoder-side conversion KERNEL SOURCE on the CPU: x264vfw_b200/csrc/decode_kernel.cuh -- kernels, host-side
tables and dispatch, the very file the sm_100a build compiles -- is compiled by g++ through tests/sim/decode_sim.cpp and run thread
by thread, against the checker and against the fixtures libswscale 9.1.100 produced.  No device needed: this is what the CPU suite
knows about the device path before it reaches the GPU box (tests/test_decode_gpu.py runs the same cases there)."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import oracle_lib as ol  # noqa: E402
from make_decode_golden import pixel_bytes  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "decode_golden.json")))
YUV_OUT = (1, 2, 3, 4, 5, 6, 7)
ALL_OUT = YUV_OUT + (8, 9, 8 | 0x1000, 9 | 0x1000)


@pytest.fixture(scope="module")
def sim():
    so = os.path.join(HERE, "sim", "_build", "libdecode_sim.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    srcs = [os.path.join(HERE, "sim", "decode_sim.cpp"), os.path.join(ROOT, "x264vfw_b200", "csrc", "decode_kernel.cuh"),
            os.path.join(ROOT, "include", "x264vfw_cuda.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so, srcs[0]], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.sim_dec_convert.restype = C.c_int
    lib.sim_dec_convert.argtypes = [C.c_int] * 6 + [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_size_t, C.c_size_t, C.c_int]
    lib.sim_last_error.restype = C.c_char_p
    return lib


def aligned(shape, align=64, offset=0):
    """uint8 array whose first byte sits `offset` bytes past an `align`-byte boundary (the kernels pick their 128-bit paths by alignment)."""
    n = int(np.prod(shape))
    raw = np.zeros(n + align + offset, np.uint8)
    o = (-raw.ctypes.data) % align + offset
    return raw[o:o + n].reshape(shape)


def run(sim, y, u, v, csp, spc, full, src, offset=0):
    h, w = y.shape
    size = ol.decode_picture_size(csp, w, h)
    ya, ua, va = aligned(y.shape, offset=offset), aligned(u.shape, offset=offset), aligned(v.shape, offset=offset)
    ya[:], ua[:], va[:] = y, u, v
    out = aligned((size,), offset=4 * (offset != 0))
    out[:] = 0
    rc = sim.sim_dec_convert(csp, w, h, src, spc, full, out.ctypes.data, ya.ctypes.data, ua.ctypes.data, va.ctypes.data,
                             ya.strides[0], ua.strides[0], va.strides[0], 0, 0, 1)
    return None if rc < 0 else out


def picture(rng, w, h, src, kind):
    cw, ch = (w if src == 3 else w // 2), (h if src >= 2 else h // 2)
    if kind == 0:
        return tuple(rng.integers(0, 256, s, dtype=np.uint8) for s in ((h, w), (ch, cw), (ch, cw)))
    return tuple(rng.choice(np.array([0, 255], np.uint8), s) for s in ((h, w), (ch, cw), (ch, cw)))


@pytest.mark.parametrize("w,h", [(16, 12), (24, 24), (64, 32), (70, 38), (136, 50)])
def test_kernel_source_matches_checker_for_every_picture_format_and_output(sim, w, h):
    rng = np.random.default_rng(w * 11 + h)
    n = 0
    for src in (1, 2, 3):
        for kind in (0, 1):
            y, u, v = picture(rng, w, h, src, kind)
            for csp in ALL_OUT:
                for spc, full in ((2, 0), (1, 1)) if csp & 0xff >= 8 else ((2, 0),):
                    want = ol.oracle_decode_convert(y, u, v, csp, spc, full, src_chroma=src)
                    for offset in (0, 1):                              # aligned buffers: 128-bit paths; offset by a byte: the generic ones
                        got = run(sim, y, u, v, csp, spc, full, src, offset)
                        if want is None:
                            assert got is None, (w, h, src, hex(csp))  # refused on both sides (too small for the full tap count)
                            continue
                        assert got is not None, (w, h, src, hex(csp), sim.sim_last_error())
                        assert (pixel_bytes(got, csp, w, h) == pixel_bytes(want, csp, w, h)).all(), (w, h, src, hex(csp), spc, full, kind, offset)
                        n += 1
    assert n > 100


def test_kernel_source_reproduces_the_libswscale_fixtures(sim):
    """Every fixture of up to 320x240 with at least 12 rows, without the checker in the loop."""
    n = 0
    for c in GOLDEN["cases"]:
        if c["w"] > 320 or c["h"] < 12:
            continue
        src = c.get("src", 1)
        y, u, v = ol.decode_source(c["w"], c["h"], seed=c["spc"] + c["full"], pad=24, src_chroma=src)
        got = run(sim, np.ascontiguousarray(y), np.ascontiguousarray(u), np.ascontiguousarray(v), c["csp"], c["spc"], c["full"], src)
        assert got is not None, (c, sim.sim_last_error())
        assert ol.fnv(pixel_bytes(got, c["csp"], c["w"], c["h"])) == c["fnv"], c
        n += 1
    assert n > 300


def test_batch_addressing_and_refusals(sim):
    """Two pictures per launch (blockIdx.z strides), and the same refusals as the C ABI."""
    rng = np.random.default_rng(5)
    w, h = 64, 32
    for src, csp in ((1, 9 | 0x1000), (2, 6), (3, 8), (1, 4), (3, 7), (1, 5)):
        cw, ch = (w if src == 3 else w // 2), (h if src >= 2 else h // 2)
        fb = w * h + 2 * cw * ch
        size = ol.decode_picture_size(csp, w, h)
        dfb = (size + 255) & ~255
        srcbuf, dst = aligned((2, fb)), aligned((2, dfb))
        pics = []
        for f in range(2):
            y, u, v = picture(rng, w, h, src, f)
            srcbuf[f, :w * h] = y.ravel(); srcbuf[f, w * h:w * h + cw * ch] = u.ravel(); srcbuf[f, w * h + cw * ch:] = v.ravel()
            pics.append((y, u, v))
        b = srcbuf.ctypes.data
        assert sim.sim_dec_convert(csp, w, h, src, 2, 0, dst.ctypes.data, b, b + w * h, b + w * h + cw * ch, w, cw, cw, fb, dfb, 2) == 0
        for f, (y, u, v) in enumerate(pics):
            assert (dst[f, :size] == ol.oracle_decode_convert(y, u, v, csp, 2, 0, src_chroma=src)).all(), (src, hex(csp), f)
    z = np.zeros(64, np.uint8).ctypes.data
    for args in ((6 | 0x1000, 64, 32, 1), (9, 63, 32, 1), (9, 64, 10, 1), (9, 64, 32, 4), (10, 64, 32, 1), (1, 16, 10, 2)):
        assert sim.sim_dec_convert(args[0], args[1], args[2], args[3], 2, 0, z, z, z, z, 64, 32, 32, 0, 0, 1) == -1, args
