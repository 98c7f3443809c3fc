"""GPU parity tests, stage 1: every CUDA converter, called through the C ABI, against the
CPU oracle (which tests/test_csp_oracle.py pins to the unmodified reference csp.c) and the
golden hashes generated from the reference.  Bit-exact: this is byte/integer work."""
import itertools
import json
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "csp_golden.json")))["cases"]
I420, YV12, YV16, YV24, NV12, YUYV, UYVY, BGR, BGRA, FLIP = 1, 2, 3, 4, 5, 6, 7, 8, 9, 0x1000
OUTS = [2, 4, 6, 0xc, 0xe, 0xf]


@pytest.fixture(scope="module")
def ctx():
    from x264vfw_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_table_api_matches_reference_golden(case):
    """x264vfw_cuda_csp_init table with host pointers == reference csp.c golden vectors."""
    from x264vfw_b200 import csp
    from x264vfw_b200._lib import CudaError
    w, h = case["w"], case["h"]
    src = ol.lcg_bytes(ol.layout_bytes(ol.src_layout(case["in_csp"], w, h)), w, h)
    t = csp.csp_init(case["out_csp"], case["colmatrix"], case["fullrange"])
    if case["ret"] < 0:
        with pytest.raises(CudaError):
            csp.convert_host(t, src, case["in_csp"], case["out_csp"], w, h)
        return
    dst = csp.convert_host(t, src, case["in_csp"], case["out_csp"], w, h)
    assert ol.fnv(dst) == case["dst_fnv"]


def _inputs(n, seed):
    rng = np.random.default_rng(seed)
    yield rng.integers(0, 256, n, dtype=np.uint8)
    yield np.full(n, 255, dtype=np.uint8)
    yield (rng.integers(0, 2, n, dtype=np.uint8) * 255).astype(np.uint8)


@pytest.mark.parametrize("size", [(64, 48), (66, 48), (2, 2), (130, 6), (18, 34), (1920, 1080)])
def test_every_registered_pair_matches_oracle(ctx, size):
    from x264vfw_b200 import csp
    from x264vfw_b200._lib import CudaError
    w, h = size
    big = w * h > 100000
    for in_csp, flip, out_csp in itertools.product(range(1, 10), (0, FLIP), OUTS):
        variants = [(2, 0), (2, 1), (1, 0), (1, 1)] if in_csp in (BGR, BGRA) and out_csp == 2 else [(2, 0)]
        n = ol.layout_bytes(ol.src_layout(in_csp, w, h))
        for cm, fr in variants:
            for k, src in enumerate(_inputs(n, 77 * in_csp + w)):
                if big and k:
                    break
                want = ol.oracle_convert(src, in_csp | flip, out_csp, cm, fr, w, h)
                if want is None:
                    with pytest.raises(CudaError):
                        csp.convert_ctx(ctx, src, in_csp | flip, out_csp, cm, fr, w, h)
                    break
                got = csp.convert_ctx(ctx, src, in_csp | flip, out_csp, cm, fr, w, h)
                assert np.array_equal(got, want), (in_csp, flip, out_csp, cm, fr, size)


def test_extensions_match_oracle(ctx):
    from x264vfw_b200 import csp
    for (w, h) in [(64, 48), (66, 6), (1920, 1080)]:
        for in_csp, flip in itertools.product((BGR, BGRA), (0, FLIP)):
            src = ol.lcg_bytes(ol.layout_bytes(ol.src_layout(in_csp, w, h)), w, h)
            want = ol.oracle_convert(src, in_csp | flip, 4, 1, 0, w, h, ext=1)
            got = csp.convert_ctx(ctx, src, in_csp | flip, 4, 1, 0, w, h, ext=csp.EXT_RGB_TO_NV12)
            assert np.array_equal(got, want)
        for in_csp in (YUYV, UYVY):
            src = ol.lcg_bytes(ol.layout_bytes(ol.src_layout(in_csp, w, h)), w, h)
            want = ol.oracle_convert(src, in_csp, 0xc, 2, 0, w, h, ext=2)
            got = csp.convert_ctx(ctx, src, in_csp, 0xc, 2, 0, w, h, ext=csp.EXT_422_TO_I444)
            assert np.array_equal(got, want)


@pytest.mark.parametrize("in_csp,out_csp,w,h", [(BGRA | FLIP, 2, 1920, 1080), (BGR | FLIP, 2, 1920, 1080),
                                                (YUYV, 2, 1280, 720), (UYVY, 6, 3840, 2160),
                                                (YV12, 2, 640, 480), (BGR | FLIP, 2, 66, 48)])
def test_device_resident_batch_matches_oracle(ctx, in_csp, out_csp, w, h):
    """x264vfw_cuda_csp_convert_batch: n frames per launch, device buffers (the roofline entry)."""
    import torch
    from x264vfw_b200 import csp
    nf = 3
    sfb, dfb = csp.frame_bytes(in_csp, out_csp, w, h)
    rng = np.random.default_rng(5)
    host = rng.integers(0, 256, nf * sfb, dtype=np.uint8)
    d_src = torch.from_numpy(host).cuda()
    d_dst = torch.zeros(nf * dfb, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    csp.convert_batch(ctx, d_src.data_ptr(), d_dst.data_ptr(), in_csp, out_csp, 2, 0, w, h, nf)
    ctx.sync()
    out = d_dst.cpu().numpy()
    nd = ol.layout_bytes(ol.dst_layout(out_csp, w, h))
    for f in range(nf):
        want = ol.oracle_convert(host[f * sfb:(f + 1) * sfb].copy(), in_csp, out_csp, 2, 0, w, h)
        assert np.array_equal(out[f * dfb:f * dfb + nd], want), f


def test_unaligned_buffers_take_the_scalar_path_and_stay_exact(ctx):
    import ctypes as C
    import torch
    from x264vfw_b200 import csp
    from x264vfw_b200._lib import lib
    w, h = 64, 48
    for in_csp, out_csp in [(BGRA | FLIP, 2), (BGR, 2), (YUYV, 2), (UYVY, 6), (YV24, 2), (YV16, 2), (I420, 2)]:
        n = ol.layout_bytes(ol.src_layout(in_csp, w, h))
        host = np.random.default_rng(9).integers(0, 256, n, dtype=np.uint8)
        want = ol.oracle_convert(host, in_csp, out_csp, 2, 0, w, h)
        d_src = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
        d_src[1:1 + n] = torch.from_numpy(host).cuda()
        d_dst = torch.zeros(want.size + 64, dtype=torch.uint8, device="cuda")
        src, _ = csp.img_fill(d_src.data_ptr() + 1, in_csp, w, h)
        dst, _ = csp.picture_layout(d_dst.data_ptr() + 3, out_csp, w, h)
        torch.cuda.synchronize()
        rc = lib.x264vfw_cuda_csp_convert_batch(ctx.handle, out_csp, 2, 0, 0, C.byref(dst), C.byref(src), w, h, 0, 0, 1)
        assert rc == 0
        ctx.sync()
        assert np.array_equal(d_dst.cpu().numpy()[3:3 + want.size], want), (in_csp, out_csp)


def test_full_size_properties(ctx):
    """Size-independent checks at BASELINE sizes: flipping the input rows and toggling VFLIP
    give identical planes; converting twice is deterministic."""
    from x264vfw_b200 import csp
    w, h = 1920, 1080
    rng = np.random.default_rng(11)
    src = rng.integers(0, 256, 4 * w * h, dtype=np.uint8)
    a = csp.convert_ctx(ctx, src, BGRA | FLIP, 2, 2, 0, w, h)
    flipped = np.ascontiguousarray(src.reshape(h, 4 * w)[::-1]).reshape(-1)
    b = csp.convert_ctx(ctx, flipped, BGRA, 2, 2, 0, w, h)
    assert np.array_equal(a, b)
    assert np.array_equal(a, csp.convert_ctx(ctx, src, BGRA | FLIP, 2, 2, 0, w, h))
