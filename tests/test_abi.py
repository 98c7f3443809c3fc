"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/x264vfw_cuda.h declares, keeps the reference's geometry and error behaviour,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    syms = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            text = open(os.path.join(ROOT, "include", fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            syms |= set(re.findall(r"\b(x264vfw_cuda_[a-z0-9_]+)\s*\(", text))
    return sorted(syms)


def test_library_exports_every_declared_symbol():
    import x264vfw_b200
    lib = C.CDLL(x264vfw_b200._lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 12
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert "sm_100a" in x264vfw_b200.version()


def test_img_fill_matches_codec_c_geometry():
    from x264vfw_b200 import csp
    # codec.c:365 BGR stride rounds up to 4 bytes; :359 packed 4:2:2 is 2*w; :371 BGRA 4*w
    img, n = csp.img_fill(0, csp.X264VFW_CSP_BGR, 66, 48)
    assert img.i_stride[0] == 200 and n == 200 * 48 and img.i_plane == 1
    img, n = csp.img_fill(0, csp.X264VFW_CSP_YUYV, 1280, 720)
    assert img.i_stride[0] == 2560 and n == 1843200
    img, n = csp.img_fill(0, csp.X264VFW_CSP_BGRA | csp.X264VFW_CSP_VFLIP, 1920, 1080)
    assert img.i_stride[0] == 7680 and n == 8294400
    img, n = csp.img_fill(0, csp.X264VFW_CSP_YV12, 64, 48)
    assert (img.i_plane, img.i_stride[0], img.i_stride[1], n) == (3, 64, 32, 64 * 48 * 3 // 2)
    img, n = csp.img_fill(0, csp.X264VFW_CSP_NV12, 64, 48)
    assert (img.i_plane, img.i_stride[1], n) == (2, 64, 64 * 48 * 3 // 2)
    with pytest.raises(ValueError):
        csp.img_fill(0, 0, 64, 48)


def test_get_csp_and_output_csp_follow_codec_c():
    from x264vfw_b200 import csp
    assert csp.get_csp(csp.BI_RGB, 32, 1080) == csp.X264VFW_CSP_BGRA | csp.X264VFW_CSP_VFLIP   # codec.c:220
    assert csp.get_csp(csp.BI_RGB, 32, -1080) == csp.X264VFW_CSP_BGRA
    assert csp.get_csp(csp.BI_RGB, 24, 1) == csp.X264VFW_CSP_BGR | csp.X264VFW_CSP_VFLIP
    assert csp.get_csp(csp.BI_RGB, 16, 1) == csp.X264VFW_CSP_NONE
    assert csp.get_csp(csp.fourcc("YUY2"), 16, 720) == csp.X264VFW_CSP_YUYV                    # never flipped, :189
    assert csp.get_csp(csp.fourcc("HDYC"), 16, 720) == csp.X264VFW_CSP_UYVY
    assert csp.choose_output_csp(csp.X264VFW_CSP_UYVY, False) == csp.X264_CSP_I420
    assert csp.choose_output_csp(csp.X264VFW_CSP_UYVY, True) == csp.X264_CSP_I422
    assert csp.choose_output_csp(csp.X264VFW_CSP_BGRA | csp.X264VFW_CSP_VFLIP, True) == csp.X264_CSP_BGRA
    assert csp.choose_output_csp(csp.X264VFW_CSP_NV12, False) == csp.X264_CSP_NV12


def test_unregistered_pairs_return_minus_one_without_touching_the_gpu():
    """csp.c:443-444: every slot the reference leaves at convert_fail must return -1."""
    from x264vfw_b200 import csp
    from x264vfw_b200._lib import Image
    registered = {
        csp.X264_CSP_I420: {1, 2, 3, 4, 6, 7, 8, 9}, csp.X264_CSP_NV12: {5}, csp.X264_CSP_I422: {3, 6, 7},
        csp.X264_CSP_I444: {4}, csp.X264_CSP_BGR: {8}, csp.X264_CSP_BGRA: {9},
    }
    a, b = Image(), Image()
    for out, ok in registered.items():
        t = csp.csp_init(out, 2, 0)
        for i in range(csp.X264VFW_CSP_MAX):
            if i not in ok:
                assert t.convert[i](C.byref(a), C.byref(b), 64, 48) == -1, (out, i)
    t = csp.csp_init(0x1234, 2, 0)       # unknown encoder csp: whole table fails
    assert all(t.convert[i](C.byref(a), C.byref(b), 64, 48) == -1 for i in range(csp.X264VFW_CSP_MAX))


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from x264vfw_b200._lib import Context, CudaError
    with pytest.raises(CudaError):
        Context()


def test_product_does_not_reference_the_oracle():
    """The product path (package + csrc + include) must never import/link oracle/."""
    bad = []
    for base in ("x264vfw_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c", "Makefile")):
                    if re.search(r"oracle", open(os.path.join(dp, fn), errors="ignore").read()):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_decompress_query_and_picture_sizes_without_a_device():
    """codec.c:1930-1980 and x264vfw_picture_get_size (codec.c:505-508) through the Python mirror: pure host logic of the library."""
    from x264vfw_b200 import decode
    from x264vfw_b200.csp import fourcc
    assert decode.picture_get_size(8, 70, 38) == 212 * 38                    # BGR24 rows padded to 4 bytes (codec.c:489-492)
    assert decode.picture_get_size(9, 64, 32) == 64 * 32 * 4
    assert decode.picture_get_size(5, 64, 32) == decode.picture_get_size(1, 64, 32) == 64 * 32 * 3 // 2
    assert decode.picture_get_size(6, 64, 32) == decode.picture_get_size(3, 64, 32) == 64 * 32 * 2
    assert decode.picture_get_size(4, 64, 32) == 64 * 32 * 3 and decode.picture_get_size(0, 64, 32) == -1
    ok, bad = decode.ICERR_OK, decode.ICERR_BADFORMAT
    assert decode.decompress_query(64, 32, 0, 32, 64, 32) == ok and decode.decompress_query(64, 32, 0, 24, 64, -32) == ok
    assert decode.decompress_query(64, 32, fourcc("YV12"), 12, 64, 32) == ok and decode.decompress_query(64, 32, fourcc("UYVY"), 16, 64, 32) == ok
    assert decode.decompress_query(64, 32, 0, 32, 32, 32) == bad            # output size must equal the stream's
    assert decode.decompress_query(64, 32, 0, 16, 64, 32) == bad            # no csp for 16-bit RGB
    assert decode.decompress_query(64, 33, 0, 32, 64, 33) == bad            # odd height
    assert decode.decompress_query(64, 32, 0, 32, 64, 32, out_size_image=64 * 32 * 4 - 1) == bad


def test_new_entry_points_reject_null_arguments_without_a_device():
    """Argument checks of the round-2 entry points happen before any CUDA call: -1 and a message, no crash, no device needed."""
    import ctypes as C
    from x264vfw_b200 import lib, last_error
    from x264vfw_b200 import decode, b3  # noqa: F401  (registers the signatures)
    h = C.c_void_p()
    assert lib.x264vfw_cuda_dec_open(C.byref(h), None, 9, 64, 32, 1, 2, 0) == -1 and "null" in last_error()
    assert lib.x264vfw_cuda_dec_convert(None, None, None, None) == -1
    assert lib.x264vfw_cuda_dec_convert_batch(None, None, 0, None, None, 0, 1) == -1
    lib.x264vfw_cuda_dec_close(None)                                        # like sws_freeContext(NULL)
    out = (C.c_int32 * 4 * 3)()
    assert lib.x264vfw_cuda_weights_analyse(None, None, C.byref(out), None) == -1
    assert lib.x264vfw_cuda_la_weights_analyse(None, 1, 0, None, None, 0, C.byref(out), None) == -1
    assert lib.x264vfw_cuda_integral_init(None, None, None, None, 64, 64, 0, 0, 1) == -1


def test_decoder_filter_tables_of_the_product_equal_the_checkers():
    """Host logic of the decoder-side conversion without a device: the tap tables the kernels are handed (the product's own initFilter
    restatement) against the checker's, for 2x up, 2:1 down and identity, both coefficient scales; and the packed writers' row table
    with libswscale's coefficient-pair borrow and its C-writer rows."""
    import ctypes as C
    import numpy as np
    import oracle_lib as ol
    from x264vfw_b200 import lib
    o = ol.oracle()
    lib.x264vfw_cuda_dec_filter_taps.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p]
    lib.x264vfw_cuda_dec_packed_rows.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    o.orc_sws_filter.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_void_p]
    checked = 0
    for one, align in ((1 << 14, 4), (1 << 12, 2)):
        for n in (6, 7, 12, 13, 19, 35, 360, 540, 640, 960, 1080):
            for src, dst in ((n, 2 * n), (2 * n, n)):
                pos_p, co_p = np.zeros(dst, np.int32), np.zeros((dst, 8), np.int16)
                pos_o, co_o = np.zeros(dst, np.int32), np.zeros((dst, 8), np.int16)
                rp = lib.x264vfw_cuda_dec_filter_taps(src, dst, one, align, pos_p.ctypes.data, co_p.ctypes.data)
                if src > dst and src < 12:
                    assert rp == -1                                   # too small to halve with the full 8 taps
                    continue
                assert rp == 0 and o.orc_sws_filter(src, dst, align, one, 128, 128, co_o.ctypes.data, pos_o.ctypes.data) > 0
                assert (pos_p == pos_o).all() and (co_p == co_o).all(), (src, dst, one)
                assert (co_p.sum(1) == one).all()                       # every output sample is a weighted mean
                checked += 1
    assert checked >= 40
    pos, co = np.zeros(8, np.int32), np.zeros((8, 8), np.int16)
    assert lib.x264vfw_cuda_dec_filter_taps(8, 8, 1 << 14, 4, pos.ctypes.data, co.ctypes.data) == 0
    assert (pos == np.arange(8)).all() and (co[:, 0] == 1 << 14).all() and not co[:, 1:].any()      # same size: one tap of 1.0
    assert lib.x264vfw_cuda_dec_filter_taps(8, 24, 1 << 14, 4, pos.ctypes.data, co.ctypes.data) == -1
    # packed writers, 4:2:0 picture with 540 chroma lines
    n = 540
    pos, co, cw = np.zeros(2 * n, np.int32), np.zeros((2 * n, 4), np.int16), np.zeros(2 * n, np.int32)
    assert lib.x264vfw_cuda_dec_packed_rows(n, 0, pos.ctypes.data, co.ctypes.data, cw.ctypes.data) == 0
    pos_o, co_o = np.zeros(2 * n, np.int32), np.zeros((2 * n, 8), np.int16)
    assert o.orc_sws_filter(n, 2 * n, 2, 1 << 12, 128, 128, co_o.ctypes.data, pos_o.ctypes.data) == 4
    want = co_o[:, :4].astype(np.int32)
    simd = np.arange(2 * n) < 2 * n - 2                                 # libswscale leaves its SIMD writers for the last two lines
    for k in (0, 2):
        want[:, k + 1] -= (simd & (want[:, k] < 0))                     # f[k] + f[k+1] * 65536 as ONE int: the borrow
    assert (pos == pos_o).all() and (co == want).all() and (cw == ~simd).all()
    assert (pos == np.clip(((np.arange(2 * n) + 1) >> 1) - 2, 0, n - 4)).all()      # what dec_packed_kernel derives instead of loading
    assert lib.x264vfw_cuda_dec_packed_rows(n, 1, pos.ctypes.data, co.ctypes.data, cw.ctypes.data) == 0
    assert cw.all() and (co == co_o[:, :4]).all()                          # UYVY: the C writer everywhere, plain coefficients
