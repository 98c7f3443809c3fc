"""ctypes access to libswscale 9.1.100 (FFmpeg 8.0 line) as bundled in this image's opencv wheel, driven exactly the
way the reference drives it on its decompress path (codec.c:2075-2152 context, codec.c:2292 call).  Test
infrastructure: tests/golden/make_decode_golden.py makes the committed fixtures with it, and
tests/test_decode_oracle.py re-checks the oracle against it live wherever the wheel is importable."""
import ctypes as C
import glob
import os

import numpy as np

SWS_BICUBIC, SWS_FULL_CHR_H_INT, SWS_FULL_CHR_H_INP, SWS_ACCURATE_RND = 4, 0x2000, 0x4000, 0x40000
# AVPixelFormat values of libavutil 60
PIX = {"yuv420p": 0, "yuyv": 1, "bgr": 3, "yuv422p": 4, "yuv444p": 5, "uyvy": 15, "nv12": 23, "bgra": 28}
# output csp codes of the reference (csp.h:30-44) -> (AVPixelFormat, swap U/V of the output picture)
CSP_I420, CSP_YV12, CSP_YV16, CSP_YV24, CSP_NV12, CSP_YUYV, CSP_UYVY, CSP_BGR, CSP_BGRA, CSP_VFLIP = 1, 2, 3, 4, 5, 6, 7, 8, 9, 0x1000
CSP_TO_PIX = {CSP_I420: "yuv420p", CSP_YV12: "yuv420p", CSP_YV16: "yuv422p", CSP_YV24: "yuv444p", CSP_NV12: "nv12", CSP_YUYV: "yuyv", CSP_UYVY: "uyvy",
              CSP_BGR: "bgr", CSP_BGRA: "bgra"}
# codec.c:2114-2140: AVCOL_SPC_* -> SWS_CS_*
SPC_TO_CS = {1: 1, 4: 4, 5: 5, 6: 6, 7: 7, 9: 9, 10: 9}

_libs = None


def libs():
    global _libs
    if _libs is None:
        import cv2                                      # resolves the wheel's private library directory
        d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
        avu = C.CDLL(glob.glob(os.path.join(d, "libavutil-*.so*"))[0])
        sws = C.CDLL(glob.glob(os.path.join(d, "libswscale-*.so*"))[0])
        P = C.c_void_p
        sws.swscale_version.restype = C.c_uint
        sws.sws_alloc_context.restype = P
        sws.sws_init_context.argtypes = [P, P, P]
        sws.sws_freeContext.argtypes = [P]
        sws.sws_getCoefficients.restype = C.POINTER(C.c_int)
        sws.sws_getCoefficients.argtypes = [C.c_int]
        sws.sws_setColorspaceDetails.argtypes = [P, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.c_int,
                                                 C.c_int, C.c_int, C.c_int]
        sws.sws_scale.argtypes = [P, C.POINTER(P), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(P), C.POINTER(C.c_int)]
        avu.av_opt_set_int.argtypes = [P, C.c_char_p, C.c_int64, C.c_int]
        _libs = (sws, avu)
    return _libs


def available():
    try:
        sws, _ = libs()
        return sws.swscale_version() >> 16 == 9
    except Exception:
        return False


def version():
    v = libs()[0].swscale_version()
    return "%d.%d.%d" % (v >> 16, (v >> 8) & 255, v & 255)


def picture_size(csp, w, h):
    """x264vfw_picture_get_size (codec.c:505-508)."""
    csp &= 0xff
    if csp in (CSP_I420, CSP_YV12, CSP_NV12):
        return w * h + 2 * (w // 2) * (h // 2)
    if csp in (CSP_YV16, CSP_YUYV, CSP_UYVY):
        return w * 2 * h
    if csp == CSP_YV24:
        return w * 3 * h
    return ((w * 3 + 3) & ~3) * h if csp == CSP_BGR else w * 4 * h


def decompress_convert(y, u, v, out_csp, avcol_spc=2, fullrange=0, pad_tail=64, repeat=1, timing=None, src_chroma=1):
    """What x264vfw_decompress does with one decoded yuv420p (src_chroma 1) / yuv422p (2) / yuv444p (3) picture (y, u, v: 2-D uint8, any row stride):
    returns the output DIB bytes (picture_size long).  The buffer handed to libswscale carries pad_tail spare bytes
    because its SIMD writers store whole groups of 8 pixels (see oracle/decode_oracle.c header)."""
    sws, avu = libs()
    h, w = y.shape[0], y.shape[1]
    fmt, flip = out_csp & 0xff, bool(out_csp & CSP_VFLIP)
    ctx = sws.sws_alloc_context()
    flags = SWS_BICUBIC | SWS_FULL_CHR_H_INP | SWS_ACCURATE_RND                      # codec.c:2081-2082
    for k, val in ((b"sws_flags", flags), (b"srcw", w), (b"srch", h), (b"src_format", PIX[{1: "yuv420p", 2: "yuv422p", 3: "yuv444p"}[src_chroma]]),
                   (b"src_range", fullrange), (b"dstw", w), (b"dsth", h), (b"dst_format", PIX[CSP_TO_PIX[fmt]]),
                   (b"dst_range", fullrange)):                                     # codec.c:2097-2107
        avu.av_opt_set_int(ctx, k, val, 0)
    # codec.c:2110-2111 ORs SWS_FULL_CHR_H_INT into the local only: the context never receives it
    co = sws.sws_getCoefficients(SPC_TO_CS.get(avcol_spc, 5))
    sws.sws_setColorspaceDetails(ctx, co, fullrange, co, fullrange, 0, 1 << 16, 1 << 16)   # codec.c:2141-2144
    if sws.sws_init_context(ctx, None, None) < 0:
        sws.sws_freeContext(ctx)
        raise RuntimeError("sws_init_context failed")
    size = picture_size(out_csp, w, h)
    out = np.zeros(size + pad_tail, np.uint8)
    base = out.ctypes.data
    cw, ch = w // 2, h // 2
    if fmt in (CSP_I420, CSP_YV12):                                                  # x264vfw_picture_fill, codec.c:425-439
        data, ls = [base, base + w * h, base + w * h + cw * ch], [w, cw, cw]
        if fmt == CSP_YV12:                                                          # codec.c:2263-2274
            data[1], data[2] = data[2], data[1]
    elif fmt == CSP_YV16:                                                            # codec.c:441-454, swapped like YV12
        data, ls = [base, base + w * h + cw * h, base + w * h], [w, cw, cw]
    elif fmt == CSP_YV24:                                                            # codec.c:456-467, swapped like YV12
        data, ls = [base, base + 2 * w * h, base + w * h], [w, w, w]
    elif fmt == CSP_NV12:
        data, ls = [base, base + w * h], [w, w]
    elif fmt in (CSP_YUYV, CSP_UYVY):
        data, ls = [base], [w * 2]
    else:
        stride = (w * 3 + 3) & ~3 if fmt == CSP_BGR else w * 4
        data, ls = [base], [stride]
        if flip:                                                                     # codec.c:510-520
            data, ls = [base + stride * (h - 1)], [-stride]
    data += [None] * (4 - len(data))
    ls += [0] * (4 - len(ls))
    src = (C.c_void_p * 4)(y.ctypes.data, u.ctypes.data, v.ctypes.data, None)
    ss = (C.c_int * 4)(y.strides[0], u.strides[0], v.strides[0], 0)
    dp, dl = (C.c_void_p * 4)(*data), (C.c_int * 4)(*ls)
    import time
    t0 = time.perf_counter()
    for _ in range(repeat):                     # repeat > 1: bench.py times the call alone, context built once like codec.c:2282
        r = sws.sws_scale(ctx, src, ss, 0, h, dp, dl)                                    # codec.c:2292
    if timing is not None:
        timing.append((time.perf_counter() - t0) / repeat)
    sws.sws_freeContext(ctx)
    if r != h:
        raise RuntimeError("sws_scale returned %d" % r)
    return out[:size].copy()
