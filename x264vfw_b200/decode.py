"""Host-side mirror of the reference's DECOMPRESS output stage (codec.c:1982-2310), SURVEY 8(f) row 4.

The reference decodes with libavcodec and converts each decoded picture to the application's DIB with libswscale:
  x264vfw_decompress_query    codec.c:1930-1980   which output headers are accepted
  x264vfw_decompress_begin    codec.c:1982-2060   output csp -> pix_fmt, vflip, U/V swap
  x264vfw_init_sws_context    codec.c:2075-2152   the conversion context (lazily, codec.c:2282-2290)
  sws_scale                   codec.c:2292        one picture
  x264vfw_decompress_end      codec.c:2298-2309   sws_freeContext
Here the context is an x264vfw_cuda_dec handle and every picture is converted by libx264vfw_cuda.so (no CPU path).
The H.264 decode itself (libavcodec) is outside the hot path and stays where it is.
"""
import ctypes as C

import numpy as np

from ._lib import lib, Context, CudaError, last_error
from .csp import (X264VFW_CSP_MASK, X264VFW_CSP_VFLIP, X264VFW_CSP_I420, X264VFW_CSP_YV12, X264VFW_CSP_NV12,
                  X264VFW_CSP_YUYV, X264VFW_CSP_UYVY, X264VFW_CSP_BGR, X264VFW_CSP_BGRA, get_csp)

# AVCOL_SPC_* values the reference switches on (codec.c:2114-2140)
AVCOL_SPC_BT709, AVCOL_SPC_UNSPECIFIED, AVCOL_SPC_FCC, AVCOL_SPC_BT470BG = 1, 2, 4, 5
AVCOL_SPC_SMPTE170M, AVCOL_SPC_SMPTE240M, AVCOL_SPC_BT2020_NCL, AVCOL_SPC_BT2020_CL = 6, 7, 9, 10

ICERR_OK, ICERR_BADFORMAT, ICERR_ERROR = 0, -2, -100

_P = C.POINTER
lib.x264vfw_cuda_dec_open.restype = C.c_int
lib.x264vfw_cuda_dec_open.argtypes = [_P(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
lib.x264vfw_cuda_dec_close.restype = None
lib.x264vfw_cuda_dec_close.argtypes = [C.c_void_p]
lib.x264vfw_cuda_dec_picture_size.restype = C.c_int64
lib.x264vfw_cuda_dec_picture_size.argtypes = [C.c_int, C.c_int, C.c_int]
lib.x264vfw_cuda_dec_convert.restype = C.c_int
lib.x264vfw_cuda_dec_convert.argtypes = [C.c_void_p, C.c_void_p, _P(C.c_void_p), _P(C.c_int)]
lib.x264vfw_cuda_dec_convert_batch.restype = C.c_int
lib.x264vfw_cuda_dec_convert_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, _P(C.c_void_p), _P(C.c_int),
                                               C.c_size_t, C.c_int]


def picture_get_size(i_csp: int, width: int, height: int) -> int:
    """x264vfw_picture_get_size (codec.c:505-508) for the covered output formats; -1 otherwise."""
    return int(lib.x264vfw_cuda_dec_picture_size(i_csp, width, height))


def decompress_query(width: int, height: int, out_compression: int, out_bit_count: int, out_width: int,
                     out_height: int, out_size_image: int = 0) -> int:
    """codec.c:1930-1980 for an input header the codec accepts: positive even size, same size out (|biHeight|), a
    known output csp, biSizeImage either 0 or large enough."""
    if width <= 0 or height <= 0 or width % 2 or height % 2:
        return ICERR_BADFORMAT
    if width != out_width or height != abs(out_height):
        return ICERR_BADFORMAT
    i_csp = get_csp(out_compression, out_bit_count, out_height)
    size = picture_get_size(i_csp & X264VFW_CSP_MASK, width, height) if i_csp else -1
    if size < 0 or (out_size_image != 0 and out_size_image < size):
        return ICERR_BADFORMAT
    return ICERR_OK


class Decompressor:
    """decompress_begin .. decompress_end for one output format.  `i_csp` is get_csp() of the OUTPUT header
    (codec.c:1994): its VFLIP bit makes RGB bottom-up, YV12 swaps U and V (codec.c:1995-1998)."""

    def __init__(self, i_csp: int, width: int, height: int, colorspace: int = AVCOL_SPC_UNSPECIFIED,
                 fullrange: bool = False, ctx: Context = None, src_chroma: int = 1):
        """src_chroma: 1 = the decoder delivers yuv420p, 2 = yuv422p (High 4:2:2 streams), 3 = yuv444p (High 4:4:4 Predictive)."""
        self.ctx = ctx or Context()
        self.i_csp, self.width, self.height, self.src_chroma = i_csp, width, height, src_chroma
        self.picture_size = picture_get_size(i_csp, width, height)
        h = C.c_void_p()
        if lib.x264vfw_cuda_dec_open(C.byref(h), self.ctx.handle, i_csp, width, height, src_chroma, colorspace, int(bool(fullrange))) < 0:
            raise CudaError(last_error())
        self.handle = h

    def decompress(self, y: np.ndarray, u: np.ndarray, v: np.ndarray, out: np.ndarray = None) -> np.ndarray:
        """The sws_scale call of codec.c:2292 on one decoded yuv420p picture held in HOST memory (2-D uint8 arrays,
        any row stride, like AVFrame data[]/linesize[]).  Returns the output DIB bytes."""
        ch = self.height if self.src_chroma >= 2 else self.height // 2
        cw = self.width if self.src_chroma == 3 else self.width // 2
        if y.shape != (self.height, self.width) or u.shape != (ch, cw) or v.shape != u.shape:
            raise ValueError("plane shapes do not match the context")
        for p in (y, u, v):
            if p.dtype != np.uint8 or p.strides[1] != 1:
                raise ValueError("planes must be uint8 with contiguous rows")
        if out is None:
            out = np.zeros(self.picture_size, np.uint8)
        src = (C.c_void_p * 3)(y.ctypes.data, u.ctypes.data, v.ctypes.data)
        ss = (C.c_int * 3)(y.strides[0], u.strides[0], v.strides[0])
        if lib.x264vfw_cuda_dec_convert(self.handle, out.ctypes.data, src, ss) < 0:
            raise CudaError(last_error())
        return out

    def decompress_batch(self, dst_dev: int, dst_frame_bytes: int, src_dev, src_stride, src_frame_bytes: int, n_frames: int):
        """n_frames pictures resident in DEVICE memory, one launch on the context's stream (asynchronous)."""
        src = (C.c_void_p * 3)(*src_dev)
        ss = (C.c_int * 3)(*src_stride)
        if lib.x264vfw_cuda_dec_convert_batch(self.handle, dst_dev, dst_frame_bytes, src, ss, src_frame_bytes, n_frames) < 0:
            raise CudaError(last_error())

    def close(self):
        if getattr(self, "handle", None):
            lib.x264vfw_cuda_dec_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
