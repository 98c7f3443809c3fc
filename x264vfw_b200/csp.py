"""Host-side mirror of the reference's colour-space interface (csp.h / csp.c / codec.c).

Names and argument meaning follow the reference so parity tests read like tests of csp.c:
  X264VFW_CSP_*            csp.h:30-44
  get_csp                  codec.c:187-231   (BITMAPINFOHEADER -> input csp | VFLIP)
  choose_output_csp        codec.c:269-302
  img_fill                 codec.c:304-379   (borrowed input buffer geometry)
  csp_init / convert[]     csp.c:440-514, call site codec.c:1774
All compute goes through libx264vfw_cuda.so.
"""
import ctypes as C
import numpy as np

from ._lib import lib, Image, CspFunctionTable, Context, CudaError, last_error

X264VFW_CSP_MASK = 0x00ff
X264VFW_CSP_NONE = 0
X264VFW_CSP_I420 = 1
X264VFW_CSP_YV12 = 2
X264VFW_CSP_YV16 = 3
X264VFW_CSP_YV24 = 4
X264VFW_CSP_NV12 = 5
X264VFW_CSP_YUYV = 6
X264VFW_CSP_UYVY = 7
X264VFW_CSP_BGR = 8
X264VFW_CSP_BGRA = 9
X264VFW_CSP_MAX = 10
X264VFW_CSP_VFLIP = 0x1000

X264_CSP_I420 = 0x0002
X264_CSP_NV12 = 0x0004
X264_CSP_I422 = 0x0006
X264_CSP_I444 = 0x000c
X264_CSP_BGR = 0x000e
X264_CSP_BGRA = 0x000f

EXT_NONE, EXT_RGB_TO_NV12, EXT_422_TO_I444 = 0, 1, 2

BI_RGB = 0


def fourcc(s: str) -> int:
    return ord(s[0]) | (ord(s[1]) << 8) | (ord(s[2]) << 16) | (ord(s[3]) << 24)


_FOURCC_CSP = {
    fourcc("I420"): X264VFW_CSP_I420, fourcc("IYUV"): X264VFW_CSP_I420,
    fourcc("YV12"): X264VFW_CSP_YV12, fourcc("YV16"): X264VFW_CSP_YV16,
    fourcc("YV24"): X264VFW_CSP_YV24, fourcc("NV12"): X264VFW_CSP_NV12,
    fourcc("YUYV"): X264VFW_CSP_YUYV, fourcc("YUY2"): X264VFW_CSP_YUYV,
    fourcc("UYVY"): X264VFW_CSP_UYVY, fourcc("HDYC"): X264VFW_CSP_UYVY,
}


def get_csp(bi_compression: int, bi_bit_count: int, bi_height: int) -> int:
    """codec.c:187-231: only BI_RGB with non-negative biHeight is bottom-up."""
    if bi_compression in _FOURCC_CSP:
        return _FOURCC_CSP[bi_compression]
    if bi_compression == BI_RGB:
        flip = 0 if bi_height < 0 else X264VFW_CSP_VFLIP
        if bi_bit_count == 24:
            return X264VFW_CSP_BGR | flip
        if bi_bit_count == 32:
            return X264VFW_CSP_BGRA | flip
    return X264VFW_CSP_NONE


def choose_output_csp(i_csp: int, b_keep_input_csp: bool) -> int:
    """codec.c:269-302."""
    c = i_csp & X264VFW_CSP_MASK
    if c in (X264VFW_CSP_I420, X264VFW_CSP_YV12):
        return X264_CSP_I420
    if c == X264VFW_CSP_YV16:
        return X264_CSP_I422 if b_keep_input_csp else X264_CSP_I420
    if c == X264VFW_CSP_YV24:
        return X264_CSP_I444 if b_keep_input_csp else X264_CSP_I420
    if c == X264VFW_CSP_NV12:
        return X264_CSP_NV12
    if c in (X264VFW_CSP_YUYV, X264VFW_CSP_UYVY):
        return X264_CSP_I422 if b_keep_input_csp else X264_CSP_I420
    if c == X264VFW_CSP_BGR:
        return X264_CSP_BGR if b_keep_input_csp else X264_CSP_I420
    if c == X264VFW_CSP_BGRA:
        return X264_CSP_BGRA if b_keep_input_csp else X264_CSP_I420
    return X264_CSP_I420


def img_fill(ptr: int, i_csp: int, width: int, height: int):
    """codec.c:304-379.  Returns (Image, total_bytes); raises on unknown csp (-1 there)."""
    img = Image()
    n = lib.x264vfw_cuda_img_fill(C.byref(img), C.c_void_p(ptr), i_csp, width, height)
    if n < 0:
        raise ValueError(f"unsupported input csp {i_csp:#x}")
    return img, int(n)


def picture_layout(ptr: int, i_x264_csp: int, width: int, height: int):
    """[x264] x264_picture_alloc layout of conv_pic (codec.c:1673): tight planes."""
    img = Image()
    n = lib.x264vfw_cuda_picture_layout(C.byref(img), C.c_void_p(ptr), i_x264_csp, width, height)
    if n < 0:
        raise ValueError(f"unsupported encoder csp {i_x264_csp:#x}")
    return img, int(n)


def csp_init(i_x264_csp: int, i_colmatrix: int, b_fullrange: int) -> CspFunctionTable:
    """x264vfw_csp_init (csp.c:440): returns the filled function table."""
    t = CspFunctionTable()
    lib.x264vfw_cuda_csp_init(C.byref(t), i_x264_csp, i_colmatrix, b_fullrange)
    return t


def convert_host(table: CspFunctionTable, src_bytes: np.ndarray, i_csp: int, i_x264_csp: int,
                 width: int, height: int) -> np.ndarray:
    """codec.c:1762-1779 for one frame: img_fill + csp.convert[i_csp & MASK](conv_pic, pic).
    Returns the tight destination buffer, or raises CudaError when the converter returns <0."""
    assert src_bytes.dtype == np.uint8 and src_bytes.flags.c_contiguous
    src, sbytes = img_fill(src_bytes.ctypes.data, i_csp, width, height)
    assert src_bytes.size >= sbytes
    _, dbytes = picture_layout(0, i_x264_csp, width, height)
    out = np.zeros(dbytes, dtype=np.uint8)
    dst, _ = picture_layout(out.ctypes.data, i_x264_csp, width, height)
    rc = table.convert[i_csp & X264VFW_CSP_MASK](C.byref(dst), C.byref(src), width, height)
    if rc < 0:
        raise CudaError(f"convert returned {rc}: {last_error()}")
    return out


def frame_bytes(i_csp: int, i_x264_csp: int, width: int, height: int, align: int = 256):
    """(source, destination) bytes per frame in a device-resident batch, rounded to `align`."""
    _, s = img_fill(0, i_csp, width, height)
    _, d = picture_layout(0, i_x264_csp, width, height)
    r = lambda v: (v + align - 1) // align * align
    return r(s), r(d)


def convert_batch(ctx: Context, d_src: int, d_dst: int, i_csp: int, i_x264_csp: int,
                  i_colmatrix: int, b_fullrange: int, width: int, height: int, n_frames: int,
                  ext: int = EXT_NONE, src_frame_bytes: int = 0, dst_frame_bytes: int = 0):
    """Device-resident batch: n_frames frames, one launch, asynchronous on ctx's stream
    (x264vfw_cuda_csp_convert_batch).  d_src/d_dst are device addresses of frame 0."""
    sfb, dfb = frame_bytes(i_csp, i_x264_csp, width, height)
    sfb = src_frame_bytes or sfb
    dfb = dst_frame_bytes or dfb
    src, _ = img_fill(d_src, i_csp, width, height)
    dst, _ = picture_layout(d_dst, i_x264_csp, width, height)
    rc = lib.x264vfw_cuda_csp_convert_batch(ctx.handle, i_x264_csp, i_colmatrix, b_fullrange, ext,
                                            C.byref(dst), C.byref(src), width, height, sfb, dfb, n_frames)
    if rc < 0:
        raise CudaError(f"convert_batch returned {rc}: {last_error()}")
    return sfb, dfb


def convert_ctx(ctx: Context, src_bytes: np.ndarray, i_csp: int, i_x264_csp: int, i_colmatrix: int,
                b_fullrange: int, width: int, height: int, ext: int = EXT_NONE, out: np.ndarray = None):
    """One frame, host buffers, explicit context (x264vfw_cuda_csp_convert)."""
    src, sbytes = img_fill(src_bytes.ctypes.data, i_csp, width, height)
    assert src_bytes.size >= sbytes
    _, dbytes = picture_layout(0, i_x264_csp, width, height)
    if out is None:
        out = np.zeros(dbytes, dtype=np.uint8)
    dst, _ = picture_layout(out.ctypes.data, i_x264_csp, width, height)
    rc = lib.x264vfw_cuda_csp_convert(ctx.handle, i_x264_csp, i_colmatrix, b_fullrange, ext,
                                      C.byref(dst), C.byref(src), width, height)
    if rc < 0:
        raise CudaError(f"convert returned {rc}: {last_error()}")
    return out


__all__ = [n for n in dir() if not n.startswith("_")]
