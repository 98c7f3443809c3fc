"""ctypes access to the CPU oracle (oracle/liboracle.so) and, when it was built in the
authoring container, the unmodified reference csp.c (oracle/_ref/libref_csp.so).
Test infrastructure only -- the product never imports this module."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_CSP_SO = os.path.join(ROOT, "oracle", "_ref", "libref_csp.so")


class OrcImage(C.Structure):
    _fields_ = [("i_csp", C.c_int), ("i_plane", C.c_int), ("i_stride", C.c_int * 4),
                ("plane", C.c_void_p * 4)]


_orc = None


def oracle():
    global _orc
    if _orc is None:
        o = C.CDLL(ORACLE_SO)
        P = C.POINTER
        o.orc_csp_convert.restype = C.c_int
        o.orc_csp_convert.argtypes = [C.c_int, C.c_int, C.c_int, P(OrcImage), P(OrcImage), C.c_int, C.c_int]
        o.orc_ext_rgb_to_nv12.restype = C.c_int
        o.orc_ext_rgb_to_nv12.argtypes = [C.c_int, C.c_int, P(OrcImage), P(OrcImage), C.c_int, C.c_int]
        o.orc_ext_422_to_i444.restype = C.c_int
        o.orc_ext_422_to_i444.argtypes = [P(OrcImage), P(OrcImage), C.c_int, C.c_int]
        o.orc_rgb_coefficients.argtypes = [C.c_int, C.c_int, P(C.c_uint32)]
        o.orc_fnv1a64.restype = C.c_uint64
        o.orc_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        o.orc_lcg_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        _orc = o
    return _orc


def have_ref_csp():
    return os.path.exists(REF_CSP_SO)


_REF_FN = C.CFUNCTYPE(C.c_int, C.POINTER(OrcImage), C.POINTER(OrcImage), C.c_int, C.c_int)


class _RefTable(C.Structure):
    _fields_ = [("convert", _REF_FN * 10)]


_ref = None


def ref_csp_table(out_csp, colmatrix, fullrange):
    """x264vfw_csp_init of the UNMODIFIED reference object (x264_image_t has the OrcImage layout)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_CSP_SO)
        _ref.x264vfw_csp_init.argtypes = [C.POINTER(_RefTable), C.c_int, C.c_int, C.c_int]
    t = _RefTable()
    _ref.x264vfw_csp_init(C.byref(t), out_csp, colmatrix, fullrange)
    return t


# ---- geometry (mirrors codec.c:304-379 and x264_picture_alloc; pure python so the oracle
# tests do not depend on the CUDA library) -------------------------------------------------
def src_layout(csp, w, h):
    c = csp & 0xff
    if c in (1, 2):
        hh, ww = (h + 1) & ~1, (w + 1) & ~1
        return [(ww, hh), (ww // 2, hh // 2), (ww // 2, hh // 2)]
    if c == 3:
        ww = (w + 1) & ~1
        return [(ww, h), (ww // 2, h), (ww // 2, h)]
    if c == 4:
        return [(w, h)] * 3
    if c == 5:
        hh, ww = (h + 1) & ~1, (w + 1) & ~1
        return [(ww, hh), (ww, hh // 2)]
    if c in (6, 7):
        return [(2 * ((w + 1) & ~1), h)]
    if c == 8:
        return [((3 * w + 3) & ~3, h)]
    if c == 9:
        return [(4 * w, h)]
    raise ValueError(csp)


def dst_layout(out_csp, w, h):
    return {2: [(w, h), (w // 2, h // 2), (w // 2, h // 2)], 4: [(w, h), (w, h // 2)],
            6: [(w, h), (w // 2, h), (w // 2, h)], 0xc: [(w, h)] * 3,
            0xe: [(3 * w, h)], 0xf: [(4 * w, h)]}[out_csp]


def make_image(buf: np.ndarray, layout, csp=0):
    img = OrcImage()
    img.i_csp = csp
    img.i_plane = len(layout)
    off = 0
    for i, (stride, rows) in enumerate(layout):
        img.i_stride[i] = stride
        img.plane[i] = buf.ctypes.data + off
        off += stride * rows
    return img, off


def layout_bytes(layout):
    return sum(s * r for s, r in layout)


def lcg_bytes(n, w, h):
    buf = np.empty(n, dtype=np.uint8)
    oracle().orc_lcg_fill(buf.ctypes.data, n, w, h)
    return buf


def fnv(buf: np.ndarray) -> str:
    buf = np.ascontiguousarray(buf)
    return "%016x" % oracle().orc_fnv1a64(buf.ctypes.data, buf.size)


def oracle_convert(src: np.ndarray, in_csp, out_csp, colmat, full, w, h, ext=0):
    """Returns the tight destination buffer or None when the pair is unsupported (-1)."""
    sl, dl = src_layout(in_csp, w, h), dst_layout(out_csp, w, h)
    simg, _ = make_image(src, sl, in_csp)
    out = np.zeros(layout_bytes(dl), dtype=np.uint8)
    dimg, _ = make_image(out, dl)
    o = oracle()
    if ext == 1:
        rc = o.orc_ext_rgb_to_nv12(colmat, full, C.byref(dimg), C.byref(simg), w, h)
    elif ext == 2:
        rc = o.orc_ext_422_to_i444(C.byref(dimg), C.byref(simg), w, h)
    else:
        rc = o.orc_csp_convert(out_csp, colmat, full, C.byref(dimg), C.byref(simg), w, h)
    return out if rc == 0 else None


def ref_convert(src: np.ndarray, in_csp, out_csp, colmat, full, w, h):
    t = ref_csp_table(out_csp, colmat, full)
    sl, dl = src_layout(in_csp, w, h), dst_layout(out_csp, w, h)
    simg, _ = make_image(src, sl, in_csp)
    out = np.zeros(layout_bytes(dl), dtype=np.uint8)
    dimg, _ = make_image(out, dl)
    rc = t.convert[in_csp & 0xff](C.byref(dimg), C.byref(simg), w, h)
    return out if rc == 0 else None


# ---- stage 2a: lowres --------------------------------------------------------------------
def lowres_geometry(w, h):
    g = (C.c_int * 10)()
    o = oracle()
    o.orc_lowres_geometry.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    o.orc_lowres_geometry(w, h, g)
    names = ("mb_w", "mb_h", "luma_w", "luma_h", "luma_stride", "lw", "lh", "lstride", "lplane_bytes", "lorigin")
    return dict(zip(names, list(g)))


def oracle_lowres_init(y: np.ndarray, w, h):
    """y: tight (h, w) uint8.  Returns the 4 padded planes as one flat buffer."""
    g = lowres_geometry(w, h)
    y = np.ascontiguousarray(y, dtype=np.uint8)
    out = np.zeros(4 * g["lplane_bytes"], dtype=np.uint8)
    o = oracle()
    o.orc_lowres_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    o.orc_lowres_init(out.ctypes.data, y.ctypes.data, w, w, h)
    return out


def oracle_luma_pad(y: np.ndarray, w, h):
    g = lowres_geometry(w, h)
    y = np.ascontiguousarray(y, dtype=np.uint8)
    out = np.zeros(g["luma_stride"] * (g["luma_h"] + 1), dtype=np.uint8)
    o = oracle()
    o.orc_luma_pad.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    o.orc_luma_pad(out.ctypes.data, g["luma_stride"], y.ctypes.data, w, w, h)
    return out
