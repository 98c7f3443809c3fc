"""ctypes access to the CPU oracle (oracle/liboracle.so) and, when it was built in the
authoring container, the unmodified reference csp.c (oracle/_ref/libref_csp.so).
Test infrastructure only -- the product never imports this module."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# bench.py's CPU arms point these at the -march=native builds made on the box (oracle/Makefile: native)
ORACLE_SO = os.environ.get("X264VFW_ORACLE_SO") or os.path.join(ROOT, "oracle", "liboracle.so")
REF_CSP_SO = os.environ.get("X264VFW_REF_CSP_SO") or os.path.join(ROOT, "oracle", "_ref", "libref_csp.so")


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
        # test hooks of the lookahead checker (pointers must never travel as default C ints)
        V, I = C.c_void_p, C.c_int
        o.orc_test_get_ref_8x8.argtypes = [V, V, V, V, V, I, I, I]
        o.orc_test_get_ref_8x8.restype = None
        o.orc_test_get_ref_8x8_weighted.argtypes = [V, V, V, V, V, I, I, I, I, I, I]
        o.orc_test_get_ref_8x8_weighted.restype = None
        o.orc_test_intra_pred_8x8.argtypes = [V, I, V, I]
        o.orc_test_intra_pred_8x8.restype = None
        o.orc_test_pixel_avg_8x8.argtypes = [V, V, V, I]
        o.orc_test_pixel_avg_8x8.restype = None
        o.orc_test_bipred_weight.argtypes = [I, I, I, I]
        o.orc_test_bipred_weight.restype = I
        o.orc_test_sad_8x8.argtypes = [V, V]
        o.orc_test_sad_8x8.restype = I
        o.orc_test_satd_8x8.argtypes = [V, V]
        o.orc_test_satd_8x8.restype = I
        o.orc_test_predict_picture.argtypes = [V, V, I, C.c_size_t] + [I] * 7
        o.orc_test_predict_picture.restype = None
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


# ---- next row f3: half-pel reference planes ---------------------------------------------------
def hpel_geometry(w, h):
    g = (C.c_int * 3)()
    o = oracle()
    o.orc_hpel_geometry.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    o.orc_hpel_geometry(w, h, g)
    return dict(zip(("stride", "plane_bytes", "origin"), list(g)))


def oracle_hpel_planes(y: np.ndarray, w, h):
    """y: tight (h, w) uint8 reconstructed plane.  Returns (4, h+64, stride): P0, H, V, C, all padded."""
    g = hpel_geometry(w, h)
    y = np.ascontiguousarray(y, dtype=np.uint8)
    out = np.zeros(4 * g["plane_bytes"], dtype=np.uint8)
    o = oracle()
    o.orc_hpel_planes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    o.orc_hpel_planes(out.ctypes.data, y.ctypes.data, w, w, h)
    return out.reshape(4, h + 64, g["stride"])


def numpy_hpel_planes(y: np.ndarray):
    """Independent formulation of the same planes (no borders materialised, no loops over pixels):
    Pi(x,y) = Fi(clamp(x,-4,w+3), clamp(y,-8,h+7)) with Fi the 6-tap filter on the edge-clamped frame."""
    h, w = y.shape
    P = 32
    ext = np.pad(y.astype(np.int64), P + 8, mode="edge")        # frame with an (ample) replicated border
    o = P + 8                                                    # index of pixel (0,0) in ext

    def tap(a, axis):
        r = lambda k: np.roll(a, -k, axis=axis)
        return r(-2) + r(3) - 5 * (r(-1) + r(2)) + 20 * (r(0) + r(1))

    v16 = tap(ext, 0)
    fh = np.clip((tap(ext, 1) + 16) >> 5, 0, 255)
    fv = np.clip((v16 + 16) >> 5, 0, 255)
    fc = np.clip((tap(v16, 1) + 512) >> 10, 0, 255)
    ys = np.clip(np.arange(-P, h + P), -8, h + 7) + o
    xs = np.clip(np.arange(-P, w + P), -4, w + 3) + o
    out = [ext[o - P:o + h + P, o - P:o + w + P]]
    out += [f[np.ix_(ys, xs)] for f in (fh, fv, fc)]
    return np.stack(out).astype(np.uint8)                        # (4, h+64, w+64)


# ---- stage 2: lookahead oracle -------------------------------------------------------------
def oracle_chroma_nv12_pad(u: np.ndarray, v: np.ndarray, w, h):
    g = lowres_geometry(w, h)
    o = oracle()
    out = np.zeros(g["luma_w"] * (g["luma_h"] // 2), dtype=np.uint8)
    o.orc_chroma_nv12_pad.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    o.orc_chroma_nv12_pad(out.ctypes.data, g["luma_w"], u.ctypes.data, v.ctypes.data, w // 2, w, h)
    return out


class LaParams(C.Structure):
    """orc_la_params / x264vfw_cuda_la_params (same field order)."""
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("chroma_format", C.c_int), ("bframes", C.c_int),
                ("b_adapt", C.c_int), ("b_pyramid", C.c_int), ("b_bias", C.c_int), ("rc_lookahead", C.c_int),
                ("b_mbtree", C.c_int), ("scenecut", C.c_int), ("keyint_max", C.c_int), ("keyint_min", C.c_int),
                ("open_gop", C.c_int), ("weightp", C.c_int), ("weightb", C.c_int), ("subme", C.c_int),
                ("me_method", C.c_int), ("me_range", C.c_int), ("mv_range", C.c_int), ("aq_mode", C.c_int),
                ("aq_strength", C.c_float), ("qcompress", C.c_float), ("frame_reference", C.c_int),
                ("lookahead_threads", C.c_int), ("fps_num", C.c_int), ("fps_den", C.c_int), ("b_psy", C.c_int)]


class LaDecision(C.Structure):
    _fields_ = [("i_frame", C.c_int), ("i_type", C.c_int), ("b_keyframe", C.c_int), ("i_bframes", C.c_int),
                ("i_cost_est", C.c_int), ("i_cost_est_aq", C.c_int), ("i_intra_mbs", C.c_int), ("mb_count", C.c_int)]


TYPE_NAMES = {0: "AUTO", 1: "IDR", 2: "I", 3: "P", 4: "BREF", 5: "B", 6: "KEY"}


def la_params(preset, w, h, **over):
    o = oracle()
    p = LaParams()
    o.orc_la_params_preset.argtypes = [C.POINTER(LaParams), C.c_char_p, C.c_int, C.c_int]
    o.orc_la_params_preset(C.byref(p), preset.encode(), w, h)
    for k, v in over.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


class OracleLookahead:
    def __init__(self, params: LaParams):
        o = self.o = oracle()
        P = C.POINTER
        o.orc_la_open.restype = C.c_void_p
        o.orc_la_open.argtypes = [P(LaParams)]
        o.orc_la_close.argtypes = [C.c_void_p]
        o.orc_la_put_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        o.orc_la_flush.argtypes = [C.c_void_p]
        o.orc_la_get_decision.argtypes = [C.c_void_p, P(LaDecision), C.c_void_p, C.c_void_p]
        o.orc_la_frame_cost.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        o.orc_la_mb_count.argtypes = [C.c_void_p]
        for name, res in (("orc_la_lowres_planes", C.c_void_p), ("orc_la_intra_cost", C.c_void_p),
                          ("orc_la_inv_qscale", C.c_void_p), ("orc_la_propagate_cost", C.c_void_p)):
            getattr(o, name).restype = res
            getattr(o, name).argtypes = [C.c_void_p, C.c_int]
        o.orc_la_qp_offset.restype = C.c_void_p
        o.orc_la_qp_offset.argtypes = [C.c_void_p, C.c_int, C.c_int]
        o.orc_la_mvs.restype = C.c_void_p
        o.orc_la_mvs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        o.orc_la_mv_costs.restype = C.c_void_p
        o.orc_la_mv_costs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        o.orc_la_lowres_costs.restype = C.c_void_p
        o.orc_la_lowres_costs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        o.orc_la_row_satds.restype = C.c_void_p
        o.orc_la_row_satds.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        o.orc_la_cost_est.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        o.orc_la_intra_mbs.argtypes = [C.c_void_p, C.c_int, C.c_int]
        o.orc_la_pixel_stats.argtypes = [C.c_void_p, C.c_int, P(C.c_uint64), P(C.c_uint64)]
        o.orc_la_weight.argtypes = [C.c_void_p, C.c_int, P(C.c_int)]
        o.orc_la_mbtree.argtypes = [C.c_void_p, P(C.c_int), P(C.c_int), C.c_int, C.c_int]
        o.orc_la_counters.argtypes = [C.c_void_p, P(C.c_uint64)]
        self.p = params
        self.h = o.orc_la_open(C.byref(params))
        self.mb_count = o.orc_la_mb_count(self.h)
        self.g = lowres_geometry(params.width, params.height)

    def close(self):
        if self.h:
            self.o.orc_la_close(self.h)
            self.h = None

    def put_i420(self, buf: np.ndarray):
        """buf: tight planar frame in the encoder csp (I420 / I422 / I444 by chroma_format)."""
        w, h = self.p.width, self.p.height
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        cf = self.p.chroma_format
        cw = w if cf == 3 else w // 2
        chh = h // 2 if cf == 1 else h
        base = buf.ctypes.data
        self._keep = buf
        return self.o.orc_la_put_frame(self.h, base, w, base + w * h, base + w * h + cw * chh, cw)

    def put_luma(self, y: np.ndarray):
        y = np.ascontiguousarray(y, dtype=np.uint8)
        return self.o.orc_la_put_frame(self.h, y.ctypes.data, self.p.width, None, None, 0)

    def flush(self):
        return self.o.orc_la_flush(self.h)

    def decisions(self):
        out = []
        d = LaDecision()
        while True:
            q = np.zeros(self.mb_count, dtype=np.float32)
            qa = np.zeros(self.mb_count, dtype=np.float32)
            if not self.o.orc_la_get_decision(self.h, C.byref(d), q.ctypes.data, qa.ctypes.data):
                break
            out.append(dict(i_frame=d.i_frame, i_type=d.i_type, b_keyframe=d.b_keyframe, i_bframes=d.i_bframes,
                            i_cost_est=d.i_cost_est, i_cost_est_aq=d.i_cost_est_aq, i_intra_mbs=d.i_intra_mbs,
                            qp_offset=q, qp_offset_aq=qa))
        return out

    def frame_cost(self, p0, p1, b):
        return self.o.orc_la_frame_cost(self.h, p0, p1, b)

    def weights_full(self, fenc, ref, fenc_uv: np.ndarray, ref_uv: np.ndarray, uv_stride):
        """[x264] x264_weights_analyse(h, fenc, ref, 0): ([[on, scale, denom, offset]] * 3, cost_delta) or None (-1)."""
        out = (C.c_int * 4 * 3)()
        delta = C.c_float(0)
        self.o.orc_la_weights_full.restype = C.c_int
        self.o.orc_la_weights_full.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int * 4 * 3), C.POINTER(C.c_float)]
        if self.o.orc_la_weights_full(self.h, fenc, ref, fenc_uv.ctypes.data, ref_uv.ctypes.data, uv_stride, C.byref(out), C.byref(delta)) < 0:
            return None
        return [[int(out[p][i]) for i in range(4)] for p in range(3)], float(delta.value)

    def weights_full_cost(self, fenc, ref, fenc_uv, ref_uv, uv_stride, plane, weight=None):
        """One score of that analysis (test hook): weight = (scale, denom, offset) or None for the unweighted score."""
        self.o.orc_test_weights_full_cost.restype = C.c_uint
        self.o.orc_test_weights_full_cost.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        w = weight or (0, 0, 0)
        return int(self.o.orc_test_weights_full_cost(self.h, fenc, ref, fenc_uv.ctypes.data, ref_uv.ctypes.data, uv_stride, plane,
                                                     int(weight is not None), w[0], w[1], w[2]))

    def _arr(self, ptr, n, dtype):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype).copy()

    def lowres_planes(self, f):
        return self._arr(self.o.orc_la_lowres_planes(self.h, f), 4 * self.g["lplane_bytes"], np.uint8)

    def intra_cost(self, f):
        return self._arr(self.o.orc_la_intra_cost(self.h, f), self.mb_count, np.uint16)

    def inv_qscale(self, f):
        return self._arr(self.o.orc_la_inv_qscale(self.h, f), self.mb_count, np.uint16)

    def propagate_cost(self, f):
        return self._arr(self.o.orc_la_propagate_cost(self.h, f), self.mb_count, np.uint16)

    def qp_offset(self, f, aq=False):
        return self._arr(self.o.orc_la_qp_offset(self.h, f, int(aq)), self.mb_count, np.float32)

    def mvs(self, f, lst, dist):
        return self._arr(self.o.orc_la_mvs(self.h, f, lst, dist), 2 * self.mb_count, np.int16).reshape(-1, 2)

    def mv_costs(self, f, lst, dist):
        return self._arr(self.o.orc_la_mv_costs(self.h, f, lst, dist), self.mb_count, np.int32)

    def lowres_costs(self, f, d0, d1):
        return self._arr(self.o.orc_la_lowres_costs(self.h, f, d0, d1), self.mb_count, np.uint16)

    def row_satds(self, f, d0, d1):
        return self._arr(self.o.orc_la_row_satds(self.h, f, d0, d1), (self.p.height + 15) >> 4, np.int32)

    def cost_est(self, f, d0, d1, aq=False):
        return self.o.orc_la_cost_est(self.h, f, d0, d1, int(aq))

    def intra_mbs(self, f, d0):
        return self.o.orc_la_intra_mbs(self.h, f, d0)

    def pixel_stats(self, f):
        s, q = (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
        self.o.orc_la_pixel_stats(self.h, f, s, q)
        return list(s), list(q)

    def weight(self, f):
        w = (C.c_int * 4)()
        self.o.orc_la_weight(self.h, f, w)
        return dict(scale=w[0], denom=w[1], offset=w[2], on=w[3])

    def mbtree(self, frame_idx, types, b_intra=0):
        n = len(frame_idx) - 1
        fi = (C.c_int * (n + 1))(*frame_idx)
        ty = (C.c_int * (n + 1))(*types)
        self.o.orc_la_mbtree(self.h, fi, ty, n, b_intra)

    def counters(self):
        c = (C.c_uint64 * 4)()
        self.o.orc_la_counters(self.h, c)
        return dict(mb_cost=c[0], searches=c[1], sad=c[2], satd=c[3])


# ---- f4: decoder-side output conversion (oracle/decode_oracle.c) ---------------------------------------------
def decode_picture_size(out_csp, w, h):
    o = oracle()
    o.orc_decode_picture_size.restype = C.c_int64
    o.orc_decode_picture_size.argtypes = [C.c_int, C.c_int, C.c_int]
    return int(o.orc_decode_picture_size(out_csp, w, h))


def oracle_decode_convert(y, u, v, out_csp, avcol_spc=2, fullrange=0, src_chroma=1):
    """y, u, v: 2-D uint8 planes of one decoded yuv420p (src_chroma 1), yuv422p (2) or yuv444p (3) picture (any row stride).  Returns the
    output DIB bytes, or None where the checker refuses (-1)."""
    o = oracle()
    o.orc_decode_convert_src.restype = C.c_int
    o.orc_decode_convert_src.argtypes = [C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int,
                                         C.c_int, C.c_int]
    h, w = y.shape
    size = decode_picture_size(out_csp, w, h)
    if size < 0:
        return None
    out = np.zeros(size, np.uint8)
    src = (C.c_void_p * 3)(y.ctypes.data, u.ctypes.data, v.ctypes.data)
    ss = (C.c_int * 3)(y.strides[0], u.strides[0], v.strides[0])
    if o.orc_decode_convert_src(src_chroma, out_csp, out.ctypes.data, src, ss, w, h, avcol_spc, fullrange) != 0:
        return None
    return out


def decode_source(w, h, seed=0, pad=0, src_chroma=1):
    """Seeded yuv420p / yuv422p / yuv444p (src_chroma 1 / 2 / 3) picture (SURVEY A.4 byte generator, seed folded into the generator's size arguments); rows
    carry `pad` spare bytes so that strides differ from widths like a decoder's AVFrame linesize."""
    cw, ch = (w if src_chroma == 3 else w // 2), (h if src_chroma >= 2 else h // 2)
    raw = lcg_bytes((w + pad) * h + 2 * (cw + pad) * ch, w + 7 * seed, h + 13 * seed)
    y = raw[:(w + pad) * h].reshape(h, w + pad)[:, :w]
    o = (w + pad) * h
    u = raw[o:o + (cw + pad) * ch].reshape(ch, cw + pad)[:, :cw]
    o += (cw + pad) * ch
    v = raw[o:o + (cw + pad) * ch].reshape(ch, cw + pad)[:, :cw]
    return y, u, v


def oracle_integral_init(plane: np.ndarray, with_sum4=True):
    """[x264] integral_init8h/8v (+4h/4v) on a padded plane (2-D uint8, contiguous rows): (sum8, sum4 or None), uint16 arrays of the
    plane's shape; valid where the 8x8 (4x4) window lies inside the plane."""
    o = oracle()
    rows, stride = plane.shape
    plane = np.ascontiguousarray(plane)
    s8 = np.zeros((rows, stride), np.uint16)
    s4 = np.zeros((rows, stride), np.uint16) if with_sum4 else None
    o.orc_integral_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    o.orc_integral_init(s8.ctypes.data, s4.ctypes.data if with_sum4 else None, plane.ctypes.data, stride, rows)
    return s8, s4
