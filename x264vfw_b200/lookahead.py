"""Host-side mirror of the lookahead boundary: what happens inside x264_encoder_encode
(reference codec.c:1693) between receiving a picture and knowing its frame type and per-MB
qp offsets ([x264] x264_adaptive_quant_frame, x264_frame_init_lowres, x264_slicetype_decide,
x264_slicetype_analyse, macroblock_tree).  All compute runs in libx264vfw_cuda.so."""
import ctypes as C

import numpy as np

from ._lib import lib, Image, CudaError, last_error
from . import csp as _csp

X264_TYPE_AUTO, X264_TYPE_IDR, X264_TYPE_I, X264_TYPE_P, X264_TYPE_BREF, X264_TYPE_B = 0, 1, 2, 3, 4, 5
TYPE_CHAR = {1: "I", 2: "i", 3: "P", 4: "b", 5: "B"}

(LA_LOWRES, LA_INTRA_COST, LA_INV_QSCALE, LA_PROPAGATE, LA_QP_OFFSET, LA_QP_OFFSET_AQ, LA_MVS, LA_MV_COSTS,
 LA_LOWRES_COSTS, LA_COST_EST, LA_PIXEL_STATS, LA_WEIGHT, LA_CONV_PLANES, LA_ROW_SATDS) = range(1, 15)


class LaParams(C.Structure):
    """x264vfw_cuda_la_params: the x264_param_t fields the lookahead reads."""
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


_P = C.POINTER
lib.x264vfw_cuda_la_params_preset.restype = C.c_int
lib.x264vfw_cuda_la_params_preset.argtypes = [_P(LaParams), C.c_char_p, C.c_int, C.c_int]
lib.x264vfw_cuda_la_open.restype = C.c_int
lib.x264vfw_cuda_la_open.argtypes = [_P(C.c_void_p), _P(LaParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
lib.x264vfw_cuda_la_close.restype = None
lib.x264vfw_cuda_la_close.argtypes = [C.c_void_p]
lib.x264vfw_cuda_la_put_frame.restype = C.c_int
lib.x264vfw_cuda_la_put_frame.argtypes = [C.c_void_p, _P(Image), C.c_int, _P(Image)]
lib.x264vfw_cuda_la_flush.restype = C.c_int
lib.x264vfw_cuda_la_flush.argtypes = [C.c_void_p]
lib.x264vfw_cuda_la_get_decision.restype = C.c_int
lib.x264vfw_cuda_la_get_decision.argtypes = [C.c_void_p, _P(LaDecision), C.c_void_p, C.c_void_p]
lib.x264vfw_cuda_la_frame_cost.restype = C.c_int
lib.x264vfw_cuda_la_frame_cost.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
lib.x264vfw_cuda_la_mbtree.restype = C.c_int
lib.x264vfw_cuda_la_mbtree.argtypes = [C.c_void_p, _P(C.c_int), _P(C.c_int), C.c_int, C.c_int]
lib.x264vfw_cuda_la_read.restype = C.c_int64
lib.x264vfw_cuda_la_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
lib.x264vfw_cuda_la_counters.restype = None
lib.x264vfw_cuda_la_counters.argtypes = [C.c_void_p, _P(C.c_uint64)]


lib.x264vfw_cuda_la_stats.restype = C.c_int
lib.x264vfw_cuda_la_stats.argtypes = [C.c_void_p, _P(C.c_uint64)]
lib.x264vfw_cuda_la_profile.restype = C.c_int
lib.x264vfw_cuda_la_profile.argtypes = [C.c_void_p, C.c_int, _P(C.c_double), _P(C.c_uint64)]
KERNEL_CLASSES = ("csp", "aq", "lowres", "intra", "me", "finalize", "weights", "mbtree", "me_pass", "frontend")


lib.x264vfw_cuda_la_params_tune.restype = C.c_int
lib.x264vfw_cuda_la_params_tune.argtypes = [_P(LaParams), C.c_char_p]


def params_tune(p: LaParams, tune: str) -> LaParams:
    """x264_param_apply_tune (codec.c:1463 passes the dialog's tuning with the preset)."""
    if lib.x264vfw_cuda_la_params_tune(C.byref(p), tune.encode()) < 0:
        raise ValueError(last_error())
    return p


def params_preset(preset: str, width: int, height: int, **over) -> LaParams:
    """x264_param_default_preset (codec.c:1463) reduced to the lookahead's fields; keyword
    overrides play the role of the extra command line (codec.c:1349)."""
    p = LaParams()
    if lib.x264vfw_cuda_la_params_preset(C.byref(p), preset.encode(), width, height) < 0:
        raise ValueError(last_error())
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class Lookahead:
    """One encoder session's lookahead (one per stream / per CODEC)."""

    def __init__(self, params: LaParams, in_csp: int = 0, out_csp: int = _csp.X264_CSP_I420,
                 colmatrix: int = 2, fullrange: int = 0, device: int = -1, keep_frames: bool = False):
        h = C.c_void_p()
        if lib.x264vfw_cuda_la_open(C.byref(h), C.byref(params), device, in_csp, out_csp, colmatrix, fullrange,
                                    int(keep_frames)) < 0:
            raise CudaError(last_error())
        self.h = h
        self.p = params
        self.in_csp, self.out_csp = in_csp, out_csp
        self.mb_w, self.mb_h = (params.width + 15) >> 4, (params.height + 15) >> 4
        self.mb_count = self.mb_w * self.mb_h

    def close(self):
        if self.h:
            lib.x264vfw_cuda_la_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _src_image(self, ptr: int):
        w, h = self.p.width, self.p.height
        if self.in_csp & 0xff:
            img, n = _csp.img_fill(ptr, self.in_csp, w, h)
        else:
            img, n = _csp.picture_layout(ptr, self.out_csp, w, h)
        return img, n

    def put_frame(self, frame, on_device: bool = False, conv_pic: np.ndarray = None) -> int:
        """frame: numpy uint8 buffer (host) or a device address (int) when on_device.
        conv_pic: optional host buffer receiving the converted planes (codec->conv_pic)."""
        if on_device:
            ptr = int(frame)
        else:
            assert frame.dtype == np.uint8 and frame.flags.c_contiguous
            ptr = frame.ctypes.data
        src, nbytes = self._src_image(ptr)
        if not on_device:
            assert frame.size >= nbytes, (frame.size, nbytes)
        dst_p = None
        if conv_pic is not None:
            dst, _ = _csp.picture_layout(conv_pic.ctypes.data, self.out_csp, self.p.width, self.p.height)
            dst_p = C.byref(dst)
        rc = lib.x264vfw_cuda_la_put_frame(self.h, C.byref(src), int(on_device), dst_p)
        if rc < 0:
            raise CudaError(last_error())
        return rc

    def flush(self) -> int:
        rc = lib.x264vfw_cuda_la_flush(self.h)
        if rc < 0:
            raise CudaError(last_error())
        return rc

    def decisions(self, with_offsets: bool = True):
        out = []
        d = LaDecision()
        while True:
            q = np.empty(self.mb_count, dtype=np.float32) if with_offsets else None
            qa = np.empty(self.mb_count, dtype=np.float32) if with_offsets else None
            rc = lib.x264vfw_cuda_la_get_decision(self.h, C.byref(d), q.ctypes.data if with_offsets else None,
                                                  qa.ctypes.data if with_offsets else None)
            if rc < 0:
                raise CudaError(last_error())
            if rc == 0:
                break
            out.append(dict(i_frame=d.i_frame, i_type=d.i_type, b_keyframe=d.b_keyframe, i_bframes=d.i_bframes,
                            i_cost_est=d.i_cost_est, i_cost_est_aq=d.i_cost_est_aq, i_intra_mbs=d.i_intra_mbs,
                            qp_offset=q, qp_offset_aq=qa))
        return out

    # ---- white-box (parity tests) ------------------------------------------------------------
    def frame_cost(self, p0: int, p1: int, b: int) -> int:
        rc = lib.x264vfw_cuda_la_frame_cost(self.h, p0, p1, b)
        if rc < 0:
            raise CudaError(last_error())
        return rc

    def mbtree(self, frame_idx, types, b_intra=0):
        n = len(frame_idx) - 1
        rc = lib.x264vfw_cuda_la_mbtree(self.h, (C.c_int * (n + 1))(*frame_idx), (C.c_int * (n + 1))(*types), n, b_intra)
        if rc < 0:
            raise CudaError(last_error())

    def read(self, frame: int, what: int, a: int = 0, b: int = 0, dtype=np.uint8, count: int = 0):
        buf = np.empty(count, dtype=dtype)
        rc = lib.x264vfw_cuda_la_read(self.h, frame, what, a, b, buf.ctypes.data, buf.nbytes)
        if rc < 0:
            raise CudaError(last_error())
        return buf[: rc // buf.itemsize]

    def intra_cost(self, f): return self.read(f, LA_INTRA_COST, dtype=np.uint16, count=self.mb_count)
    def inv_qscale(self, f): return self.read(f, LA_INV_QSCALE, dtype=np.uint16, count=self.mb_count)
    def propagate(self, f): return self.read(f, LA_PROPAGATE, dtype=np.int32, count=self.mb_count)
    def qp_offset(self, f, aq=False): return self.read(f, LA_QP_OFFSET_AQ if aq else LA_QP_OFFSET, dtype=np.float32, count=self.mb_count)
    def mvs(self, f, lst, dist): return self.read(f, LA_MVS, lst, dist, dtype=np.int16, count=2 * self.mb_count).reshape(-1, 2)
    def mv_costs(self, f, lst, dist): return self.read(f, LA_MV_COSTS, lst, dist, dtype=np.int32, count=self.mb_count)
    def lowres_costs(self, f, d0, d1): return self.read(f, LA_LOWRES_COSTS, d0, d1, dtype=np.uint16, count=self.mb_count)
    def row_satds(self, f, d0, d1): return self.read(f, LA_ROW_SATDS, d0, d1, dtype=np.int32, count=self.mb_h)
    def cost_est(self, f, d0, d1): return [int(v) for v in self.read(f, LA_COST_EST, d0, d1, dtype=np.int32, count=3)]
    def pixel_stats(self, f):
        v = self.read(f, LA_PIXEL_STATS, dtype=np.uint64, count=6)
        return [int(x) for x in v[:3]], [int(x) for x in v[3:]]
    def weight(self, f):
        v = self.read(f, LA_WEIGHT, dtype=np.int32, count=4)
        return dict(scale=int(v[0]), denom=int(v[1]), offset=int(v[2]), on=int(v[3]))
    def lowres_planes(self, f, nbytes): return self.read(f, LA_LOWRES, dtype=np.uint8, count=nbytes)
    def conv_planes(self, nbytes): return self.read(0, LA_CONV_PLANES, dtype=np.uint8, count=nbytes)

    def profile(self, enable: int = -1):
        """Per-kernel-class device time (ms) and launch counts since the last reset."""
        ms, n = (C.c_double * 16)(), (C.c_uint64 * 16)()
        if lib.x264vfw_cuda_la_profile(self.h, enable, ms, n) < 0:
            raise CudaError(last_error())
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_CLASSES)}

    def stats(self):
        """Device-side work counters, valid while profile(1) is on (x264vfw_cuda_la_stats)."""
        c = (C.c_uint64 * 16)()
        if lib.x264vfw_cuda_la_stats(self.h, c) < 0:
            raise CudaError(last_error())
        names = ("kept", "researched", "pass0", "pass1", "pass2", "pass3", "sad8x8", "satd8x8", "tree_steps", "tree_walks",
                 "spec_jobs", "ondemand_jobs", "ondemand_launches", "searches_asked_for", "frames_since_open")
        return {k: int(c[i]) for i, k in enumerate(names)}

    def counters(self):
        c = (C.c_uint64 * 8)()
        lib.x264vfw_cuda_la_counters(self.h, c)
        return dict(frame_costs=int(c[0]), mb_searches=int(c[1]), launches=int(c[2]), syncs=int(c[3]),
                    put_us=int(c[4]), decide_us=int(c[5]), sync_us=int(c[6]), frames=int(c[7]))


def smoke_check(frames, w=128, h=96):
    """Tiny end-to-end lookahead run on cuda:0 over packed bottom-up BGRA frames;
    __graft_entry__.smoke() checks the returned decisions against its CPU checker."""
    p = params_preset("medium", w, h, rc_lookahead=6, keyint_max=50, keyint_min=2)
    la = Lookahead(p, in_csp=_csp.X264VFW_CSP_BGRA | _csp.X264VFW_CSP_VFLIP, device=0)
    out = []
    for f in frames:
        la.put_frame(f)
        out += la.decisions()
    la.flush()
    out += la.decisions()
    la.close()
    assert sorted(d["i_frame"] for d in out) == list(range(len(frames)))
    return out
