"""ctypes mirror of the B3 entry points: upstream libx264's function-table shapes on device pointers
(include/x264vfw_cuda.h "B3"; [x264] common/mc.c, common/pixel.c, encoder/slicetype-cl.c hook names)."""
import ctypes as C

from ._lib import lib, Context, CudaError, last_error, Image

_V, _I = C.c_void_p, C.c_int
_sig = {
    "x264vfw_cuda_frame_init_lowres_core": [_V, _V, _V, _V, _V, _V, C.c_ssize_t, C.c_ssize_t, _I, _I],
    "x264vfw_cuda_mbtree_propagate_cost": [_V, _V, _V, _V, _V, _V, C.c_float, _I],
    "x264vfw_cuda_mbtree_propagate_list": [_V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _I],
    "x264vfw_cuda_pixel_cmp_8x8": [_V, _I, _V, C.c_ssize_t, _V, C.c_ssize_t, _V, _V, _V, _I],
    "x264vfw_cuda_pixel_sad_xn_8x8": [_V, _I, _V, C.c_ssize_t, _V, _V, C.c_ssize_t, _V, _V, _I],
    "x264vfw_cuda_intra_mbcmp_x3_8x8c": [_V, _I, _V, _I, _I, _I, _V],
    "x264vfw_cuda_opencl_lowres_init": [_V, C.POINTER(Image), _I],
    "x264vfw_cuda_opencl_motionsearch": [_V, _I, _I, _I],
    "x264vfw_cuda_opencl_finalize_cost": [_V, _I, _I, _I, C.POINTER(C.c_int)],
    "x264vfw_cuda_opencl_flush": [_V],
    "x264vfw_cuda_opencl_slicetype_prep": [_V, _I, _I],
    "x264vfw_cuda_opencl_slicetype_end": [_V],
}


class WeightsIn(C.Structure):
    """x264vfw_cuda_weights_in: inputs of [x264] x264_weights_analyse(h, fenc, ref, 0) as device pointers."""
    _fields_ = [("width", _I), ("height", _I), ("fenc_lowres", _V), ("ref_lowres", _V), ("lowres_mvs", _V), ("intra_cost", _V),
                ("fenc_uv", _V), ("ref_uv", _V), ("uv_stride", _I),
                ("fenc_sum", C.c_uint64 * 3), ("fenc_ssd", C.c_uint64 * 3), ("ref_sum", C.c_uint64 * 3), ("ref_ssd", C.c_uint64 * 3),
                ("subme", _I), ("weightp", _I)]


_sig["x264vfw_cuda_weights_analyse"] = [_V, C.POINTER(WeightsIn), C.POINTER(C.c_int32 * 4 * 3), C.POINTER(C.c_float)]
_sig["x264vfw_cuda_la_weights_analyse"] = [_V, _I, _I, _V, _V, _I, C.POINTER(C.c_int32 * 4 * 3), C.POINTER(C.c_float)]
_sig["x264vfw_cuda_integral_init"] = [_V, _V, _V, _V, _I, _I, C.c_size_t, C.c_size_t, _I]
for _n, _a in _sig.items():
    getattr(lib, _n).restype = C.c_int
    getattr(lib, _n).argtypes = _a


def _ck(rc):
    if rc < 0:
        raise CudaError(last_error())
    return rc


def frame_init_lowres_core(ctx: Context, src0, dst0, dsth, dstv, dstc, src_stride, dst_stride, width, height):
    _ck(lib.x264vfw_cuda_frame_init_lowres_core(ctx.handle, src0, dst0, dsth, dstv, dstc, src_stride, dst_stride, width, height))


def mbtree_propagate_cost(ctx: Context, dst, propagate_in, intra_costs, inter_costs, inv_qscales, fps_factor, n):
    _ck(lib.x264vfw_cuda_mbtree_propagate_cost(ctx.handle, dst, propagate_in, intra_costs, inter_costs, inv_qscales, fps_factor, n))


def mbtree_propagate_list(ctx: Context, ref_costs, mvs, propagate_amount, lowres_costs, bipred_weight, mb_y, n, lst, mb_w, mb_h):
    _ck(lib.x264vfw_cuda_mbtree_propagate_list(ctx.handle, ref_costs, mvs, propagate_amount, lowres_costs, bipred_weight, mb_y, n, lst, mb_w, mb_h))


def pixel_cmp_8x8(ctx: Context, satd, pix1, stride1, pix2, stride2, off1, off2, scores, n):
    _ck(lib.x264vfw_cuda_pixel_cmp_8x8(ctx.handle, int(satd), pix1, stride1, pix2, stride2, off1, off2, scores, n))


def pixel_sad_xn_8x8(ctx: Context, n_ref, fenc, fenc_stride, off_fenc, ref, ref_stride, off_ref, scores, n):
    _ck(lib.x264vfw_cuda_pixel_sad_xn_8x8(ctx.handle, n_ref, fenc, fenc_stride, off_fenc, ref, ref_stride, off_ref, scores, n))


def intra_mbcmp_x3_8x8c(ctx: Context, satd, plane, stride, mb_w, mb_h, res):
    _ck(lib.x264vfw_cuda_intra_mbcmp_x3_8x8c(ctx.handle, int(satd), plane, stride, mb_w, mb_h, res))


class OpenclHooks:
    """The x264_opencl_* hook names on a cost-engine session (lookahead.Lookahead opened with keep_frames=True and
    rc_lookahead = 250)."""

    def __init__(self, la):
        self.la = la

    def lowres_init(self, frame, on_device=False):
        ptr = int(frame) if on_device else frame.ctypes.data
        src, _ = self.la._src_image(ptr)
        return _ck(lib.x264vfw_cuda_opencl_lowres_init(self.la.h, C.byref(src), int(on_device)))

    def motionsearch(self, b, ref, b_islist1):
        _ck(lib.x264vfw_cuda_opencl_motionsearch(self.la.h, b, ref, int(b_islist1)))

    def finalize_cost(self, p0, p1, b):
        out = (C.c_int * 3)()
        score = _ck(lib.x264vfw_cuda_opencl_finalize_cost(self.la.h, p0, p1, b, out))
        return score, list(out)

    def flush(self):
        _ck(lib.x264vfw_cuda_opencl_flush(self.la.h))

    def slicetype_prep(self, first, num_frames):
        _ck(lib.x264vfw_cuda_opencl_slicetype_prep(self.la.h, first, num_frames))

    def slicetype_end(self):
        _ck(lib.x264vfw_cuda_opencl_slicetype_end(self.la.h))


def weights_analyse(ctx: Context, win: WeightsIn):
    """[x264] x264_weights_analyse(h, fenc, ref, 0) on device buffers.  Returns ([[on, scale, denom, offset]] * 3, cost_delta)."""
    out = (C.c_int32 * 4 * 3)()
    delta = C.c_float(0)
    _ck(lib.x264vfw_cuda_weights_analyse(ctx.handle, C.byref(win), C.byref(out), C.byref(delta)))
    return [[int(out[p][i]) for i in range(4)] for p in range(3)], float(delta.value)


def la_weights_analyse(la, fenc: int, ref: int, fenc_uv: int, ref_uv: int, uv_stride: int):
    """The same for display indices of a lookahead session opened with keep_frames; fenc_uv / ref_uv: device addresses of the
    frames' NV12 chroma planes padded to mod 16 (x264vfw_cuda_chroma_nv12_pad)."""
    out = (C.c_int32 * 4 * 3)()
    delta = C.c_float(0)
    _ck(lib.x264vfw_cuda_la_weights_analyse(la.h, fenc, ref, fenc_uv, ref_uv, uv_stride, C.byref(out), C.byref(delta)))
    return [[int(out[p][i]) for i in range(4)] for p in range(3)], float(delta.value)


def integral_init(ctx: Context, sum8: int, sum4: int, plane: int, stride: int, rows: int, plane_bytes: int = 0, sum_elems: int = 0,
                  n_frames: int = 1):
    """[x264] integral_init8h/8v (+4h/4v): 8x8 (and 4x4) box sums of a padded plane, device addresses."""
    _ck(lib.x264vfw_cuda_integral_init(ctx.handle, sum8, sum4 or None, plane, stride, rows, plane_bytes, sum_elems, n_frames))
