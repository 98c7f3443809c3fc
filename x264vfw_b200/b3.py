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
