"""Host-side mirror of the lookahead's frame preparation: [x264] x264_frame_init_lowres
(common/mc.c) and x264_frame_expand_border_mod16/_lowres (common/frame.c), which the
reference reaches only through x264_encoder_encode (codec.c:1693)."""
import ctypes as C

from ._lib import lib, LowresGeom, Context, CudaError, last_error


def geometry(width: int, height: int) -> LowresGeom:
    g = LowresGeom()
    lib.x264vfw_cuda_lowres_geometry(C.byref(g), width, height)
    return g


def lowres_init(ctx: Context, d_dst: int, d_y: int, y_stride: int, width: int, height: int,
                src_frame_bytes: int = 0, dst_frame_bytes: int = 0, n_frames: int = 1):
    """Tight luma (device) -> 4 padded half-pel phase planes (device), n_frames per launch."""
    rc = lib.x264vfw_cuda_lowres_init(ctx.handle, C.c_void_p(d_dst), C.c_void_p(d_y), y_stride, width, height,
                                      src_frame_bytes, dst_frame_bytes, n_frames)
    if rc < 0:
        raise CudaError(last_error())


def luma_pad(ctx: Context, d_dst: int, d_y: int, y_stride: int, width: int, height: int,
             src_frame_bytes: int = 0, dst_frame_bytes: int = 0, n_frames: int = 1):
    rc = lib.x264vfw_cuda_luma_pad(ctx.handle, C.c_void_p(d_dst), C.c_void_p(d_y), y_stride, width, height,
                                   src_frame_bytes, dst_frame_bytes, n_frames)
    if rc < 0:
        raise CudaError(last_error())


def chroma_nv12_pad(ctx: Context, d_dst: int, dst_stride: int, d_u: int, d_v: int, c_stride: int, width: int, height: int,
                    src_frame_bytes: int = 0, dst_frame_bytes: int = 0, n_frames: int = 1):
    """[x264] x264_frame_copy_picture chroma (planar U, V -> NV12) + x264_frame_expand_border_mod16."""
    rc = lib.x264vfw_cuda_chroma_nv12_pad(ctx.handle, C.c_void_p(d_dst), dst_stride, C.c_void_p(d_u), C.c_void_p(d_v), c_stride,
                                          width, height, src_frame_bytes, dst_frame_bytes, n_frames)
    if rc < 0:
        raise CudaError(last_error())


# ---- fused front end: csp.convert[] + [x264] x264_adaptive_quant_frame + x264_frame_init_lowres in one kernel -----
from ._lib import Image  # noqa: E402

lib.x264vfw_cuda_frontend_batch.restype = C.c_int
lib.x264vfw_cuda_frontend_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Image), C.POINTER(Image), C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int]


def frontend_batch(ctx: Context, d_src: int, d_dst: int, d_lowres: int, d_qp_aq: int, d_invq: int, d_stats: int, in_csp: int,
                   width: int, height: int, n_frames: int, src_frame_bytes: int, dst_frame_bytes: int,
                   colmatrix: int = 2, fullrange: int = 0, aq_strength: float = 1.0):
    """x264vfw_cuda_frontend_batch on device addresses (packed BGRA in; I420 planes, 4 lowres planes, per-MB
    qp_offset_aq / inv_qscale and 6 frame sums per frame out), asynchronous on ctx's stream."""
    from . import csp as _csp
    src, _ = _csp.img_fill(d_src, in_csp, width, height)
    dst, _ = _csp.picture_layout(d_dst, _csp.X264_CSP_I420, width, height)
    rc = lib.x264vfw_cuda_frontend_batch(ctx.handle, colmatrix, fullrange, C.byref(dst), C.byref(src), C.c_void_p(d_lowres),
                                         C.c_void_p(d_qp_aq), C.c_void_p(d_invq), C.c_void_p(d_stats), aq_strength,
                                         width, height, src_frame_bytes, dst_frame_bytes, n_frames)
    if rc < 0:
        raise CudaError(last_error())


class FusedBatch:
    """Scratch for n_frames of AQ outputs + a bound frontend_batch call (bench.py stage-1 probe)."""

    def __init__(self, ctx: Context, width: int, height: int, n_frames: int, in_csp: int = 9 | 0x1000):
        import torch
        from . import csp as _csp
        self.ctx, self.w, self.h, self.n, self.in_csp = ctx, width, height, n_frames, in_csp
        g = geometry(width, height)
        mb = g.mb_w * g.mb_h
        self.qp = torch.empty(n_frames * mb, dtype=torch.float32, device="cuda")
        self.invq = torch.empty(n_frames * mb, dtype=torch.int16, device="cuda")
        self.stats = torch.empty(n_frames * 6, dtype=torch.int64, device="cuda")
        self.sfb, self.dfb = _csp.frame_bytes(in_csp, _csp.X264_CSP_I420, width, height)

    def run(self, d_src: int, d_dst: int, d_lowres: int):
        frontend_batch(self.ctx, d_src, d_dst, d_lowres, self.qp.data_ptr(), self.invq.data_ptr(), self.stats.data_ptr(), self.in_csp,
                       self.w, self.h, self.n, self.sfb, self.dfb)
