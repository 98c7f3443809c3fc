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
