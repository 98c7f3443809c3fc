"""Host-side mirror of the encoder's reference-frame preparation: [x264]
x264_frame_expand_border, x264_frame_filter -> x264_mc_functions_t.hpel_filter (common/mc.c) and
x264_frame_expand_border_filtered (common/frame.c), which the reference reaches only through
x264_encoder_encode (codec.c:1693)."""
import ctypes as C

from ._lib import lib, HpelGeom, Context, CudaError, last_error


def geometry(width: int, height: int) -> HpelGeom:
    g = HpelGeom()
    lib.x264vfw_cuda_hpel_geometry(C.byref(g), width, height)
    return g


def hpel_filter(ctx: Context, d_dst: int, d_src: int, src_stride: int, width: int, height: int,
                src_frame_bytes: int = 0, dst_frame_bytes: int = 0, n_frames: int = 1):
    """Tight reconstructed plane (device) -> the four padded planes filtered[0][0..3] (device): the frame with
    its 32-pixel border, then the H, V and centre half-pel planes; n_frames per launch."""
    rc = lib.x264vfw_cuda_hpel_filter(ctx.handle, C.c_void_p(d_dst), C.c_void_p(d_src), src_stride, width, height,
                                      src_frame_bytes, dst_frame_bytes, n_frames)
    if rc < 0:
        raise CudaError(last_error())
