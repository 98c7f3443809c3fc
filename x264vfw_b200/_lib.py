"""Loader for libx264vfw_cuda.so (built in-tree by x264vfw_b200/csrc/Makefile)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("X264VFW_CUDA_LIB") or os.path.join(_HERE, "libx264vfw_cuda.so")


class LibraryMissing(RuntimeError):
    pass


class Image(C.Structure):
    """x264vfw_cuda_image_t == libx264's public x264_image_t layout (csp.h:46)."""
    _fields_ = [("i_csp", C.c_int), ("i_plane", C.c_int), ("i_stride", C.c_int * 4),
                ("plane", C.c_void_p * 4)]


CSP_FN = C.CFUNCTYPE(C.c_int, C.POINTER(Image), C.POINTER(Image), C.c_int, C.c_int)


class CspFunctionTable(C.Structure):
    """x264vfw_cuda_csp_function_t == x264vfw_csp_function_t (csp.h:48-51)."""
    _fields_ = [("convert", CSP_FN * 10)]


class LowresGeom(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("mb_w", "mb_h", "luma_w", "luma_h", "luma_stride",
                                       "lw", "lh", "lstride", "lplane_bytes", "lorigin")]


class HpelGeom(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("stride", "plane_bytes", "origin")]


def _load():
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for this package)")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    sig = {
        "x264vfw_cuda_last_error": (C.c_char_p, []),
        "x264vfw_cuda_version": (C.c_char_p, []),
        "x264vfw_cuda_launch_count": (C.c_uint64, []),
        "x264vfw_cuda_ctx_create": (C.c_int, [P(C.c_void_p), C.c_int]),
        "x264vfw_cuda_ctx_destroy": (None, [C.c_void_p]),
        "x264vfw_cuda_ctx_stream": (C.c_void_p, [C.c_void_p]),
        "x264vfw_cuda_ctx_sync": (C.c_int, [C.c_void_p]),
        "x264vfw_cuda_csp_init": (None, [P(CspFunctionTable), C.c_int, C.c_int, C.c_int]),
        "x264vfw_cuda_csp_convert": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                               P(Image), P(Image), C.c_int, C.c_int]),
        "x264vfw_cuda_csp_convert_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                     P(Image), P(Image), C.c_int, C.c_int,
                                                     C.c_size_t, C.c_size_t, C.c_int]),
        "x264vfw_cuda_img_fill": (C.c_int64, [P(Image), C.c_void_p, C.c_int, C.c_int, C.c_int]),
        "x264vfw_cuda_picture_layout": (C.c_int64, [P(Image), C.c_void_p, C.c_int, C.c_int, C.c_int]),
        "x264vfw_cuda_lowres_geometry": (None, [P(LowresGeom), C.c_int, C.c_int]),
        "x264vfw_cuda_luma_pad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_size_t, C.c_size_t, C.c_int]),
        "x264vfw_cuda_chroma_nv12_pad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                   C.c_size_t, C.c_size_t, C.c_int]),
        "x264vfw_cuda_hpel_geometry": (None, [P(HpelGeom), C.c_int, C.c_int]),
        "x264vfw_cuda_hpel_filter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                               C.c_size_t, C.c_size_t, C.c_int]),
        "x264vfw_cuda_lowres_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                               C.c_size_t, C.c_size_t, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error() -> str:
    return (lib.x264vfw_cuda_last_error() or b"").decode()


def version() -> str:
    return lib.x264vfw_cuda_version().decode()


def launch_count() -> int:
    return int(lib.x264vfw_cuda_launch_count())


class CudaError(RuntimeError):
    pass


class Context:
    """x264vfw_cuda_ctx: one device + one stream + staging buffers."""

    def __init__(self, device: int = -1):
        h = C.c_void_p()
        if lib.x264vfw_cuda_ctx_create(C.byref(h), device) < 0:
            raise CudaError(last_error())
        self.handle = h

    @property
    def stream(self) -> int:
        return int(lib.x264vfw_cuda_ctx_stream(self.handle) or 0)

    def sync(self):
        if lib.x264vfw_cuda_ctx_sync(self.handle) < 0:
            raise CudaError(last_error())

    def close(self):
        if self.handle:
            lib.x264vfw_cuda_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
