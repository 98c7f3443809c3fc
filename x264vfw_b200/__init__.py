"""x264vfw_b200 -- B200 (sm_100a) front end for the x264vfw encoder path.

Thin ctypes mirror of the C ABI in include/x264vfw_cuda.h.  The product is the CUDA
library (x264vfw_b200/libx264vfw_cuda.so, sources in x264vfw_b200/csrc); Python is only
the test / benchmark harness language.  There is no CPU fallback: importing works without
a GPU (symbols can be inspected) but every compute call fails without a CUDA device.
"""
from ._lib import lib, LibraryMissing, last_error, version, launch_count  # noqa: F401
from . import csp, lowres, hpel, lookahead, decode  # noqa: F401
