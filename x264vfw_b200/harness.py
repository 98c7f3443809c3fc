"""ctypes binding of host/libx264vfw_harness.so: the plain-C host side above the C ABI
(host/x264vfw_harness.c, mirrors codec.c's compress_begin / compress / compress_end call order).

StreamSet drives several open sessions at once with ONE NATIVE HOST THREAD PER STREAM -- the way
the reference is driven (codec.c:1728 runs on the application thread, one CODEC per stream).
bench.py and the tests use it so that the per-frame loop is C, not Python threads contending for
the interpreter lock."""
import ctypes as C
import os
import subprocess

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "host", "libx264vfw_harness.so")


class _Stream(C.Structure):
    _fields_ = [("la", C.c_void_p), ("frames", C.POINTER(C.c_void_p)), ("n_frames", C.c_int), ("on_device", C.c_int),
                ("in_csp", C.c_int), ("out_csp", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("conv", C.POINTER(C.c_void_p)), ("n_conv", C.c_int),
                ("pos", C.c_long), ("decided", C.c_long), ("checksum", C.c_double),
                ("mb_count", C.c_int), ("error", C.c_int), ("qp", C.c_void_p), ("qp_aq", C.c_void_p), ("count", C.c_int),
                ("log", C.c_void_p), ("log_cap", C.c_int), ("log_n", C.c_int)]


class _Decision(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("i_frame", "i_type", "b_keyframe", "i_bframes", "i_cost_est", "i_cost_est_aq",
                                       "i_intra_mbs", "mb_count")]


class _LogRec(C.Structure):
    _fields_ = [("d", _Decision), ("qp_fnv", C.c_uint64), ("qp_aq_fnv", C.c_uint64)]


def _load():
    if not os.path.exists(_SO):
        subprocess.run(["make", "-C", os.path.join(_ROOT, "host")], check=True, capture_output=True)
    lib = C.CDLL(_SO)
    lib.harness_run_streams.restype = C.c_int
    lib.harness_run_streams.argtypes = [C.POINTER(_Stream), C.c_int, C.c_int]
    lib.harness_stream_free.argtypes = [C.POINTER(_Stream)]
    lib.harness_last_error.restype = C.c_char_p
    lib.harness_flush_streams.restype = C.c_int
    lib.harness_flush_streams.argtypes = [C.POINTER(_Stream), C.c_int]
    return lib


class StreamSet:
    """sessions: open lookahead.Lookahead objects; frames[s]: that stream's clip as device
    addresses (on_device) or uint8 numpy buffers (host); conv[s]: optional list of host numpy
    buffers that receive conv_pic."""

    def __init__(self, sessions, frames, on_device, conv=None, log_decisions=0):
        """on_device: 0 host buffers, 1 device buffers borrowed for the call, 2 device buffers that stay
        resident and unmodified (x264vfw_cuda.h: X264VFW_CUDA_SRC_*).  log_decisions: keep up to that many
        decisions per stream (frame, type, costs, FNV-1a-64 of the qp offset arrays) for parity tests."""
        self.lib = _load()
        self.n = len(sessions)
        self.arr = (_Stream * self.n)()
        self._keep = [frames, conv, sessions]
        for s, la in enumerate(sessions):
            st = self.arr[s]
            ptrs = [int(f) if on_device else f.ctypes.data for f in frames[s]]
            fa = (C.c_void_p * len(ptrs))(*ptrs)
            self._keep.append(fa)
            st.la = la.h
            st.frames = fa
            st.n_frames = len(ptrs)
            st.on_device = int(on_device)
            st.in_csp, st.out_csp = la.in_csp, la.out_csp
            st.width, st.height = la.p.width, la.p.height
            if conv:
                ca = (C.c_void_p * len(conv[s]))(*[b.ctypes.data for b in conv[s]])
                self._keep.append(ca)
                st.conv = ca
                st.n_conv = len(conv[s])
            if log_decisions:
                buf = (_LogRec * log_decisions)()
                self._keep.append(buf)
                st.log = C.cast(buf, C.c_void_p)
                st.log_cap = log_decisions
        self._logs = log_decisions

    def run(self, count: int) -> None:
        """Feed `count` frames to every stream concurrently (native threads; returns when all did)."""
        if self.lib.harness_run_streams(self.arr, self.n, count) < 0:
            from ._lib import last_error
            raise RuntimeError("harness_run_streams failed: " + (self.lib.harness_last_error() or b"").decode() + " " + last_error())

    def flush(self) -> None:
        """x264vfw_cuda_la_flush on every session, decisions drained into the counters / the log."""
        if self.lib.harness_flush_streams(self.arr, self.n) < 0:
            from ._lib import last_error
            raise RuntimeError("harness_flush_streams failed: " + (self.lib.harness_last_error() or b"").decode() + " " + last_error())

    def log(self, s):
        """Logged decisions of stream s in coded order: dicts with the decision fields and the two hashes."""
        st = self.arr[s]
        recs = C.cast(st.log, C.POINTER(_LogRec))
        out = []
        for i in range(st.log_n):
            r = recs[i]
            d = {n: int(getattr(r.d, n)) for n, _ in _Decision._fields_}
            d["qp_fnv"] = "%016x" % r.qp_fnv
            d["qp_aq_fnv"] = "%016x" % r.qp_aq_fnv
            out.append(d)
        return out

    @property
    def decided(self):
        return [int(self.arr[s].decided) for s in range(self.n)]

    def close(self):
        for s in range(self.n):
            self.lib.harness_stream_free(C.byref(self.arr[s]))
