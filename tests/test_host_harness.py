"""The plain-C host harness (host/x264vfw_harness.c) compiles against the C ABI without CUDA
headers; on a GPU it runs a clip end to end through compress_begin/compress/compress_end."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    return os.path.join(ROOT, "host", "x264vfw_harness")


def test_harness_builds_with_plain_gcc():
    exe = build()
    assert os.path.exists(exe)
    # object code must not need anything but the C ABI library
    out = subprocess.run(["gcc", "-std=c99", "-pedantic", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "x264vfw_cuda.h")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


@pytest.mark.gpu
def test_harness_runs_a_clip():
    exe = build()
    out = subprocess.run([exe, "320", "192", "40", "medium"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("frame ")]
    assert sorted(int(l.split()[1]) for l in lines) == list(range(40))
    assert lines[0].split()[1] == "0" and lines[0].split()[3] == "1"       # first coded frame is the IDR


@pytest.mark.gpu
def test_native_stream_runner_matches_the_python_driver():
    """harness_run_streams (one native thread per stream) produces the same decisions as feeding
    the same clips frame by frame from Python."""
    import numpy as np
    from x264vfw_b200 import lookahead
    from clipgen import SyntheticClip
    from x264vfw_b200.harness import StreamSet
    build()
    w, h, n = 256, 144, 36
    in_csp = 9 | 0x1000
    clips = [[SyntheticClip(w, h, n_frames=n, stream_id=s, cuts=(20,), flash=None).packed(i, "bgra") for i in range(n)] for s in range(2)]

    def params():
        return lookahead.params_preset("medium", w, h, rc_lookahead=10, keyint_max=30, keyint_min=3)

    ref = []
    for s in range(2):
        la = lookahead.Lookahead(params(), in_csp=in_csp, device=0)
        out = []
        for f in clips[s]:
            la.put_frame(f)
            out += la.decisions()
        la.flush(); out += la.decisions()
        la.close()
        ref.append([(d["i_frame"], d["i_type"]) for d in out])

    sessions = [lookahead.Lookahead(params(), in_csp=in_csp, device=0) for _ in range(2)]
    conv = [[np.empty(w * h * 3 // 2, dtype=np.uint8) for _ in range(2)] for _ in range(2)]
    ss = StreamSet(sessions, clips, False, conv)
    ss.run(n)                      # n < 2n-2: the ping-pong playback has not turned around yet
    got = []
    for s, la in enumerate(sessions):
        early = ss.decided[s]
        la.flush()
        tail = la.decisions()
        assert early + len(tail) == n
        got.append([(d["i_frame"], d["i_type"]) for d in tail])
        la.close()
    ss.close()
    for s in range(2):
        assert got[s] == ref[s][len(ref[s]) - len(got[s]):]


class _Bih(C.Structure):
    """harness_bih: the BITMAPINFOHEADER fields the path looks at."""
    _fields_ = [("biWidth", C.c_int), ("biHeight", C.c_int), ("biBitCount", C.c_int), ("biCompression", C.c_uint)]


def _harness_lib():
    build()
    lib = C.CDLL(os.path.join(ROOT, "host", "libx264vfw_harness.so"))
    lib.harness_decompress_query.argtypes = [C.POINTER(_Bih), C.POINTER(_Bih), C.c_uint]
    lib.harness_decompress_begin.argtypes = [C.c_void_p, C.POINTER(_Bih), C.POINTER(_Bih), C.c_int, C.c_int, C.c_int]
    lib.harness_decompress.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_void_p]
    lib.harness_decompress_end.argtypes = [C.c_void_p]
    return lib


def test_harness_decompress_query_follows_the_reference_rules():
    """codec.c:1930-1980 through the C harness; no device needed."""
    lib = _harness_lib()
    fcc = lambda s: ord(s[0]) | ord(s[1]) << 8 | ord(s[2]) << 16 | ord(s[3]) << 24
    h264 = _Bih(64, 32, 24, fcc("H264"))
    q = lambda out, size=0, inp=h264: lib.harness_decompress_query(C.byref(inp), C.byref(out) if out else None, size)
    assert q(None) == 0
    assert q(_Bih(64, 32, 32, 0)) == 0 and q(_Bih(64, -32, 24, 0)) == 0 and q(_Bih(64, 32, 16, fcc("YUY2"))) == 0
    assert q(_Bih(64, 32, 12, fcc("YV12"))) == 0 and q(_Bih(64, 32, 12, fcc("NV12"))) == 0
    assert q(_Bih(32, 32, 32, 0)) == -2                      # size must match
    assert q(_Bih(64, 32, 16, 0)) == -2                      # RGB565: no csp
    assert q(_Bih(64, 32, 32, 0), size=100) == -2            # biSizeImage too small
    assert q(_Bih(64, 32, 32, 0), size=64 * 32 * 4) == 0
    assert q(_Bih(64, 33, 32, 0), inp=_Bih(64, 33, 24, fcc("H264"))) == -2   # odd height


@pytest.mark.gpu
def test_harness_decompress_matches_the_checker():
    """decompress_begin / decompress / decompress_end of the C harness: bottom-up RGB32, RGB24 and YV12 output DIBs of one
    decoded picture equal the checker's (and so libswscale's, tests/test_decode_oracle.py)."""
    import numpy as np
    import oracle_lib as ol
    lib = _harness_lib()
    fcc = lambda s: ord(s[0]) | ord(s[1]) << 8 | ord(s[2]) << 16 | ord(s[3]) << 24
    w, h = 320, 176
    y, u, v = ol.decode_source(w, h, seed=5, pad=32)
    inp = _Bih(w, h, 24, fcc("H264"))
    for out, csp in ((_Bih(w, h, 32, 0), 9 | 0x1000), (_Bih(w, -h, 24, 0), 8), (_Bih(w, h, 12, fcc("YV12")), 2), (_Bih(w, h, 16, fcc("UYVY")), 7)):
        dec = C.create_string_buffer(256)
        assert lib.harness_decompress_begin(dec, C.byref(inp), C.byref(out), 1, 1, 0) == 0
        dib = np.zeros(ol.decode_picture_size(csp, w, h), np.uint8)
        data = (C.c_void_p * 3)(y.ctypes.data, u.ctypes.data, v.ctypes.data)
        ls = (C.c_int * 3)(y.strides[0], u.strides[0], v.strides[0])
        for _ in range(2):                                   # the second call reuses the context (codec.c:2282)
            assert lib.harness_decompress(dec, data, ls, dib.ctypes.data) == 0
        lib.harness_decompress_end(dec)
        assert (dib == ol.oracle_decode_convert(y, u, v, csp, 1, 0)).all(), hex(csp)
