"""The plain-C host harness (host/x264vfw_harness.c) compiles against the C ABI without CUDA
headers; on a GPU it runs a clip end to end through compress_begin/compress/compress_end."""
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
