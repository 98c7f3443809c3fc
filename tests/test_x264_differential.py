"""Differential test against a REAL libx264 -- skipped while none exists (libx264 is neither vendored by
the reference nor installed in the image; baseline/_ref/ is reserved for one).  `make -C oracle x264probe` builds
oracle/_ref/x264_ref_probe when it finds x264.h + libx264 there; this test then compares libx264's own frame types
(pic_out.i_type through the public API, codec.c:1693) with the CPU checker's decisions on the same planes:
north_star's ">= 99.9 % of frames" criterion, and the first external pin of rows a13-a15."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

PROBE = os.path.join(ol.ROOT, "oracle", "_ref", "x264_ref_probe")


def probe_available():
    subprocess.run(["make", "-C", os.path.join(ol.ROOT, "oracle"), "x264probe"], capture_output=True)
    return os.path.exists(PROBE)


@pytest.mark.skipif(not probe_available(), reason="no libx264 under baseline/_ref or oracle/_ref (stage 2 stays 'parity unpinned')")
@pytest.mark.parametrize("preset,opts,over", [
    ("medium", ["rc-lookahead=12", "keyint=50", "min-keyint=5"], dict(rc_lookahead=12, keyint_max=50, keyint_min=5)),
    ("veryfast", ["rc-lookahead=8", "keyint=50", "min-keyint=5"], dict(rc_lookahead=8, keyint_max=50, keyint_min=5)),
    ("slower", ["rc-lookahead=12", "keyint=50", "min-keyint=5"], dict(rc_lookahead=12, keyint_max=50, keyint_min=5)),
])
def test_frame_types_agree_with_libx264(preset, opts, over):
    from clipgen import SyntheticClip
    w, h, n = 320, 192, 120
    clip = SyntheticClip(w, h, n_frames=n, cuts=(40, 85), flash=60, flash_len=1)
    planes = [ol.oracle_convert(clip.packed(i, "bgra"), 9 | 0x1000, 2, 2, 0, w, h) for i in range(n)]
    out = subprocess.run([PROBE, str(w), str(h), str(n), preset] + opts, input=b"".join(p.tobytes() for p in planes),
                         capture_output=True, check=True).stdout.decode().split("\n")
    ref = {int(a): int(b) for a, b, _ in (line.split() for line in out if line.strip())}
    orc = ol.OracleLookahead(ol.la_params(preset, w, h, **over))
    got = {}
    try:
        for p in planes:
            orc.put_i420(p)
            got.update({d["i_frame"]: d["i_type"] for d in orc.decisions()})
        orc.flush()
        got.update({d["i_frame"]: d["i_type"] for d in orc.decisions()})
    finally:
        orc.close()
    assert sorted(ref) == list(range(n))
    agree = np.mean([ref[i] == got[i] for i in range(n)])
    assert agree >= 0.999, (agree, [(i, ref[i], got[i]) for i in range(n) if ref[i] != got[i]][:10])
