"""Row a15 without self-comparison: the C checker's decision logic (and with it the product's la_host.cu, which
is compared with the checker on the GPU) against tests/py_slicetype.py, a second implementation of [x264]'s
x264_slicetype_decide / x264_slicetype_analyse / scenecut / slicetype_path written separately in Python and driven
only by the checker's COST table (no decision code of the checker runs in the engine session)."""
import numpy as np
import pytest

import oracle_lib as ol
from py_slicetype import CostEngine, PyLookahead

BGRA_FLIP = 9 | 0x1000
TYPE_CH = "?IiPbB"


def clip_planes(w, h, n, **kw):
    from clipgen import SyntheticClip
    clip = SyntheticClip(w, h, n_frames=n, **kw)
    return [ol.oracle_convert(clip.packed(i, "bgra"), BGRA_FLIP, 2, 2, 0, w, h) for i in range(n)]


def c_session(params, planes):
    orc = ol.OracleLookahead(params)
    out = []
    try:
        for p in planes:
            orc.put_i420(p)
            out += orc.decisions()
        orc.flush()
        out += orc.decisions()
    finally:
        orc.close()
    return out


def py_session(params, planes):
    eng = CostEngine(params)
    try:
        la = PyLookahead(params, eng)
        for p in planes:
            la.put(p)
        la.flush()
        return la.decisions
    finally:
        eng.close()


def types_of(ds):
    return "".join(TYPE_CH[d["i_type"]] for d in sorted(ds, key=lambda d: d["i_frame"]))


CASES = [
    ("medium", dict(rc_lookahead=12, keyint_max=50, keyint_min=5), dict(cuts=(25,), flash=36, flash_len=1)),
    ("medium", dict(rc_lookahead=40, keyint_max=250, keyint_min=25), dict(cuts=(20, 45), flash=30, flash_len=2)),
    ("medium", dict(rc_lookahead=10, keyint_max=16, keyint_min=2), dict(cuts=(), flash=None)),
    ("slower", dict(rc_lookahead=12, keyint_max=50, keyint_min=5), dict(cuts=(25,), flash=36, flash_len=2)),
    ("veryslow", dict(rc_lookahead=16, keyint_max=60, keyint_min=5), dict(cuts=(31,), flash=12, flash_len=1)),
    ("slow", dict(b_adapt=2, rc_lookahead=30, keyint_max=24, keyint_min=2), dict(cuts=(40,), flash=20, flash_len=3)),
    ("veryfast", dict(rc_lookahead=8, keyint_max=50, keyint_min=5), dict(cuts=(25,), flash=36, flash_len=1)),
    ("superfast", dict(keyint_max=50, keyint_min=5), dict(cuts=(25,), flash=36, flash_len=1)),
    ("ultrafast", dict(), dict(cuts=(25,), flash=None)),
    ("medium", dict(rc_lookahead=12, b_pyramid=0, b_adapt=0, keyint_max=20, keyint_min=2), dict(cuts=(25,), flash=36, flash_len=1)),
    ("medium", dict(rc_lookahead=10, open_gop=1, keyint_max=24, keyint_min=2, bframes=5), dict(cuts=(25,), flash=36, flash_len=1)),
    ("medium", dict(rc_lookahead=12, b_pyramid=1, keyint_max=40, keyint_min=4), dict(cuts=(25,), flash=36, flash_len=1)),
    ("medium", dict(rc_lookahead=12, scenecut=0, keyint_max=40, keyint_min=4), dict(cuts=(25,), flash=36, flash_len=1)),
    ("medium", dict(rc_lookahead=12, b_mbtree=0, keyint_max=40, keyint_min=4), dict(cuts=(25,), flash=36, flash_len=1)),
    ("medium", dict(rc_lookahead=12, b_psy=0, aq_mode=2, keyint_max=30, keyint_min=30), dict(cuts=(25,), flash=36, flash_len=1)),
    ("medium", dict(rc_lookahead=12, bframes=0, keyint_max=40, keyint_min=4), dict(cuts=(25,), flash=None)),
    ("medium", dict(rc_lookahead=2, bframes=3, keyint_max=40, keyint_min=4), dict(cuts=(25,), flash=36, flash_len=1)),
    ("medium", dict(rc_lookahead=12, weightp=0, weightb=0, keyint_max=40, keyint_min=4), dict(cuts=(25,), flash=36, flash_len=1)),
]


@pytest.mark.parametrize("preset,over,clip_kw", CASES, ids=[f"{c[0]}-{i}" for i, c in enumerate(CASES)])
def test_c_decision_logic_equals_the_python_restatement(preset, over, clip_kw):
    w, h, n = 128, 96, 60
    planes = clip_planes(w, h, n, **clip_kw)
    params = ol.la_params(preset, w, h, **over)
    want = py_session(params, planes)
    got = c_session(params, planes)
    assert types_of(got) == types_of(want)
    assert [d["i_frame"] for d in got] == [d["i_frame"] for d in want]                 # coded order
    for a, b in zip(got, want):
        for k in ("i_type", "b_keyframe", "i_bframes", "i_cost_est", "i_cost_est_aq", "i_intra_mbs"):
            assert a[k] == b[k], (k, a["i_frame"], a[k], b[k])
        # the same evaluations in the same order: even the mb-tree offsets come out bit-identical
        assert np.array_equal(a["qp_offset"].view(np.uint32), b["qp_offset"].view(np.uint32)), a["i_frame"]
        assert np.array_equal(a["qp_offset_aq"].view(np.uint32), b["qp_offset_aq"].view(np.uint32)), a["i_frame"]


def test_python_restatement_on_a_static_clip_and_a_fade():
    """Corner content: identical frames (every cost ties) and a fade (weights, X264_WEIGHTP_FAKE)."""
    w, h, n = 128, 96, 40
    still = clip_planes(w, h, 1, cuts=(), flash=None) * n
    base = clip_planes(w, h, n, cuts=(), flash=None)
    fade = [((f.astype(np.int32) * (256 - 5 * i)) >> 8).astype(np.uint8) for i, f in enumerate(base)]
    for planes in (still, fade):
        for over in (dict(rc_lookahead=10, keyint_max=250, keyint_min=5), dict(rc_lookahead=10, keyint_max=250, keyint_min=5, weightp=0, b_adapt=2)):
            params = ol.la_params("medium", w, h, **over)
            want, got = py_session(params, planes), c_session(params, planes)
            assert [(d["i_frame"], d["i_type"], d["i_cost_est"]) for d in got] == [(d["i_frame"], d["i_type"], d["i_cost_est"]) for d in want]
            assert all(np.array_equal(a["qp_offset"].view(np.uint32), b["qp_offset"].view(np.uint32)) for a, b in zip(got, want))
