"""GPU parity at the BASELINE.json configurations (full frame sizes, shortened clips): the whole
path -- packed input -> device csp -> AQ/lowres -> lookahead decisions -- against the CPU oracle
fed with oracle-converted planes.  Frame types, coded order, rate-control costs and per-MB qp
offsets must be identical."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

FLIP = 0x1000
YUYV, UYVY, BGR, BGRA = 6, 7, 8, 9


def run_config(w, h, n_frames, fmt, in_csp, out_csp, chroma_format, preset, over, ext=0, cuts=None):
    from x264vfw_b200 import lookahead
    from x264vfw_b200.clipgen import SyntheticClip
    clip = SyntheticClip(w, h, n_frames=n_frames, cuts=cuts if cuts is not None else (n_frames * 5 // 8,), flash=n_frames // 3, flash_len=1)
    over = dict(over, chroma_format=chroma_format)
    po, pg = ol.la_params(preset, w, h, **over), lookahead.params_preset(preset, w, h, **over)
    orc = ol.OracleLookahead(po)
    gpu = lookahead.Lookahead(pg, in_csp=in_csp, out_csp=out_csp, device=0)
    do, dg = [], []
    try:
        for n in range(n_frames):
            packed = clip.packed(n, fmt)
            planes = ol.oracle_convert(packed, in_csp, out_csp, 2, 0, w, h, ext=ext)
            assert planes is not None
            orc.put_i420(planes)
            do += orc.decisions()
            gpu.put_frame(packed)
            dg += gpu.decisions()
        orc.flush(); do += orc.decisions()
        gpu.flush(); dg += gpu.decisions()
    finally:
        orc.close(); gpu.close()
    assert len(dg) == n_frames
    assert [d["i_frame"] for d in dg] == [d["i_frame"] for d in do]
    for a, b in zip(dg, do):
        for k in ("i_type", "b_keyframe", "i_bframes", "i_cost_est", "i_cost_est_aq", "i_intra_mbs"):
            assert a[k] == b[k], (k, a["i_frame"], a[k], b[k])
        assert np.array_equal(a["qp_offset"].view(np.uint32), b["qp_offset"].view(np.uint32)), a["i_frame"]
        assert np.array_equal(a["qp_offset_aq"].view(np.uint32), b["qp_offset_aq"].view(np.uint32)), a["i_frame"]
    return "".join({1: "I", 2: "i", 3: "P", 4: "b", 5: "B"}[d["i_type"]] for d in sorted(dg, key=lambda d: d["i_frame"]))


def test_config1_1080p_rgb32_bottom_up_to_i420_medium():
    types = run_config(1920, 1080, 46, "bgra", BGRA | FLIP, 2, 1, "medium", {})
    assert types[0] == "I" and len(types) == 46


def test_config2_720p_yuy2_to_i420_veryfast_lookahead20():
    types = run_config(1280, 720, 40, "yuyv", YUYV, 2, 1, "veryfast", {"rc_lookahead": 20})
    assert types[0] == "I"


def test_config3_1080p_rgb24_to_i420_slow_badapt2_lookahead60():
    """The lookahead of config 3 on the reference-defined I420 planes (b-adapt 2 trellis)."""
    types = run_config(1920, 1080, 24, "bgr", BGR | FLIP, 2, 1, "slow", {"b_adapt": 2, "rc_lookahead": 60})
    assert types[0] == "I"


def test_config3_rgb24_to_nv12_layout_is_the_interleaved_reference_i420():
    """'RGB24 -> NV12' is not a csp.c conversion (csp.c:490-492); the NV12 output is defined as
    the reference I420 result with U/V interleaved (x264_frame_copy_picture)."""
    from x264vfw_b200 import csp
    from x264vfw_b200._lib import Context
    from x264vfw_b200.clipgen import SyntheticClip
    w, h = 1920, 1080
    src = SyntheticClip(w, h, n_frames=1).packed(0, "bgr")
    ctx = Context(0)
    try:
        nv12 = csp.convert_ctx(ctx, src, BGR | FLIP, 4, 2, 0, w, h, ext=csp.EXT_RGB_TO_NV12)
    finally:
        ctx.close()
    i420 = ol.ref_convert(src, BGR | FLIP, 2, 2, 0, w, h) if ol.have_ref_csp() else ol.oracle_convert(src, BGR | FLIP, 2, 2, 0, w, h)
    y, u, v = i420[:w * h], i420[w * h:w * h * 5 // 4], i420[w * h * 5 // 4:]
    assert np.array_equal(nv12[:w * h], y)
    assert np.array_equal(nv12[w * h::2], u) and np.array_equal(nv12[w * h + 1::2], v)


def test_config4_2160p_uyvy_to_i422_medium():
    """Reference-defined 4:2:2 target of config 4 (UYVY -> I422, csp.c:498), High 4:2:2 AQ chroma."""
    types = run_config(3840, 2160, 12, "uyvy", UYVY, 6, 2, "medium", {})
    assert types[0] == "I"


def test_config4_2160p_uyvy_to_i444_extension():
    """UYVY -> I444 has no reference path (csp.c:501-504); extension = I422 samples, chroma doubled."""
    types = run_config(3840, 2160, 6, "uyvy", UYVY, 0xc, 3, "medium", {}, ext=2)
    assert types[0] == "I"
