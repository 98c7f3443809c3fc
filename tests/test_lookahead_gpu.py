"""GPU parity tests, stage 2: the CUDA lookahead (through the C ABI) against the CPU oracle
(oracle/lookahead_oracle.c -- PARITY UNPINNED: a restatement of upstream libx264, which the
reference does not vendor).  Integer arrays (planes, costs, MVs, frame types) must be
bit-exact; the float qp offsets are compared bit-exactly too (both sides follow the C
operation order without FMA contraction), far inside north_star's 1e-5 relative bound."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

BGRA_FLIP = 9 | 0x1000


def make_clip(w, h, n, **kw):
    from clipgen import SyntheticClip
    clip = SyntheticClip(w, h, n_frames=n, **kw)
    return [clip.packed(i, "bgra") for i in range(n)]


def to_i420(frames, w, h):
    return [ol.oracle_convert(f, BGRA_FLIP, 2, 2, 0, w, h) for f in frames]


def params_pair(preset, w, h, **over):
    from x264vfw_b200 import lookahead
    return ol.la_params(preset, w, h, **over), lookahead.params_preset(preset, w, h, **over)


def open_pair(preset, w, h, keep=True, in_csp=0, **over):
    from x264vfw_b200 import lookahead
    po, pg = params_pair(preset, w, h, **over)
    return ol.OracleLookahead(po), lookahead.Lookahead(pg, in_csp=in_csp, device=0, keep_frames=keep)


def test_params_presets_agree_with_oracle():
    from x264vfw_b200 import lookahead
    for preset in ("ultrafast", "superfast", "veryfast", "faster", "fast", "medium", "slow", "slower", "veryslow", "placebo"):
        a, b = ol.la_params(preset, 1920, 1080), lookahead.params_preset(preset, 1920, 1080)
        for name, _ in a._fields_:
            assert getattr(a, name) == getattr(b, name), (preset, name)


def test_tunes_only_touch_the_documented_fields():
    from x264vfw_b200 import lookahead
    base = lookahead.params_preset("medium", 1920, 1080)
    want = {"film": {}, "animation": {"frame_reference": 6, "aq_strength": 0.6, "bframes": 5},
            "grain": {"aq_strength": 0.5, "qcompress": 0.8}, "stillimage": {"aq_strength": 1.2},
            "psnr": {"aq_mode": 0, "b_psy": 0}, "ssim": {"aq_mode": 2, "b_psy": 0},
            "fastdecode": {"weightb": 0, "weightp": 0}, "zerolatency": {"rc_lookahead": 0, "bframes": 0, "b_mbtree": 0},
            "touhou": {"frame_reference": 6, "aq_strength": 1.3}}
    for tune, delta in want.items():
        p = lookahead.params_tune(lookahead.params_preset("medium", 1920, 1080), tune)
        for name, _ in p._fields_:
            exp = delta.get(name, getattr(base, name))
            got = getattr(p, name)
            assert (abs(got - exp) < 1e-6) if isinstance(exp, float) else got == exp, (tune, name, got, exp)
    with pytest.raises(ValueError):
        lookahead.params_tune(lookahead.params_preset("medium", 64, 64), "nosuchtune")
    # the reference joins the dialog's tunings with commas (codec.c:1430-1445)
    p = lookahead.params_tune(lookahead.params_preset("medium", 1920, 1080), "film,fastdecode,zerolatency")
    assert (p.weightb, p.weightp, p.rc_lookahead, p.bframes, p.b_mbtree) == (0, 0, 0, 0, 0)
    with pytest.raises(ValueError):
        lookahead.params_tune(lookahead.params_preset("medium", 64, 64), "film,nosuchtune")


def test_open_refuses_what_is_not_restated():
    """Kept-RGB encoder formats and lookaheadless mb-tree fail loudly instead of returning other numbers."""
    from x264vfw_b200 import lookahead
    from x264vfw_b200._lib import CudaError
    p = lookahead.params_preset("medium", 64, 64)
    for out_csp in (0xe, 0xf):
        with pytest.raises(CudaError, match="kept-RGB"):
            lookahead.Lookahead(p, in_csp=9, out_csp=out_csp, device=0)
    with pytest.raises(CudaError, match="rc-lookahead 0"):
        lookahead.Lookahead(lookahead.params_preset("medium", 64, 64, rc_lookahead=0), device=0)


@pytest.mark.parametrize("size,over", [((128, 96), {}), ((320, 192), {}), ((330, 186), {}),
                                       ((320, 192), {"aq_mode": 2}), ((330, 186), {"aq_mode": 3, "aq_strength": 0.8}),
                                       ((320, 192), {"aq_mode": 0})])
def test_frame_preparation_matches_oracle(size, over):
    """AQ statistics, qp offsets, inverse qscale and lowres planes of every put frame (aq-mode 0, 1
    and the auto-variance modes 2 / 3, whose frame averages are float sums in MB order)."""
    w, h = size
    frames = to_i420(make_clip(w, h, 4, cuts=(2,), flash=None), w, h)
    orc, gpu = open_pair("medium", w, h, rc_lookahead=10, **over)
    try:
        for f in frames:
            orc.put_i420(f)
            gpu.put_frame(f)
        g = ol.lowres_geometry(w, h)
        for i in range(len(frames)):
            assert gpu.pixel_stats(i) == orc.pixel_stats(i)
            assert np.array_equal(gpu.inv_qscale(i), orc.inv_qscale(i))
            assert np.array_equal(gpu.qp_offset(i, aq=True).view(np.uint32), orc.qp_offset(i, aq=True).view(np.uint32))
            a = gpu.lowres_planes(i, 4 * g["lplane_bytes"]).reshape(4, g["lh"] + 64, g["lstride"])[:, :, :g["lw"] + 64]
            b = orc.lowres_planes(i).reshape(4, g["lh"] + 64, g["lstride"])[:, :, :g["lw"] + 64]
            assert np.array_equal(a, b)
    finally:
        orc.close(); gpu.close()


def compare_cost(orc, gpu, p0, p1, b):
    co, cg = orc.frame_cost(p0, p1, b), gpu.frame_cost(p0, p1, b)
    d0, d1 = b - p0, p1 - b
    assert np.array_equal(gpu.intra_cost(b), orc.intra_cost(b)), ("intra", p0, p1, b)
    for lst, dist in ((0, d0), (1, d1)):
        if dist:
            mg, mo = gpu.mvs(b, lst, dist), orc.mvs(b, lst, dist)
            bad = np.nonzero((mg != mo).any(axis=1))[0]
            assert bad.size == 0, ("mvs", p0, p1, b, lst, bad[:5], mg[bad[:5]], mo[bad[:5]])
            assert np.array_equal(gpu.mv_costs(b, lst, dist), orc.mv_costs(b, lst, dist)), ("mv_costs", p0, p1, b, lst)
    if p0 != p1:
        lg, lo = gpu.lowres_costs(b, d0, d1), orc.lowres_costs(b, d0, d1)
        bad = np.nonzero(lg != lo)[0]
        assert bad.size == 0, ("lowres_costs", p0, p1, b, bad[:5], lg[bad[:5]], lo[bad[:5]])
    assert np.array_equal(gpu.row_satds(b, d0, d1), orc.row_satds(b, d0, d1)), ("row_satds", p0, p1, b)
    assert np.array_equal(gpu.row_satds(b, 0, 0), orc.row_satds(b, 0, 0)), ("intra row_satds", b)
    assert gpu.cost_est(b, d0, d1)[:2] == [orc.cost_est(b, d0, d1), orc.cost_est(b, d0, d1, aq=True)]
    if b == p1 and p0 != p1:
        assert gpu.cost_est(b, d0, d1)[2] == orc.intra_mbs(b, d0)
    assert gpu.weight(b) == orc.weight(b), ("weight", p0, p1, b)
    assert cg == co, ("score", p0, p1, b, cg, co)


COST_SEQUENCE = [(0, 0, 0), (0, 1, 1), (1, 2, 2), (0, 2, 2), (0, 2, 1), (2, 3, 3), (1, 3, 3), (1, 3, 2), (0, 3, 3),
                 (0, 3, 1), (0, 3, 2), (3, 4, 4), (0, 4, 4), (0, 4, 2), (2, 4, 3), (4, 5, 5), (3, 5, 4), (5, 5, 5)]


@pytest.mark.parametrize("preset,size,over", [
    ("medium", (128, 96), {}),
    ("medium", (320, 192), {}),
    ("medium", (330, 186), {"lookahead_threads": 3}),
    ("veryfast", (320, 192), {}),
    ("superfast", (320, 192), {"b_mbtree": 1}),
    ("medium", (320, 192), {"weightb": 0, "aq_mode": 0, "weightp": 0}),
    ("medium", (64, 48), {"mv_range": 32}),
    # geometry corners of the search kernels: 1x1 / 2x2 / 3x2 MB frames (every MB is an edge MB), one
    # very wide row band, a tall narrow frame (more rows than columns: deep row pipeline, short rows)
    ("medium", (16, 16), {}),
    ("medium", (32, 32), {}),
    ("medium", (48, 32), {}),
    ("medium", (1280, 48), {}),
    ("medium", (48, 400), {}),
    ("veryslow", (320, 192), {}),
    ("medium", (320, 192), {"b_mbtree": 0}),
    ("superfast", (330, 186), {"lookahead_threads": 4}),
])
def test_frame_cost_matches_oracle(preset, size, over):
    """slicetype_frame_cost on explicit (p0,p1,b) triples in a fixed order: MVs, MV costs,
    per-MB lowres costs, intra costs, frame sums, weights."""
    w, h = size
    frames = to_i420(make_clip(w, h, 6, cuts=(4,), flash=2, flash_len=1), w, h)
    orc, gpu = open_pair(preset, w, h, rc_lookahead=20, **over)
    try:
        for f in frames:
            orc.put_i420(f)
            gpu.put_frame(f)
        for (p0, p1, b) in COST_SEQUENCE:
            compare_cost(orc, gpu, p0, p1, b)
    finally:
        orc.close(); gpu.close()


def test_frame_cost_with_fade_exercises_weights():
    """A linear fade makes the lookahead weight analysis pick a non-trivial weight."""
    w, h = 320, 192
    base = make_clip(w, h, 6, cuts=(), flash=None)
    faded = []
    for i, f in enumerate(base):
        k = 256 - 30 * i
        faded.append(((f.astype(np.int32) * k) >> 8).astype(np.uint8))
    frames = to_i420(faded, w, h)
    orc, gpu = open_pair("medium", w, h, rc_lookahead=20)
    try:
        for f in frames:
            orc.put_i420(f)
            gpu.put_frame(f)
        seen = 0
        for (p0, p1, b) in [(0, 1, 1), (1, 2, 2), (0, 2, 2), (0, 2, 1), (2, 3, 3), (3, 5, 5), (3, 5, 4)]:
            compare_cost(orc, gpu, p0, p1, b)
            seen += orc.weight(b)["on"]
        assert seen > 0, "fade clip did not trigger weighted prediction"
    finally:
        orc.close(); gpu.close()


def fade_clip(w, h, n, step=30):
    base = make_clip(w, h, n, cuts=(), flash=None)
    return [((f.astype(np.int32) * (256 - step * i)) >> 8).astype(np.uint8) for i, f in enumerate(base)]


def test_weightp_fake_on_a_fade():
    """tune fastdecode (weightp 0) with mb-tree and psy is X264_WEIGHTP_FAKE: the lookahead still analyses and
    uses luma weights, and macroblock_tree_finish consumes f_weighted_cost_delta."""
    w, h = 320, 192
    frames = to_i420(fade_clip(w, h, 6), w, h)
    orc, gpu = open_pair("medium", w, h, rc_lookahead=20, weightp=0, weightb=0)
    try:
        for f in frames:
            orc.put_i420(f)
            gpu.put_frame(f)
        seen = 0
        for (p0, p1, b) in [(0, 1, 1), (1, 2, 2), (0, 2, 2), (0, 2, 1), (2, 3, 3), (3, 5, 5), (3, 5, 4)]:
            compare_cost(orc, gpu, p0, p1, b)
            seen += orc.weight(b)["on"]
        assert seen > 0, "fade clip did not trigger the weight analysis under X264_WEIGHTP_FAKE"
    finally:
        orc.close(); gpu.close()


def run_session(la, frames, put):
    out = []
    for f in frames:
        put(la, f)
        out += la.decisions()
    la.flush()
    out += la.decisions()
    return out


def compare_sessions(preset, w, h, n, clip_kw, over, in_csp=0):
    packed = make_clip(w, h, n, **clip_kw)
    i420 = to_i420(packed, w, h)
    orc, gpu = open_pair(preset, w, h, keep=False, in_csp=in_csp, **over)
    try:
        do = run_session(orc, i420, lambda la, f: la.put_i420(f))
        dg = run_session(gpu, packed if in_csp else i420, lambda la, f: la.put_frame(f))
        assert [d["i_frame"] for d in dg] == [d["i_frame"] for d in do]          # coded order
        to = "".join(ol.TYPE_NAMES[d["i_type"]][0] if d["i_type"] != 4 else "b" for d in sorted(do, key=lambda d: d["i_frame"]))
        tg = "".join(ol.TYPE_NAMES[d["i_type"]][0] if d["i_type"] != 4 else "b" for d in sorted(dg, key=lambda d: d["i_frame"]))
        assert tg == to
        for a, b in zip(dg, do):
            for k in ("i_type", "b_keyframe", "i_bframes", "i_cost_est", "i_cost_est_aq", "i_intra_mbs"):
                assert a[k] == b[k], (k, a["i_frame"], a[k], b[k])
            assert np.array_equal(a["qp_offset_aq"].view(np.uint32), b["qp_offset_aq"].view(np.uint32)), a["i_frame"]
            # north_star: mb-tree float propagation within 1e-5 relative; we get bit-exact
            assert np.allclose(a["qp_offset"], b["qp_offset"], rtol=1e-5, atol=1e-6), a["i_frame"]
            assert np.array_equal(a["qp_offset"].view(np.uint32), b["qp_offset"].view(np.uint32)), a["i_frame"]
        return to
    finally:
        orc.close(); gpu.close()


@pytest.mark.parametrize("preset,over", [
    ("medium", {"rc_lookahead": 12, "keyint_max": 50, "keyint_min": 5}),
    ("veryfast", {"rc_lookahead": 8, "keyint_max": 50, "keyint_min": 5}),
    ("slower", {"rc_lookahead": 12, "keyint_max": 50, "keyint_min": 5}),
    ("superfast", {"keyint_max": 50, "keyint_min": 5}),
    ("ultrafast", {}),
    ("medium", {"rc_lookahead": 12, "b_pyramid": 0, "b_adapt": 0, "keyint_max": 20, "keyint_min": 2}),
    ("medium", {"rc_lookahead": 10, "open_gop": 1, "keyint_max": 24, "keyint_min": 2, "bframes": 5}),
    ("veryslow", {"rc_lookahead": 16, "keyint_max": 60, "keyint_min": 5}),
    ("medium", {"rc_lookahead": 12, "b_pyramid": 1, "keyint_max": 40, "keyint_min": 4}),
    # tune zerolatency: no lookahead, no B-frames, no mb-tree
    ("medium", {"rc_lookahead": 0, "bframes": 0, "b_mbtree": 0, "keyint_max": 30, "keyint_min": 3}),
    ("medium", {"rc_lookahead": 12, "scenecut": 0, "keyint_max": 40, "keyint_min": 4}),
    ("fast", {"rc_lookahead": 12, "lookahead_threads": 3, "keyint_max": 40, "keyint_min": 4}),
    # tune ssim (aq-mode 2, no psy) and the biased auto-variance mode
    ("medium", {"rc_lookahead": 12, "aq_mode": 2, "b_psy": 0, "keyint_max": 40, "keyint_min": 4}),
    ("medium", {"rc_lookahead": 12, "aq_mode": 3, "keyint_max": 40, "keyint_min": 4}),
])
def test_session_decisions_match_oracle(preset, over):
    """Whole sessions: frame types, coded order, rate-control costs and per-MB qp offsets."""
    types = compare_sessions(preset, 320, 192, 48, dict(cuts=(25,), flash=36, flash_len=1), over)
    assert types[0] == "I"


def test_session_fastdecode_fade_matches_oracle():
    """Whole session under X264_WEIGHTP_FAKE on a fading clip: the qp offsets carry the weight delta."""
    w, h, n = 320, 192, 40
    frames = to_i420(fade_clip(w, h, n, step=5), w, h)
    over = {"rc_lookahead": 12, "keyint_max": 50, "keyint_min": 5, "weightp": 0, "weightb": 0}
    orc, gpu = open_pair("medium", w, h, keep=False, **over)
    try:
        do = run_session(orc, frames, lambda la, f: la.put_i420(f))
        dg = run_session(gpu, frames, lambda la, f: la.put_frame(f))
    finally:
        orc.close(); gpu.close()
    assert [(d["i_frame"], d["i_type"], d["i_cost_est"], d["i_cost_est_aq"]) for d in dg] == \
           [(d["i_frame"], d["i_type"], d["i_cost_est"], d["i_cost_est_aq"]) for d in do]
    for a, b in zip(dg, do):
        assert np.array_equal(a["qp_offset"].view(np.uint32), b["qp_offset"].view(np.uint32)), a["i_frame"]


def test_nv12_session_uses_the_chroma_planes():
    """NV12 in -> NV12 out (csp.c:490-492): the AQ energies include U and V read out of the interleaved plane,
    so every decision equals the I420 session's on the same samples."""
    from x264vfw_b200 import lookahead
    w, h, n = 320, 192, 30
    i420 = to_i420(make_clip(w, h, n, cuts=(17,), flash=None), w, h)
    nv12 = []
    for f in i420:
        y, u, v = f[:w * h], f[w * h:w * h * 5 // 4], f[w * h * 5 // 4:]
        uv = np.empty(w * h // 2, dtype=np.uint8)
        uv[0::2] = u; uv[1::2] = v
        nv12.append(np.concatenate([y, uv]))
    over = {"rc_lookahead": 10, "keyint_max": 50, "keyint_min": 5}
    po, pg = params_pair("medium", w, h, **over)
    orc = ol.OracleLookahead(po)
    gpu = lookahead.Lookahead(pg, in_csp=5, out_csp=4, device=0)
    try:
        do = run_session(orc, i420, lambda la, f: la.put_i420(f))
        dg = run_session(gpu, nv12, lambda la, f: la.put_frame(f))
    finally:
        orc.close(); gpu.close()
    assert [(d["i_frame"], d["i_type"], d["i_cost_est"], d["i_cost_est_aq"]) for d in dg] == \
           [(d["i_frame"], d["i_type"], d["i_cost_est"], d["i_cost_est_aq"]) for d in do]
    for a, b in zip(dg, do):
        assert np.array_equal(a["qp_offset_aq"].view(np.uint32), b["qp_offset_aq"].view(np.uint32)), a["i_frame"]
        assert np.array_equal(a["qp_offset"].view(np.uint32), b["qp_offset"].view(np.uint32)), a["i_frame"]


def test_session_with_device_csp_front_end_matches_oracle():
    """Packed bottom-up BGRA in, stage 1 + stage 2 on the device, decisions identical to the
    oracle fed with the oracle-converted planes; conv_pic D2H equals the oracle planes."""
    from x264vfw_b200 import lookahead
    w, h = 320, 192
    packed = make_clip(w, h, 8, cuts=(5,), flash=None)
    compare_sessions("medium", w, h, 24, dict(cuts=(13,), flash=None), {"rc_lookahead": 8, "keyint_max": 50, "keyint_min": 5},
                     in_csp=BGRA_FLIP)
    pg = lookahead.params_preset("medium", w, h, rc_lookahead=4)
    la = lookahead.Lookahead(pg, in_csp=BGRA_FLIP, device=0)
    try:
        conv = np.zeros(w * h * 3 // 2, dtype=np.uint8)
        la.put_frame(packed[0], conv_pic=conv)
        assert np.array_equal(conv, ol.oracle_convert(packed[0], BGRA_FLIP, 2, 2, 0, w, h))
    finally:
        la.close()


def test_mbtree_whitebox_matches_oracle():
    """macroblock_tree over an explicit type pattern: propagate costs (saturated view) and the
    resulting qp offsets."""
    w, h = 320, 192
    frames = to_i420(make_clip(w, h, 8, cuts=(), flash=None), w, h)
    orc, gpu = open_pair("medium", w, h, rc_lookahead=20)
    try:
        for f in frames:
            orc.put_i420(f)
            gpu.put_frame(f)
        idx = list(range(8))
        for types, b_intra in (([1, 5, 4, 5, 3, 5, 3, 3], 1), ([3, 5, 5, 3, 3, 5, 5, 3], 0)):
            orc.mbtree(idx, types, b_intra)
            gpu.mbtree(idx, types, b_intra)
            for i in idx:
                assert np.array_equal(np.minimum(gpu.propagate(i), 32767).astype(np.uint16), orc.propagate_cost(i)), i
                assert np.array_equal(gpu.qp_offset(i).view(np.uint32), orc.qp_offset(i).view(np.uint32)), i
    finally:
        orc.close(); gpu.close()


def test_full_size_properties_1080p():
    """BASELINE size (1080p, preset medium): size-independent properties instead of a slow
    full oracle pass -- a frame predicted from itself costs exactly the zero-residual floor,
    MVs of a static clip are zero, and one P evaluation is still bit-exact vs the oracle."""
    w, h = 1920, 1080
    frames = to_i420(make_clip(w, h, 2, cuts=(), flash=None), w, h)
    orc, gpu = open_pair("medium", w, h, rc_lookahead=20, weightp=0)
    try:
        for f in (frames[0], frames[0], frames[1]):
            orc.put_i420(f)
            gpu.put_frame(f)
        gpu.frame_cost(0, 1, 1)
        assert not gpu.mvs(1, 0, 1).any()
        lc = gpu.lowres_costs(1, 1, 0)
        assert ((lc & 0x3fff) == 4).all() and ((lc >> 14) == 1).all()     # SATD 0 + lowres_penalty, list0
        compare_cost(orc, gpu, 1, 2, 2)
    finally:
        orc.close(); gpu.close()


def test_concurrent_sessions_are_deterministic_and_match_oracle():
    """Eight sessions driven from eight host threads on one GPU (the bench configuration):
    every one must produce exactly the oracle's decisions, whatever the interleaving of the
    streams (side-stream searches, adaptive speculation, recycled frame slots)."""
    import threading
    from x264vfw_b200 import lookahead
    w, h, n = 320, 192, 60
    packed = make_clip(w, h, n, cuts=(25,), flash=40, flash_len=1)
    i420 = to_i420(packed, w, h)
    over = {"rc_lookahead": 12, "keyint_max": 50, "keyint_min": 5}
    orc = ol.OracleLookahead(ol.la_params("medium", w, h, **over))
    want = run_session(orc, i420, lambda la, f: la.put_i420(f))
    orc.close()
    results, errors = {}, []

    def work(k):
        try:
            la = lookahead.Lookahead(lookahead.params_preset("medium", w, h, **over), in_csp=BGRA_FLIP, device=0)
            results[k] = run_session(la, packed, lambda s, f: s.put_frame(f))
            la.close()
        except Exception as e:          # noqa: BLE001
            errors.append(e)

    ths = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors
    for k in range(8):
        got = results[k]
        assert [(d["i_frame"], d["i_type"], d["i_cost_est"], d["i_cost_est_aq"]) for d in got] == \
               [(d["i_frame"], d["i_type"], d["i_cost_est"], d["i_cost_est_aq"]) for d in want], k
        assert all(np.array_equal(a["qp_offset"].view(np.uint32), b["qp_offset"].view(np.uint32)) for a, b in zip(got, want)), k


def test_gop_segmented_clip_equals_one_reference_session_per_segment():
    """a single clip shards only as GOP segments, each through its own session (closed GOP, own IDR).
    Two 'ranks' share the device here; the stitched result must be what one CPU reference session per segment gives."""
    from x264vfw_b200 import lookahead, sharding
    w, h, n, seg = 192, 128, 30, 12
    packed = make_clip(w, h, n, cuts=(17,))
    i420 = to_i420(packed, w, h)
    po, pg = params_pair("medium", w, h, rc_lookahead=8, keyint_max=50, keyint_min=5)
    parts = {}
    for rank in range(2):
        parts.update(sharding.run_clip_segments(lambda: lookahead.Lookahead(pg, device=0), i420, seg, rank, 2))
    got = sharding.stitch_segments(parts)
    want = []
    for a, b in sharding.gop_segments(n, seg):
        orc = ol.OracleLookahead(po)
        try:
            want += [dict(d, i_frame=d["i_frame"] + a) for d in run_session(orc, i420[a:b], lambda la, f: la.put_i420(f))]
        finally:
            orc.close()
    assert [d["i_frame"] for d in got] == [d["i_frame"] for d in want]
    assert sorted(d["i_frame"] for d in got) == list(range(n))
    for a, b in zip(got, want):
        for k in ("i_type", "b_keyframe", "i_bframes", "i_cost_est", "i_cost_est_aq", "i_intra_mbs"):
            assert a[k] == b[k], (k, a["i_frame"], a[k], b[k])
        assert np.array_equal(a["qp_offset"].view(np.uint32), b["qp_offset"].view(np.uint32)), a["i_frame"]
    starts = {a for a, _ in sharding.gop_segments(n, seg)}
    assert all(d["i_type"] == 1 and d["b_keyframe"] for d in got if d["i_frame"] in starts)   # X264_TYPE_IDR


def _golden_cases():
    import json, os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lookahead_golden.json")))


@pytest.mark.parametrize("case", _golden_cases(), ids=[c["name"] for c in _golden_cases()])
def test_device_path_matches_the_committed_fingerprints(case):
    """The whole device path (packed BGRA in, csp + lookahead) against tests/golden/lookahead_golden.json,
    without the checker in the loop."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_lookahead_golden as mk
    fp = mk.fingerprint(case["preset"], case["w"], case["h"], case["frames"], case["over"], backend="gpu")
    for k in ("types", "coded_order", "costs", "qp_offset_fnv", "qp_offset_aq_fnv"):
        assert fp[k] == case[k], (case["name"], k)
