"""The CPU lookahead checker against its own frozen fingerprints (tests/golden/lookahead_golden.json).
Not a reference pin -- libx264 is absent from the reference tree -- but it keeps the target of the
GPU parity tests from moving unnoticed."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_lookahead_golden as mk  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "lookahead_golden.json")))


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_checker_reproduces_its_golden_fingerprints(case):
    fp = mk.fingerprint(case["preset"], case["w"], case["h"], case["frames"], case["over"])
    for k in ("types", "coded_order", "costs", "qp_offset_fnv", "qp_offset_aq_fnv"):
        assert fp[k] == case[k], (case["name"], k)


def test_block_metrics_against_a_matrix_formulation():
    """SAD and SATD of the checker against numpy: SATD(8x8) = sum over the four 4x4 blocks of sum|H D H^T|, halved
    per 8x4 half ([x264] pixel_satd_8x4), H the 4x4 Hadamard matrix -- an independent formulation, not a reference."""
    import numpy as np
    import oracle_lib as ol
    o = ol.oracle()
    H = np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, -1, 1], [1, -1, 1, -1]], dtype=np.int64)
    rng = np.random.default_rng(264)
    for k in range(200):
        a = rng.integers(0, 256, (8, 8), dtype=np.uint8)
        b = rng.integers(0, 256, (8, 8), dtype=np.uint8) if k % 4 else (a if k % 8 else 255 - a)
        if k % 5 == 0:
            a, b = (a > 127).astype(np.uint8) * 255, (b > 127).astype(np.uint8) * 255
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        d = a.astype(np.int64) - b.astype(np.int64)
        sad = int(np.abs(d).sum())
        satd = 0
        for half in range(2):
            s = 0
            for blk in range(2):
                s += int(np.abs(H @ d[4 * half:4 * half + 4, 4 * blk:4 * blk + 4] @ H.T).sum())
            satd += s >> 1
        assert o.orc_test_sad_8x8(a.ctypes.data, b.ctypes.data) == sad
        assert o.orc_test_satd_8x8(a.ctypes.data, b.ctypes.data) == satd


def test_intra_costs_of_a_session_rebuilt_from_pinned_pieces():
    """Row a11 end to end on the CPU: the checker's per-MB intra costs of a frame against a Python restatement of
    [x264] slicetype_mb_cost's intra part -- neighbours from the padded lowres plane, the ten predictions through
    the decoder-pinned hook (tests/test_h264_pins.py), SATD through the matrix formulation (SAD and three modes only
    at subme <= 1), min + 5*lambda + 4."""
    import ctypes as C
    import numpy as np
    import oracle_lib as ol
    o = ol.oracle()
    H = np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, -1, 1], [1, -1, 1, -1]], dtype=np.int64)

    def satd(a, b):
        d = a.astype(np.int64) - b.astype(np.int64)
        return sum(sum(int(np.abs(H @ d[4 * hf:4 * hf + 4, 4 * bl:4 * bl + 4] @ H.T).sum()) for bl in range(2)) >> 1 for hf in range(2))

    w, h = 112, 80                                      # 7 x 5 macroblocks, lowres 56 x 40
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:h, 0:w]
    y = ((xx * 3 + yy * 2) % 200 + rng.integers(0, 40, (h, w))).astype(np.uint8)     # gradients + noise: several modes win
    for subme, modes in ((7, [0, 1, 2, 3, 13, 14, 15, 16, 17, 18]), (1, [0, 1, 2])):
        la = ol.OracleLookahead(ol.la_params("medium", w, h, subme=subme))
        try:
            la.put_luma(y)
            la.frame_cost(0, 0, 0)
            got = la.intra_cost(0)
            g = la.g
            plane = np.ascontiguousarray(la.lowres_planes(0)[:g["lplane_bytes"]])
            st, org = g["lstride"], g["lorigin"]
            blk = np.zeros(64, dtype=np.uint8)
            winners = set()
            for mby in range(g["mb_h"]):
                for mbx in range(g["mb_w"]):
                    off = org + 8 * mby * st + 8 * mbx
                    src = np.stack([plane[off + r * st:off + r * st + 8] for r in range(8)])
                    best = None
                    for m in modes:
                        o.orc_test_intra_pred_8x8(blk.ctypes.data, m, C.c_void_p(plane.ctypes.data + off), st)
                        # [x264] mbcmp_init: SATD when subme > 1, SAD otherwise
                        c = satd(blk.reshape(8, 8), src) if subme > 1 else int(np.abs(blk.reshape(8, 8).astype(np.int64) - src).sum())
                        if best is None or c < best[0]:
                            best = (c, m)
                    winners.add(best[1])
                    assert got[mby * g["mb_w"] + mbx] == best[0] + 5 + 4, (subme, mbx, mby)
            assert len(winners) >= (3 if subme > 1 else 2)   # the picture exercises more than one predictor
        finally:
            la.close()


def test_inter_costs_at_the_chosen_vectors_rebuilt_from_pinned_pieces():
    """Row a12, the part that is arithmetic rather than search order: for every MB of a P search the checker's stored
    cost must be SATD(source, get_ref(reference planes, chosen vector)) + cost_mv(vector - predictor) - cost_mv(0)
    [+ 5 when the vector is not zero], with the predictor rebuilt here from the final vector field (right, below,
    below-left, below-right neighbours of the reverse-raster scan) -- get_ref through the decoder-pinned hook, SATD
    through the matrix formulation, cost_mv from its defining formula.  Zero-predictor MBs whose co-located SATD is
    below 64 keep the zero vector and that SATD."""
    import ctypes as C
    import math
    import numpy as np
    import oracle_lib as ol
    o = ol.oracle()
    H = np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, -1, 1], [1, -1, 1, -1]], dtype=np.int64)

    def satd(a, b):
        d = a.astype(np.int64) - b.astype(np.int64)
        return sum(sum(int(np.abs(H @ d[4 * hf:4 * hf + 4, 4 * bl:4 * bl + 4] @ H.T).sum()) for bl in range(2)) >> 1 for hf in range(2))

    def cost_mv(d):                                     # [x264] x264_analyse_init_costs at lambda 1
        d = abs(d)
        return int(np.float32(0.718) + np.float32(0.5)) if d == 0 else \
            int(np.float32(np.float32(math.log2(d + 1)) * np.float32(2.0) + np.float32(1.718)) + np.float32(0.5))

    def median3(a, b, c):
        return sorted((a, b, c))[1]

    w, h = 160, 96                                      # 10 x 6 macroblocks
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (h + 32, w + 32), dtype=np.uint8)
    base = ((base.astype(np.int32) + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) // 4).astype(np.uint8)
    f0, f1 = base[8:8 + h, 8:8 + w], base[5:5 + h, 11:11 + w]          # a global shift of (+3, -3) full-res pixels
    la = ol.OracleLookahead(ol.la_params("medium", w, h, weightp=0))
    try:
        la.put_luma(np.ascontiguousarray(f0))
        la.put_luma(np.ascontiguousarray(f1))
        la.frame_cost(0, 1, 1)
        g = la.g
        mvs, costs = la.mvs(1, 0, 1), la.mv_costs(1, 0, 1)
        st, org, pb = g["lstride"], g["lorigin"], g["lplane_bytes"]
        ref = np.ascontiguousarray(la.lowres_planes(0))
        cur = np.ascontiguousarray(la.lowres_planes(1)[:pb])
        mbw, mbh = g["mb_w"], g["mb_h"]
        blk = np.zeros(64, dtype=np.uint8)
        moved = 0
        for mby in range(mbh):
            for mbx in range(mbw):
                i = mby * mbw + mbx
                cand = []
                if mbx < mbw - 1:
                    cand.append(mvs[i + 1])
                if mby < mbh - 1:
                    cand.append(mvs[i + mbw])
                    if mbx > 0:
                        cand.append(mvs[i + mbw - 1])
                    if mbx < mbw - 1:
                        cand.append(mvs[i + mbw + 1])
                cand = [(int(c[0]), int(c[1])) for c in cand] + [(0, 0)] * 4
                n = sum(1 for _ in cand) - 4
                mvp = cand[0] if n <= 1 else (median3(cand[0][0], cand[1][0], cand[2][0]), median3(cand[0][1], cand[1][1], cand[2][1]))
                off = org + 8 * mby * st + 8 * mbx
                src = np.stack([cur[off + r * st:off + r * st + 8] for r in range(8)])
                mx, my = int(mvs[i][0]), int(mvs[i][1])
                o.orc_test_get_ref_8x8(blk.ctypes.data, *[C.c_void_p(ref.ctypes.data + p * pb + off) for p in range(4)], st, mx, my)
                s = satd(src, blk.reshape(8, 8))
                if mvp == (0, 0):
                    o.orc_test_get_ref_8x8(blk.ctypes.data, *[C.c_void_p(ref.ctypes.data + p * pb + off) for p in range(4)], st, 0, 0)
                    s0 = satd(src, blk.reshape(8, 8))
                    if s0 < 64:
                        assert (mx, my) == (0, 0) and costs[i] == s0, (mbx, mby)
                        continue
                want = s + cost_mv(mx - mvp[0]) + cost_mv(my - mvp[1]) - cost_mv(0) + (5 if (mx or my) else 0)
                assert costs[i] == want, (mbx, mby, (mx, my), mvp, int(costs[i]), want)
                moved += (mx, my) != (0, 0)
        assert moved > mbw * mbh // 2                    # the clip really moves: most MBs carry a vector
    finally:
        la.close()


def test_behavioural_anchors_of_the_decision_logic():
    """What any x264 does on these clips, whatever its version (documented behaviour, not a reference pin): the
    first frame is an IDR; a hard scene cut starts with an I-frame that is not a keyframe-interval IDR; a short
    white flash is not mistaken for a cut (how short depends on b-adapt); a static clip uses its B-frames; keyint forces an IDR; with bframes 0
    there are no B-frames; decisions come out in coded order (every B after the P/I that closes its mini-GOP)."""
    import numpy as np
    import oracle_lib as ol
    from clipgen import SyntheticClip
    BGRA_FLIP = 9 | 0x1000
    w, h = 128, 96

    def run(n, over, **clip_kw):
        clip = SyntheticClip(w, h, n_frames=n, **clip_kw)
        la = ol.OracleLookahead(ol.la_params("medium", w, h, **over))
        out = []
        try:
            for i in range(n):
                la.put_i420(ol.oracle_convert(clip.packed(i, "bgra"), BGRA_FLIP, 2, 2, 0, w, h))
                out += la.decisions()
            la.flush()
            out += la.decisions()
        finally:
            la.close()
        assert sorted(d["i_frame"] for d in out) == list(range(n))
        types = {d["i_frame"]: d["i_type"] for d in out}
        # coded order: a B-frame follows the next non-B frame in display order
        pos = {d["i_frame"]: k for k, d in enumerate(out)}
        for f, t in types.items():
            if t in (4, 5):
                nxt = min(g for g, tg in types.items() if g > f and tg not in (4, 5))
                assert pos[nxt] < pos[f], (f, nxt)
        return types, {d["i_frame"]: d["b_keyframe"] for d in out}

    IDR, I, P, BREF, B = 1, 2, 3, 4, 5
    base = dict(rc_lookahead=10, keyint_max=250, keyint_min=5)
    types, key = run(24, base, cuts=(13,), flash=None)
    assert types[0] == IDR and key[0]
    assert types[13] in (I, IDR), types                                  # the cut
    assert all(types[f] not in (I, IDR) for f in range(1, 24) if f != 13), types
    # flash rejection looks ahead to p0 + 2 with b-adapt 1 and to p0 + 1 + bframes with b-adapt 2 ([x264] scenecut):
    # a one-frame flash is never a cut, a two-frame flash is none under b-adapt 2 (and is one under b-adapt 1)
    types, _ = run(24, base, cuts=(), flash=11, flash_len=1)
    assert all(types[f] not in (I, IDR) for f in range(1, 24)), types
    types, _ = run(24, dict(base, b_adapt=2), cuts=(), flash=11, flash_len=2)
    assert all(types[f] not in (I, IDR) for f in range(1, 24)), types
    types, _ = run(24, base, cuts=(), flash=11, flash_len=2)
    assert types[11] in (I, IDR), types
    types, _ = run(16, dict(base, keyint_max=8, keyint_min=2), cuts=(), flash=None)
    assert types[0] == IDR and types[8] == IDR, types                    # keyint
    types, _ = run(12, dict(base, bframes=0), cuts=(), flash=None)
    assert all(t in (IDR, I, P) for t in types.values()), types
    still = SyntheticClip(w, h, n_frames=1, cuts=(), flash=None).packed(0, "bgra")
    la = ol.OracleLookahead(ol.la_params("medium", w, h, **base))
    out = []
    try:
        f = ol.oracle_convert(still, BGRA_FLIP, 2, 2, 0, w, h)
        for _ in range(14):
            la.put_i420(f)
            out += la.decisions()
        la.flush()
        out += la.decisions()
    finally:
        la.close()
    nb = sum(d["i_type"] in (BREF, B) for d in out)
    assert nb >= 6, [d["i_type"] for d in sorted(out, key=lambda d: d["i_frame"])]      # a static clip is mostly B-frames


def test_behavioural_anchors_of_the_qp_offsets():
    """Without mb-tree the offsets handed to the encoder are the AQ offsets; without AQ and mb-tree they are zero;
    with mb-tree on a static clip every MB of an early frame is referenced by the whole window, so its offset is well
    below the AQ offset, and the last P-frame of the clip (nothing references it) keeps the AQ offset."""
    import numpy as np
    import oracle_lib as ol
    from clipgen import SyntheticClip
    w, h, n = 128, 96, 12
    f = ol.oracle_convert(SyntheticClip(w, h, n_frames=1, cuts=(), flash=None).packed(0, "bgra"), 9 | 0x1000, 2, 2, 0, w, h)

    def run(**over):
        la = ol.OracleLookahead(ol.la_params("medium", w, h, rc_lookahead=8, keyint_max=250, keyint_min=5, **over))
        out = []
        try:
            for _ in range(n):
                la.put_i420(f)
                out += la.decisions()
            la.flush()
            out += la.decisions()
        finally:
            la.close()
        return {d["i_frame"]: d for d in out}

    d = run(b_mbtree=0)
    assert all(np.array_equal(x["qp_offset"].view(np.uint32), x["qp_offset_aq"].view(np.uint32)) for x in d.values())
    d = run(b_mbtree=0, aq_mode=0)
    assert all(not x["qp_offset"].any() and not x["qp_offset_aq"].any() for x in d.values())
    d = run()
    assert float((d[0]["qp_offset"] - d[0]["qp_offset_aq"]).max()) < -1.0        # importance raises quality: lower qp
    last_ref = max(k for k, x in d.items() if x["i_type"] in (1, 2, 3))
    assert np.array_equal(d[last_ref]["qp_offset"].view(np.uint32), d[last_ref]["qp_offset_aq"].view(np.uint32))


def test_the_search_recovers_a_global_translation():
    """A picture shifted by (dx, dy) full-resolution pixels is found at (2 dx, 2 dy) quarter samples of the lowres
    plane by (almost) every macroblock: whole-, half- and odd shifts."""
    from collections import Counter
    import numpy as np
    import oracle_lib as ol
    w, h = 160, 96
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (h + 32, w + 32), dtype=np.uint8)
    base = ((base.astype(np.int32) + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) // 4).astype(np.uint8)
    for dx, dy in ((3, -3), (4, 0), (0, -2), (0, 0), (-6, 5)):
        la = ol.OracleLookahead(ol.la_params("medium", w, h, weightp=0))
        try:
            la.put_luma(np.ascontiguousarray(base[8:8 + h, 8:8 + w]))
            la.put_luma(np.ascontiguousarray(base[8 + dy:8 + dy + h, 8 + dx:8 + dx + w]))
            la.frame_cost(0, 1, 1)
            (mv, cnt), = Counter(map(tuple, la.mvs(1, 0, 1).tolist())).most_common(1)
            assert mv == (2 * dx, 2 * dy) and cnt >= 0.9 * la.mb_count, ((dx, dy), mv, cnt)
        finally:
            la.close()


def test_the_weight_analysis_recovers_a_fade():
    """frame1 = frame0 * scale + offset: the lookahead's weight analysis finds scale / 2^denom and the offset (to
    within one level), with the smallest denominator; an unchanged frame gets no weight."""
    import numpy as np
    import oracle_lib as ol
    w, h = 160, 96
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (h, w), dtype=np.uint8)
    base = ((base.astype(np.int32) + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) // 4).astype(np.uint8)
    for sc, of, want in ((0.75, 10, (3, 2)), (0.5, 0, (1, 1)), (1.0, 20, (1, 0)), (1.25, -20, (5, 2)), (1.0, 0, None)):
        la = ol.OracleLookahead(ol.la_params("medium", w, h))
        try:
            la.put_luma(base)
            la.put_luma(np.clip(base.astype(np.float64) * sc + of + 0.5, 0, 255).astype(np.uint8))
            la.frame_cost(0, 1, 1)
            wt = la.weight(1)
            if want is None:
                assert not wt["on"], wt
            else:
                assert wt["on"] and (wt["scale"], wt["denom"]) == want and abs(wt["offset"] - of) <= 1, ((sc, of), wt)
        finally:
            la.close()


def test_aq_offsets_against_a_numpy_formulation():
    """f1 (adaptive quant, aq-mode 1): per MB, energy = AC energy of the 16x16 luma block + both 8x8 chroma blocks
    (ssd - sum^2 >> 8 resp. >> 6) on the frame replicated out to whole macroblocks; offset = 1.0397 * strength * (log2(max(
    energy, 1)) - 14.427).  [x264] x264_log2 is a 128-entry table, so the comparison allows its 0.0113 step; the
    point is the energy (blocks, replication, chroma) rather than the last bit."""
    import numpy as np
    import oracle_lib as ol
    from clipgen import SyntheticClip
    w, h = 136, 88                                      # not multiples of 16: the replicated part counts
    f = ol.oracle_convert(SyntheticClip(w, h, n_frames=1, cuts=(), flash=None).packed(0, "bgra"), 9 | 0x1000, 2, 2, 0, w, h)
    p = ol.la_params("medium", w, h)
    la = ol.OracleLookahead(p)
    try:
        la.put_i420(f)
        got = la.qp_offset(0, aq=True)
    finally:
        la.close()
    y = f[:w * h].reshape(h, w).astype(np.int64)
    u = f[w * h:w * h * 5 // 4].reshape(h // 2, w // 2).astype(np.int64)
    v = f[w * h * 5 // 4:].reshape(h // 2, w // 2).astype(np.int64)
    mbw, mbh = (w + 15) // 16, (h + 15) // 16
    y = np.pad(y, ((0, 16 * mbh - h), (0, 16 * mbw - w)), mode="edge")
    u = np.pad(u, ((0, 8 * mbh - h // 2), (0, 8 * mbw - w // 2)), mode="edge")
    v = np.pad(v, ((0, 8 * mbh - h // 2), (0, 8 * mbw - w // 2)), mode="edge")

    def ac(b, shift):
        s, q = int(b.sum()), int((b * b).sum())
        return q - ((s * s) >> shift)

    strength = float(np.float32(p.aq_strength) * np.float32(1.0397))       # [x264] aq-mode 1: f_aq_strength * 1.0397f
    for mby in range(mbh):
        for mbx in range(mbw):
            e = ac(y[16 * mby:16 * mby + 16, 16 * mbx:16 * mbx + 16], 8) + ac(u[8 * mby:8 * mby + 8, 8 * mbx:8 * mbx + 8], 6) + \
                ac(v[8 * mby:8 * mby + 8, 8 * mbx:8 * mbx + 8], 6)
            want = strength * (np.log2(max(e, 1)) - 14.427)
            assert abs(float(got[mby * mbw + mbx]) - want) <= strength * 0.0115 + 1e-4, (mbx, mby, float(got[mby * mbw + mbx]), want)


def test_one_mbtree_step_against_a_python_formulation():
    """Row a16: frames (I, P, P); the tree walk propagates frame 2 into frame 1 once and finishes frame 1.  The
    propagated amounts ([x264] mbtree_propagate_cost, float32 in C's order) and their split over up to four MBs
    along the lowres vector ([x264] mbtree_propagate_list: x>>5, 32-step bilinear weights, +512 >> 10, saturating
    add, frame-edge cases) are recomputed here from the checker's own inputs; the finished qp offsets are compared
    within the step of the log2 table."""
    import numpy as np
    import oracle_lib as ol
    f32 = np.float32
    w, h = 160, 96
    rng = np.random.default_rng(9)
    base = rng.integers(0, 256, (h + 32, w + 64), dtype=np.uint8)
    base = ((base.astype(np.int32) + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) // 4).astype(np.uint8)
    p = ol.la_params("medium", w, h, weightp=0)
    la = ol.OracleLookahead(p)
    try:
        for k in range(3):                               # a pan of (+5, -3) pixels per frame: vectors (10, -6), a 4-way split
            la.put_luma(np.ascontiguousarray(base[8 + 3 * (2 - k):8 + 3 * (2 - k) + h, 8 + 5 * k:8 + 5 * k + w]))
        la.mbtree([0, 1, 2], [2, 3, 3], 0)               # I, P, P
        got = la.propagate_cost(1).astype(np.int64)
        intra, invq = la.intra_cost(2).astype(np.int64), la.inv_qscale(2).astype(np.int64)
        lc, mvs = la.lowres_costs(2, 1, 0).astype(np.int64), la.mvs(2, 0, 1).astype(np.int64)
        mbw, mbh = la.g["mb_w"], la.g["mb_h"]
        fps_factor = f32(f32(p.fps_den) / f32(p.fps_num)) / (f32(f32(p.fps_den) / f32(p.fps_num)) * f32(256.0)) * f32(0.5)
        want = np.zeros(mbw * mbh, dtype=np.int64)

        def add(idx, v):
            want[idx] = min(want[idx] + v, 32767)

        split = 0
        for mby in range(mbh):
            for i in range(mbw):
                k = mby * mbw + i
                if not (lc[k] >> 14) & 1 or not intra[k]:
                    continue
                inter = min(intra[k], lc[k] & 16383)
                amount = f32(0) + f32(intra[k] * invq[k]) * fps_factor
                amount = min(int(amount * f32(intra[k] - inter) / f32(intra[k]) + f32(0.5)), 32767)
                x, y = int(mvs[k][0]), int(mvs[k][1])
                if not (x or y):
                    add(k, amount)
                    continue
                mbx, mby2 = (x >> 5) + i, (y >> 5) + mby
                x &= 31
                y &= 31
                ws = [(32 - y) * (32 - x), (32 - y) * x, y * (32 - x), y * x]
                ws = [(v * amount + 512) >> 10 for v in ws]
                split += all(ws)
                for (dx, dy), v in zip(((0, 0), (1, 0), (0, 1), (1, 1)), ws):
                    if 0 <= mbx + dx < mbw and 0 <= mby2 + dy < mbh:
                        add((mby2 + dy) * mbw + mbx + dx, v)
        assert split > mbw * mbh // 3                    # the clip exercises the 4-way split
        assert np.array_equal(got, want), np.argwhere(got != want)[:5]
        # finish: qp_offset = qp_offset_aq - 5 (1 - qcomp) (log2(intra' + propagate') - log2(intra'))
        i1, q1 = la.intra_cost(1).astype(np.int64), la.inv_qscale(1).astype(np.int64)
        qp, qa = la.qp_offset(1), la.qp_offset(1, aq=True)
        strength = 5.0 * (1.0 - float(p.qcompress))
        for k in range(mbw * mbh):
            ic = (i1[k] * q1[k] + 128) >> 8
            if ic:
                pc = (got[k] * 512 + 128) >> 8          # fps_factor of the finish = 256 / MBTREE_PRECISION
                ratio = np.log2(ic + pc) - np.log2(ic)
                assert abs(float(qp[k]) - (float(qa[k]) - strength * ratio)) <= strength * 0.0115 + 1e-4, k
    finally:
        la.close()


def test_frame_sums_are_the_sums_of_the_per_mb_results():
    """Row a13: [x264] slicetype_slice_cost / slicetype_frame_cost accumulators rebuilt from the per-MB results --
    cost estimate and intra-MB count over the interior macroblocks (the frame's border MBs are not scored), the AQ
    cost with each MB scaled by its inv_qscale (+128 >> 8), row sums over all macroblocks."""
    import numpy as np
    import oracle_lib as ol
    from clipgen import SyntheticClip
    w, h = 160, 112
    clip = SyntheticClip(w, h, n_frames=2, cuts=(), flash=None)
    la = ol.OracleLookahead(ol.la_params("medium", w, h))
    try:
        for i in range(2):
            la.put_i420(ol.oracle_convert(clip.packed(i, "bgra"), 9 | 0x1000, 2, 2, 0, w, h))
        score = la.frame_cost(0, 1, 1)
        mbw, mbh = la.g["mb_w"], la.g["mb_h"]
        lc = la.lowres_costs(1, 1, 0).astype(np.int64).reshape(mbh, mbw)
        invq = la.inv_qscale(1).astype(np.int64).reshape(mbh, mbw)
        cost, used = lc & 16383, lc >> 14
        cost_aq = (cost * invq + 128) >> 8
        inner = (slice(1, mbh - 1), slice(1, mbw - 1))
        assert la.cost_est(1, 1, 0) == int(cost[inner].sum()) == score
        assert la.cost_est(1, 1, 0, aq=True) == int(cost_aq[inner].sum())
        assert la.intra_mbs(1, 1) == int((used[inner] == 0).sum())
        assert np.array_equal(la.row_satds(1, 1, 0), cost_aq.sum(axis=1))
        assert 0 < (used == 0).sum() < mbw * mbh or (used == 1).all()
    finally:
        la.close()


def test_weightp_fake_is_analysed_and_feeds_the_tree_finish():
    """[x264] validate_parameters turns weightp 0 into X264_WEIGHTP_FAKE when mb-tree and psy are on (tune
    fastdecode on the default presets): the lookahead still finds a weight on a fade, searches with it, and
    macroblock_tree_finish lowers log2_ratio by 1 - minscore/origscore.  With psy off there is no analysis at all."""
    import numpy as np
    import oracle_lib as ol
    from clipgen import SyntheticClip
    w, h, n = 128, 96, 16
    clip = SyntheticClip(w, h, n_frames=n, cuts=(), flash=None)
    base = [ol.oracle_convert(clip.packed(i, "bgra"), 9 | 0x1000, 2, 2, 0, w, h) for i in range(n)]
    fade = [((f.astype(np.int32) * (256 - 12 * i)) >> 8).astype(np.uint8) for i, f in enumerate(base)]

    def run(**over):
        la = ol.OracleLookahead(ol.la_params("medium", w, h, rc_lookahead=8, keyint_max=250, keyint_min=5, bframes=0, **over))
        out = []
        try:
            for f in fade:
                la.put_i420(f)
                out += la.decisions()
            weights = [la.weight(i)["on"] for i in range(n)]
            la.flush()
            out += la.decisions()
        finally:
            la.close()
        return {d["i_frame"]: d for d in out}, weights

    fake, w_fake = run(weightp=0)
    real, w_real = run(weightp=2)
    none, w_none = run(weightp=0, b_psy=0)
    assert sum(w_fake) > 0 and w_fake == w_real            # the same analysis as with real weighted prediction
    assert sum(w_none) == 0
    # fake and real share the costs (the weight is used in the lookahead either way) ...
    assert [fake[i]["i_cost_est"] for i in range(n)] == [real[i]["i_cost_est"] for i in range(n)]
    # ... but only the fake mode hands the gain to the tree finish: offsets differ from the real mode's on weighted frames
    differ = [i for i in range(n) if not np.array_equal(fake[i]["qp_offset"], real[i]["qp_offset"])]
    assert differ and all(w_fake[i] for i in differ), (differ, w_fake)
    # log2_ratio grows by weightdelta = 1 - minscore/origscore > 0 on every MB: the offsets drop by 5 (1 - qcomp) weightdelta
    i = differ[0]
    drop = real[i]["qp_offset"] - fake[i]["qp_offset"]
    assert (drop > 0).all() and np.allclose(drop, drop[0], atol=1e-5) and 0 < drop[0] < 2.0 * 0.5
