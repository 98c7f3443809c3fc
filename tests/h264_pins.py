"""Shared by tests/golden/make_h264_pins.py (generator) and tests/test_h264_pins.py: the pictures, motion vectors
and intra test layout of the H.264-decoder pins, and the checker-side evaluation of the same predictions.

What is pinned: five restated [x264] pieces whose results the H.264 standard fixes, because an encoder's
prediction must be the decoder's -- (1) the half-pel planes (oracle/hpel_oracle.c: border expansion, 6-tap H / V /
centre), (2) get_ref's choice and averaging of two planes per quarter-sample position (oracle/lookahead_oracle.c),
(3) the ten intra predictors the lookahead scores (predict_8x8c_{dc,h,v,p}, predict_8x8_filter +
predict_8x8_{ddl,ddr,vr,hd,vl,hu}), (4) mc_weight, the explicit weighted prediction of a weighted reference,
(5) the bidirectional average: pixel_avg and the lookahead's bipred weight for every (p0, b, p1) up to 16 B-frames
(implicit weights).  The reference is libavcodec's H.264 decoder (the FFmpeg build inside this
image's opencv wheel), driven by bitstreams from tests/golden/h264_mini.py; its outputs are frozen as FNV-1a-64
hashes in tests/golden/h264_pins.json.  libx264 itself stays absent: everything else in the lookahead checker
(search order, costs, decisions, mb-tree) remains unpinned."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import oracle_lib as ol  # noqa: E402

MC_W, MC_H = 64, 48
# every quarter-sample phase at the origin, then vectors that reach up to 22 samples outside the picture (the
# decoder clamps coordinates; the planes carry a 32-sample replicated border)
MC_MVS = [(fx, fy) for fy in range(4) for fx in range(4)] + [(-37, 22), (50, -61), (3, -90), (-85, 1), (-6, -7), (83, 86)]
MC_KINDS = ("noise", "extreme")
# the same at BASELINE's frame size (1080p padded to whole macroblocks): the kernel's tiles and strips
HD_W, HD_H = 1920, 1088
HD_MVS = [(1, 1), (2, 2), (3, 1), (2, 0), (0, 2), (-61, 30)]
# explicit weighted prediction (mc_weight): (vector, (log2 denominator, scale, offset))
WP_CASES = [((0, 0), (0, 1, 5)), ((0, 0), (5, 40, -3)), ((2, 1), (6, 100, 20)), ((3, 3), (7, 127, -128)), ((1, 2), (2, 7, 127)),
            ((0, 0), (0, 2, -100)), ((-37, 22), (5, 31, 0)), ((2, 2), (7, 127, 1)), ((0, 0), (1, 1, 0)), ((1, 0), (6, 64, -1))]
# bidirectional prediction: (distance p1 - p0, position b - p0, list-0 vector, list-1 vector, weighted_bipred_idc);
# idc 2 = implicit weights (x264's weightb), 0 = plain average.  Then every (distance, position) the lookahead can
# meet with up to 16 B-frames, to pin its own distance scale against the standard's.
BI_CASES = [(2, 1, (0, 0), (0, 0), 0), (2, 1, (0, 0), (0, 0), 2), (3, 1, (2, 1), (-3, 2), 2), (4, 1, (5, -6), (1, 1), 2),
            (4, 3, (3, 3), (-9, 7), 2), (4, 3, (3, 3), (-9, 7), 0), (5, 2, (-37, 22), (50, -61), 2), (8, 7, (1, 2), (3, 0), 2)]
BI_CASES += [(d, b, (1, 0), (0, 1), 2) for d in range(2, 18) for b in range(1, d)]
INTRA_MBW, INTRA_MBH = 4, 24


def _lcg(n, w, h):
    buf = np.zeros(n, dtype=np.uint8)
    ol.oracle().orc_lcg_fill(buf.ctypes.data, buf.size, w, h)        # SURVEY A.4 generator, seeded by (w, h)
    return buf


def mc_picture(kind):
    w, h = MC_W, MC_H
    b = _lcg(w * h * 3 // 2, w + (7 if kind == "extreme" else 0), h)
    if kind == "extreme":                                   # 0 / 255 only: every tap sum reaches its limits
        b = np.where(b > 127, 255, 0).astype(np.uint8)
    return b[:w * h].reshape(h, w), b[w * h:w * h * 5 // 4].reshape(h // 2, w // 2), b[w * h * 5 // 4:].reshape(h // 2, w // 2)


def intra_picture():
    w, h = 16 * INTRA_MBW, 16 * INTRA_MBH
    b = _lcg(w * h * 3 // 2, w, h)
    return b[:w * h].reshape(h, w), b[w * h:w * h * 5 // 4].reshape(h // 2, w // 2), b[w * h * 5 // 4:].reshape(h // 2, w // 2)


def intra_tests():
    """{(mbx, mby): (Intra_8x8 luma mode 3..8, chroma mode 0..3)}: odd rows, odd columns; every neighbour I_PCM."""
    t, k = {}, 0
    for mby in range(1, INTRA_MBH, 2):
        for mbx in (1, 3):
            t[(mbx, mby)] = (3 + k % 6, (k // 2) % 4)
            k += 1
    return t


def hd_picture():
    b = _lcg(HD_W * HD_H, HD_W, HD_H)
    return b.reshape(HD_H, HD_W)


def hd_stream():
    import h264_mini as hm
    y = hd_picture()
    c = np.full((HD_H // 2, HD_W // 2), 128, dtype=np.uint8)
    aus = [hm.sps(HD_W // 16, HD_H // 16) + hm.pps() + hm.idr_pcm_picture(y, c, c)]
    for k, (mx, my) in enumerate(HD_MVS):
        aus.append(hm.p_picture_uniform_mv(HD_W // 16, HD_H // 16, mx, my, 1, 2 * (k + 1)))
    return aus


def checker_hd_hashes(planes=None):
    g = ol.hpel_geometry(HD_W, HD_H)
    if planes is None:
        planes = ol.oracle_hpel_planes(hd_picture(), HD_W, HD_H)
    return [fnv(predict_from_planes(planes, HD_W, HD_H, g["stride"], mx, my)) for mx, my in HD_MVS]


def decoder_hd_hashes():
    import avdec
    pics = avdec.decode_h264(hd_stream())
    assert len(pics) == 1 + len(HD_MVS) and np.array_equal(pics[0][0], hd_picture())
    return [fnv(p[0]) for p in pics[1:]]


def mc_stream(kind):
    import h264_mini as hm
    y, u, v = mc_picture(kind)
    aus = [hm.sps(MC_W // 16, MC_H // 16) + hm.pps() + hm.idr_pcm_picture(y, u, v)]
    for k, (mx, my) in enumerate(MC_MVS):
        aus.append(hm.p_picture_uniform_mv(MC_W // 16, MC_H // 16, mx, my, 1, 2 * (k + 1)))
    return aus


def wp_stream():
    import h264_mini as hm
    y, u, v = mc_picture("noise")
    aus = [hm.sps(MC_W // 16, MC_H // 16) + hm.pps() + hm.pps(1, 1) + hm.idr_pcm_picture(y, u, v)]
    for k, ((mx, my), wt) in enumerate(WP_CASES):
        aus.append(hm.p_picture_uniform_mv(MC_W // 16, MC_H // 16, mx, my, 1, 2 * (k + 1), weight=wt))
    return aus


def bi_stream(case):
    import h264_mini as hm
    d, b, mv0, mv1, idc = case
    ya, ua, va = mc_picture("noise")
    yb, ub, vb = mc_picture("extreme")
    mw, mh = MC_W // 16, MC_H // 16
    return [hm.sps(mw, mh, 2, 1) + hm.pps() + hm.pps(2, 0, 2) + hm.idr_pcm_picture(ya, ua, va),
            hm.p_pcm_reference_picture(yb, ub, vb, 1, 2 * d),
            hm.b_picture_uniform_mvs(mw, mh, mv0, mv1, 2, 2 * b, 2 if idc else 0)]


def intra_stream():
    import h264_mini as hm
    y, u, v = intra_picture()
    return [hm.sps(INTRA_MBW, INTRA_MBH) + hm.pps() + hm.idr_pcm_picture(y, u, v, intra_tests())]


def fnv(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return f"{ol.oracle().orc_fnv1a64(a.ctypes.data, a.size):016x}"


def predict_from_planes(planes, w, h, stride, mvx, mvy, weight=None):
    """The whole picture predicted with one quarter-sample vector: get_ref (checker's own code) per 8x8 block on
    four padded half-pel planes of shape (4, h + 64, stride) -- the checker's or the device's.  weight = (log2
    denominator, scale, offset) applies mc_weight after the interpolation, as get_ref does for a weighted reference."""
    o = ol.oracle()
    o.orc_test_predict_picture.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t] + [C.c_int] * 7
    planes = np.ascontiguousarray(planes)
    out = np.zeros((h, w), dtype=np.uint8)
    dn, sc, of = weight if weight is not None else (0, -1, 0)
    o.orc_test_predict_picture(out.ctypes.data, planes.ctypes.data, stride, planes.shape[1] * stride, w, h, mvx, mvy, sc, dn, of)
    return out


def checker_mc_hashes(kind, planes=None):
    y, _, _ = mc_picture(kind)
    g = ol.hpel_geometry(MC_W, MC_H)
    if planes is None:
        planes = ol.oracle_hpel_planes(y, MC_W, MC_H)
    return [fnv(predict_from_planes(planes, MC_W, MC_H, g["stride"], mx, my)) for mx, my in MC_MVS]


def checker_mc_chroma_hashes(kind):
    """[x264] mc_chroma (the checker's, used by the encoder-side weight analysis) on the picture's chroma planes, one hash per
    vector over U then V: H.264 8.4.2.2.2, the chroma vector is the luma vector in eighth samples."""
    _, u, v = mc_picture(kind)
    ch, cw = u.shape
    uv = np.empty((ch, 2 * cw), np.uint8)
    uv[:, 0::2], uv[:, 1::2] = u, v
    o = ol.oracle()
    o.orc_test_mc_chroma_picture.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5
    out = []
    for mx, my in MC_MVS:
        pu, pv = np.zeros((ch, cw), np.uint8), np.zeros((ch, cw), np.uint8)
        o.orc_test_mc_chroma_picture(pu.ctypes.data, pv.ctypes.data, uv.ctypes.data, 2 * cw, cw, ch, mx, my)
        out.append(fnv(np.concatenate([pu.ravel(), pv.ravel()])))
    return out


def decoder_mc_chroma_hashes(kind):
    import avdec
    pics = avdec.decode_h264(mc_stream(kind))
    return [fnv(np.concatenate([p[1].ravel(), p[2].ravel()])) for p in pics[1:]]


def checker_wp_hashes():
    y, _, _ = mc_picture("noise")
    g = ol.hpel_geometry(MC_W, MC_H)
    planes = ol.oracle_hpel_planes(y, MC_W, MC_H)
    return [fnv(predict_from_planes(planes, MC_W, MC_H, g["stride"], mx, my, wt)) for (mx, my), wt in WP_CASES]


def decoder_wp_hashes():
    import avdec
    pics = avdec.decode_h264(wp_stream())
    assert len(pics) == 1 + len(WP_CASES)
    return [fnv(p[0]) for p in pics[1:]]


def checker_bi_hashes():
    """get_ref on both references, the lookahead's bipred weight for (p0, b, p1) = (0, b, d), pixel_avg."""
    o = ol.oracle()
    g = ol.hpel_geometry(MC_W, MC_H)
    pa = ol.oracle_hpel_planes(mc_picture("noise")[0], MC_W, MC_H)
    pb = ol.oracle_hpel_planes(mc_picture("extreme")[0], MC_W, MC_H)
    cache, out = {}, []
    blk = np.zeros(64, dtype=np.uint8)
    for d, b, mv0, mv1, idc in BI_CASES:
        if (0, mv0) not in cache:
            cache[(0, mv0)] = predict_from_planes(pa, MC_W, MC_H, g["stride"], *mv0)
        if (1, mv1) not in cache:
            cache[(1, mv1)] = predict_from_planes(pb, MC_W, MC_H, g["stride"], *mv1)
        a, c = cache[(0, mv0)], cache[(1, mv1)]
        wt = o.orc_test_bipred_weight(0, d, b, 1 if idc else 0)
        pr = np.zeros((MC_H, MC_W), dtype=np.uint8)
        for by in range(0, MC_H, 8):
            for bx in range(0, MC_W, 8):
                x = np.ascontiguousarray(a[by:by + 8, bx:bx + 8]).reshape(-1)
                z = np.ascontiguousarray(c[by:by + 8, bx:bx + 8]).reshape(-1)
                o.orc_test_pixel_avg_8x8(blk.ctypes.data, x.ctypes.data, z.ctypes.data, wt)
                pr[by:by + 8, bx:bx + 8] = blk.reshape(8, 8)
        out.append(fnv(pr))
    return out


def decoder_bi_hashes(cases=None):
    import avdec
    ya, yb = mc_picture("noise")[0], mc_picture("extreme")[0]
    out = []
    for case in (BI_CASES if cases is None else cases):
        pics = avdec.decode_h264(bi_stream(case))        # output order: the IDR, the B picture, the later reference
        assert len(pics) == 3 and np.array_equal(pics[0][0], ya) and np.array_equal(pics[2][0], yb), case
        out.append(fnv(pics[1][0]))
    return out


def checker_intra_hashes():
    """Per test macroblock: hashes of the predicted top-left luma 8x8 and of the two chroma 8x8 blocks, from the
    SOURCE picture's neighbours (I_PCM neighbours decode to the source)."""
    o = ol.oracle()
    y, u, v = [np.ascontiguousarray(p) for p in intra_picture()]
    w = y.shape[1]
    blk = np.zeros(64, dtype=np.uint8)
    out = []
    for (mbx, mby), (m, c) in sorted(intra_tests().items()):
        o.orc_test_intra_pred_8x8(blk.ctypes.data, 10 + m, C.c_void_p(y.ctypes.data + 16 * mby * w + 16 * mbx), w)
        hy = fnv(blk)
        hc = []
        for pl in (u, v):
            o.orc_test_intra_pred_8x8(blk.ctypes.data, c, C.c_void_p(pl.ctypes.data + 8 * mby * (w // 2) + 8 * mbx), w // 2)
            hc.append(fnv(blk))
        out.append({"mb": [mbx, mby], "luma_mode": m, "chroma_mode": c, "luma": hy, "u": hc[0], "v": hc[1]})
    return out


def decoder_mc_hashes(kind):
    import avdec
    pics = avdec.decode_h264(mc_stream(kind))
    y, _, _ = mc_picture(kind)
    assert len(pics) == 1 + len(MC_MVS) and np.array_equal(pics[0][0], y)
    return [fnv(p[0]) for p in pics[1:]]


def decoder_intra_hashes():
    import avdec
    (Y, U, V), = avdec.decode_h264(intra_stream())
    out = []
    for (mbx, mby), (m, c) in sorted(intra_tests().items()):
        out.append({"mb": [mbx, mby], "luma_mode": m, "chroma_mode": c,
                    "luma": fnv(Y[16 * mby:16 * mby + 8, 16 * mbx:16 * mbx + 8]),
                    "u": fnv(U[8 * mby:8 * mby + 8, 8 * mbx:8 * mbx + 8]), "v": fnv(V[8 * mby:8 * mby + 8, 8 * mbx:8 * mbx + 8])})
    return out
