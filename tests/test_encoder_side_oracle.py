"""SURVEY 8(f) row 3, remainder -- CPU checks of the checker's encoder-side functions ([x264] x264_weights_analyse(.., 0) and the
integral image).  libx264 is not in the tree (PARITY UNPINNED): these tests hold the C restatement against independent numpy
formulations of its pieces and against behaviour the algorithm must show."""
import numpy as np
import pytest

import oracle_lib as ol


# ---- integral image ---------------------------------------------------------------------------------------------------
def box_sums(plane, n):
    p = plane.astype(np.int64)
    c = np.zeros((p.shape[0] + 1, p.shape[1] + 1), np.int64)
    c[1:, 1:] = p.cumsum(0).cumsum(1)
    return (c[n:, n:] - c[:-n, n:] - c[n:, :-n] + c[:-n, :-n])          # (rows-n+1, cols-n+1)


@pytest.mark.parametrize("rows,stride", [(8, 16), (40, 64), (72, 136), (200, 320)])
def test_integral_image_is_the_box_sum_of_every_window(rows, stride):
    rng = np.random.default_rng(rows * stride)
    plane = rng.integers(0, 256, (rows, stride), dtype=np.uint8)
    plane[: rows // 3] = 255                                           # upstream's running sums wrap in uint16 here
    s8, s4 = ol.oracle_integral_init(plane)
    b8, b4 = box_sums(plane, 8), box_sums(plane, 4)
    assert (s8[: rows - 7, : stride - 8] == (b8[:, : stride - 8] & 0xffff)).all()
    assert (s4[: rows - 3, : stride - 4] == (b4[:, : stride - 4] & 0xffff)).all()
    s8b, none = ol.oracle_integral_init(plane, with_sum4=False)
    assert none is None and (s8b == s8).all()


# ---- encoder-side weight analysis -------------------------------------------------------------------------------------
def fade_frames(w, h, n, step, chroma_step=0, offset=0):
    """I420 frames of a moving synthetic scene whose luma (and optionally chroma saturation) fades by `step` / 256 per frame."""
    from clipgen import SyntheticClip
    clip = SyntheticClip(w, h, n_frames=n, cuts=(), flash=None)
    out = []
    for i in range(n):
        f = ol.oracle_convert(clip.packed(i, "bgra"), 9 | 0x1000, 2, 2, 0, w, h).astype(np.int32)
        y, c = f[: w * h], f[w * h:]
        y = np.clip(((y * (256 - step * i)) >> 8) + offset * i, 0, 255)
        c = np.clip((((c - 128) * (256 - chroma_step * i)) >> 8) + 128, 0, 255)
        out.append(np.concatenate([y, c]).astype(np.uint8))
    return out


def nv12(frame, w, h):
    u = frame[w * h: w * h + (w // 2) * (h // 2)].reshape(h // 2, w // 2)
    v = frame[w * h + (w // 2) * (h // 2):].reshape(h // 2, w // 2)
    return ol.oracle_chroma_nv12_pad(np.ascontiguousarray(u), np.ascontiguousarray(v), w, h)


def session(frames, w, h, **over):
    orc = ol.OracleLookahead(ol.la_params("medium", w, h, **over))
    for f in frames:
        orc.put_i420(f)
    return orc


def mc_chroma_numpy(uv, cw, ch, bx, by, mvx, mvy, comp):
    """[x264] mc_chroma of one 8x8 block, straight from H.264 8.4.2.2.2 (bilinear, eighth-sample), edge pairs replicated."""
    xs = np.clip(bx + (mvx >> 3) + np.arange(9), 0, cw - 1)
    ys = np.clip(by + (mvy >> 3) + np.arange(9), 0, ch - 1)
    blk = uv[np.ix_(ys, 2 * xs + comp)].astype(np.int64)
    dx, dy = mvx & 7, mvy & 7
    return ((8 - dx) * (8 - dy) * blk[:8, :8] + dx * (8 - dy) * blk[:8, 1:] + (8 - dx) * dy * blk[1:, :8] + dx * dy * blk[1:, 1:] + 32) >> 6


def weight_px(p, scale, denom, offset):
    if denom >= 1:
        return np.clip(((p * scale + (1 << (denom - 1))) >> denom) + offset, 0, 255)
    return np.clip(p * scale + offset, 0, 255)


def bits_ue(v):
    return 2 * int(np.floor(np.log2(v + 1))) + 1


def bits_se(v):
    return bits_ue(2 * v - 1 if v > 0 else -2 * v) if v else 1


def test_chroma_score_against_a_numpy_formulation():
    """weight_cost_chroma (mc_chroma with the MB's lowres vector, asd8 per 8x8 block, header bits x 4) recomputed in numpy, with
    and without a weight, for a (frame, reference) pair the lookahead has searched."""
    w, h = 160, 96
    frames = fade_frames(w, h, 4, 25, chroma_step=20)
    orc = session(frames, w, h, rc_lookahead=10)
    try:
        orc.frame_cost(0, 2, 2)                                        # runs the list-0 search of frame 2 at distance 2
        uv = [nv12(f, w, h) for f in frames]
        g = ol.lowres_geometry(w, h)
        stride, cw, ch = g["luma_w"], g["luma_w"] // 2, g["luma_h"] // 2
        mvs = orc.mvs(2, 0, 2)
        assert np.abs(mvs).sum() > 0
        fe, rf = uv[2].reshape(ch, stride), uv[0].reshape(ch, stride)
        for plane in (1, 2):
            for weight in (None, (57, 6, 3), (100, 7, -2), (31, 5, 0)):
                cost = 0
                for i, (mvx, mvy) in enumerate(mvs):
                    bx, by = 8 * (i % g["mb_w"]), 8 * (i // g["mb_w"])
                    blk = mc_chroma_numpy(rf, cw, ch, bx, by, int(mvx), int(mvy), plane - 1)
                    if weight:
                        blk = weight_px(blk, *weight)
                    src = fe[by:by + 8, 2 * bx + plane - 1: 2 * bx + 16: 2].astype(np.int64)
                    cost += abs(int(blk.sum() - src.sum()))
                if weight:
                    cost += 4 * (10 + bits_ue(weight[1]) + 2 * (bits_se(weight[0]) + bits_se(weight[2])))
                assert orc.weights_full_cost(2, 0, uv[2], uv[0], stride, plane, weight) == cost, (plane, weight)
    finally:
        orc.close()


def test_luma_score_without_vectors_is_the_lookahead_score():
    """A pair the lookahead never searched (upstream's 0x7FFF sentinel): weight_cost_init_luma returns the plain lowres plane, so the
    unweighted luma score is the lookahead's own weight_cost_luma -- the intra-capped SATD sum against the uncompensated reference."""
    w, h = 160, 96
    frames = fade_frames(w, h, 3, 30)
    orc = session(frames, w, h, rc_lookahead=10)
    try:
        uv = [nv12(f, w, h) for f in frames]
        g = ol.lowres_geometry(w, h)
        a = orc.weights_full_cost(2, 1, uv[2], uv[1], g["luma_w"], 0, None)
        plane0 = lambda f: orc.lowres_planes(f)[: g["lplane_bytes"]].reshape(-1, g["lstride"])     # padded plane 0, origin at (32, 32)
        lr = plane0(2), plane0(1)
        intra = orc.intra_cost(2)
        o = ol.oracle()
        cost = 0
        for i in range(orc.mb_count):
            x, y = 8 * (i % g["mb_w"]) + 32, 8 * (i // g["mb_w"]) + 32
            blk_f = np.ascontiguousarray(lr[0][y:y + 8, x:x + 8]); blk_r = np.ascontiguousarray(lr[1][y:y + 8, x:x + 8])
            cost += min(int(o.orc_test_satd_8x8(blk_r.ctypes.data, blk_f.ctypes.data)), int(intra[i]))
        assert a == cost
    finally:
        orc.close()


def test_fade_gets_luma_and_chroma_weights_and_a_still_scene_none():
    w, h = 320, 192
    frames = fade_frames(w, h, 4, 28, chroma_step=24)
    orc = session(frames, w, h, rc_lookahead=10)
    try:
        uv = [nv12(f, w, h) for f in frames]
        stride = ol.lowres_geometry(w, h)["luma_w"]
        orc.frame_cost(1, 2, 2)
        wts, delta = orc.weights_full(2, 1, uv[2], uv[1], stride)
        assert wts[0][0] == 1 and wts[0][1] < (1 << wts[0][2])          # luma scale below one: the picture gets darker
        assert wts[1][0] == 1 and wts[2][0] == 1 and wts[1][2] == wts[2][2]      # both chroma planes, one denominator
        assert wts[1][1] < (1 << wts[1][2]) and wts[2][1] < (1 << wts[2][2])
        assert delta == 0.0                                             # only X264_WEIGHTP_FAKE records it
        # no fade at all: nothing to find
        same = session([frames[0], frames[0]], w, h, rc_lookahead=10)
        wts2, _ = same.weights_full(1, 0, uv[0], uv[0], stride)
        same.close()
        assert wts2 == [[0, 1, 0, 0]] * 3
    finally:
        orc.close()


def test_search_window_grows_with_subme():
    """weight_check_distance: subme 2 scores 1 scale x up to 3 offsets per plane, subme 7 up to 3 x 3, subme 11 up to 9 x 5 -- seen in
    the number of luma block comparisons the analysis performs (early exits only ever shorten an offset row)."""
    w, h = 320, 192
    frames = fade_frames(w, h, 3, 33, offset=3)
    uv = [nv12(f, w, h) for f in frames]
    stride = ol.lowres_geometry(w, h)["luma_w"]
    evals = {}
    for subme in (2, 7, 11):
        orc = session(frames, w, h, rc_lookahead=10, subme=subme)
        try:
            orc.frame_cost(1, 2, 2)
            before = orc.counters()["satd"]
            wts, _ = orc.weights_full(2, 1, uv[2], uv[1], stride)
            evals[subme] = (orc.counters()["satd"] - before) // orc.mb_count      # luma scores (chroma uses asd8, not SATD)
            assert wts[0][0] == 1
        finally:
            orc.close()
    assert 2 <= evals[2] <= 1 + 3 and evals[2] < evals[7] <= 1 + 9 and evals[7] < evals[11] <= 1 + 45, evals


def test_weights_full_refuses_other_chroma_formats_and_distances():
    w, h = 64, 48
    frames = fade_frames(w, h, 2, 30)
    orc = session(frames, w, h, rc_lookahead=10)
    uv = nv12(frames[0], w, h)
    assert orc.weights_full(0, 1, uv, uv, 64) is None                   # reference must precede the frame
    orc.close()


# ---- the integral kernel's source on the CPU (lockstep warps, tests/sim/integral_sim.cpp) ------------------------------
@pytest.fixture(scope="module")
def integral_sim():
    import ctypes as C
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "sim", "_build", "libintegral_sim.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    srcs = [os.path.join(here, "sim", "integral_sim.cpp"), os.path.join(here, "sim", "warp_sim.h"),
            os.path.join(os.path.dirname(here), "x264vfw_b200", "csrc", "integral_kernel.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so, srcs[0], "-lpthread"], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.sim_integral.restype = C.c_int
    lib.sim_integral.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int]
    return lib


@pytest.mark.parametrize("rows,stride,with4", [(8, 16, True), (40, 64, False), (72, 136, True), (150, 260, True), (71, 128, False)])
def test_integral_kernel_source_on_cpu_matches_checker(integral_sim, rows, stride, with4):
    """x264vfw_b200/csrc/integral_kernel.cuh compiled by g++, every warp of the grid run as 32 threads in lockstep (shuffles as
    exchanges): 8x8 / 4x4 box sums of two planes, and nothing written outside the valid domain."""
    rng = np.random.default_rng(rows * 7 + stride)
    nf = 2
    planes = rng.integers(0, 256, (nf, rows, stride), dtype=np.uint8)
    planes[0, : rows // 2] = 255
    s8 = np.full((nf, rows, stride), 0x5a5a, np.uint16)
    s4 = np.full((nf, rows, stride), 0x5a5a, np.uint16)
    warps = integral_sim.sim_integral(s8.ctypes.data, s4.ctypes.data if with4 else None, planes.ctypes.data, stride, rows, rows * stride, rows * stride, nf)
    assert warps > 0
    for f in range(nf):
        w8, w4 = ol.oracle_integral_init(planes[f], with_sum4=with4)
        assert (s8[f, : rows - 7, : stride - 8] == w8[: rows - 7, : stride - 8]).all()
        assert (s8[f, rows - 7:] == 0x5a5a).all() and (s8[f, :, stride - 8:] == 0x5a5a).all()
        if with4:
            assert (s4[f, : rows - 3, : stride - 4] == w4[: rows - 3, : stride - 4]).all()
            assert (s4[f, rows - 3:] == 0x5a5a).all() and (s4[f, :, stride - 4:] == 0x5a5a).all()
        else:
            assert (s4[f] == 0x5a5a).all()
