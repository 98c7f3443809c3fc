"""GPU parity of the B3 entry points (upstream libx264's function-table shapes on device pointers), each through the
C ABI against the checker's pieces: frame_init_lowres_core, mbtree_propagate_cost / _list, sad / satd / sad_x3 / sad_x4
8x8, intra_mbcmp_x3_8x8c, and the x264_opencl_*-named hooks on a cost-engine session."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture()
def ctx():
    from x264vfw_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("w,h", [(64, 48), (330, 186), (1920, 1080)])
def test_frame_init_lowres_core(ctx, w, h):
    """Upstream's calling convention: the caller duplicates the last column / row of the mod-16 frame, the function
    writes lw x lh pixels of four planes, no border.  Checked against the interior of the lowres checker."""
    import torch
    from x264vfw_b200 import b3
    rng = np.random.default_rng(w + h)
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    g = ol.lowres_geometry(w, h)
    want = ol.oracle_lowres_init(y, w, h).reshape(4, g["lh"] + 64, g["lstride"])[:, 32:32 + g["lh"], 32:32 + g["lw"]]
    padded = ol.oracle_luma_pad(y, w, h).reshape(g["luma_h"] + 1, g["luma_stride"])      # mod-16 frame + duplicated row / column
    src = dev(padded)
    out = torch.zeros((4, g["lh"], g["lw"]), dtype=torch.uint8, device="cuda")
    plane = g["lh"] * g["lw"]
    b3.frame_init_lowres_core(ctx, src.data_ptr(), out.data_ptr(), out.data_ptr() + plane, out.data_ptr() + 2 * plane, out.data_ptr() + 3 * plane,
                              g["luma_stride"], g["lw"], g["lw"], g["lh"])
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), want)


def test_mbtree_propagate_cost_and_list(ctx):
    """One mb-tree step -- the unreferenced B-frame 1 between frames 0 and 2 -- rebuilt from the two function-table
    entries, row by row like upstream's macroblock_tree_propagate, on the checker's inputs (intra costs, inverse
    qscale, lowres costs, both MV fields).  Amounts against the float32 formula, the scatter against a Python
    formulation of mbtree_propagate_list."""
    import torch
    from clipgen import SyntheticClip
    from x264vfw_b200 import b3
    w, h, n = 320, 192, 3
    clip = SyntheticClip(w, h, n_frames=n, cuts=(), flash=None)
    orc = ol.OracleLookahead(ol.la_params("medium", w, h, rc_lookahead=250))
    try:
        for i in range(n):
            orc.put_i420(ol.oracle_convert(clip.packed(i, "bgra"), 9 | 0x1000, 2, 2, 0, w, h))
        orc.mbtree([0, 1, 2], [1, 5, 3], 1)                       # types I B P, keyframe walk: evaluates (0,2,2) and (0,2,1)
        intra, invq, lc = orc.intra_cost(1), orc.inv_qscale(1), orc.lowres_costs(1, 1, 1)
        mv0, mv1 = orc.mvs(1, 0, 1), orc.mvs(1, 1, 1)
    finally:
        orc.close()
    mb_w, mb_h = (w + 15) >> 4, (h + 15) >> 4
    fps = np.float32(1.0 / 512.0)                                 # CFR: clip(dur) / (clip(avg) * 256) * MBTREE_PRECISION
    bipred = 64 - ((((1 << 8) + 1) // 2) >> 2)                    # weightb, b halfway between p0 and p1: 32
    d_intra, d_invq, d_lc = dev(intra), dev(invq), dev(lc)
    d_zero = dev(np.zeros(mb_w * mb_h, dtype=np.uint16))
    d_amount = torch.zeros(mb_w * mb_h, dtype=torch.int16, device="cuda")
    b3.mbtree_propagate_cost(ctx, d_amount.data_ptr(), d_zero.data_ptr(), d_intra.data_ptr(), d_lc.data_ptr(), d_invq.data_ptr(), float(fps), mb_w * mb_h)
    ctx.sync()
    amount = d_amount.cpu().numpy().astype(np.int64)
    ic = intra.astype(np.int64)
    inter = np.minimum(ic, lc.astype(np.int64) & 0x3fff)
    pa = (ic * invq.astype(np.int64)).astype(np.float32) * fps
    want_amount = np.minimum(((pa * (ic - inter).astype(np.float32)) / ic.astype(np.float32) + np.float32(0.5)).astype(np.int64), 32767)
    assert np.array_equal(amount, want_amount) and amount.any()
    ref = [torch.zeros(mb_w * mb_h, dtype=torch.int16, device="cuda") for _ in range(2)]
    d_mv = [dev(mv0), dev(mv1)]
    for my in range(mb_h):
        for lst in range(2):
            b3.mbtree_propagate_list(ctx, ref[lst].data_ptr(), d_mv[lst].data_ptr() + 4 * my * mb_w, d_amount.data_ptr() + 2 * my * mb_w,
                                     d_lc.data_ptr() + 2 * my * mb_w, bipred if lst == 0 else 64 - bipred, my, mb_w, lst, mb_w, mb_h)
    ctx.sync()
    got = [r.cpu().numpy().view(np.uint16).astype(np.int64) for r in ref]
    exp = [np.zeros(mb_w * mb_h, dtype=np.int64) for _ in range(2)]
    for lst, mv in enumerate((mv0, mv1)):
        bw = bipred if lst == 0 else 64 - bipred
        for idx in range(mb_w * mb_h):
            used = int(lc[idx]) >> 14
            if not used & (1 << lst):
                continue
            la = int(amount[idx])
            if used == 3:
                la = (la * bw + 32) >> 6
            i, my = idx % mb_w, idx // mb_w
            x, y = int(mv[idx][0]), int(mv[idx][1])
            if not (x | y):
                exp[lst][idx] = min(exp[lst][idx] + la, 32767)
                continue
            mbx, mby = (x >> 5) + i, (y >> 5) + my
            fx, fy = x & 31, y & 31
            for dx, dy, wgt in ((0, 0, (32 - fy) * (32 - fx)), (1, 0, (32 - fy) * fx), (0, 1, fy * (32 - fx)), (1, 1, fy * fx)):
                xx, yy = mbx + dx, mby + dy
                if 0 <= xx < mb_w and 0 <= yy < mb_h:
                    k = xx + yy * mb_w
                    exp[lst][k] = min(exp[lst][k] + ((wgt * la + 512) >> 10), 32767)
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
    assert exp[0].any() and exp[1].any()


def test_pixel_sad_satd_and_xn(ctx):
    import torch
    from x264vfw_b200 import b3
    o = ol.oracle()
    rng = np.random.default_rng(7)
    h, w = 96, 160
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    b = rng.integers(0, 256, (h, w), dtype=np.uint8)
    b[:48] = np.clip(a[:48].astype(np.int32) + rng.integers(-3, 4, (48, w)), 0, 255).astype(np.uint8)     # small residuals too
    n = 300
    ys, xs = rng.integers(0, h - 8, n), rng.integers(0, w - 8, n)
    ys2, xs2 = rng.integers(0, h - 8, (n, 4)), rng.integers(0, w - 8, (n, 4))
    off1 = (ys * w + xs).astype(np.int32)
    off2 = (ys2[:, 0] * w + xs2[:, 0]).astype(np.int32)
    offr = (ys2 * w + xs2).astype(np.int32)
    da, db, d1, d2, dr = dev(a), dev(b), dev(off1), dev(off2), dev(offr)
    def score(fn, p, y, x, q, y2, x2):
        b1, b2 = np.ascontiguousarray(p[y:y + 8, x:x + 8]), np.ascontiguousarray(q[y2:y2 + 8, x2:x2 + 8])      # both stay alive for the call
        return fn(b1.ctypes.data, b2.ctypes.data)

    for satd in (0, 1):
        sc = torch.zeros(n, dtype=torch.int32, device="cuda")
        b3.pixel_cmp_8x8(ctx, satd, da.data_ptr(), w, db.data_ptr(), w, d1.data_ptr(), d2.data_ptr(), sc.data_ptr(), n)
        ctx.sync()
        fn = o.orc_test_satd_8x8 if satd else o.orc_test_sad_8x8
        want = [score(fn, a, ys[i], xs[i], b, ys2[i, 0], xs2[i, 0]) for i in range(n)]
        assert sc.cpu().numpy().tolist() == want, satd
    for nref in (3, 4):
        sc = torch.zeros(n * nref, dtype=torch.int32, device="cuda")
        dro = dev(np.ascontiguousarray(offr[:, :nref]))
        b3.pixel_sad_xn_8x8(ctx, nref, da.data_ptr(), w, d1.data_ptr(), db.data_ptr(), w, dro.data_ptr(), sc.data_ptr(), n)
        ctx.sync()
        want = [score(o.orc_test_sad_8x8, a, ys[i], xs[i], b, ys2[i, k], xs2[i, k]) for i in range(n) for k in range(nref)]
        assert sc.cpu().numpy().tolist() == want, nref


def test_intra_mbcmp_x3_8x8c(ctx):
    """predict_8x8c_{dc,h,v} (pinned to the H.264 decoder in tests/test_h264_pins.py) scored with the checker's SAD / SATD."""
    import torch
    from x264vfw_b200 import b3
    o = ol.oracle()
    rng = np.random.default_rng(11)
    mb_w, mb_h, stride = 9, 5, 128
    plane = rng.integers(0, 256, ((mb_h * 8 + 16), stride), dtype=np.uint8)
    org = 8 * stride + 16                                                  # block (0,0) has a row above and a column to its left
    d = dev(plane)
    for satd in (0, 1):
        res = torch.zeros(mb_w * mb_h * 3, dtype=torch.int32, device="cuda")
        b3.intra_mbcmp_x3_8x8c(ctx, satd, d.data_ptr() + org, stride, mb_w, mb_h, res.data_ptr())
        ctx.sync()
        got = res.cpu().numpy().reshape(-1, 3)
        fn = o.orc_test_satd_8x8 if satd else o.orc_test_sad_8x8
        flat = plane.reshape(-1)
        for my in range(mb_h):
            for mx in range(mb_w):
                p = org + 8 * (mx + my * stride)
                src = np.ascontiguousarray(plane[8 + 8 * my:16 + 8 * my, 16 + 8 * mx:24 + 8 * mx])
                for k in range(3):                                            # kinds 0, 1, 2 = dc, h, v
                    pred = np.zeros(64, dtype=np.uint8)
                    o.orc_test_intra_pred_8x8(pred.ctypes.data, k, flat.ctypes.data + p, stride)
                    assert got[mx + my * mb_w][k] == fn(src.ctypes.data, pred.ctypes.data), (satd, mx, my, k)


def test_opencl_named_hooks_on_a_cost_engine_session():
    """lowres_init / slicetype_prep / motionsearch / finalize_cost / flush drive a cost-engine session (keep_frames,
    maximum lookahead: it never decides) to the checker's frame costs, whatever the order the searches were enqueued in."""
    from clipgen import SyntheticClip
    from x264vfw_b200 import lookahead, b3
    w, h, n = 320, 192, 7
    clip = SyntheticClip(w, h, n_frames=n, cuts=(4,), flash=None)
    packed = [clip.packed(i, "bgra") for i in range(n)]
    orc = ol.OracleLookahead(ol.la_params("medium", w, h, rc_lookahead=250))
    gpu = lookahead.Lookahead(lookahead.params_preset("medium", w, h, rc_lookahead=250), in_csp=9 | 0x1000, device=0, keep_frames=True)
    hooks = b3.OpenclHooks(gpu)
    try:
        for i, f in enumerate(packed):
            orc.put_i420(ol.oracle_convert(f, 9 | 0x1000, 2, 2, 0, w, h))
            assert hooks.lowres_init(f) == i
        hooks.slicetype_prep(0, n - 1)
        hooks.motionsearch(5, 1, 0)                               # distance 4, list 0: not covered by prep (bframes = 3)
        hooks.flush()
        for (p0, p1, b) in [(0, 0, 0), (0, 1, 1), (0, 2, 2), (0, 2, 1), (1, 5, 5), (1, 5, 3), (3, 6, 4), (4, 5, 5), (6, 6, 6)]:
            score, (ce, ce_aq, imb) = hooks.finalize_cost(p0, p1, b)
            assert score == orc.frame_cost(p0, p1, b), (p0, p1, b)
            assert (ce, ce_aq) == (orc.cost_est(b, b - p0, p1 - b), orc.cost_est(b, b - p0, p1 - b, aq=True)), (p0, p1, b)
            if b == p1 and p0 != p1:
                assert imb == orc.intra_mbs(b, b - p0)
            for lst, dist in ((0, b - p0), (1, p1 - b)):
                if dist:
                    assert np.array_equal(gpu.mvs(b, lst, dist), orc.mvs(b, lst, dist)), (p0, p1, b, lst)
        hooks.slicetype_end()
        assert not gpu.decisions()                                # the cost-engine session decided nothing itself
    finally:
        orc.close(); gpu.close()
