"""SURVEY 8(f) row 3, remainder, on the device: [x264] x264_weights_analyse(h, fenc, ref, 0) (x264vfw_cuda_la_weights_analyse /
x264vfw_cuda_weights_analyse) and the integral image (x264vfw_cuda_integral_init) against the CPU checker
(PARITY UNPINNED like the rest of stage 2: the checker restates upstream libx264)."""
import numpy as np
import pytest

import oracle_lib as ol
from test_encoder_side_oracle import fade_frames, nv12

pytestmark = pytest.mark.gpu


def open_pair(w, h, **over):
    from x264vfw_b200 import lookahead
    return (ol.OracleLookahead(ol.la_params("medium", w, h, **over)),
            lookahead.Lookahead(lookahead.params_preset("medium", w, h, **over), device=0, keep_frames=True))


@pytest.mark.parametrize("size,subme,kw", [((320, 192), 7, dict(step=28, chroma_step=24)), ((320, 192), 11, dict(step=33, offset=3)),
                                           ((330, 186), 2, dict(step=20, chroma_step=30)), ((320, 192), 9, dict(step=-20, chroma_step=-25)),
                                           ((1280, 720), 7, dict(step=30, chroma_step=20)), ((320, 192), 7, dict(step=0)),
                                           ((320, 192), 1, dict(step=30, chroma_step=20)), ((320, 192), 0, dict(step=25))])
def test_weights_analyse_matches_checker(size, subme, kw):
    """Every (P frame, reference) pair of a fading clip, with the lookahead's vectors present (after slicetype_frame_cost) and absent:
    the three weights and the X264_WEIGHTP_FAKE ratio are identical."""
    import torch
    from x264vfw_b200 import b3
    w, h = size
    n = 5
    frames = fade_frames(w, h, n, **kw)
    if kw.get("step", 0) < 0:
        frames = frames[::-1]                                            # fade in
    orc, gpu = open_pair(w, h, rc_lookahead=10, subme=subme)
    try:
        for f in frames:
            orc.put_i420(f)
            gpu.put_frame(f)
        uv = [nv12(f, w, h) for f in frames]
        d_uv = [torch.from_numpy(u).cuda() for u in uv]
        stride = ol.lowres_geometry(w, h)["luma_w"]
        seen = 0
        for ref, fenc in ((0, 1), (1, 3), (2, 3), (3, 4)):               # (1, 3) / (3, 4) first without vectors ...
            for searched in (False, True):
                if searched:                                             # ... then with the lookahead's list-0 search of that distance
                    assert orc.frame_cost(ref, fenc, fenc) == gpu.frame_cost(ref, fenc, fenc)
                want = orc.weights_full(fenc, ref, uv[fenc], uv[ref], stride)
                got = b3.la_weights_analyse(gpu, fenc, ref, d_uv[fenc].data_ptr(), d_uv[ref].data_ptr(), stride)
                assert got[0] == want[0], (ref, fenc, searched, got, want)
                assert np.float32(got[1]).view(np.uint32) == np.float32(want[1]).view(np.uint32)
                seen += sum(p[0] for p in want[0])
        if kw.get("step", 0):
            assert seen > 0, "the fade did not produce a single weight"
    finally:
        orc.close(); gpu.close()


def test_weightp_fake_reports_the_cost_ratio():
    import torch
    from x264vfw_b200 import b3
    w, h = 320, 192
    frames = fade_frames(w, h, 3, 30)
    orc, gpu = open_pair(w, h, rc_lookahead=10, weightp=0, weightb=0)      # mb-tree + psy: X264_WEIGHTP_FAKE
    try:
        for f in frames:
            orc.put_i420(f); gpu.put_frame(f)
        uv = [nv12(f, w, h) for f in frames]
        d_uv = [torch.from_numpy(u).cuda() for u in uv]
        stride = ol.lowres_geometry(w, h)["luma_w"]
        want = orc.weights_full(2, 1, uv[2], uv[1], stride)
        got = b3.la_weights_analyse(gpu, 2, 1, d_uv[2].data_ptr(), d_uv[1].data_ptr(), stride)
        assert got[0] == want[0] and want[0][0][0] == 1
        assert 0 < want[1] < 0.998 and np.float32(got[1]).view(np.uint32) == np.float32(want[1]).view(np.uint32)
    finally:
        orc.close(); gpu.close()


def test_stateless_entry_on_device_buffers():
    """x264vfw_cuda_weights_analyse: the same analysis with every input handed over as a device pointer (what a libx264 patch that
    keeps its frames on the device would call)."""
    import torch
    from x264vfw_b200 import b3, lookahead
    from x264vfw_b200._lib import Context
    w, h = 320, 192
    frames = fade_frames(w, h, 3, 26, chroma_step=22)
    orc, gpu = open_pair(w, h, rc_lookahead=10)
    try:
        for f in frames:
            orc.put_i420(f); gpu.put_frame(f)
        assert orc.frame_cost(1, 2, 2) == gpu.frame_cost(1, 2, 2)
        uv = [nv12(f, w, h) for f in frames]
        g = ol.lowres_geometry(w, h)
        want = orc.weights_full(2, 1, uv[2], uv[1], g["luma_w"])
        keep = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in
                (gpu.lowres_planes(2, 4 * g["lplane_bytes"]), gpu.lowres_planes(1, 4 * g["lplane_bytes"]), gpu.mvs(2, 0, 1), gpu.intra_cost(2),
                 uv[2], uv[1])]
        win = b3.WeightsIn()
        win.width, win.height = w, h
        win.fenc_lowres, win.ref_lowres, win.lowres_mvs, win.intra_cost, win.fenc_uv, win.ref_uv = (t.data_ptr() for t in keep)
        win.uv_stride = g["luma_w"]
        fs, rs = gpu.pixel_stats(2), gpu.pixel_stats(1)
        for i in range(3):
            win.fenc_sum[i], win.fenc_ssd[i] = fs[0][i], fs[1][i]
            win.ref_sum[i], win.ref_ssd[i] = rs[0][i], rs[1][i]
        win.subme, win.weightp = 7, 2
        ctx = Context(0)
        got = b3.weights_analyse(ctx, win)
        assert got[0] == want[0] and want[0][0][0] == 1
        win.lowres_mvs = None                                            # upstream's 0x7FFF sentinel
        got2 = b3.weights_analyse(ctx, win)
        assert got2[0][0][0] in (0, 1)                                   # runs; the value is covered by the session test above
        ctx.close()
    finally:
        orc.close(); gpu.close()


@pytest.mark.parametrize("rows,stride,with4", [(8, 16, True), (72, 136, True), (200, 320, False), (1152, 1984, True)])
def test_integral_image_matches_checker(rows, stride, with4):
    import torch
    from x264vfw_b200 import b3
    from x264vfw_b200._lib import Context
    rng = np.random.default_rng(rows + stride)
    nf = 2
    planes = rng.integers(0, 256, (nf, rows, stride), dtype=np.uint8)
    planes[0, : rows // 2] = 255
    d_p = torch.from_numpy(planes).cuda()
    d_8 = torch.full((nf, rows, stride), 0x5a5a, dtype=torch.int16, device="cuda")
    d_4 = torch.full((nf, rows, stride), 0x5a5a, dtype=torch.int16, device="cuda")
    ctx = Context(0)
    b3.integral_init(ctx, d_8.data_ptr(), d_4.data_ptr() if with4 else 0, d_p.data_ptr(), stride, rows, rows * stride, rows * stride, nf)
    ctx.sync()
    g8, g4 = d_8.cpu().numpy().view(np.uint16), d_4.cpu().numpy().view(np.uint16)
    for f in range(nf):
        s8, s4 = ol.oracle_integral_init(planes[f], with_sum4=with4)
        assert (g8[f, : rows - 7, : stride - 8] == s8[: rows - 7, : stride - 8]).all()
        assert (g8[f, rows - 7:] == 0x5a5a).all() and (g8[f, :, stride - 8:] == 0x5a5a).all()      # nothing written outside the valid domain
        if with4:
            assert (g4[f, : rows - 3, : stride - 4] == s4[: rows - 3, : stride - 4]).all()
            assert (g4[f, rows - 3:] == 0x5a5a).all() and (g4[f, :, stride - 4:] == 0x5a5a).all()
        else:
            assert (g4[f] == 0x5a5a).all()
    ctx.close()


def test_integral_image_on_the_half_pel_plane_of_a_frame():
    """The way upstream uses it: over plane 0 of the padded half-pel planes of a reference frame (x264vfw_cuda_hpel_filter)."""
    import torch
    from x264vfw_b200 import b3, hpel
    from x264vfw_b200._lib import Context
    w, h = 320, 192
    y = np.random.default_rng(3).integers(0, 256, (h, w), dtype=np.uint8)
    g = hpel.geometry(w, h)
    ctx = Context(0)
    d_y = torch.from_numpy(y).cuda()
    d_hp = torch.zeros(4 * g.plane_bytes, dtype=torch.uint8, device="cuda")
    hpel.hpel_filter(ctx, d_hp.data_ptr(), d_y.data_ptr(), w, w, h, w * h, 4 * g.plane_bytes, 1)
    rows = g.plane_bytes // g.stride
    d_8 = torch.zeros(rows * g.stride, dtype=torch.int16, device="cuda")
    b3.integral_init(ctx, d_8.data_ptr(), 0, d_hp.data_ptr(), g.stride, rows)
    ctx.sync()
    plane0 = d_hp[: g.plane_bytes].cpu().numpy().reshape(rows, g.stride)
    s8, _ = ol.oracle_integral_init(plane0, with_sum4=False)
    got = d_8.cpu().numpy().view(np.uint16).reshape(rows, g.stride)
    assert (got[: rows - 7, : g.stride - 8] == s8[: rows - 7, : g.stride - 8]).all()
    # the window at the picture's origin sums the picture's first 8x8 pixels
    oy, ox = divmod(g.origin, g.stride)
    assert int(got[oy, ox]) == int(y[:8, :8].astype(np.int64).sum())
    ctx.close()
