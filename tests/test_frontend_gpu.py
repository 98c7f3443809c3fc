"""GPU parity of the fused front end (frontend_kernel.cu: TMA-staged tiles of the packed rows -> I420 planes +
adaptive-quant arrays + the four lowres planes in ONE kernel) against the three things it replaces, each checked
with its own oracle: the reference csp.c object / port (planes), the lowres checker (planes incl. border), the
lookahead checker's x264_adaptive_quant_frame (per-MB offsets, inverse qscale, frame sums)."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
BGRA = 9
FLIP = 0x1000


def run_frontend(frames, in_csp, w, h, aq_strength=1.0):
    import torch
    from x264vfw_b200 import csp, lowres
    from x264vfw_b200._lib import Context
    n = len(frames)
    ctx = Context(0)
    try:
        g = lowres.geometry(w, h)
        mb = g.mb_w * g.mb_h
        sfb, dfb = csp.frame_bytes(in_csp, 2, w, h)
        src = torch.zeros(n * sfb, dtype=torch.uint8, device="cuda")
        for i, f in enumerate(frames):
            src[i * sfb:i * sfb + f.size] = torch.from_numpy(f).cuda()
        dst = torch.zeros(n * dfb, dtype=torch.uint8, device="cuda")
        lr = torch.zeros(n * 4 * g.lplane_bytes, dtype=torch.uint8, device="cuda")
        qp = torch.zeros(n * mb, dtype=torch.float32, device="cuda")
        invq = torch.zeros(n * mb, dtype=torch.int16, device="cuda")
        stats = torch.zeros(n * 6, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        lowres.frontend_batch(ctx, src.data_ptr(), dst.data_ptr(), lr.data_ptr(), qp.data_ptr(), invq.data_ptr(), stats.data_ptr(),
                              in_csp, w, h, n, sfb, dfb, aq_strength=aq_strength)
        ctx.sync()
        return (dst.cpu().numpy().reshape(n, dfb)[:, :w * h * 3 // 2], lr.cpu().numpy().reshape(n, 4, g.lh + 64, g.lstride),
                qp.cpu().numpy().reshape(n, mb), invq.cpu().numpy().view(np.uint16).reshape(n, mb), stats.cpu().numpy().reshape(n, 6), g)
    finally:
        ctx.close()


@pytest.mark.parametrize("w,h,flip", [(1920, 1080, True), (1920, 1080, False), (1280, 720, True), (64, 48, True), (16, 16, True), (16, 2, False),
                                      (336, 190, True), (336, 190, False), (144, 34, True), (128, 32, True), (256, 66, False), (3840, 2160, True)])
def test_fused_front_end_matches_the_three_oracles(w, h, flip):
    from clipgen import SyntheticClip
    in_csp = BGRA | (FLIP if flip else 0)
    n = 2 if w * h > 1e6 else 3
    clip = SyntheticClip(w, h, n_frames=n + 1, cuts=(1,), flash=None)
    frames = [clip.packed(i, "bgra") for i in range(n)]
    rng = np.random.default_rng(w * 31 + h)
    frames[-1] = rng.integers(0, 256, frames[-1].size, dtype=np.uint8)            # noise: every byte matters
    if n > 2:
        frames[1] = np.where(rng.random(frames[1].size) < 0.5, 0, 255).astype(np.uint8)   # saturation
    planes, lr, qp, invq, stats, g = run_frontend(frames, in_csp, w, h)
    conv = ol.ref_convert if ol.have_ref_csp() else ol.oracle_convert
    orc = ol.OracleLookahead(ol.la_params("medium", w, h, rc_lookahead=10))
    try:
        for i, f in enumerate(frames):
            want = conv(f, in_csp, 2, 2, 0, w, h)
            assert np.array_equal(planes[i], want), ("planes", i)
            lw = ol.oracle_lowres_init(want[:w * h].reshape(h, w), w, h).reshape(4, g.lh + 64, g.lstride)
            bad = np.argwhere(lr[i][:, :, :g.lw + 64] != lw[:, :, :g.lw + 64])
            assert bad.size == 0, ("lowres", i, bad[:6])
            orc.put_i420(want)
            assert np.array_equal(invq[i], orc.inv_qscale(i)), ("inv_qscale", i)
            assert np.array_equal(qp[i].view(np.uint32), orc.qp_offset(i, aq=True).view(np.uint32)), ("qp_offset_aq", i)
            # frame sums: the checker keeps them after mean removal ([x264] "Remove mean from SSD calculation")
            s_want, q_want = orc.pixel_stats(i)
            for pl in range(3):
                pw, ph = (16 * g.mb_w) >> (pl > 0), (16 * g.mb_h) >> (pl > 0)
                s, q = int(stats[i][pl]), int(stats[i][3 + pl])
                assert s == s_want[pl], ("sum", i, pl)
                assert q - (s * s + pw * ph // 2) // (pw * ph) == q_want[pl], ("ssd", i, pl)
    finally:
        orc.close()


def test_fused_front_end_refuses_what_it_cannot_tile():
    import torch
    from x264vfw_b200 import lowres
    from x264vfw_b200._lib import Context, CudaError
    ctx = Context(0)
    try:
        t = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
        with pytest.raises(CudaError, match="width"):
            lowres.frontend_batch(ctx, t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), BGRA, 72, 48, 1, 72 * 48 * 4, 72 * 48 * 2)
    finally:
        ctx.close()


@pytest.mark.parametrize("fused", ["1", "0"])
def test_session_is_identical_with_and_without_the_fused_front_end(fused, monkeypatch):
    """The session uses the fused kernel for eligible frames (X264VFW_CUDA_FUSED=0 keeps the three separate kernels):
    both must give the checker's decisions; aq-mode 2 exercises the auto-variance tail after the fused kernel."""
    from clipgen import SyntheticClip
    from x264vfw_b200 import lookahead
    monkeypatch.setenv("X264VFW_CUDA_FUSED", fused)
    w, h, n = 320, 192, 40
    clip = SyntheticClip(w, h, n_frames=n, cuts=(23,), flash=31, flash_len=1)
    packed = [clip.packed(i, "bgra") for i in range(n)]
    for over in ({"rc_lookahead": 10, "keyint_max": 50, "keyint_min": 5}, {"rc_lookahead": 10, "aq_mode": 2, "keyint_max": 50, "keyint_min": 5},
                 {"rc_lookahead": 10, "aq_mode": 0, "keyint_max": 50, "keyint_min": 5}):
        orc = ol.OracleLookahead(ol.la_params("medium", w, h, **over))
        gpu = lookahead.Lookahead(lookahead.params_preset("medium", w, h, **over), in_csp=BGRA | FLIP, device=0)
        do, dg = [], []
        try:
            for f in packed:
                orc.put_i420(ol.oracle_convert(f, BGRA | FLIP, 2, 2, 0, w, h))
                do += orc.decisions()
                conv = np.zeros(w * h * 3 // 2, dtype=np.uint8)
                gpu.put_frame(f, conv_pic=conv)
                assert np.array_equal(conv, ol.oracle_convert(f, BGRA | FLIP, 2, 2, 0, w, h))
                dg += gpu.decisions()
            orc.flush(); do += orc.decisions()
            gpu.flush(); dg += gpu.decisions()
        finally:
            orc.close(); gpu.close()
        assert [(d["i_frame"], d["i_type"], d["i_cost_est"], d["i_cost_est_aq"]) for d in dg] == \
               [(d["i_frame"], d["i_type"], d["i_cost_est"], d["i_cost_est_aq"]) for d in do], over
        for a, b in zip(dg, do):
            assert np.array_equal(a["qp_offset"].view(np.uint32), b["qp_offset"].view(np.uint32)), (over, a["i_frame"])
            assert np.array_equal(a["qp_offset_aq"].view(np.uint32), b["qp_offset_aq"].view(np.uint32)), (over, a["i_frame"])
