"""Round-2 kernels at small sizes for compute-sanitizer (memcheck): decoder output conversion (every picture format x output
format, aligned and ragged), integral image, encoder-side weight analysis, lowres_init's two-row path.  Results are also checked
against the CPU checker so that a sanitizer run doubles as a parity run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import x264vfw_b200 as xv
from x264vfw_b200 import decode, b3, lookahead, lowres
import oracle_lib as ol
from test_encoder_side_oracle import fade_frames, nv12

ctx = xv._lib.Context(0)
n = 0
for src, csps in ((1, (1, 2, 3, 4, 5, 6, 7, 8, 9, 8 | 0x1000, 9 | 0x1000)), (2, (1, 2, 3, 4, 5, 6, 7, 8, 9 | 0x1000)), (3, (1, 2, 3, 4, 5, 6, 7, 8, 9, 9 | 0x1000))):
    for w, h in ((64, 32), (70, 38), (16, 12)):
        y, u, v = ol.decode_source(w, h, seed=src, pad=8, src_chroma=src)
        for csp in csps:
            if ol.oracle_decode_convert(y, u, v, csp, 1, 0, src_chroma=src) is None:
                continue                                    # too small for libswscale's full tap count: refused on both sides
            d = decode.Decompressor(csp, w, h, 1, 0, ctx=ctx, src_chroma=src)
            got = d.decompress(y, u, v)
            d.close()
            assert (got == ol.oracle_decode_convert(y, u, v, csp, 1, 0, src_chroma=src)).all(), (src, w, h, hex(csp))
            n += 1
print("decode cases", n)

rows, stride = 72, 136
plane = np.random.default_rng(0).integers(0, 256, (rows, stride), dtype=np.uint8)
d_p = torch.from_numpy(plane).cuda()
d_8 = torch.zeros((rows, stride), dtype=torch.int16, device="cuda"); d_4 = torch.zeros((rows, stride), dtype=torch.int16, device="cuda")
b3.integral_init(ctx, d_8.data_ptr(), d_4.data_ptr(), d_p.data_ptr(), stride, rows)
ctx.sync()
s8, s4 = ol.oracle_integral_init(plane)
assert (d_8.cpu().numpy().view(np.uint16)[: rows - 7, : stride - 8] == s8[: rows - 7, : stride - 8]).all()
assert (d_4.cpu().numpy().view(np.uint16)[: rows - 3, : stride - 4] == s4[: rows - 3, : stride - 4]).all()
print("integral ok")

w, h = 176, 112
frames = fade_frames(w, h, 3, 30, chroma_step=20)
orc = ol.OracleLookahead(ol.la_params("medium", w, h, rc_lookahead=10, subme=9))
gpu = lookahead.Lookahead(lookahead.params_preset("medium", w, h, rc_lookahead=10, subme=9), device=0, keep_frames=True)
for f in frames:
    orc.put_i420(f); gpu.put_frame(f)
assert orc.frame_cost(1, 2, 2) == gpu.frame_cost(1, 2, 2)
uv = [nv12(f, w, h) for f in frames]
d_uv = [torch.from_numpy(a).cuda() for a in uv]
st = ol.lowres_geometry(w, h)["luma_w"]
assert b3.la_weights_analyse(gpu, 2, 1, d_uv[2].data_ptr(), d_uv[1].data_ptr(), st)[0] == orc.weights_full(2, 1, uv[2], uv[1], st)[0]
orc.close(); gpu.close()
print("weights ok")

for w, h in ((64, 48), (330, 186)):
    yy = np.random.default_rng(1).integers(0, 256, (h, w), dtype=np.uint8)
    g = lowres.geometry(w, h)
    d_y = torch.from_numpy(yy).cuda(); d_lr = torch.zeros(4 * g.lplane_bytes, dtype=torch.uint8, device="cuda")
    lowres.lowres_init(ctx, d_lr.data_ptr(), d_y.data_ptr(), w, w, h)
    ctx.sync()
    want = ol.oracle_lowres_init(yy, w, h).reshape(4, g.lh + 64, g.lstride)[:, :, :g.lw + 64]
    assert (d_lr.cpu().numpy().reshape(4, g.lh + 64, g.lstride)[:, :, :g.lw + 64] == want).all()
print("lowres ok")
ctx.close()
