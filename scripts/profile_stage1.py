"""Few device-resident batch launches of the stage-1 kernels for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x264vfw_b200 as xv
from x264vfw_b200 import csp, lowres
ctx = xv._lib.Context(0)
W, H, nf = 1920, 1080, 64
for ic, oc in ((9 | 0x1000, 2), (8 | 0x1000, 2)):
    sfb, dfb = csp.frame_bytes(ic, oc, W, H)
    src = torch.randint(0, 256, (nf * sfb,), dtype=torch.uint8, device="cuda")
    dst = torch.empty(nf * dfb, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        csp.convert_batch(ctx, src.data_ptr(), dst.data_ptr(), ic, oc, 2, 0, W, H, nf)
    ctx.sync()
g = lowres.geometry(W, H)
lr = torch.empty(nf * 4 * g.lplane_bytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    lowres.lowres_init(ctx, lr.data_ptr(), dst.data_ptr(), W, W, H, dfb, 4 * g.lplane_bytes, nf)
ctx.sync()
