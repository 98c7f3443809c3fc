"""Fused front end alone: n frames of 1080p BGRA per launch, a few launches (for ncu captures and timing)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x264vfw_b200 as xv
from x264vfw_b200 import csp, lowres

W, H = (int(v) for v in os.environ.get("SIZE", "1920x1080").split("x"))
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 96
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = xv._lib.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
sfb, dfb = csp.frame_bytes(9 | 0x1000, 2, W, H)
g = lowres.geometry(W, H)
src = torch.randint(0, 256, (nf * sfb,), dtype=torch.uint8, device="cuda")
dst = torch.empty(nf * dfb, dtype=torch.uint8, device="cuda")
lr = torch.empty(nf * 4 * g.lplane_bytes, dtype=torch.uint8, device="cuda")
fe = lowres.FusedBatch(ctx, W, H, nf)
torch.cuda.synchronize()
for _ in range(3):
    fe.run(src.data_ptr(), dst.data_ptr(), lr.data_ptr())
ctx.sync()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st)
for _ in range(iters):
    fe.run(src.data_ptr(), dst.data_ptr(), lr.data_ptr())
b.record(st)
b.synchronize()
t = a.elapsed_time(b) / iters * 1e-3
algo = W * H * 4 + W * H * 3 // 2 + 4 * g.lw * g.lh
print(json.dumps({"size": f"{W}x{H}", "frames_per_launch": nf, "us_per_launch": t * 1e6, "us_per_frame": t * 1e6 / nf,
                  "algorithmic_bytes_per_frame": algo, "gbs": algo * nf / t / 1e9}))
