"""Quick device-resident throughput probe for the stage-1 kernels (not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x264vfw_b200 as xv
from x264vfw_b200 import csp, lowres

ctx = xv._lib.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
PEAK = 6453.7


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters):
        fn()
    e1.record(st)
    e1.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


cases = [("BGRA|flip->I420 1080p", 9 | 0x1000, 2, 1920, 1080, 0), ("BGR|flip->I420 1080p", 8 | 0x1000, 2, 1920, 1080, 0),
         ("BGR|flip->NV12ext 1080p", 8 | 0x1000, 4, 1920, 1080, 1),
         ("YUYV->I420 720p", 6, 2, 1280, 720, 0), ("UYVY->I420 2160p", 7, 2, 3840, 2160, 0),
         ("UYVY->I422 2160p", 7, 6, 3840, 2160, 0), ("UYVY->I444ext 2160p", 7, 0xc, 3840, 2160, 2),
         ("YV12->I420 1080p", 2, 2, 1920, 1080, 0)]
for name, ic, oc, w, h, ext in cases:
    sfb, dfb = csp.frame_bytes(ic, oc, w, h)
    _, sb = csp.img_fill(0, ic, w, h)
    _, db = csp.picture_layout(0, oc, w, h)
    nf = max(4, int(1.2e9 // (sfb + dfb)))
    src = torch.randint(0, 256, (nf * sfb,), dtype=torch.uint8, device="cuda")
    dst = torch.empty(nf * dfb, dtype=torch.uint8, device="cuda")
    t = timeit(lambda: csp.convert_batch(ctx, src.data_ptr(), dst.data_ptr(), ic, oc, 2, 0, w, h, nf, ext))
    gbs = (sb + db) * nf / t / 1e9
    print(f"{name:28s} nf={nf:4d} {t*1e3:8.3f} ms  {nf/t:10.0f} fps  {gbs:7.1f} GB/s  {gbs/PEAK*100:5.1f}% of measured peak")
    del src, dst

for w, h in [(1920, 1080), (1280, 720), (3840, 2160)]:
    g = lowres.geometry(w, h)
    sfb = (w * h + 255) // 256 * 256
    dfb = 4 * g.lplane_bytes
    nf = max(4, int(1.2e9 // (sfb + dfb)))
    y = torch.randint(0, 256, (nf * sfb,), dtype=torch.uint8, device="cuda")
    out = torch.empty(nf * dfb, dtype=torch.uint8, device="cuda")
    t = timeit(lambda: lowres.lowres_init(ctx, out.data_ptr(), y.data_ptr(), w, w, h, sfb, dfb, nf))
    alg = g.luma_w * g.luma_h + 4 * g.lw * g.lh
    gbs = alg * nf / t / 1e9
    print(f"lowres_init {w}x{h:5d}      nf={nf:4d} {t*1e3:8.3f} ms  {nf/t:10.0f} fps  {gbs:7.1f} GB/s  {gbs/PEAK*100:5.1f}% (algorithmic bytes)")
    del y, out
