"""Decoder-side output conversion alone: n resident 1080p yuv420p pictures per launch, each output
format, CUDA events on the launching stream; also the one-picture HOST entry (the sws_scale call) end to end."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import x264vfw_b200 as xv
from x264vfw_b200 import decode

W, H = (int(v) for v in os.environ.get("SIZE", "1920x1080").split("x"))
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 48
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
ctx = xv._lib.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
sfb = W * H * 3 // 2
src = torch.randint(0, 256, (nf * sfb,), dtype=torch.uint8, device="cuda")
b0 = src.data_ptr()
FORMATS = {"bgra_bottom_up": (9 | 0x1000, 1), "bgra": (9, 1), "bgr24_bottom_up": (8 | 0x1000, 1), "yuyv": (6, 1), "uyvy": (7, 1), "nv12": (5, 1),
           "yv12": (2, 1), "422_to_bgra_bottom_up": (9 | 0x1000, 2), "422_to_yuyv": (6, 2), "444_to_bgra_bottom_up": (9 | 0x1000, 3),
           "444_to_bgr24_bottom_up": (8 | 0x1000, 3)}
res = {}
src444 = torch.randint(0, 256, (nf * W * H * 3,), dtype=torch.uint8, device="cuda")
for name, (csp, chroma) in FORMATS.items():
    if only and name not in only:
        continue
    d = decode.Decompressor(csp, W, H, decode.AVCOL_SPC_BT709, False, ctx=ctx, src_chroma=chroma)
    dfb = (d.picture_size + 255) & ~255
    dst = torch.empty(nf * dfb, dtype=torch.uint8, device="cuda")
    if chroma == 1:
        fb, planes, strides = sfb, (b0, b0 + W * H, b0 + W * H * 5 // 4), (W, W // 2, W // 2)
    elif chroma == 2:
        b4 = src444.data_ptr()
        fb, planes, strides = W * H * 2, (b4, b4 + W * H, b4 + W * H * 3 // 2), (W, W // 2, W // 2)
    else:
        b4 = src444.data_ptr()
        fb, planes, strides = W * H * 3, (b4, b4 + W * H, b4 + 2 * W * H), (W, W, W)
    run = lambda: d.decompress_batch(dst.data_ptr(), dfb, planes, strides, fb, nf)
    for _ in range(3):
        run()
    ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(iters):
        run()
    b.record(st)
    b.synchronize()
    t = a.elapsed_time(b) / iters * 1e-3
    algo = fb + d.picture_size
    res[name] = {"us_per_frame": t * 1e6 / nf, "algorithmic_bytes_per_frame": algo, "gbs": algo * nf / t / 1e9}
    if name == "bgra_bottom_up" and not only:
        # host entry: pageable numpy planes in, DIB out, synchronous (what one ICM_DECOMPRESS call costs)
        y = np.random.randint(0, 256, (H, W), dtype=np.uint8); u = np.random.randint(0, 256, (H // 2, W // 2), dtype=np.uint8); v = u.copy()
        out = np.zeros(d.picture_size, np.uint8)
        for _ in range(3):
            d.decompress(y, u, v, out)
        t0 = time.perf_counter()
        for _ in range(20):
            d.decompress(y, u, v, out)
        res["host_entry_bgra_ms_per_picture"] = (time.perf_counter() - t0) / 20 * 1e3
    d.close()
    del dst
print(json.dumps({"size": f"{W}x{H}", "frames_per_launch": nf, "formats": res}))
