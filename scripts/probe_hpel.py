"""Device-resident timing of the half-pel reference plane kernel at the BASELINE frame sizes:
CUDA events on the launching stream, batches larger than L2.  Prints one JSON line per case; with --once it
launches each case exactly once (for ncu)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import x264vfw_b200 as xv  # noqa: E402
from x264vfw_b200 import hpel  # noqa: E402

PEAK = 6453.7
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def main():
    once = "--once" in sys.argv
    ctx = xv._lib.Context(0)
    st = torch.cuda.ExternalStream(ctx.stream)
    cases = [(1920, 1088, 48, None), (1920, 1088, 96, None), (1920, 1088, 1, None), (1280, 720, 96, None), (3840, 2160, 12, None)]
    for (w, h, nf, variant) in cases:
        g = hpel.geometry(w, h)
        src = torch.randint(0, 256, (nf * w * h,), dtype=torch.uint8, device="cuda")
        dst = torch.empty(nf * 4 * g.plane_bytes, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        fn = lambda: hpel.hpel_filter(ctx, dst.data_ptr(), src.data_ptr(), w, w, h, w * h, 4 * g.plane_bytes, nf)
        if once:
            fn(); ctx.sync()
            continue
        for _ in range(3):
            fn()
        ctx.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        a.record(st)
        for _ in range(iters):
            fn()
        b.record(st)
        b.synchronize()
        t = a.elapsed_time(b) / iters * 1e-3
        algo = 5 * w * h * nf
        print(json.dumps({"kernel": "hpel_kernel", "w": w, "h": h, "frames_per_launch": nf, "us_per_launch": t * 1e6,
                          "algorithmic_bytes": algo, "gbs": algo / t / 1e9, "frac_of_measured_peak": algo / t / 1e9 / PEAK,
                          "written_bytes_incl_border": 4 * g.plane_bytes * nf}))
        del src, dst
    ctx.close()


if __name__ == "__main__":
    main()
