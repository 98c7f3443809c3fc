"""SURVEY 8(f) row 3 remainder, timing: the integral image on batches of padded 1080p half-pel planes (CUDA events) and one
encoder-side weight analysis of a 1080p fade (wall clock around the call: 2 launches + 2 synchronisations), with the CPU checker
beside it."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import x264vfw_b200 as xv
from x264vfw_b200 import b3, hpel, lookahead
import oracle_lib as ol
from test_encoder_side_oracle import fade_frames, nv12

res = {}
ctx = xv._lib.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
g = hpel.geometry(1920, 1088)
rows, stride, nf = g.plane_bytes // g.stride, g.stride, 48
planes = torch.randint(0, 256, (nf * rows * stride,), dtype=torch.uint8, device="cuda")
s8 = torch.empty(nf * rows * stride, dtype=torch.int16, device="cuda")
s4 = torch.empty(nf * rows * stride, dtype=torch.int16, device="cuda")
for name, p4 in (("sum8", 0), ("sum8_sum4", s4.data_ptr())):
    run = lambda: b3.integral_init(ctx, s8.data_ptr(), p4, planes.data_ptr(), stride, rows, rows * stride, rows * stride, nf)
    for _ in range(3):
        run()
    ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(10):
        run()
    b.record(st); b.synchronize()
    t = a.elapsed_time(b) / 10 * 1e-3
    algo = rows * stride * (1 + 2 + (2 if p4 else 0))
    res["integral_" + name] = {"plane": f"{stride}x{rows}", "planes_per_launch": nf, "us_per_plane": t * 1e6 / nf,
                               "algorithmic_bytes_per_plane": algo, "gbs": algo * nf / t / 1e9}
del planes, s8, s4

w, h = 1920, 1080
frames = fade_frames(w, h, 3, 28, chroma_step=24)
uv = [nv12(f, w, h) for f in frames]
d_uv = [torch.from_numpy(u).cuda() for u in uv]
stride = ol.lowres_geometry(w, h)["luma_w"]
for subme in (7, 11):
    orc = ol.OracleLookahead(ol.la_params("medium", w, h, rc_lookahead=10, subme=subme))
    gpu = lookahead.Lookahead(lookahead.params_preset("medium", w, h, rc_lookahead=10, subme=subme), device=0, keep_frames=True)
    for f in frames:
        orc.put_i420(f); gpu.put_frame(f)
    assert orc.frame_cost(1, 2, 2) == gpu.frame_cost(1, 2, 2)
    t0 = time.perf_counter(); want = orc.weights_full(2, 1, uv[2], uv[1], stride); t_cpu = time.perf_counter() - t0
    got = b3.la_weights_analyse(gpu, 2, 1, d_uv[2].data_ptr(), d_uv[1].data_ptr(), stride)
    t0 = time.perf_counter()
    for _ in range(10):
        got = b3.la_weights_analyse(gpu, 2, 1, d_uv[2].data_ptr(), d_uv[1].data_ptr(), stride)
    t_gpu = (time.perf_counter() - t0) / 10
    assert got[0] == want[0]
    res[f"weights_analyse_subme{subme}"] = {"weights": got[0], "gpu_ms_per_call": t_gpu * 1e3, "cpu_checker_ms_per_call": t_cpu * 1e3}
    orc.close(); gpu.close()
print(json.dumps(res))
