"""Throughput of the other BASELINE.json configurations (parity-test cases, not bench lines):
8 concurrent device-resident streams each, native host threads (host/x264vfw_harness.c)."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x264vfw_b200 import lookahead
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from clipgen import SyntheticClip
from x264vfw_b200.harness import StreamSet

FLIP = 0x1000
CONFIGS = [
    ("C1/C5 1080p RGB32 bottom-up -> I420, medium", 1920, 1080, "bgra", 9 | FLIP, 2, 1, "medium", {}, 0),
    ("C2 720p YUY2 -> I420, veryfast, rc-lookahead 20", 1280, 720, "yuyv", 6, 2, 1, "veryfast", {"rc_lookahead": 20}, 0),
    ("C3 1080p RGB24 bottom-up -> I420 planes, slow, b-adapt 2, rc-lookahead 60", 1920, 1080, "bgr", 8 | FLIP, 2, 1, "slow", {"b_adapt": 2, "rc_lookahead": 60}, 0),
    ("C4 2160p UYVY -> I422, medium", 3840, 2160, "uyvy", 7, 6, 2, "medium", {}, 0),
]
S = int(os.environ.get("STREAMS", "8"))
for name, w, h, fmt, in_csp, out_csp, cf, preset, over, ext in CONFIGS:
    n = 24
    clips = []
    for s in range(S):
        clip = SyntheticClip(w, h, n_frames=n, stream_id=s % 2, cuts=(15,), flash=None)
        clips.append([torch.from_numpy(clip.packed(i, fmt)).cuda() for i in range(n)] if s < 2 else clips[s % 2])
    ptrs = [[t.data_ptr() for t in c] for c in clips]
    over = dict(over, chroma_format=cf)
    las = [lookahead.Lookahead(lookahead.params_preset(preset, w, h, **over), in_csp=in_csp, out_csp=out_csp, device=0) for _ in range(S)]
    ss = StreamSet(las, ptrs, True, None)
    ss.run(80)                      # fill the lookahead, warm up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ss.run(96)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:80s} {S} streams: {S * 96 / dt:8.0f} frames/s")
    ss.close()
    for la in las:
        la.close()
    del clips, ptrs
    torch.cuda.empty_cache()
