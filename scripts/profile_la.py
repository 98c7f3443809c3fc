"""Small single-stream 1080p lookahead run for ncu captures (not a bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x264vfw_b200 import lookahead, csp
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from clipgen import SyntheticClip

W, H = 1920, 1080
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
clip = SyntheticClip(W, H, n_frames=n, cuts=(n * 5 // 8,), flash=None)
frames = [torch.from_numpy(clip.packed(i, "bgra")).cuda() for i in range(n)]
la = lookahead.Lookahead(lookahead.params_preset("medium", W, H, rc_lookahead=int(os.environ.get("RC_LOOKAHEAD", "40"))),
                         in_csp=9 | 0x1000, device=0)
out = []
for f in frames:
    la.put_frame(f.data_ptr(), on_device=True)
    out += la.decisions()
la.flush()
out += la.decisions()
print(len(out), la.counters())
la.close()
