"""Per-launch hit rates / times of the speculative search on one 1080p stream (diagnostics)."""
import os, sys
os.environ["X264VFW_CUDA_STATS"] = "2"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x264vfw_b200 import lookahead
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from clipgen import SyntheticClip

W, H = 1920, 1080
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
clip = SyntheticClip(W, H, n_frames=48, cuts=(30,), flash=None)
frames = [torch.from_numpy(clip.packed(i, "bgra")).cuda() for i in range(48)]
la = lookahead.Lookahead(lookahead.params_preset("medium", W, H), in_csp=9 | 0x1000, device=0)
for i in range(n):
    print(f"--- put {i}", file=sys.stderr)
    la.put_frame(frames[i % 48].data_ptr(), on_device=True)
    la.decisions()
la.close()
