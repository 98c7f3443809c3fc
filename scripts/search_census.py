"""Which (list, distance) searches the decision logic asks for vs which are launched (speculation waste), one 1080p
stream on the bench clip (first N frames), X264VFW_CUDA_STATS=1 summary at close."""
import os, sys
os.environ["X264VFW_CUDA_STATS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from x264vfw_b200 import lookahead
from clipgen import SyntheticClip

W, H = 1920, 1080
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
clip = SyntheticClip(W, H, n_frames=300, cuts=(100, 200), flash=150, flash_len=2)
la = lookahead.Lookahead(lookahead.params_preset("medium", W, H), in_csp=9 | 0x1000, device=0)
for i in range(n):
    f = torch.from_numpy(clip.packed(i, "bgra")).cuda()
    la.put_frame(f.data_ptr(), on_device=True)
    la.decisions()
la.flush()
la.close()
