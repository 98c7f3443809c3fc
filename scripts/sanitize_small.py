"""Small lookahead session for compute-sanitizer runs (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from x264vfw_b200 import lookahead
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from clipgen import SyntheticClip

w, h, n = 336, 192, 22
clip = SyntheticClip(w, h, n_frames=n, cuts=(12,), flash=None)
for preset, over in (("medium", dict(rc_lookahead=8, keyint_max=30, keyint_min=3)), ("superfast", dict(keyint_max=30, keyint_min=3))):
    la = lookahead.Lookahead(lookahead.params_preset(preset, w, h, **over), in_csp=9 | 0x1000, device=0)
    conv = np.empty(w * h * 3 // 2, dtype=np.uint8)
    out = []
    for i in range(n):
        la.put_frame(clip.packed(i, "bgra"), conv_pic=conv)
        out += la.decisions()
    la.flush(); out += la.decisions()
    la.close()
    print(preset, len(out), "".join("?IiPbB"[d["i_type"]] for d in sorted(out, key=lambda d: d["i_frame"])))
