"""Regenerates tests/golden/lookahead_golden.json: fingerprints of what the CPU lookahead checker
(oracle/lookahead_oracle.c) produces on small synthetic clips.

libx264 is not part of the reference tree, so these are NOT reference outputs (the checker stays
"parity unpinned", see oracle/lookahead_oracle.h): the file freezes the checker itself, so that an
edit to it -- or to the clip generator -- that changes any frame type, cost, MV or qp offset is
caught by `pytest -m "not gpu"` before it silently moves the target of the GPU parity tests.
Run in the authoring container: `python tests/golden/make_lookahead_golden.py`."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle_lib as ol  # noqa: E402
from clipgen import SyntheticClip  # noqa: E402

BGRA_FLIP = 9 | 0x1000
CASES = [
    ("medium 128x96", "medium", 128, 96, 14, dict(rc_lookahead=6, keyint_max=50, keyint_min=2)),
    ("veryfast 160x96", "veryfast", 160, 96, 12, dict(rc_lookahead=5, keyint_max=50, keyint_min=2)),
    ("slower b-adapt 2 128x96", "slower", 128, 96, 12, dict(rc_lookahead=6, keyint_max=50, keyint_min=2)),
    ("superfast 128x80", "superfast", 128, 80, 10, dict(keyint_max=50, keyint_min=2)),
    ("medium aq-mode 2 128x96", "medium", 128, 96, 10, dict(rc_lookahead=5, aq_mode=2, keyint_max=50, keyint_min=2)),
    ("medium aq-mode 3 112x96", "medium", 112, 96, 10, dict(rc_lookahead=5, aq_mode=3, keyint_max=50, keyint_min=2)),
]


def fingerprint(preset, w, h, n, over, backend="oracle"):
    """backend "oracle": the CPU checker; "gpu": the device path through the C ABI (packed frames in)."""
    clip = SyntheticClip(w, h, n_frames=n, cuts=(n * 5 // 8,), flash=None)
    out = []
    if backend == "gpu":
        from x264vfw_b200 import lookahead
        la = lookahead.Lookahead(lookahead.params_preset(preset, w, h, **over), in_csp=BGRA_FLIP, device=0)
        for i in range(n):
            la.put_frame(clip.packed(i, "bgra"))
            out += la.decisions()
        la.flush()
        out += la.decisions()
        la.close()
    else:
        orc = ol.OracleLookahead(ol.la_params(preset, w, h, **over))
        for i in range(n):
            orc.put_i420(ol.oracle_convert(clip.packed(i, "bgra"), BGRA_FLIP, 2, 2, 0, w, h))
            out += orc.decisions()
        orc.flush()
        out += orc.decisions()
        orc.close()
    types = "".join("?IiPbB"[d["i_type"]] for d in sorted(out, key=lambda d: d["i_frame"]))
    order = [int(d["i_frame"]) for d in out]
    costs = [[int(d["i_cost_est"]), int(d["i_cost_est_aq"]), int(d["i_intra_mbs"])] for d in out]
    qp = ol.fnv(np.concatenate([d["qp_offset"].view(np.uint8) for d in out]))
    qp_aq = ol.fnv(np.concatenate([d["qp_offset_aq"].view(np.uint8) for d in out]))
    return {"types": types, "coded_order": order, "costs": costs, "qp_offset_fnv": qp, "qp_offset_aq_fnv": qp_aq}


def main():
    out = []
    for name, preset, w, h, n, over in CASES:
        fp = fingerprint(preset, w, h, n, over)
        out.append({"name": name, "preset": preset, "w": w, "h": h, "frames": n, "over": over, **fp})
        print(name, fp["types"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lookahead_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
