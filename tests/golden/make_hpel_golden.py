"""Regenerates tests/golden/hpel_golden.json: FNV-1a-64 fingerprints of the four half-pel reference planes
the CPU checker (oracle/hpel_oracle.c) produces for seeded inputs (SURVEY A.4 byte generator).

libx264 is not part of the reference tree, so these are NOT reference outputs (the checker stays "parity
unpinned"): the file freezes the checker, so that an edit which changes any byte of any plane is caught by
`pytest -m "not gpu"` before it silently moves the target of the GPU parity tests, and it lets the GPU suite
check the device path without the checker in the loop.
Run in the authoring container: `python tests/golden/make_hpel_golden.py`."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as ol  # noqa: E402

CASES = [(16, 16), (64, 48), (248, 12), (480, 32), (1280, 720), (1920, 1088)]


def source_plane(w, h):
    """SURVEY A.4 generator: s = 0x264 + 31w + h; s = s*1664525 + 1013904223; byte = s >> 24."""
    buf = np.zeros(w * h, dtype=np.uint8)
    ol.oracle().orc_lcg_fill(buf.ctypes.data, buf.size, w, h)
    return buf.reshape(h, w)


def fnv(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return f"{ol.oracle().orc_fnv1a64(a.ctypes.data, a.size):016x}"


def fingerprint(planes: np.ndarray, w: int):
    """planes: (4, h+64, stride) -> hashes of the (h+64) x (w+64) bytes of each plane."""
    return [fnv(planes[p, :, :w + 64]) for p in range(4)]


def main():
    out = []
    for w, h in CASES:
        y = source_plane(w, h)
        out.append({"w": w, "h": h, "src_fnv": fnv(y), "planes_fnv": fingerprint(ol.oracle_hpel_planes(y, w, h), w)})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hpel_golden.json")
    json.dump(out, open(path, "w"), indent=1)
    print(f"wrote {path}: {len(out)} cases")


if __name__ == "__main__":
    main()
