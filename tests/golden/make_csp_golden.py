"""Regenerates tests/golden/csp_golden.json from the UNMODIFIED reference csp.c
(oracle/_ref/libref_csp.so, built by oracle/Makefile from /root/reference/csp.c).
Run in the authoring container only: `python tests/golden/make_csp_golden.py`.
Input bytes and hash are the SURVEY.md appendix A.4 definitions."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as ol  # noqa: E402

I420, YV12, YV16, YV24, NV12, YUYV, UYVY, BGR, BGRA, FLIP = 1, 2, 3, 4, 5, 6, 7, 8, 9, 0x1000
O420, ONV12, O422, O444, OBGR, OBGRA = 2, 4, 6, 0xc, 0xe, 0xf
CASES = [
    ("BGRA|VFLIP->I420 1080p", BGRA | FLIP, O420, 1920, 1080, 2, 0),
    ("BGRA|VFLIP->I420", BGRA | FLIP, O420, 64, 48, 2, 0),
    ("BGRA topdown->I420", BGRA, O420, 64, 48, 2, 0),
    ("BGRA|VFLIP->I420 601pc", BGRA | FLIP, O420, 64, 48, 2, 1),
    ("BGRA|VFLIP->I420 709tv", BGRA | FLIP, O420, 64, 48, 1, 0),
    ("BGRA|VFLIP->I420 709pc", BGRA | FLIP, O420, 64, 48, 1, 1),
    ("BGR|VFLIP->I420 1080p", BGR | FLIP, O420, 1920, 1080, 2, 0),
    ("BGR|VFLIP->I420 w66", BGR | FLIP, O420, 66, 48, 2, 0),
    ("BGR|VFLIP->I420 709tv w66", BGR | FLIP, O420, 66, 48, 1, 0),
    ("YUYV->I420 720p", YUYV, O420, 1280, 720, 2, 0),
    ("YUYV->I420", YUYV, O420, 64, 48, 2, 0),
    ("YUYV->I422", YUYV, O422, 64, 48, 2, 0),
    ("UYVY->I420 2160p", UYVY, O420, 3840, 2160, 2, 0),
    ("UYVY->I420", UYVY, O420, 64, 48, 2, 0),
    ("UYVY->I422 2160p", UYVY, O422, 3840, 2160, 2, 0),
    ("UYVY->I422", UYVY, O422, 64, 48, 2, 0),
    ("UYVY->I444 (unregistered)", UYVY, O444, 64, 48, 2, 0),
    ("BGR->NV12 (unregistered)", BGR, ONV12, 64, 48, 2, 0),
    ("YV12->I420", YV12, O420, 64, 48, 2, 0),
    ("I420->I420", I420, O420, 64, 48, 2, 0),
    ("NV12->NV12", NV12, ONV12, 64, 48, 2, 0),
    ("YV16->I420", YV16, O420, 64, 48, 2, 0),
    ("YV16->I422", YV16, O422, 64, 48, 2, 0),
    ("YV24->I420", YV24, O420, 64, 48, 2, 0),
    ("YV24->I444", YV24, O444, 64, 48, 2, 0),
    ("BGR|VFLIP->BGR w66", BGR | FLIP, OBGR, 66, 48, 0, 1),
    ("BGRA|VFLIP->BGRA", BGRA | FLIP, OBGRA, 64, 48, 0, 1),
    ("I420|VFLIP->I420", I420 | FLIP, O420, 64, 48, 2, 0),
    ("YV16|VFLIP->I420", YV16 | FLIP, O420, 64, 48, 2, 0),
    ("YV24|VFLIP->I420", YV24 | FLIP, O420, 64, 48, 2, 0),
    ("YUYV|VFLIP->I420", YUYV | FLIP, O420, 64, 48, 2, 0),
]


def main():
    assert ol.have_ref_csp(), "build oracle/_ref first: make -C oracle"
    out = []
    for name, icsp, ocsp, w, h, cm, fr in CASES:
        n = ol.layout_bytes(ol.src_layout(icsp, w, h))
        src = ol.lcg_bytes(n, w, h)
        dst = ol.ref_convert(src, icsp, ocsp, cm, fr, w, h)
        out.append({"name": name, "in_csp": icsp, "out_csp": ocsp, "w": w, "h": h, "colmatrix": cm,
                    "fullrange": fr, "src_fnv": ol.fnv(src),
                    "dst_fnv": None if dst is None else ol.fnv(dst),
                    "ret": -1 if dst is None else 0,
                    "dst_head": None if dst is None else [int(v) for v in dst[:4]]})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csp_golden.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_csp_golden.py", "source": "reference csp.c via oracle/_ref/libref_csp.so",
                   "cases": out}, f, indent=1)
    print("wrote", path, len(out), "cases")


if __name__ == "__main__":
    main()
