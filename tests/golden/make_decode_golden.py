"""Regenerates tests/golden/decode_golden.json: FNV-1a-64 fingerprints of what libswscale 9.1.100 (the copy inside
this image's opencv wheel) writes when it is driven the way the reference's decompress path drives it
(golden/swsref.py restates codec.c:2075-2152 + :2292 call for call).  These ARE outputs of the third-party
library the reference delegates this stage to -- the pin of oracle/decode_oracle.c and, through it, of the CUDA path.

Inputs: oracle_lib.decode_source (SURVEY A.4 byte generator, seeded).  The fingerprint covers the picture's pixel
bytes only (BGR24 rows carry 0-3 alignment bytes which libswscale's 8-pixel stores scribble on).
Run in the authoring container: `python tests/golden/make_decode_golden.py`."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import oracle_lib as ol  # noqa: E402
import swsref as sr  # noqa: E402

SIZES = [(16, 10), (64, 32), (70, 38), (72, 40), (320, 240), (1280, 720), (1920, 1080)]
FORMATS = [sr.CSP_I420, sr.CSP_YV12, sr.CSP_NV12, sr.CSP_YUYV, sr.CSP_UYVY, sr.CSP_BGR, sr.CSP_BGRA,
           sr.CSP_BGR | sr.CSP_VFLIP, sr.CSP_BGRA | sr.CSP_VFLIP]
MATRICES = [(2, 0), (1, 0), (5, 1), (9, 0), (7, 1), (4, 0), (6, 1)]      # (AVCOL_SPC_*, full range)


def pixel_bytes(out, csp, w, h):
    if csp & 0xff == sr.CSP_BGR:
        st = (w * 3 + 3) & ~3
        return np.ascontiguousarray(out.reshape(h, st)[:, :w * 3])
    return out


def cases_422():
    """High 4:2:2 decoder pictures (what "keep input colorspace" makes of YUY2 / UYVY input): the outputs that keep or raise the
    chroma height."""
    for w, h in [(16, 12), (70, 38), (72, 40), (320, 240), (1920, 1080)]:
        for csp in [sr.CSP_YV16, sr.CSP_YUYV, sr.CSP_UYVY, sr.CSP_BGR, sr.CSP_BGRA, sr.CSP_BGRA | sr.CSP_VFLIP, sr.CSP_BGR | sr.CSP_VFLIP]:
            if (csp & sr.CSP_VFLIP) and w % 8:
                continue
            for spc, full in MATRICES:
                if w >= 320 and (spc, full) not in ((2, 0), (1, 0)):
                    continue
                if csp & 0xff in (sr.CSP_YV16, sr.CSP_YUYV, sr.CSP_UYVY) and (spc, full) != (2, 0):
                    continue
                yield w, h, csp, spc, full


def cases_444():
    """High 4:4:4 decoder pictures: RGB output (libswscale's full-chroma writer) and the YV24 plane copy."""
    for w, h in [(16, 12), (70, 38), (72, 40), (320, 240), (1920, 1080)]:
        for csp in [sr.CSP_YV24, sr.CSP_BGR, sr.CSP_BGRA, sr.CSP_BGRA | sr.CSP_VFLIP, sr.CSP_BGR | sr.CSP_VFLIP]:
            for spc, full in MATRICES:
                if w >= 320 and (spc, full) not in ((2, 0), (1, 0)):
                    continue
                if csp == sr.CSP_YV24 and (spc, full) != (2, 0):
                    continue
                yield w, h, csp, spc, full


def cases_up():
    """YUV outputs with ANOTHER chroma resolution than the decoder picture (libswscale's scaler on the chroma planes: 4 taps up, 8 taps
    down), packed 4:2:2 from 4:4:4 included."""
    for w, h in [(24, 24), (70, 38), (72, 40), (320, 240), (1920, 1080)]:
        for src, csps in ((1, (sr.CSP_YV16, sr.CSP_YV24)),
                          (2, (sr.CSP_I420, sr.CSP_YV12, sr.CSP_NV12, sr.CSP_YV24)),
                          (3, (sr.CSP_I420, sr.CSP_YV12, sr.CSP_NV12, sr.CSP_YV16, sr.CSP_YUYV, sr.CSP_UYVY))):
            for csp in csps:
                yield w, h, csp, src


def cases():
    for w, h in SIZES:
        for csp in FORMATS:
            if (csp & sr.CSP_VFLIP) and w % 8:
                continue      # libswscale's 8-pixel stores then land in a row it already wrote (documented in the oracle)
            for spc, full in MATRICES:
                if w >= 320 and (spc, full) not in ((2, 0), (1, 0), (5, 1)):
                    continue
                if csp & 0xff in (sr.CSP_I420, sr.CSP_YV12, sr.CSP_NV12) and (spc, full) != (2, 0):
                    continue  # plane copies do not look at the matrix
                yield w, h, csp, spc, full


def main():
    out = {"library": "libswscale " + sr.version(), "cases": []}
    for w, h, csp, spc, full in cases():
        y, u, v = ol.decode_source(w, h, seed=spc + full, pad=24)
        dib = sr.decompress_convert(y, u, v, csp, spc, full)
        out["cases"].append({"w": w, "h": h, "csp": csp, "spc": spc, "full": full,
                             "fnv": ol.fnv(pixel_bytes(dib, csp, w, h))})
    for w, h, csp, spc, full in cases_422():
        y, u, v = ol.decode_source(w, h, seed=spc + full, pad=24, src_chroma=2)
        dib = sr.decompress_convert(y, u, v, csp, spc, full, src_chroma=2)
        out["cases"].append({"w": w, "h": h, "csp": csp, "spc": spc, "full": full, "src": 2,
                             "fnv": ol.fnv(pixel_bytes(dib, csp, w, h))})
    for w, h, csp, spc, full in cases_444():
        y, u, v = ol.decode_source(w, h, seed=spc + full, pad=24, src_chroma=3)
        dib = sr.decompress_convert(y, u, v, csp, spc, full, src_chroma=3)
        out["cases"].append({"w": w, "h": h, "csp": csp, "spc": spc, "full": full, "src": 3,
                             "fnv": ol.fnv(pixel_bytes(dib, csp, w, h))})
    for w, h, csp, src in cases_up():
        y, u, v = ol.decode_source(w, h, seed=2, pad=24, src_chroma=src)
        dib = sr.decompress_convert(y, u, v, csp, 2, 0, src_chroma=src)
        out["cases"].append({"w": w, "h": h, "csp": csp, "spc": 2, "full": 0, "src": src, "fnv": ol.fnv(dib)})
    # one small picture in full, for debugging a mismatch by eye
    y, u, v = ol.decode_source(16, 10, seed=2, pad=24)
    out["sample_16x10_bgra"] = sr.decompress_convert(y, u, v, sr.CSP_BGRA, 2, 0).tolist()
    with open(os.path.join(HERE, "decode_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
        f.write("\n")
    print(len(out["cases"]), "cases from", out["library"])


if __name__ == "__main__":
    main()
