"""Regenerates tests/golden/h264_pins.json from libavcodec's H.264 decoder (the FFmpeg build inside the opencv wheel
of this image): hashes of the pictures it decodes from the bitstreams of tests/golden/h264_mini.py -- motion-
compensated pictures for every quarter-sample phase (and vectors far outside the picture), and Intra_8x8 / intra
chroma predictions.  See tests/h264_pins.py for what this pins and what it does not.
Run in the authoring container: `python tests/golden/make_h264_pins.py`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import h264_pins as hp  # noqa: E402


def main():
    out = {"source": "libavcodec h264 decoder (opencv_python_headless wheel, avcodec 62.11), bitstreams from h264_mini.py",
           "mc": {"w": hp.MC_W, "h": hp.MC_H, "mvs": [list(m) for m in hp.MC_MVS],
                  "pictures": {k: hp.decoder_mc_hashes(k) for k in hp.MC_KINDS},
                  "chroma": {k: hp.decoder_mc_chroma_hashes(k) for k in hp.MC_KINDS}},
           "hd": {"w": hp.HD_W, "h": hp.HD_H, "mvs": [list(m) for m in hp.HD_MVS], "pictures": hp.decoder_hd_hashes()},
           "wp": {"cases": [[list(m), list(w)] for m, w in hp.WP_CASES], "pictures": hp.decoder_wp_hashes()},
           "bi": {"cases": [[c[0], c[1], list(c[2]), list(c[3]), c[4]] for c in hp.BI_CASES], "pictures": hp.decoder_bi_hashes()},
           "intra": hp.decoder_intra_hashes()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "h264_pins.json")
    json.dump(out, open(path, "w"), indent=1)
    print(f"wrote {path}: {sum(len(v) for v in out['mc']['pictures'].values())} motion-compensated pictures, {len(out['intra'])} intra macroblocks")


if __name__ == "__main__":
    main()
