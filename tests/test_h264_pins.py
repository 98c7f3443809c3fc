"""The checker's half-pel planes, quarter-sample get_ref and intra predictors against an independent H.264 DECODER
(libavcodec, frozen in tests/golden/h264_pins.json by tests/golden/make_h264_pins.py).  The standard fixes these
results -- an encoder's prediction has to be the decoder's -- so this is a reference pin for those three pieces of
the restated [x264] code (tests/h264_pins.py says what stays unpinned)."""
import json
import os

import numpy as np
import pytest

import h264_pins as hp

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h264_pins.json")))


def test_fixture_matches_the_committed_layout():
    assert (GOLD["mc"]["w"], GOLD["mc"]["h"]) == (hp.MC_W, hp.MC_H)
    assert [tuple(m) for m in GOLD["mc"]["mvs"]] == hp.MC_MVS
    assert {tuple(t["mb"]): (t["luma_mode"], t["chroma_mode"]) for t in GOLD["intra"]} == hp.intra_tests()
    assert {t["luma_mode"] for t in GOLD["intra"]} == set(range(3, 9))
    assert {t["chroma_mode"] for t in GOLD["intra"]} == set(range(4))


@pytest.mark.parametrize("kind", hp.MC_KINDS)
def test_checker_hpel_planes_and_get_ref_equal_the_decoders_motion_compensation(kind):
    """hpel_oracle.c's four planes + lookahead_oracle.c's get_ref, whole picture, 22 vectors: all 16 quarter-sample
    phases and vectors up to 22 samples outside the picture (border replication)."""
    got = hp.checker_mc_hashes(kind)
    want = GOLD["mc"]["pictures"][kind]
    bad = [hp.MC_MVS[i] for i in range(len(want)) if got[i] != want[i]]
    assert not bad, bad


@pytest.mark.parametrize("kind", hp.MC_KINDS)
def test_checker_mc_chroma_equals_the_decoders_chroma_prediction(kind):
    """[x264] mc_chroma as restated for the encoder-side weight analysis: both chroma planes of the same 22
    pictures -- every eighth-sample phase the 16 quarter-sample luma vectors produce, and vectors far outside the picture."""
    got = hp.checker_mc_chroma_hashes(kind)
    want = GOLD["mc"]["chroma"][kind]
    bad = [hp.MC_MVS[i] for i in range(len(want)) if got[i] != want[i]]
    assert not bad, bad


def test_checker_equals_the_decoder_at_1080p():
    """1920x1088 (BASELINE's frame padded to whole macroblocks), six vectors."""
    assert (GOLD["hd"]["w"], GOLD["hd"]["h"], [tuple(m) for m in GOLD["hd"]["mvs"]]) == (hp.HD_W, hp.HD_H, hp.HD_MVS)
    assert hp.checker_hd_hashes() == GOLD["hd"]["pictures"]


def test_checker_weighted_get_ref_equals_the_decoders_explicit_weighted_prediction():
    """get_ref + mc_weight (a weighted reference of the lookahead) against P pictures with a pred_weight_table:
    denominators 0..7, scales 1..127, offsets -128..127, integer and fractional vectors."""
    assert [[list(m), list(w)] for m, w in hp.WP_CASES] == GOLD["wp"]["cases"]
    assert hp.checker_wp_hashes() == GOLD["wp"]["pictures"]


def test_checker_bidirectional_average_equals_the_decoders():
    """pixel_avg with the lookahead's bipred weight against B pictures (plain and implicitly weighted), incl. every
    (distance, position) the lookahead can meet with up to 16 B-frames: its own distance scale gives the
    standard's implicit weights."""
    assert [[c[0], c[1], list(c[2]), list(c[3]), c[4]] for c in hp.BI_CASES] == GOLD["bi"]["cases"]
    got, want = hp.checker_bi_hashes(), GOLD["bi"]["pictures"]
    bad = [hp.BI_CASES[i] for i in range(len(want)) if got[i] != want[i]]
    assert not bad, bad


def test_checker_intra_predictors_equal_the_decoders():
    """predict_8x8c_{dc,h,v,p} (as the decoder's intra chroma prediction) and predict_8x8_filter +
    predict_8x8_{ddl,ddr,vr,hd,vl,hu} (as its Intra_8x8 prediction), neighbours from I_PCM macroblocks."""
    got = hp.checker_intra_hashes()
    assert len(got) == len(GOLD["intra"])
    for a, b in zip(got, GOLD["intra"]):
        assert a == b, (a, b)


def test_numpy_formulation_of_the_planes_gives_the_same_pictures():
    """The independent numpy formulation of the half-pel planes (oracle_lib.numpy_hpel_planes) through the same
    get_ref: a second path to the decoder's pictures."""
    import oracle_lib as ol
    y, _, _ = hp.mc_picture("noise")
    planes = ol.numpy_hpel_planes(y)
    g = ol.hpel_geometry(hp.MC_W, hp.MC_H)
    full = np.zeros((4, hp.MC_H + 64, g["stride"]), dtype=np.uint8)
    full[:, :, :hp.MC_W + 64] = planes
    assert hp.checker_mc_hashes("noise", planes=full) == GOLD["mc"]["pictures"]["noise"]


def test_live_decoder_reproduces_the_fixture():
    """Only where the wheel's libavcodec is loadable (this image): the generator's path, end to end."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import avdec
    if not avdec.available():
        pytest.skip("no loadable libavcodec")
    assert hp.decoder_mc_hashes("noise") == GOLD["mc"]["pictures"]["noise"]
    assert hp.decoder_intra_hashes() == GOLD["intra"]
    assert hp.decoder_wp_hashes() == GOLD["wp"]["pictures"]
    assert hp.decoder_bi_hashes(hp.BI_CASES[:8]) == GOLD["bi"]["pictures"][:8]
