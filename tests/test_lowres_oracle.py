"""CPU tests of the lowres oracle (parity unpinned: upstream libx264 is absent).  The C
restatement is cross-checked against an independent numpy formulation of the same published
filter, plus size-independent properties."""
import numpy as np
import pytest

import oracle_lib as ol


def numpy_lowres(y, w, h):
    g = ol.lowres_geometry(w, h)
    P = np.pad(y, ((0, g["luma_h"] + 1 - h), (0, g["luma_w"] + 1 - w)), mode="edge").astype(np.int32)
    avg = lambda a, b: (a + b + 1) >> 1
    lw, lh = g["lw"], g["lh"]
    r0, r1, r2 = P[0:2 * lh:2], P[1:2 * lh + 1:2], P[2:2 * lh + 2:2]
    v01, v12 = avg(r0, r1), avg(r1, r2)
    e = lambda v: v[:, 0:2 * lw:2]
    o = lambda v: v[:, 1:2 * lw + 1:2]
    n = lambda v: v[:, 2:2 * lw + 2:2]
    planes = [avg(e(v01), o(v01)), avg(o(v01), n(v01)), avg(e(v12), o(v12)), avg(o(v12), n(v12))]
    out = []
    for p in planes:
        q = np.pad(p, 32, mode="edge").astype(np.uint8)
        full = np.zeros((lh + 64, g["lstride"]), dtype=np.uint8)
        full[:, :lw + 64] = q
        out.append(full)
    return np.stack(out), g


@pytest.mark.parametrize("size", [(64, 48), (66, 50), (1280, 720), (1920, 1080), (18, 2)])
def test_lowres_oracle_matches_numpy_formulation(size):
    w, h = size
    y = np.random.default_rng(w + h).integers(0, 256, (h, w), dtype=np.uint8)
    want, g = numpy_lowres(y, w, h)
    got = ol.oracle_lowres_init(y, w, h).reshape(4, g["lh"] + 64, g["lstride"])
    assert np.array_equal(got[:, :, :g["lw"] + 64], want[:, :, :g["lw"] + 64])


def test_lowres_geometry_of_baseline_configs():
    g = ol.lowres_geometry(1920, 1080)
    assert (g["mb_w"], g["mb_h"], g["luma_h"], g["lw"], g["lh"]) == (120, 68, 1088, 960, 544)   # C1
    g = ol.lowres_geometry(1280, 720)
    assert (g["mb_w"], g["mb_h"], g["lw"], g["lh"]) == (80, 45, 640, 360)
    g = ol.lowres_geometry(3840, 2160)
    assert (g["mb_w"], g["mb_h"], g["lw"], g["lh"]) == (240, 135, 1920, 1080)


def test_lowres_constant_and_padding_properties():
    w, h = 130, 70
    g = ol.lowres_geometry(w, h)
    out = ol.oracle_lowres_init(np.full((h, w), 77, np.uint8), w, h).reshape(4, g["lh"] + 64, g["lstride"])
    assert (out[:, :, :g["lw"] + 64] == 77).all()
    y = np.random.default_rng(3).integers(0, 256, (h, w), dtype=np.uint8)
    out = ol.oracle_lowres_init(y, w, h).reshape(4, g["lh"] + 64, g["lstride"])
    # every padding pixel equals the nearest interior pixel
    for k in range(4):
        inner = out[k, 32:32 + g["lh"], 32:32 + g["lw"]]
        assert np.array_equal(out[k, :, :g["lw"] + 64], np.pad(inner, 32, mode="edge"))


@pytest.mark.parametrize("size", [(64, 48), (66, 50), (34, 18), (1920, 1080)])
def test_chroma_nv12_pad_oracle_against_numpy(size):
    """x264_frame_copy_picture (4:2:0 chroma -> NV12) + expand_border_mod16, independent numpy formulation."""
    w, h = size
    g = ol.lowres_geometry(w, h)
    rng = np.random.default_rng(w * 31 + h)
    u = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
    v = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
    got = ol.oracle_chroma_nv12_pad(u, v, w, h).reshape(g["luma_h"] // 2, g["luma_w"])
    rows = np.minimum(np.arange(g["luma_h"] // 2), h // 2 - 1)
    cols = np.minimum(np.arange(g["luma_w"] // 2), w // 2 - 1)
    want = np.empty_like(got)
    want[:, 0::2] = u[rows][:, cols]
    want[:, 1::2] = v[rows][:, cols]
    assert np.array_equal(got, want)
