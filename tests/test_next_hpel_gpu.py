"""(SURVEY 8f rows run after the main path: this file sorts behind test_lowres_gpu / test_lookahead_gpu.)
GPU parity: the half-pel reference plane kernel through the C ABI vs the CPU checker,
bit-exact on all four padded planes including every border byte."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from x264vfw_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


def _content(w, h, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "extreme":                      # drives every 6-tap sum and the centre intermediate to its limits
        return (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)
    return np.tile((np.arange(w) % 2 * 255).astype(np.uint8), (h, 1))


@pytest.mark.parametrize("size", [(16, 16), (64, 48), (112, 18), (240, 64), (1280, 720), (1920, 1088), (3840, 2160), (248, 12), (24, 6)])
@pytest.mark.parametrize("kind", ["noise", "extreme"])
def test_hpel_filter_matches_checker(ctx, size, kind):
    import torch
    from x264vfw_b200 import hpel
    w, h = size
    nf = 2 if w * h < 4_000_000 else 1
    g = hpel.geometry(w, h)
    og = ol.hpel_geometry(w, h)
    assert (g.stride, g.plane_bytes, g.origin) == (og["stride"], og["plane_bytes"], og["origin"])
    sfb = (w * h + 255) // 256 * 256
    dfb = 4 * g.plane_bytes
    frames = [_content(w, h, kind, 11 * w + h + f) for f in range(nf)]
    host = np.zeros(nf * sfb, dtype=np.uint8)
    for f, y in enumerate(frames):
        host[f * sfb:f * sfb + w * h] = y.reshape(-1)
    d_src = torch.from_numpy(host).cuda()
    d_out = torch.full((nf * dfb,), 0x5A, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    hpel.hpel_filter(ctx, d_out.data_ptr(), d_src.data_ptr(), w, w, h, sfb, dfb, nf)
    ctx.sync()
    got = d_out.cpu().numpy().reshape(nf, 4, h + 64, g.stride)
    for f, y in enumerate(frames):
        want = ol.oracle_hpel_planes(y, w, h)
        for p in range(4):
            assert np.array_equal(got[f, p, :, :w + 64], want[p, :, :w + 64]), (size, kind, f, p)
    assert np.all(got[:, :, :, w + 64:] == 0x5A)


def test_hpel_filter_unaligned_source_and_big_batch(ctx):
    """Source plane at an odd address / odd stride (byte gathers) and a batch large enough for 24-row strips."""
    import torch
    from x264vfw_b200 import hpel
    w, h, ss, off = 64, 32, 67, 1
    g = hpel.geometry(w, h)
    nf = 40
    sfb = ss * h + 5
    rng = np.random.default_rng(3)
    host = rng.integers(0, 256, nf * sfb + 8, dtype=np.uint8)
    d_src = torch.from_numpy(host).cuda()
    d_out = torch.zeros(nf * 4 * g.plane_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    hpel.hpel_filter(ctx, d_out.data_ptr(), d_src.data_ptr() + off, ss, w, h, sfb, 4 * g.plane_bytes, nf)
    ctx.sync()
    got = d_out.cpu().numpy().reshape(nf, 4, h + 64, g.stride)
    for f in (0, 17, nf - 1):
        y = np.stack([host[off + f * sfb + r * ss: off + f * sfb + r * ss + w] for r in range(h)])
        want = ol.oracle_hpel_planes(y, w, h)
        assert np.array_equal(got[f, :, :, :w + 64], want[:, :, :w + 64]), f

    # 1080p batch: 24-row strips
    w, h = 1920, 1088
    g = hpel.geometry(w, h)
    nf = 8
    sfb = w * h
    host = rng.integers(0, 256, nf * sfb, dtype=np.uint8)
    d_src = torch.from_numpy(host).cuda()
    d_out = torch.zeros(nf * 4 * g.plane_bytes, dtype=torch.uint8, device="cuda")
    hpel.hpel_filter(ctx, d_out.data_ptr(), d_src.data_ptr(), w, w, h, sfb, 4 * g.plane_bytes, nf)
    ctx.sync()
    got = d_out.cpu().numpy().reshape(nf, 4, h + 64, g.stride)
    for f in (0, nf - 1):
        want = ol.oracle_hpel_planes(host[f * sfb:(f + 1) * sfb].reshape(h, w), w, h)
        assert np.array_equal(got[f, :, :, :w + 64], want[:, :, :w + 64]), f


def test_device_path_matches_the_committed_fingerprints(ctx):
    """The device path against tests/golden/hpel_golden.json, without the checker in the loop."""
    import importlib.util
    import json
    import os
    import torch
    from x264vfw_b200 import hpel
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_hpel_golden", os.path.join(gdir, "make_hpel_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for case in json.load(open(os.path.join(gdir, "hpel_golden.json"))):
        w, h = case["w"], case["h"]
        y = mg.source_plane(w, h)
        g = hpel.geometry(w, h)
        d_src = torch.from_numpy(y.reshape(-1)).cuda()
        d_out = torch.zeros(4 * g.plane_bytes, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        hpel.hpel_filter(ctx, d_out.data_ptr(), d_src.data_ptr(), w, w, h)
        ctx.sync()
        got = d_out.cpu().numpy().reshape(4, h + 64, g.stride)
        assert mg.fingerprint(got, w) == case["planes_fnv"], (w, h)


@pytest.mark.parametrize("kind", ["noise", "extreme"])
def test_device_planes_reproduce_the_h264_decoders_motion_compensation(ctx, kind):
    """The device's four half-pel planes, read at every quarter-sample phase (get_ref), against the pictures
    libavcodec's H.264 decoder produced from the same source (tests/golden/h264_pins.json): a reference that is
    not the checker."""
    import json
    import os
    import torch
    import h264_pins as hp
    from x264vfw_b200 import hpel
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h264_pins.json")))
    w, h = hp.MC_W, hp.MC_H
    y, _, _ = hp.mc_picture(kind)
    g = hpel.geometry(w, h)
    d_src = torch.from_numpy(np.ascontiguousarray(y).reshape(-1)).cuda()
    d_out = torch.zeros(4 * g.plane_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    hpel.hpel_filter(ctx, d_out.data_ptr(), d_src.data_ptr(), w, w, h)
    ctx.sync()
    planes = d_out.cpu().numpy().reshape(4, h + 64, g.stride)
    assert hp.checker_mc_hashes(kind, planes=planes) == gold["mc"]["pictures"][kind]


def test_device_planes_reproduce_the_h264_decoder_at_1080p(ctx):
    """1920x1088, six vectors: every tile and strip of the kernel at BASELINE's frame size against the decoder's
    pictures (tests/golden/h264_pins.json)."""
    import json
    import os
    import torch
    import h264_pins as hp
    from x264vfw_b200 import hpel
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h264_pins.json")))
    w, h = hp.HD_W, hp.HD_H
    g = hpel.geometry(w, h)
    d_src = torch.from_numpy(np.ascontiguousarray(hp.hd_picture()).reshape(-1)).cuda()
    d_out = torch.zeros(4 * g.plane_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    hpel.hpel_filter(ctx, d_out.data_ptr(), d_src.data_ptr(), w, w, h)
    ctx.sync()
    planes = d_out.cpu().numpy().reshape(4, h + 64, g.stride)
    assert hp.checker_hd_hashes(planes=planes) == gold["hd"]["pictures"]


def test_hpel_filter_rejects_bad_geometry(ctx):
    from x264vfw_b200 import hpel
    from x264vfw_b200._lib import CudaError
    with pytest.raises(CudaError):
        hpel.hpel_filter(ctx, 256, 256, 20, 20, 16)
