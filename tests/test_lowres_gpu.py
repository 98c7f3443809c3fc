"""GPU parity: lowres_init / luma_pad kernels through the C ABI vs the CPU oracle, bit-exact,
including all padding bytes."""
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


@pytest.mark.parametrize("size", [(64, 48), (66, 50), (18, 2), (1280, 720), (1920, 1080), (3840, 2160), (1366, 768)])
def test_lowres_init_matches_oracle(ctx, size):
    import torch
    from x264vfw_b200 import lowres
    w, h = size
    nf = 2
    g = lowres.geometry(w, h)
    og = ol.lowres_geometry(w, h)
    assert all(getattr(g, k) == v for k, v in og.items())
    sfb = (w * h + 255) // 256 * 256
    dfb = 4 * g.lplane_bytes
    rng = np.random.default_rng(w)
    host = rng.integers(0, 256, nf * sfb, dtype=np.uint8)
    d_y = torch.from_numpy(host).cuda()
    d_out = torch.full((nf * dfb,), 0xAA, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    lowres.lowres_init(ctx, d_out.data_ptr(), d_y.data_ptr(), w, w, h, sfb, dfb, nf)
    ctx.sync()
    got = d_out.cpu().numpy().reshape(nf, 4, g.lh + 64, g.lstride)
    for f in range(nf):
        y = host[f * sfb:f * sfb + w * h].reshape(h, w)
        want = ol.oracle_lowres_init(y, w, h).reshape(4, g.lh + 64, g.lstride)
        assert np.array_equal(got[f][:, :, :g.lw + 64], want[:, :, :g.lw + 64]), f


@pytest.mark.parametrize("size", [(64, 48), (66, 50), (1920, 1080)])
def test_luma_pad_matches_oracle(ctx, size):
    import torch
    from x264vfw_b200 import lowres
    w, h = size
    g = lowres.geometry(w, h)
    y = np.random.default_rng(h).integers(0, 256, (h, w), dtype=np.uint8)
    d_y = torch.from_numpy(y).cuda()
    d_out = torch.zeros(g.luma_stride * (g.luma_h + 1), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    lowres.luma_pad(ctx, d_out.data_ptr(), d_y.data_ptr(), w, w, h)
    ctx.sync()
    got = d_out.cpu().numpy().reshape(g.luma_h + 1, g.luma_stride)[:, :g.luma_w + 1]
    want = ol.oracle_luma_pad(y, w, h).reshape(g.luma_h + 1, g.luma_stride)[:, :g.luma_w + 1]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("size", [(64, 48), (66, 50), (34, 18), (1920, 1080)])
def test_chroma_nv12_pad_matches_oracle(ctx, size):
    """[x264] x264_frame_copy_picture chroma (planar U, V -> NV12) + expand_border_mod16 on the device."""
    import torch
    from x264vfw_b200 import lowres
    w, h = size
    g = lowres.geometry(w, h)
    rng = np.random.default_rng(w + 7 * h)
    u = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
    v = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
    d_u, d_v = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
    d_out = torch.zeros(g.luma_w * (g.luma_h // 2), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    lowres.chroma_nv12_pad(ctx, d_out.data_ptr(), g.luma_w, d_u.data_ptr(), d_v.data_ptr(), w // 2, w, h)
    ctx.sync()
    got = d_out.cpu().numpy()
    want = ol.oracle_chroma_nv12_pad(u, v, w, h)
    assert np.array_equal(got, want)
