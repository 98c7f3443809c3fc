"""Deterministic synthetic clips for the csp + lookahead path.

Integer-only content so every platform produces the same bytes:
  * per scene s: a value-noise texture (4 octaves, cells 64/32/16/4 px) plus a 32-px checker,
    panned by (2, 1) px/frame;
  * 6 moving 96x64 rectangles (speeds +-1..+-5 px/frame) per scene;
  * uniform noise of +-2 per sample (seeded per frame);
  * hard scene cuts at the frame indices in `cuts` and a 2-frame white flash at `flash`
    (exercises scenecut and its flash rejection).
RGB is then packed into the input layouts the reference accepts (codec.c:187-231):
bottom-up BGRA/BGR DIBs, YUY2, UYVY, or planar YV12 (BT.601 TV-range via integer maths).
"""
import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def _hash32(a):
    """xorshift-multiply integer hash on uint64 arrays, reduced to 32 bits (vectorised)."""
    a = (a.astype(np.uint64) * np.uint64(0x9E3779B1)) & _M32
    a ^= a >> np.uint64(15)
    a = (a * np.uint64(0x85EBCA77)) & _M32
    a ^= a >> np.uint64(13)
    a = (a * np.uint64(0xC2B2AE3D)) & _M32
    a ^= a >> np.uint64(16)
    return a


def _value_noise(h, w, cell, seed):
    """8-bit value noise: lattice of hashed bytes, integer bilinear interpolation."""
    gy, gx = h // cell + 2, w // cell + 2
    iy, ix = np.meshgrid(np.arange(gy, dtype=np.uint64), np.arange(gx, dtype=np.uint64), indexing="ij")
    lat = (_hash32(iy * np.uint64(7919) + ix * np.uint64(104729) + np.uint64(seed)) & np.uint64(0xFF)).astype(np.int32)
    y = np.arange(h, dtype=np.int32)
    x = np.arange(w, dtype=np.int32)
    y0, fy = y // cell, (y % cell)[:, None]
    x0, fx = x // cell, (x % cell)[None, :]
    a = lat[y0][:, x0]
    b = lat[y0][:, x0 + 1]
    c = lat[y0 + 1][:, x0]
    d = lat[y0 + 1][:, x0 + 1]
    top = a * (cell - fx) + b * fx
    bot = c * (cell - fx) + d * fx
    return ((top * (cell - fy) + bot * fy) // (cell * cell)).astype(np.int32)


class SyntheticClip:
    """frame(n) -> (h, w, 3) uint8 RGB; packers produce the reference's input layouts."""

    def __init__(self, width, height, n_frames=300, stream_id=0, cuts=(100, 200), flash=150, flash_len=2):
        self.w, self.h, self.n = width, height, n_frames
        self.seed = 0x264 + stream_id
        self.cuts = tuple(c for c in cuts if 0 < c < n_frames)
        self.flash = flash
        self.flash_len = flash_len
        self._tex = {}

    def scene_of(self, n):
        return sum(1 for c in self.cuts if n >= c)

    def _scene_start(self, s):
        return 0 if s == 0 else self.cuts[s - 1]

    def _texture(self, s):
        if s not in self._tex:
            self._tex.clear()               # keep one scene resident
            start = self._scene_start(s)
            end = self.cuts[s] if s < len(self.cuts) else self.n
            span = max(1, end - start)
            th, tw = self.h + span + 1, self.w + 2 * span + 2
            chans = []
            for ch in range(3):
                base = self.seed * 977 + s * 131 + ch * 17
                v = (_value_noise(th, tw, 64, base) >> 1) + (_value_noise(th, tw, 32, base + 1) >> 2) + \
                    (_value_noise(th, tw, 16, base + 2) >> 3) + (_value_noise(th, tw, 4, base + 3) >> 2)
                yy, xx = np.meshgrid(np.arange(th), np.arange(tw), indexing="ij")
                v = v + ((((xx + 16 * s) >> 5) + ((yy + 48 * (s & 1)) >> 5) + s) & 1) * 40 + 8
                chans.append(np.clip(v, 0, 255).astype(np.uint8))
            rects = []
            hh = _hash32(np.arange(6 * 8, dtype=np.uint64) + np.uint64(self.seed * 31 + s * 7))
            for r in range(6):
                q = [int(v) for v in hh[r * 8:(r + 1) * 8]]
                rects.append(dict(x0=q[0] % max(1, self.w), y0=q[1] % max(1, self.h),
                                  vx=(q[2] % 5 + 1) * (1 if q[3] & 1 else -1), vy=(q[4] % 5 + 1) * (1 if q[3] & 2 else -1),
                                  col=(q[5] & 0xFF, q[6] & 0xFF, q[7] & 0xFF)))
            self._tex[s] = (np.stack(chans, axis=-1), rects, start)
        return self._tex[s]

    def frame(self, n):
        s = self.scene_of(n)
        tex, rects, start = self._texture(s)
        k = n - start
        img = tex[k:k + self.h, 2 * k:2 * k + self.w].astype(np.int16)
        rw, rh = min(96, self.w // 2), min(64, self.h // 2)
        for r in rects:
            x = (r["x0"] + r["vx"] * k) % self.w
            y = (r["y0"] + r["vy"] * k) % self.h
            img[y:min(self.h, y + rh), x:min(self.w, x + rw)] = r["col"]
        rng = np.random.Generator(np.random.PCG64(self.seed * 100003 + n))
        img += rng.integers(-2, 3, img.shape, dtype=np.int16)
        if self.flash is not None and self.flash <= n < self.flash + self.flash_len:
            img = img // 8 + 224
        return np.clip(img, 0, 255).astype(np.uint8)

    # ---- packers ---------------------------------------------------------------------------
    @staticmethod
    def pack_bgra_bottom_up(rgb):
        h, w, _ = rgb.shape
        out = np.empty((h, w, 4), dtype=np.uint8)
        out[..., 0] = rgb[..., 2]; out[..., 1] = rgb[..., 1]; out[..., 2] = rgb[..., 0]; out[..., 3] = 0
        return np.ascontiguousarray(out[::-1]).reshape(-1)

    @staticmethod
    def pack_bgr_bottom_up(rgb):
        h, w, _ = rgb.shape
        stride = (3 * w + 3) & ~3
        out = np.zeros((h, stride), dtype=np.uint8)
        out[:, :3 * w] = rgb[::-1, :, ::-1].reshape(h, 3 * w)
        return out.reshape(-1)

    @staticmethod
    def _yuv601tv(rgb):
        r, g, b = (rgb[..., i].astype(np.int32) for i in range(3))
        y = (66 * r + 129 * g + 25 * b + 128 >> 8) + 16
        u = (-38 * r - 74 * g + 112 * b + 128 >> 8) + 128
        v = (112 * r - 94 * g - 18 * b + 128 >> 8) + 128
        return (np.clip(a, 0, 255).astype(np.uint8) for a in (y, u, v))

    @classmethod
    def pack_422(cls, rgb, uyvy=False):
        y, u, v = cls._yuv601tv(rgb)
        h, w = y.shape
        u2 = ((u[:, 0::2].astype(np.int32) + u[:, 1::2] + 1) >> 1).astype(np.uint8)
        v2 = ((v[:, 0::2].astype(np.int32) + v[:, 1::2] + 1) >> 1).astype(np.uint8)
        out = np.empty((h, w // 2, 4), dtype=np.uint8)
        if uyvy:
            out[..., 0] = u2; out[..., 1] = y[:, 0::2]; out[..., 2] = v2; out[..., 3] = y[:, 1::2]
        else:
            out[..., 0] = y[:, 0::2]; out[..., 1] = u2; out[..., 2] = y[:, 1::2]; out[..., 3] = v2
        return out.reshape(-1)

    @classmethod
    def pack_yv12(cls, rgb):
        y, u, v = cls._yuv601tv(rgb)
        box = lambda c: ((c[0::2, 0::2].astype(np.int32) + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2] + 2) >> 2).astype(np.uint8)
        return np.concatenate([y.reshape(-1), box(v).reshape(-1), box(u).reshape(-1)])

    def packed(self, n, fmt):
        rgb = self.frame(n)
        if fmt == "bgra":
            return self.pack_bgra_bottom_up(rgb)
        if fmt == "bgr":
            return self.pack_bgr_bottom_up(rgb)
        if fmt == "yuyv":
            return self.pack_422(rgb, False)
        if fmt == "uyvy":
            return self.pack_422(rgb, True)
        if fmt == "yv12":
            return self.pack_yv12(rgb)
        raise ValueError(fmt)
