"""Half-pel reference planes, CPU side:
 * the C checker (oracle/hpel_oracle.c: upstream's literal sequence) against an independent numpy formulation
   (clamped coordinates, no materialised borders);
 * the CUDA warp program itself (x264vfw_b200/csrc/hpel_kernel.cuh), compiled by g++ against the lockstep warp
   shim in tests/sim/ and run on the CPU, against the checker -- same source as the sm_100a build, so index
   arithmetic, packed-lane biases and border ownership are verified before the kernel reaches a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frames(w, h, seed):
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, (h, w), dtype=np.uint8)
    # saturating content: 0/255 blocks drive the 6-tap sums to both clip() limits and the centre
    # plane's intermediate to its extremes
    extreme = (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)
    cols = np.tile((np.arange(w) % 2 * 255).astype(np.uint8), (h, 1))
    return [noise, extreme, cols]


@pytest.mark.parametrize("size", [(16, 16), (64, 48), (128, 32), (240, 64), (20, 6)])
def test_checker_matches_numpy_formulation(size):
    w, h = size
    for i, y in enumerate(_frames(w, h, 7 * w + h)):
        want = ol.numpy_hpel_planes(y)
        got = ol.oracle_hpel_planes(y, w, h)[:, :, :w + 64]
        assert np.array_equal(got, want), (size, i)


def test_checker_plane0_is_the_border_expanded_frame():
    w, h = 32, 16
    y = np.arange(w * h, dtype=np.uint32).astype(np.uint8).reshape(h, w)
    p0 = ol.oracle_hpel_planes(y, w, h)[0, :, :w + 64]
    assert np.array_equal(p0, np.pad(y, 32, mode="edge"))


def test_checker_matches_the_committed_fingerprints():
    """tests/golden/hpel_golden.json freezes the checker (it is not a reference output: libx264 is absent)."""
    import json
    sys_path = os.path.join(ROOT, "tests", "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_hpel_golden", os.path.join(sys_path, "make_hpel_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for case in json.load(open(os.path.join(sys_path, "hpel_golden.json"))):
        w, h = case["w"], case["h"]
        y = mg.source_plane(w, h)
        assert mg.fnv(y) == case["src_fnv"], (w, h)
        assert mg.fingerprint(ol.oracle_hpel_planes(y, w, h), w) == case["planes_fnv"], (w, h)


@pytest.fixture(scope="module")
def sim():
    so = os.path.join(ROOT, "tests", "sim", "_build", "libhpel_sim.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    srcs = [os.path.join(ROOT, "tests", "sim", "hpel_sim.cpp"), os.path.join(ROOT, "tests", "sim", "warp_sim.h"),
            os.path.join(ROOT, "x264vfw_b200", "csrc", "hpel_kernel.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so, srcs[0],
                        "-lpthread"], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.sim_hpel.restype = C.c_int
    lib.sim_hpel.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int,
                             C.c_size_t, C.c_size_t, C.c_int]
    return lib


# (w, h, rows_per_strip, source stride, source offset): one tile; the word right of the frame exactly in the halo
# lane 31 (240, 480); a second tile that holds one word (248); strips that end inside / beyond the frame; an
# unaligned source (byte gathers); frames smaller than the filter's support; three tiles (the middle one runs the
# interior-tile loop) with odd heights / strip lengths (the two-rows-per-trip loop computes one row too many)
@pytest.mark.parametrize("case", [(16, 16, 12, 16, 0), (64, 48, 12, 64, 0), (112, 18, 6, 112, 0),
                                  (128, 32, 24, 128, 0), (240, 20, 12, 240, 0), (248, 12, 12, 248, 0),
                                  (480, 8, 6, 480, 0), (64, 16, 12, 67, 1), (24, 6, 12, 24, 0), (8, 3, 12, 8, 0),
                                  (600, 10, 12, 600, 0), (720, 9, 5, 720, 0), (520, 7, 24, 523, 3)])
def test_warp_program_on_cpu_matches_checker(sim, case):
    w, h, rps, ss, off = case
    g = ol.hpel_geometry(w, h)
    nf = 2
    frames = _frames(w, h, w + 3 * h)[:nf]
    sfb = ss * h + 16
    src = np.zeros(nf * sfb + 8, dtype=np.uint8)
    for f, y in enumerate(frames):
        for r in range(h):
            src[off + f * sfb + r * ss: off + f * sfb + r * ss + w] = y[r]
    dfb = 4 * g["plane_bytes"]
    dst = np.full(nf * dfb, 0x5A, dtype=np.uint8)
    units = sim.sim_hpel(dst.ctypes.data, src.ctypes.data + off, ss, w, h, g["stride"], g["plane_bytes"], rps,
                         sfb, dfb, nf)
    assert units > 0
    got = dst.reshape(nf, 4, h + 64, g["stride"])
    for f, y in enumerate(frames):
        want = ol.oracle_hpel_planes(y, w, h)
        for p in range(4):
            assert np.array_equal(got[f, p, :, :w + 64], want[p, :, :w + 64]), (case, f, p)
        # nothing outside the w+64 columns is touched
        assert np.all(got[f, :, :, w + 64:] == 0x5A)


def test_warp_program_on_cpu_random_geometry(sim):
    """Random widths (multiples of 8, up to four tiles), heights, strip lengths and source alignments.
    (A one-off soak of 300 such cases, up to five tiles and two frames per launch, also matched.)"""
    rng = np.random.default_rng(20261017)
    for _ in range(32):
        w = 8 * int(rng.integers(1, 100))
        h = int(rng.integers(1, 40))
        rps = int(rng.integers(1, 30))
        off = int(rng.choice([0, 0, 0, 8, 3]))
        ss = w + int(rng.choice([0, 0, 8, 24, 5]))
        if off % 8 or ss % 8:
            w = min(w, 96)                      # the byte-gather path is slow in the simulation
        g = ol.hpel_geometry(w, h)
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if rng.integers(0, 3) == 0:
            y = (y > 127).astype(np.uint8) * 255
        src = np.zeros(ss * h + 32, dtype=np.uint8)
        base = src.ctypes.data
        pad = (-base) % 8                       # make the buffer itself 8-byte aligned, then apply `off`
        for r in range(h):
            src[pad + off + r * ss: pad + off + r * ss + w] = y[r]
        dst = np.full(4 * g["plane_bytes"], 0x5A, dtype=np.uint8)
        sim.sim_hpel(dst.ctypes.data, base + pad + off, ss, w, h, g["stride"], g["plane_bytes"], rps, ss * h, 4 * g["plane_bytes"], 1)
        got = dst.reshape(4, h + 64, g["stride"])
        want = ol.oracle_hpel_planes(y, w, h)
        assert np.array_equal(got[:, :, :w + 64], want[:, :, :w + 64]), (w, h, rps, off, ss)
        assert np.all(got[:, :, w + 64:] == 0x5A)


@pytest.mark.parametrize("kind", ["noise", "extreme"])
def test_warp_program_on_cpu_reproduces_the_h264_decoders_motion_compensation(sim, kind):
    """The kernel's source, run in lockstep on the CPU, against libavcodec's H.264 decoder (tests/golden/h264_pins.json):
    its four planes read at every quarter-sample phase give the decoder's motion-compensated pictures."""
    import json
    import h264_pins as hp
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "h264_pins.json")))
    w, h = hp.MC_W, hp.MC_H
    y, _, _ = hp.mc_picture(kind)
    y = np.ascontiguousarray(y)
    g = ol.hpel_geometry(w, h)
    dst = np.zeros(4 * g["plane_bytes"], dtype=np.uint8)
    sim.sim_hpel(dst.ctypes.data, y.ctypes.data, w, w, h, g["stride"], g["plane_bytes"], 12, w * h, 4 * g["plane_bytes"], 1)
    assert hp.checker_mc_hashes(kind, planes=dst.reshape(4, h + 64, g["stride"])) == gold["mc"]["pictures"][kind]
