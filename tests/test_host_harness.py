"""The plain-C host harness (host/x264vfw_harness.c) compiles against the C ABI without CUDA
headers; on a GPU it runs a clip end to end through compress_begin/compress/compress_end."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    return os.path.join(ROOT, "host", "x264vfw_harness")


def test_harness_builds_with_plain_gcc():
    exe = build()
    assert os.path.exists(exe)
    # object code must not need anything but the C ABI library
    out = subprocess.run(["gcc", "-std=c99", "-pedantic", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "x264vfw_cuda.h")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


@pytest.mark.gpu
def test_harness_runs_a_clip():
    exe = build()
    out = subprocess.run([exe, "320", "192", "40", "medium"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("frame ")]
    assert sorted(int(l.split()[1]) for l in lines) == list(range(40))
    assert lines[0].split()[1] == "0" and lines[0].split()[3] == "1"       # first coded frame is the IDR
