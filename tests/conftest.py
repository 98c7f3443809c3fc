import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The oracle is test infrastructure: (re)build it when a compiler is around."""
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".c", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    # the product library: a stale binary must never be what a test run measures (the GPU box gets the in-tree .so as it is)
    lib = os.path.join(ROOT, "x264vfw_b200", "libx264vfw_cuda.so")
    csrc = os.path.join(ROOT, "x264vfw_b200", "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h"))] + [os.path.join(ROOT, "include", "x264vfw_cuda.h")]
    if os.path.exists(lib) and any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            try:
                subprocess.run(["make", "-j8", "-C", csrc], check=True, capture_output=True)
                subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
            except Exception as e:                      # never fatal here: the tests then run (and say so) on the binary that is there
                print(f"warning: could not rebuild the stale product library: {e}", file=sys.stderr)
    yield
