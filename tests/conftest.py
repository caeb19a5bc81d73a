import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emu_lib():
    """host build of the kernel bodies with a serial thread model (see aletsch_b200/csrc/dev.h):
    test-only, lets the CPU tier check kernel logic; the product never loads it."""
    if os.environ.get("ALETSCH_EMU_LIB"):          # e.g. an AddressSanitizer build of the same sources (tools/asan_emu.sh)
        return os.environ["ALETSCH_EMU_LIB"]
    out = os.path.join(ROOT, "tests", "emu", "libaletsch_emu.so")
    src = os.path.join(ROOT, "aletsch_b200", "csrc")
    deps = [os.path.join(src, f) for f in os.listdir(src)] + [os.path.join(ROOT, "include", "aletsch_gpu.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DAGPU_EMU", "-x", "c++", "-w", "-pthread",
                               "-o", out, os.path.join(src, "aletsch_gpu.cu")])
    return out


@pytest.fixture(scope="session")
def checkers():
    """the CPU checkers that exist here: the reference build (oracle/_ref) and the restatement."""
    import orclib
    out = {}
    for p in ("ref", "orc"):
        try:
            out[p] = orclib.Checker(p)
        except (FileNotFoundError, OSError):
            pass
    return out
