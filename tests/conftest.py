import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _ensure_built():
    """Build the product library and the oracle where a toolchain exists (the GPU box uses the shipped .so files)."""
    so = os.path.join(ROOT, "lut-renderer_b200", "liblutr.so")
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        subprocess.run([sys.executable, os.path.join(ROOT, "lut-renderer_b200", "build.py")], check=True)
    elif not os.path.exists(so):
        raise RuntimeError("liblutr.so missing and no nvcc to build it")
    if shutil.which("make") and shutil.which("gcc"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    _ensure_built()
    yield


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
