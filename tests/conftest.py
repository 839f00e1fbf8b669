"""pytest configuration: `gpu` marker, import path, in-tree build of libcspe.so."""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def libcspe_path():
    """Path of the built C-ABI library; (re)builds it when nvcc is around and sources changed."""
    from constructionsceneposeestimation_b200 import build

    try:
        return build.build_library()
    except RuntimeError:
        if build.LIB_PATH.exists():  # no nvcc on this machine: use the shipped .so
            return build.LIB_PATH
        raise


@pytest.fixture(scope="session")
def cuda_device(libcspe_path):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def reference_available() -> bool:
    from oracle import reference_extract

    return reference_extract.available()
