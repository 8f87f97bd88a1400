import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(autouse=True)
def _bind_backend(request):
    """GPU tests run on the nvcc-built library; everything else binds the CPU kernel interpreter (tests/emu)."""
    from cultionet_b200 import _lib

    if request.node.get_closest_marker("gpu") is not None:
        import torch

        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        _lib.use_library(_lib.DEFAULT_LIB)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        yield "cuda"
    else:
        from tests.emu.build_emu import build

        _lib.use_library(build())
        yield "cpu"


@pytest.fixture
def dev(_bind_backend):
    return _bind_backend
