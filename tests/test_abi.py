"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls: no GPU needed)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    names = set()
    for h in (ROOT / "include").glob("*.h"):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names.update(re.findall(r"\b(cnb_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_bound_functions():
    from cultionet_b200 import _lib

    declared = set(declared_symbols())
    bound = set(_lib.EXPORTED_SYMBOLS)
    assert bound <= declared, sorted(bound - declared)
    assert declared <= bound, sorted(declared - bound)


def test_cuda_library_exports_every_declared_symbol():
    from cultionet_b200 import _lib
    from cultionet_b200.build import build

    try:
        path = build()
    except Exception as e:  # no nvcc on this box
        if _lib.DEFAULT_LIB.is_file():
            path = _lib.DEFAULT_LIB
        else:
            pytest.skip(f"cannot build the CUDA library here: {e}")
    lib = ctypes.CDLL(str(path))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/ but not exported by {path.name}"
    lib.cnb_version.restype = ctypes.c_int
    lib.cnb_sm_arch.restype = ctypes.c_int
    assert lib.cnb_version() == 100
    assert lib.cnb_sm_arch() == 100


def test_product_path_fails_loudly_without_cuda_tensors():
    import torch

    from cultionet_b200 import _lib
    from cultionet_b200 import functional as F

    if not _lib.DEFAULT_LIB.is_file():
        pytest.skip("CUDA library not built")
    _lib.use_library(_lib.DEFAULT_LIB)
    with pytest.raises(_lib.CnbError, match="no CPU fallback"):
        F.layernorm(torch.zeros(1, 2, 2, 4), torch.ones(4), torch.zeros(4))


def test_missing_library_raises(tmp_path):
    from cultionet_b200 import _lib

    with pytest.raises(_lib.CnbError):
        _lib.use_library(tmp_path / "nope.so")
