"""TEST INFRASTRUCTURE ONLY -- import the *real* reference hot-path modules from /root/reference.

Works only in the authoring container (``/root/reference`` does not exist on the GPU box).
Used by ``oracle/make_golden.py`` to produce the committed fixtures under ``tests/golden/`` and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference tree is absent) to pin the
torch restatement in ``oracle/towerunet_port.py``.

Recipe:
  1. register an empty ``cultionet`` package whose ``__path__`` points at the reference sources so the
     heavyweight ``cultionet/__init__.py`` (lightning, rasterio, geowombat ...) never executes;
  2. register a minimal ``cultionet.data`` exposing a plain ``Data`` container (the real
     ``data/data.py:1-19`` imports geowombat / pyproj / rasterio / xarray, none installed);
  3. register ``natten`` / ``natten.functional`` stubs backed by ``oracle/natten_ref.py``;
  4. ``TORCHDYNAMO_DISABLE=1`` -- ``nunet.py:141`` wraps ``PreTimeReduction`` in ``torch.compile``
     and inductor cannot build on this image's CPU toolchain.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from pathlib import Path

REFERENCE_SRC = Path(os.environ.get("CULTIONET_REFERENCE_SRC", "/root/reference/src/cultionet"))


def reference_available() -> bool:
    return (REFERENCE_SRC / "models" / "nunet.py").is_file()


class _Data:
    """Field contract of ``cultionet.data.Data`` (``data/data.py:51-139``) without the geo imports."""

    def __init__(self, x, y=None, **kwargs):
        self.x = x
        self.y = y
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_samples(self) -> int:
        return self.x.shape[0]


def load_reference():
    """Returns a namespace with the reference's TowerUNet, CultioNet, losses and enums."""
    if not reference_available():
        raise FileNotFoundError(f"reference sources not found at {REFERENCE_SRC}")
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    here = Path(__file__).resolve().parent
    if str(here.parent) not in sys.path:
        sys.path.insert(0, str(here.parent))
    from oracle import natten_ref

    if "cultionet" not in sys.modules or not getattr(sys.modules["cultionet"], "_oracle_shell", False):
        pkg = types.ModuleType("cultionet")
        pkg.__path__ = [str(REFERENCE_SRC)]
        pkg._oracle_shell = True
        sys.modules["cultionet"] = pkg

        data_mod = types.ModuleType("cultionet.data")
        data_mod.Data = _Data
        data_mod.__path__ = []
        sys.modules["cultionet.data"] = data_mod
        pkg.data = data_mod

        nat = types.ModuleType("natten")
        nat.NeighborhoodAttention2D = natten_ref.NeighborhoodAttention2D
        natf = types.ModuleType("natten.functional")
        natf.na2d = natten_ref.na2d
        natf.na2d_qk = natten_ref.na2d_qk
        natf.na2d_av = natten_ref.na2d_av
        nat.functional = natf
        sys.modules["natten"] = nat
        sys.modules["natten.functional"] = natf

    ns = types.SimpleNamespace()
    ns.nunet = importlib.import_module("cultionet.models.nunet")
    ns.cultionet_model = importlib.import_module("cultionet.models.cultionet")
    ns.losses = importlib.import_module("cultionet.losses.losses")
    ns.enums = importlib.import_module("cultionet.enums")
    ns.nn = importlib.import_module("cultionet.nn")
    ns.unet_parts = importlib.import_module("cultionet.nn.modules.unet_parts")
    ns.TowerUNet = ns.nunet.TowerUNet
    ns.CultioNet = ns.cultionet_model.CultioNet
    ns.TanimotoComplementLoss = ns.losses.TanimotoComplementLoss
    ns.TanimotoDistLoss = ns.losses.TanimotoDistLoss
    ns.Data = _Data
    return ns
