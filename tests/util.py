"""Shared helpers for the parity tests."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"

# north_star tolerances (BASELINE.json): relative L2 error ||a-b|| / ||b||
TOL_OUT_FP32 = 1e-3
TOL_GRAD_FP32 = 5e-3
TOL_OUT_BF16 = 2e-2
MASK_AGREEMENT = 0.999


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def load_golden(name: str):
    z = np.load(GOLDEN_DIR / f"towerunet_{name}.npz")
    cfg = {k: int(v) for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
    cfg["dilations"] = [int(d) for d in z["dilations"]]
    return cfg, z


def golden_spec(z):
    """Parameter inventory stored in a variant fixture (None for the default architecture: the port's inventory applies)."""
    if "spec_names" not in z:
        return None
    return [(str(n), tuple(int(v) for v in str(s).split(",")) if str(s) else ()) for n, s in zip(z["spec_names"], z["spec_shapes"])]


def mine_from_state_dict(cfg: dict, sd: dict, device: str, dtype=torch.float32):
    import cultionet_b200 as cb
    from oracle.make_golden import variant_kwargs

    m = cb.TowerUNet(in_channels=cfg["C"], in_time=cfg["T"], hidden_channels=cfg["hidden"], dilations=cfg["dilations"], compute_dtype=dtype,
                     **variant_kwargs(cfg))
    m.load_state_dict(sd, strict=True)
    return m.to(device)
