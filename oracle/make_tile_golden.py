"""TEST INFRASTRUCTURE ONLY -- write tests/golden/tile_reference.npz by running the REAL reference code either side of ``predict_step``
(``oracle/ref_tile_loader.py``: ``BatchStore.__setitem__`` / ``write_batch``, ``LightningGTiffWriter.write_on_batch_end``,
``NormValues.transform``) on seeded inputs.  Run in the authoring container:  python -m oracle.make_tile_golden

The fixture pins ``oracle/tile_port.py`` (and through it the ``cnb_window_load`` / ``cnb_predict_pack`` kernels) on machines where
``/root/reference`` does not exist.  Inputs are re-drawn from the seed by ``tile_golden_inputs``.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CASE = dict(T=3, C=2, H=53, W=47, window_size=16, padding=4, seed=77)


def tile_golden_inputs(case=CASE):
    rng = np.random.default_rng(case["seed"])
    tile = rng.integers(-200, 11000, size=(case["T"], case["C"], case["H"], case["W"])).astype(np.int16)
    mean = torch.from_numpy(rng.uniform(0.2, 0.6, size=case["C"]).astype(np.float32))
    std = torch.from_numpy(rng.uniform(0.1, 0.4, size=case["C"]).astype(np.float32))
    return tile, mean, std


def prediction_for(windows, size, seed):
    """Seeded [n,1,size,size] predictions with values beyond [0, 1] so that the clip of callbacks.py:220 is exercised."""
    rng = np.random.default_rng(seed)
    n = len(windows)
    return {k: torch.from_numpy(rng.uniform(-0.05, 1.05, size=(n, 1, size, size)).astype(np.float32)) for k in ("distance", "edge", "crop")}


def main() -> None:
    from oracle import tile_port
    from oracle.ref_tile_loader import load_tile_reference, reference_store_window, reference_write_windows

    ns = load_tile_reference()
    tile, mean, std = tile_golden_inputs()
    ws, pad = CASE["window_size"], CASE["padding"]
    H, W = tile.shape[-2:]
    size = ws + 2 * pad
    # chunks with their halo (dask map_overlap(depth=padding, boundary=0, trim=False) -- third-party, restated in tile_port) handed to
    # the REAL BatchStore together with the chunk's region in TILE coordinates
    padded = np.pad(tile, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    xs, fields = [], []
    for r0, h in tile_port.chunk_starts(H, ws):
        for c0, w in tile_port.chunk_starts(W, ws):
            item = padded[:, :, r0:r0 + h + 2 * pad, c0:c0 + w + 2 * pad]
            b = reference_store_window(ns, item, slice(r0, r0 + h), slice(c0, c0 + w), ws, pad)
            xs.append(b.x.numpy())
            fields.append([b.window_row_off[0], b.window_col_off[0], b.window_height[0], b.window_width[0], b.padding[0]])
    x_int = np.concatenate(xs, axis=0)  # [n, C, T, size, size] int32
    fields = np.array(fields, dtype=np.int64)
    # load-time arithmetic: datasets.py:443 (literal) then the REAL NormValues.transform
    x = (torch.from_numpy(x_int) / ns.constant.SCALE_FACTOR).clip(1e-9, 1)
    if ns.normalize is not None:
        nv = object.__new__(ns.normalize.NormValues)
        nv.dataset_mean = mean.reshape(1, -1, 1, 1, 1)
        nv.dataset_std = std.reshape(1, -1, 1, 1, 1)
        x_norm = nv.transform(ns.TileData(x=x)).x
    else:
        x_norm = (x - mean.reshape(1, -1, 1, 1, 1)) / std.reshape(1, -1, 1, 1, 1)
    # the REAL writer over all windows as one batch
    windows = [dict(window_row_off=int(f[0]), window_col_off=int(f[1]), window_height=int(f[2]), window_width=int(f[3]), padding=int(f[4]))
               for f in fields]
    pred = prediction_for(windows, size, CASE["seed"] + 1)
    batch = ns.TileData(x=torch.zeros(len(windows), 1), window_row_off=fields[:, 0].tolist(), window_col_off=fields[:, 1].tolist(),
                        window_height=fields[:, 2].tolist(), window_width=fields[:, 3].tolist(), padding=fields[:, 4].tolist())
    mosaic = reference_write_windows(ns, (3, H, W), pred, batch)
    out = ROOT / "tests" / "golden" / "tile_reference.npz"
    np.savez_compressed(out, cfg_keys=np.array(sorted(CASE)), cfg_vals=np.array([CASE[k] for k in sorted(CASE)], dtype=np.int64),
                        x_int=x_int.astype(np.int16), fields=fields, x_norm=x_norm.numpy().astype(np.float32), mosaic=mosaic,
                        normalize_from_reference=np.array(int(ns.normalize is not None)))
    print("windows", len(windows), "x_int", x_int.shape, "mosaic", mosaic.shape, "bytes", out.stat().st_size,
          "NormValues from reference:", ns.normalize is not None)


if __name__ == "__main__":
    main()
