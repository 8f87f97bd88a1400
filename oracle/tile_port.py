"""TEST INFRASTRUCTURE ONLY (imported by tests/, smoke() and bench.py's CPU legs; never by the product path).

CPU restatement, in the reference's own order of operations, of the two steps either side of ``predict_step``:

* prediction windowing -- ``create_predict_dataset`` (``src/cultionet/data/create.py:114-246``) + ``BatchStore.write_batch``
  (``data/store.py:68-144``) + the load-time arithmetic of ``EdgeDataset.get`` (``data/datasets.py:443``) and
  ``NormValues.transform`` (``utils/normalize.py:78-80``);
* prediction writing -- ``LightningGTiffWriter.write_on_batch_end`` (``callbacks.py:148-227``).

Pinning status.  The arithmetic lines are the reference's literal torch / numpy expressions evaluated by the same libraries (torch fp32
division / clip, numpy fp32 product / clip / ``astype``), so they are pinned by construction.  The *window geometry* is "parity
unpinned": dask, rasterio and geowombat are absent from this image, the reference's only test of ``create_predict_dataset`` is
disabled (``tests/_test_create_dataset.py``), and in the snapshot ``BatchStore.__setitem__`` receives ``da.store`` regions of the
overlapped array (offsets advance by window_size + 2*padding).  This port states the geometry the writer needs -- window offsets in
tile coordinates -- and re-implements dask's documented ``map_overlap(depth, boundary=0, trim=False)`` semantics with ``np.pad``.
The uint16 conversion is numpy's C truncation: rasterio's ``DatasetWriter.write`` brings the array to the dataset dtype with
``np.require(arr, dtype=...)`` (rasterio is third-party and absent; stated from its published source).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

SCALE_FACTOR = 10_000.0  # data/constant.py:1


def chunk_starts(length: int, window_size: int) -> List[Tuple[int, int]]:
    """(start, size) of dask's regular chunks of ``window_size`` along one axis (``create.py:174-181``)."""
    return [(s, min(window_size, length - s)) for s in range(0, length, window_size)]


def create_predict_windows(tile: np.ndarray, window_size: int = 100, padding: int = 20) -> List[Dict]:
    """``tile``: int16 ``[T, C, H, W]``.  Returns one dict per window in chunk order with the fields ``BatchStore.write_batch`` stores
    (``store.py:118-136``): ``x`` int32 ``[1, C, T, ws+2p, ws+2p]``, ``window_row_off/col_off/height/width``, ``padding``."""
    T, C, H, W = tile.shape
    # map_overlap(depth=padding, boundary=0, trim=False): every chunk grows by `padding` on both sides of y and x, filled from the
    # neighbouring chunks or with 0 beyond the array (create.py:209-214)
    padded = np.pad(tile, ((0, 0), (0, 0), (padding, padding), (padding, padding)), mode="constant", constant_values=0)
    size = window_size + 2 * padding
    out = []
    for r0, h in chunk_starts(H, window_size):
        for c0, w in chunk_starts(W, window_size):
            item = padded[:, :, r0:r0 + h + 2 * padding, c0:c0 + w + 2 * padding]
            # store.py:73-90: pad ragged chunks after the data to the full window
            item = np.pad(item, ((0, 0), (0, 0), (0, size - item.shape[-2]), (0, size - item.shape[-1])), mode="constant",
                          constant_values=0)
            x = torch.from_numpy(item.astype("int32")).permute(1, 0, 2, 3)[None]  # 't c h w -> 1 c t h w' (store.py:92-95)
            out.append(dict(x=x, window_row_off=r0, window_col_off=c0, window_height=h, window_width=w, padding=padding))
    return out


def load_window(x_int: torch.Tensor, mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``EdgeDataset.get`` for a prediction window: ``datasets.py:443`` then ``normalize.py:78-80`` (mean / std ``[1,C,1,1,1]``)."""
    x = (x_int / SCALE_FACTOR).clip(1e-9, 1)
    if mean is not None:
        x = (x - mean.reshape(1, -1, 1, 1, 1)) / std.reshape(1, -1, 1, 1, 1)
    return x


def write_windows(mosaic: np.ndarray, prediction: Dict[str, torch.Tensor], windows: List[Dict]) -> None:
    """``LightningGTiffWriter.write_on_batch_end`` (``callbacks.py:176-227``) into ``mosaic`` uint16 ``[3, H, W]`` (the GeoTIFF)."""
    height, width = mosaic.shape[-2:]
    distance, edge, crop = prediction["distance"], prediction["edge"], prediction["crop"]
    for i, wdw in enumerate(windows):
        row_off, col_off = int(wdw["window_row_off"]), int(wdw["window_col_off"])
        h, w = int(wdw["window_height"]), int(wdw["window_width"])
        if row_off + h > height:  # :182-185
            h = height - row_off
        if col_off + w > width:
            w = width - col_off
        pad = int(wdw["padding"])
        sl = (slice(0, None), slice(pad, pad + h), slice(pad, pad + w))  # get_batch_slice :136-146
        d, e, c = distance[i][sl], edge[i][sl], crop[i][sl]
        if c.shape[0] > 1:  # :131-132
            c = c[[1]]
        stack = torch.cat((d, e, c), dim=0).detach().cpu().numpy()  # :200-214
        stack = (stack * SCALE_FACTOR).clip(0, SCALE_FACTOR)  # :220
        mosaic[:, row_off:row_off + h, col_off:col_off + w] = np.require(stack, dtype=mosaic.dtype)  # rasterio write :222-227
