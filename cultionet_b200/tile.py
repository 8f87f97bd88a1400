"""Sliding-window prediction over a satellite tile that is RESIDENT in HBM: the steps either side of ``predict_step``
(SURVEY §8(f) N3 and N2), so that BASELINE config 5 runs tile -> windows -> TowerUNet -> 3-band uint16 mosaic without leaving the device.

What the reference does on the CPU, and where:

* ``create_predict_dataset`` (``src/cultionet/data/create.py:114-246``) rechunks the ``(time, band, y, x)`` stack into ``window_size``
  chunks, adds a ``padding``-wide halo with ``map_overlap(depth=padding, boundary=0, trim=False)`` (``:201-214``) and ``BatchStore``
  writes every padded chunk as one ``.pt`` file holding ``Data(x=int32[1,C,T,ws+2p,ws+2p], window_row_off, window_col_off,
  window_height, window_width, padding, ...)``; ragged end chunks are zero-padded at the bottom / right (``data/store.py:68-144``).
* ``EdgeDataset.get`` (``data/datasets.py:443``) turns the integers into reflectances ``(x / 10000).clip(1e-9, 1)`` and
  ``NormValues.transform`` (``utils/normalize.py:78-80``) z-scores them per band.
* ``LightningGTiffWriter.write_on_batch_end`` (``callbacks.py:148-227``) slices the halo off every prediction, stacks (distance, edge,
  crop), scales by 10000, clips, and writes the window into a 3-band uint16 GeoTIFF under a file lock.

Here the tile stays an int16 tensor on the GPU, ``cnb_window_load`` produces a batch of windows (halo, zero boundary, ragged-end
padding, scaling, clipping and z-score in one pass) and ``cnb_predict_pack`` writes the sliced, scaled uint16 result into the device
mosaic; ``TilePredictor`` replays load -> ``predict_step`` -> pack as ONE CUDA graph per window batch.  Ranks take windows
``i mod world_size`` (the reference's ``DistributedSampler`` order, ``model.py:437-467``); there is no data-path collective, the
mosaics of the ranks are disjoint and are summed onto the writer rank once per tile (``MosaicWriter.gather``).

Note on window offsets: in the reference snapshot ``BatchStore.__setitem__`` receives the slices of the *overlapped* dask array
(``da.store`` regions), so the stored ``window_row_off`` advances by ``window_size + 2*padding`` per chunk; no enabled reference test
pins that (``tests/_test_create_dataset.py`` is disabled).  This module uses tile coordinates -- ``row_off = i * window_size`` --
which is what ``LightningGTiffWriter`` needs to place a window (``callbacks.py:176-192``).
"""
from __future__ import annotations

import warnings
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, check_device, ptr, stream_ptr
from .data import Data

SCALE_FACTOR = 10_000.0  # src/cultionet/data/constant.py:1


def predict_windows(height: int, width: int, window_size: int = 100, padding: int = 20) -> np.ndarray:
    """``int32 [N, 4]`` rows ``(row_off, col_off, height, width)`` of the prediction windows of a ``height x width`` tile, in the
    reference's chunk order (row-major over the ``window_size`` chunks of ``y`` then ``x``, ``data/create.py:174-181``); the last
    row / column of windows is ragged when the tile size is not a multiple of ``window_size``."""
    if window_size <= 0 or padding < 0:
        raise ValueError("window_size must be positive and padding non-negative")
    rows = np.arange(0, height, window_size, dtype=np.int64)
    cols = np.arange(0, width, window_size, dtype=np.int64)
    rr, cc = np.meshgrid(rows, cols, indexing="ij")
    hh = np.minimum(window_size, height - rr)
    ww = np.minimum(window_size, width - cc)
    return np.stack([rr, cc, hh, ww], axis=-1).reshape(-1, 4).astype(np.int32)


class WindowLoader:
    """Cuts window batches out of a device-resident ``int16 [T, C, H, W]`` tile (``cnb_window_load``).

    ``norm_values``: an object with ``dataset_mean`` / ``dataset_std`` (the reference's ``NormValues``, any shape with C elements) or a
    ``(mean, std)`` pair, or ``None`` for no z-score (``EdgeDataset(norm_values=None)``)."""

    def __init__(self, tile: torch.Tensor, window_size: int = 100, padding: int = 20, norm_values=None):
        check_device(tile)
        if tile.dtype != torch.int16 or tile.dim() != 4:
            raise TypeError("tile must be an int16 tensor [time, band, y, x] (data/create.py:70-79)")
        if (window_size + 2 * padding) % 4 != 0:
            raise ValueError("window_size + 2 * padding must be a multiple of 4")
        self.tile = tile.contiguous()
        self.window_size, self.padding = int(window_size), int(padding)
        self.T, self.C, self.H, self.W = self.tile.shape
        self.mean = self.std = None
        if norm_values is not None:
            mean, std = (norm_values if isinstance(norm_values, (tuple, list)) else (norm_values.dataset_mean, norm_values.dataset_std))
            self.mean = torch.as_tensor(mean, dtype=torch.float32).reshape(-1).to(tile.device).contiguous()
            self.std = torch.as_tensor(std, dtype=torch.float32).reshape(-1).to(tile.device).contiguous()
            if self.mean.numel() != self.C or self.std.numel() != self.C:
                raise ValueError("norm_values must hold one mean / std per band")

    @property
    def window_shape(self) -> Tuple[int, int, int, int]:
        s = self.window_size + 2 * self.padding
        return (self.C, self.T, s, s)

    def load_into(self, win: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """``win``: device int32 ``[B, >=2]`` contiguous rows starting with (row_off, col_off); ``out``: device fp32
        ``[B, C, T, s, s]``."""
        B = out.shape[0]
        assert win.dtype == torch.int32 and win.is_contiguous() and win.shape[0] >= B and win.shape[1] >= 2
        call("cnb_window_load", ptr(self.tile), self.T, self.C, self.H, self.W, ptr(win), win.shape[1], B, self.window_size, self.padding,
             SCALE_FACTOR, 1e-9, 1.0, ptr(self.mean), ptr(self.std), ptr(out), stream_ptr(out))
        return out

    def load(self, win: torch.Tensor) -> Data:
        """A ``Data`` batch with the fields the reference's ``.pt`` windows carry (``data/store.py:118-136``)."""
        win = win.to(device=self.tile.device, dtype=torch.int32).contiguous()
        x = torch.empty((win.shape[0], *self.window_shape), dtype=torch.float32, device=self.tile.device)
        self.load_into(win, x)
        return Data(x=x, padding=[self.padding] * win.shape[0], window_row_off=win[:, 0], window_col_off=win[:, 1],
                    window_height=win[:, 2], window_width=win[:, 3])


class MosaicWriter:
    """Device-side stand-in for ``LightningGTiffWriter`` (``callbacks.py:46-227``): keeps the 3-band uint16 mosaic
    (distance, edge, crop; ``:88-100``) in HBM and fills it window by window with ``cnb_predict_pack``."""

    def __init__(self, height: int, width: int, device, window_size: int = 100):
        self.height, self.width, self.window_size = int(height), int(width), int(window_size)
        # an even row pitch so that two uint16 form one int32 for the cross-rank sum of ``gather``
        self._pitch = self.width + (self.width & 1)
        self._store = torch.zeros((3, self.height, self._pitch), dtype=torch.uint16, device=device)

    @property
    def mosaic(self) -> torch.Tensor:
        return self._store[:, :, : self.width]

    def get_batch_slice(self, padding: int, height: int, width: int) -> tuple:
        """``LightningGTiffWriter.get_batch_slice`` (``callbacks.py:136-146``)."""
        return (slice(0, None), slice(padding, padding + height), slice(padding, padding + width))

    def write_windows(self, prediction: Dict[str, torch.Tensor], win: torch.Tensor, padding: int) -> None:
        """``prediction``: the dict of ``predict_step`` (fp32 ``[B, 1 or K, Hs, Ws]`` per key); ``win``: device int32 ``[B, 4]``."""
        dist_p, edge_p, crop_p = prediction["distance"], prediction["edge"], prediction["crop"]
        check_device(dist_p, edge_p, crop_p, win)
        if crop_p.shape[1] > 1:  # callbacks.py:131-132
            crop_p = crop_p[:, 1:2]
        B, _, Hs, Ws = dist_p.shape
        strides = {t.stride(0) for t in (dist_p, edge_p, crop_p)}
        ok = all(t.dtype == torch.float32 and t.stride(3) == 1 and t.stride(2) == Ws for t in (dist_p, edge_p, crop_p))
        if not ok or len(strides) != 1:
            dist_p, edge_p, crop_p = (t.float().contiguous() for t in (dist_p, edge_p, crop_p))
        call("cnb_predict_pack", ptr(dist_p), ptr(edge_p), ptr(crop_p), dist_p.stride(0), Hs, Ws, int(padding), ptr(win), B,
             self.window_size, SCALE_FACTOR, ptr(self._store), self.height, self.width, self._pitch, stream_ptr(dist_p))

    def write_on_batch_end(self, prediction: Dict[str, torch.Tensor], batch: Data) -> None:
        """Reference-shaped entry (``callbacks.py:148-227``): window placement comes from the batch's ``window_*`` fields."""
        dev = prediction["distance"].device
        cols = [torch.as_tensor(getattr(batch, k), dtype=torch.int32, device=dev).reshape(-1)
                for k in ("window_row_off", "window_col_off", "window_height", "window_width")]
        pad = batch.padding[0] if isinstance(batch.padding, (list, tuple)) else int(batch.padding.reshape(-1)[0])
        self.write_windows(prediction, torch.stack(cols, dim=1).contiguous(), int(pad))

    def gather(self, dst: int = 0) -> Optional[torch.Tensor]:
        """Sum the ranks' disjoint mosaics onto rank ``dst`` (the single writer; the reference serialises ranks through a file lock,
        ``callbacks.py:222``).  Values are <= 10000 and every pixel is written by exactly one rank, so the sum of uint16 pairs viewed
        as int32 never carries.  Returns the mosaic on ``dst``, ``None`` elsewhere; without a process group it is the local mosaic."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self.mosaic
        packed = self._store.view(torch.int32)
        dist.reduce(packed, dst=dst, op=dist.ReduceOp.SUM)
        return self.mosaic if dist.get_rank() == dst else None


class TilePredictor:
    """tile -> windows -> ``predict_step`` -> mosaic for the windows of this rank.

    One window batch = ``cnb_window_load`` + the eval-mode TowerUNet forward + ``cnb_predict_pack``; with ``cuda_graph=True`` the three
    are captured once and replayed per batch (only the small window table changes between replays).  A ragged last batch is filled
    with height-0 windows, which load as valid inputs and write nothing."""

    def __init__(self, lit_model, tile: torch.Tensor, norm_values=None, window_size: int = 100, padding: int = 20,
                 batch_windows: int = 32, cuda_graph: bool = True, rank: Optional[int] = None, world_size: Optional[int] = None):
        self.model = lit_model
        self.model.eval()
        self.loader = WindowLoader(tile, window_size, padding, norm_values)
        self.writer = MosaicWriter(self.loader.H, self.loader.W, tile.device, window_size)
        self.batch_windows = int(batch_windows)
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank, self.world_size = rank, world_size
        all_windows = predict_windows(self.loader.H, self.loader.W, window_size, padding)
        mine = all_windows[rank::world_size]
        fill = (-len(mine)) % self.batch_windows
        if fill:
            mine = np.concatenate([mine, np.zeros((fill, 4), dtype=np.int32)], axis=0)
        self.num_windows = len(all_windows[rank::world_size])
        self.windows = torch.from_numpy(np.ascontiguousarray(mine)).to(tile.device)
        self.cuda_graph = bool(cuda_graph) and tile.is_cuda and not _lib.is_emulator()
        self._graph = None
        self._win = torch.zeros((self.batch_windows, 4), dtype=torch.int32, device=tile.device)
        self._x = torch.empty((self.batch_windows, *self.loader.window_shape), dtype=torch.float32, device=tile.device)
        self._warm = 0
        self.launches_per_batch: Optional[int] = None

    @property
    def num_batches(self) -> int:
        return self.windows.shape[0] // self.batch_windows

    @torch.no_grad()
    def _batch(self) -> None:
        self.loader.load_into(self._win, self._x)
        out = self.model.predict_step(Data(x=self._x), 0)
        self.writer.write_windows(out, self._win, self.loader.padding)

    @torch.no_grad()
    def step(self, index: int) -> None:
        """Enqueue window batch ``index`` of this rank (no synchronisation)."""
        b = self.batch_windows
        self._win.copy_(self.windows[index * b:(index + 1) * b], non_blocking=True)
        if not self.cuda_graph or _lib.TIMER is not None:
            return self._batch()
        if self._graph is None:
            if self._warm < 2:
                self._warm += 1
                return self._batch()
            try:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                l0 = _lib.launch_count()
                with torch.cuda.graph(graph):
                    self._batch()
                self.launches_per_batch = _lib.launch_count() - l0
                self._graph = graph
            except Exception as e:  # noqa: BLE001
                warnings.warn(f"cultionet_b200: CUDA graph capture of the tile batch failed ({e!r}); running eagerly")
                self.cuda_graph = False
                self._graph = None
                torch.cuda.synchronize()
                return self._batch()
        self._graph.replay()

    def run_streaming(self, host_tile: torch.Tensor, host_mosaic: Optional[torch.Tensor] = None,
                      batches: Optional[Sequence[int]] = None) -> Optional[torch.Tensor]:
        """End-to-end form for a tile in (pinned) HOST memory: the int16 rows a window batch needs are copied to the device tile on a
        side stream ahead of the batch that reads them -- every tile row crosses PCIe/NVLink-C2C exactly once, as 2-byte integers,
        while the previous batches compute -- and finished mosaic rows are copied back to ``host_mosaic`` (pinned uint16
        ``[3, H, pitch]``, allocated when ``None``) as soon as no later window of this rank touches them.  ``batches`` (default: all) must
        be increasing.  Returns ``host_mosaic``; the caller synchronises (``torch.cuda.synchronize()``) before reading it."""
        ld, wr = self.loader, self.writer
        if tuple(host_tile.shape) != (ld.T, ld.C, ld.H, ld.W) or host_tile.dtype != torch.int16 or not host_tile.is_contiguous():
            raise TypeError("host_tile must be a contiguous int16 tensor of the device tile's shape")
        cuda = ld.tile.is_cuda
        if host_mosaic is None:
            host_mosaic = torch.empty(tuple(wr._store.shape), dtype=torch.uint16, pin_memory=cuda)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(ld.tile.device) if cuda else None
            self._back_stream = torch.cuda.Stream(ld.tile.device) if cuda else None
            b = self.batch_windows
            w = self.windows.cpu().view(-1, b, 4)
            live = w[:, :, 2] > 0
            last_row = torch.where(live, w[:, :, 0] + w[:, :, 2], torch.zeros_like(w[:, :, 0])).amax(dim=1)
            first_row = torch.where(live, w[:, :, 0], torch.full_like(w[:, :, 0], ld.H)).amin(dim=1)
            self._need_rows = torch.clamp(last_row + ld.padding, max=ld.H).tolist()  # tile rows batch i reads: [.., need)
            # mosaic rows complete after batch i: everything above the first row of any later batch of this rank
            later = torch.cat([first_row[1:], torch.tensor([ld.H])]).flip(0).cummin(0).values.flip(0)
            self._done_rows = later.tolist()
        order = list(range(self.num_batches) if batches is None else batches)
        copied = done = 0
        planes_d, planes_h = ld.tile.view(ld.T * ld.C, ld.H, ld.W), host_tile.view(ld.T * ld.C, ld.H, ld.W)

        def h2d(upto: int):  # rows [copied, upto) of every (time, band) plane: one contiguous transfer per plane
            nonlocal copied
            if upto > copied:
                for p in range(planes_d.shape[0]):
                    planes_d[p, copied:upto].copy_(planes_h[p, copied:upto], non_blocking=True)
                copied = upto

        def d2h(upto: int):
            nonlocal done
            if upto > done:
                for band in range(3):
                    host_mosaic[band, done:upto].copy_(wr._store[band, done:upto], non_blocking=True)
                done = upto

        if not cuda:
            for i in order:
                h2d(self._need_rows[i])
                self.step(i)
                d2h(self._done_rows[i])
            return host_mosaic
        main = torch.cuda.current_stream(ld.tile.device)
        self._copy_stream.wait_stream(main)
        self._back_stream.wait_stream(main)
        for n, i in enumerate(order):
            with torch.cuda.stream(self._copy_stream):
                h2d(self._need_rows[i])
                ready = torch.cuda.Event()
                ready.record()
                if n + 1 < len(order):  # the next batch's rows travel while this batch computes
                    h2d(self._need_rows[order[n + 1]])
            main.wait_event(ready)
            self.step(i)
            upto = self._done_rows[i]
            if upto > done:
                finished = torch.cuda.Event()
                finished.record(main)
                with torch.cuda.stream(self._back_stream):
                    self._back_stream.wait_event(finished)
                    d2h(upto)
        main.wait_stream(self._back_stream)
        return host_mosaic

    def run(self, batches: Optional[Sequence[int]] = None) -> torch.Tensor:
        """Predict all (or the given) window batches of this rank; returns this rank's device mosaic ``uint16 [3, H, W]``."""
        for i in (range(self.num_batches) if batches is None else batches):
            self.step(i)
        return self.writer.mosaic
