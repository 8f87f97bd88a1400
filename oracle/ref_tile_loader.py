"""TEST INFRASTRUCTURE ONLY -- import the *real* reference code either side of ``predict_step`` from /root/reference:

* ``cultionet.callbacks.LightningGTiffWriter.write_on_batch_end`` (``callbacks.py:148-227``: window clipping, halo slice, stack,
  ``* SCALE_FACTOR`` clip, windowed write),
* ``cultionet.data.store.BatchStore.__setitem__`` / ``write_batch`` (``data/store.py:51-144``: ragged padding, ``t c h w -> 1 c t h w``,
  int32, the window fields stored with every batch),
* ``cultionet.utils.normalize.NormValues.transform`` (``utils/normalize.py:63-84``) and the load-time scaling line of
  ``EdgeDataset.get`` (``data/datasets.py:443``, a one-line expression restated here next to its citation because the method around
  it needs shapely / scikit-image / the augmenters).

Their third-party imports (lightning, geowombat, rasterio, dask, xarray, retry, rich ...) are absent from this image; they are replaced
by shells that carry no arithmetic: base classes, a pass-through ``retry`` decorator, a ``Window`` record with the four fields rasterio's
has, and an in-memory stand-in for the opened GeoTIFF whose ``write(array, indexes, window)`` stores ``array`` into a ``uint16`` mosaic
the way rasterio does (``np.require(arr, dtype=dataset dtype)`` -- stated from rasterio's published source; rasterio itself is the one
link that stays unpinned).  Works only where ``/root/reference`` exists; ``oracle/make_tile_golden.py`` used it to write
``tests/golden/tile_reference.npz``.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path
from unittest import mock

import numpy as np

from . import ref_loader


class Window:
    """The four fields of ``rasterio.windows.Window`` the reference reads."""

    def __init__(self, col_off=0, row_off=0, width=0, height=0):
        self.col_off, self.row_off, self.width, self.height = col_off, row_off, width, height


class MemoryGTiff:
    """Stand-in for ``rio.open(out_path, mode='r+')``: a ``[count, H, W]`` uint16 array."""

    def __init__(self, count: int, height: int, width: int, dtype="uint16"):
        self.array = np.zeros((count, height, width), dtype=dtype)
        self.profile = {"count": count, "height": height, "width": width, "dtype": dtype}

    def write(self, arr, indexes=None, window=None):
        idx = [i - 1 for i in indexes]
        r0, c0 = int(window.row_off), int(window.col_off)
        self.array[idx, r0:r0 + int(window.height), c0:c0 + int(window.width)] = np.require(arr, dtype=self.array.dtype)

    def close(self):
        pass


def _shell(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    return m


def _retry(*_a, **_k):
    return lambda fn: fn


def load_tile_reference():
    """Namespace with the reference's ``LightningGTiffWriter``, ``BatchStore``, ``NormValues`` and ``SCALE_FACTOR``."""
    ns = ref_loader.load_reference()  # registers the `cultionet` package shell + `cultionet.data.Data`
    Base = type("BasePredictionWriter", (), {"__init__": lambda self, write_interval="batch": None})
    anything = mock.MagicMock()
    _shell("lightning")
    _shell("lightning.pytorch")
    _shell("lightning.pytorch.callbacks", BasePredictionWriter=Base, LearningRateMonitor=anything, ModelCheckpoint=anything,
           ModelPruning=anything, RichProgressBar=anything, StochasticWeightAveraging=anything)
    _shell("lightning.pytorch.callbacks.progress")
    _shell("lightning.pytorch.callbacks.progress.rich_progress", RichProgressBarTheme=anything)
    _shell("geowombat")
    _shell("rasterio")
    _shell("rasterio.windows", Window=Window)
    _shell("dask")
    _shell("dask.array", Array=object)
    _shell("dask.delayed", Delayed=object)
    _shell("dask.utils", SerializableLock=lambda *a, **k: None)
    _shell("xarray", DataArray=object)
    _shell("retry", retry=_retry)
    try:
        import filelock  # noqa: F401
    except ImportError:  # pragma: no cover
        class _NoLock:
            def __init__(self, *_a, **_k):
                pass

            def __enter__(self):
                return self

            def __exit__(self, *exc):
                return False

        _shell("filelock", FileLock=_NoLock)
    if "rich.progress" not in sys.modules:
        try:
            importlib.import_module("rich.progress")
        except ImportError:  # pragma: no cover
            _shell("rich")
            _shell("rich.progress", BarColumn=anything, Progress=anything, TaskProgressColumn=anything, TextColumn=anything,
                   TimeElapsedColumn=anything)
            _shell("rich.style", Style=anything)

    # `cultionet.data` is the loader's shell module: give it the sub-modules the files below import
    data_pkg = sys.modules["cultionet.data"]
    data_pkg.__path__ = [str(ref_loader.REFERENCE_SRC / "data")]
    dd = types.ModuleType("cultionet.data.data")
    captured = []

    class Data(ref_loader._Data):
        """Field container + ``to_file`` / ``from_file`` that keep the batch in memory instead of a joblib file."""

        def to_file(self, path, compress=None):
            captured.append((Path(path).name, self))

        @classmethod
        def from_file(cls, path):
            return [b for n, b in captured if n == Path(path).name][-1]

        def copy(self):
            return Data(**{k: (v.clone() if hasattr(v, "clone") else v) for k, v in self.__dict__.items()})

    dd.Data = Data
    sys.modules["cultionet.data.data"] = dd
    data_pkg.Data = Data
    du = types.ModuleType("cultionet.data.utils")
    du.collate_fn = lambda x: x
    sys.modules["cultionet.data.utils"] = du

    ns.callbacks = importlib.import_module("cultionet.callbacks")
    ns.store = importlib.import_module("cultionet.data.store")
    ns.constant = importlib.import_module("cultionet.data.constant")
    try:
        ns.normalize = importlib.import_module("cultionet.utils.normalize")
    except Exception:  # noqa: BLE001 - stats.py pulls optional packages on some images
        ns.normalize = None
    ns.TileData = Data
    ns.captured_batches = captured
    ns.Window = Window
    ns.MemoryGTiff = MemoryGTiff
    return ns


def reference_write_windows(ns, mosaic_shape, prediction, batch) -> np.ndarray:
    """Run the REAL ``LightningGTiffWriter.write_on_batch_end`` against an in-memory GeoTIFF; returns the uint16 mosaic."""
    writer = object.__new__(ns.callbacks.LightningGTiffWriter)  # __init__ only opens files (geowombat / rasterio)
    count, height, width = mosaic_shape
    writer.profile = {"height": height, "width": width, "count": count}
    writer.dst = MemoryGTiff(count, height, width)
    writer.crs = None
    import os
    import tempfile

    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:  # the method takes a FileLock("./dst.lock")
        os.chdir(tmp)
        try:
            writer.write_on_batch_end(None, None, prediction, None, batch, 0, 0)
        finally:
            os.chdir(cwd)
    return writer.dst.array


def reference_store_window(ns, item: np.ndarray, y: slice, x: slice, window_size: int, padding: int):
    """Run the REAL ``BatchStore.__setitem__`` for one chunk ``item`` (``[T, C, h, w]``) at region ``(y, x)``; returns the stored batch."""
    geo = mock.MagicMock()
    geo.gw.geodataframe.to_crs.return_value.total_bounds.tolist.return_value = [0.0, 0.0, 1.0, 1.0]
    store = ns.store.BatchStore(data=geo, write_path=Path("/nonexistent"), res=10.0, resampling="nearest", region="r",
                                start_date="20200101", end_date="20210101", window_size=window_size, padding=padding,
                                compress_method="zlib")
    n0 = len(ns.captured_batches)
    store[(slice(None), slice(None), y, x)] = item
    assert len(ns.captured_batches) == n0 + 1
    return ns.captured_batches[-1][1]
