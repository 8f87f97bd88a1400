"""cultionet_b200 -- B200-native TowerUNet hot path behind jgrss/cultionet's module surface."""
__version__ = "0.1.0"

from . import enums  # noqa: F401
from .data import Data  # noqa: F401
from .losses import CombinedLoss, TanimotoComplementLoss, TanimotoDistLoss  # noqa: F401
from .models.cultionet import CultioNet  # noqa: F401
from .models.nunet import TowerUNet  # noqa: F401
from .models.lightning import CultionetLitModel  # noqa: F401,E402


def __getattr__(name):
    """Lazy entry points that pull in torch.distributed / the engine: ``fit``, ``fit_params``, ``CultionetParams``, ``predict_tile``,
    ``load_from_checkpoint``, ``save_checkpoint``, ``read_checkpoint`` (cultionet_b200.model) and ``TilePredictor``, ``WindowLoader``, ``MosaicWriter``, ``predict_windows``
    (cultionet_b200.tile)."""
    if name in ("fit", "fit_params", "CultionetParams", "predict_tile", "load_from_checkpoint", "save_checkpoint", "read_checkpoint"):
        from . import model

        return getattr(model, name)
    if name in ("TilePredictor", "WindowLoader", "MosaicWriter", "predict_windows"):
        from . import tile

        return getattr(tile, name)
    raise AttributeError(f"module 'cultionet_b200' has no attribute {name!r}")
