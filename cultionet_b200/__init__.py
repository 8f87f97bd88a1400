"""cultionet_b200 -- B200-native TowerUNet hot path behind jgrss/cultionet's module surface."""
__version__ = "0.1.0"

from . import enums  # noqa: F401
from .data import Data  # noqa: F401
from .losses import TanimotoComplementLoss  # noqa: F401
from .models.cultionet import CultioNet  # noqa: F401
from .models.nunet import TowerUNet  # noqa: F401
