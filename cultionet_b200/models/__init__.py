from .cultionet import CultioNet
from .nunet import PreTimeReduction, TowerUNet

__all__ = ["CultioNet", "PreTimeReduction", "TowerUNet"]
