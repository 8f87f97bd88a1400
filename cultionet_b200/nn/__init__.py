from .modules.attention import NeighborhoodAttention2D
from .modules.convolution import (
    ConvBlock2d,
    ConvTranspose2d,
    PoolResidualConv,
    ResConvBlock2d,
    ResidualAConv,
    ResidualConv,
)
from .modules.unet_parts import (
    NATTEN_PARAMS,
    SigmoidCrisp,
    StreamConv2d,
    TowerUNetBlock,
    TowerUNetDecoder,
    TowerUNetEncoder,
    TowerUNetFinal,
    TowerUNetFinalCombine,
    TowerUNetFusion,
    UNetUpBlock,
)

__all__ = [
    "ConvBlock2d", "ConvTranspose2d", "NeighborhoodAttention2D", "PoolResidualConv", "ResConvBlock2d", "ResidualAConv",
    "ResidualConv", "NATTEN_PARAMS", "SigmoidCrisp", "StreamConv2d", "TowerUNetFinal", "TowerUNetFinalCombine", "UNetUpBlock",
    "TowerUNetBlock", "TowerUNetEncoder", "TowerUNetDecoder", "TowerUNetFusion",
]
