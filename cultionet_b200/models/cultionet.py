"""CultioNet wrapper: ``Data`` batch in, prediction dict out (reference ``src/cultionet/models/cultionet.py:12-110``)."""
from __future__ import annotations

import typing as T

import torch
import torch.nn as nn

from ..data import Data
from ..enums import AttentionTypes, InferenceNames, ModelTypes, ResBlockTypes
from .nunet import TowerUNet


class CultioNet(nn.Module):
    def __init__(
        self,
        in_channels: int,
        in_time: int,
        hidden_channels: int = 32,
        model_type: str = ModelTypes.TOWERUNET,
        activation_type: str = "SiLU",
        dropout: float = 0.1,
        dilations: T.Union[int, T.Sequence[int]] = None,
        res_block_type: str = ResBlockTypes.RESA,
        attention_weights: str = AttentionTypes.NATTEN,
        pool_by_max: bool = False,
        batchnorm_first: bool = False,
        use_latlon: bool = False,
        compute_dtype: torch.dtype = torch.float32,
    ):
        super().__init__()
        self.in_channels, self.in_time, self.hidden_channels = in_channels, in_time, hidden_channels
        self.use_latlon = bool(use_latlon)
        assert model_type in (ModelTypes.TOWERUNET,), "The model type is not supported."
        self.mask_model = TowerUNet(
            in_channels=in_channels, in_time=in_time, hidden_channels=hidden_channels, num_classes=1,
            attention_weights=attention_weights, res_block_type=res_block_type, dropout=dropout, dilations=dilations,
            activation_type=activation_type, edge_activation=True, mask_activation=True, pool_by_max=pool_by_max,
            batchnorm_first=batchnorm_first, use_latlon=use_latlon, compute_dtype=compute_dtype,
        )

    def forward(self, batch: Data) -> T.Dict[str, torch.Tensor]:
        # latlon_coords [B,2] = (lon, lat) as the reference builds it (cultionet.py:88-94); TowerUNet reads it only when use_latlon=True
        # (CultionetLitModel never enables that, lightning.py:878-890), so a batch without lon / lat is fine otherwise
        latlon_coords = None
        if self.use_latlon:
            latlon_coords = torch.cat((batch.lon.reshape(-1, 1), batch.lat.reshape(-1, 1)), dim=1)
        out = self.mask_model(batch.x, latlon_coords=latlon_coords)
        out.update({InferenceNames.CROP_TYPE: None, InferenceNames.CLASSES_L2: None, InferenceNames.CLASSES_L3: None})
        return out
