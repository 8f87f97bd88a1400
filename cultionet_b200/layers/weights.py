"""Weight initialisation as in ``src/cultionet/layers/weights.py:24-39`` (applied to the whole TowerUNet,
``models/nunet.py:211``): Kaiming-normal (fan_in) conv/linear weights, N(0,1) biases, N(1,0.02) BatchNorm scales."""
import torch.nn as nn

_CONV_LIKE = (nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.Linear)
_BN_LIKE = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)


def init_conv_weights(module: nn.Module) -> None:
    if isinstance(module, _CONV_LIKE):
        nn.init.kaiming_normal_(module.weight.data, a=0, mode="fan_in")
        if module.bias is not None:
            nn.init.normal_(module.bias.data)
    elif isinstance(module, _BN_LIKE):
        nn.init.normal_(module.weight.data, 1.0, 0.02)
        nn.init.constant_(module.bias.data, 0.0)
