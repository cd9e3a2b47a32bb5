"""Initialisers / helpers mirrored from the reference's model/utils.py:5-28 (same RNG call order so that the
same torch seed gives the same weights)."""
import torch.nn as nn


def initialize_weights(m):
    if isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight.data)
        if m.bias is not None:
            nn.init.constant_(m.bias.data, val=0)


def load_state_dict(model, state_dict):
    """Shape-filtered load (model/utils.py:24-28); also strips a DataParallel ``module.`` prefix (train.py:369-376)."""
    state_dict = {(k[7:] if k.startswith('module.') else k): v for k, v in state_dict.items()}
    model_dict = model.state_dict()
    model_dict.update({k: v for k, v in state_dict.items() if k in model_dict and v.shape == model_dict[k].shape})
    model.load_state_dict(model_dict)
