"""Mirror of the reference package ``dncnn`` (src/dncnn/__init__.py:1-15)."""
import numpy as np
import torch

from .models.network_dncnn import DnCNN


def get_model(weights_path: str, n_channels: int, nb: int, device: torch.device):
    """Same behaviour as the reference's ``dncnn.get_model`` (src/dncnn/__init__.py:7-15)."""
    model = DnCNN(in_nc=n_channels, out_nc=n_channels, nc=64, nb=nb, act_mode='R')
    model.load_state_dict(torch.load(weights_path, map_location='cpu'), strict=True)
    model.eval()
    model.to(device)
    print(f"Successfully loaded {np.sum([p.numel() for p in model.parameters()]):,} parameters from {weights_path}")
    return model


__all__ = ["DnCNN", "get_model"]
