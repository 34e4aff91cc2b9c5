"""Mirror of the reference package ``restormer`` (src/restormer/__init__.py:1-20)."""
import numpy as np
import torch
import yaml

from .restormer import Restormer


def get_model(opt_path: str, device: torch.device):
    """YAML ``network_g`` -> Restormer(**kwargs), strict load of ``checkpoint['params']``, eval, to(device).

    Same behaviour as the reference's ``restormer.get_model`` (src/restormer/__init__.py:8-20)."""
    with open(opt_path, mode='r') as f:
        opt = yaml.load(f, Loader=yaml.Loader)
    kwargs = dict(opt['network_g'])
    kwargs.pop('type', None)
    model = Restormer(**kwargs)
    weights_path = opt['path']['pretrain_network_g']
    checkpoint = torch.load(weights_path, map_location='cpu')
    model.load_state_dict(checkpoint['params'])
    model.to(device)
    model.eval()
    print(f"Successfully loaded {np.sum([p.numel() for p in model.parameters()]):,} parameters from {weights_path}")
    return model


__all__ = ["Restormer", "get_model"]
