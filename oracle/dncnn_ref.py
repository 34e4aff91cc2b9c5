"""CPU restatement of the reference DnCNN forward (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates src/dncnn/models/network_dncnn.py:63-71 with the layer factory semantics of
src/dncnn/models/basicblock.py:61-98 (mode letters C/B/R only).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def dncnn_forward(sd, x, taps=None):
    """``x - model(x)`` (network_dncnn.py:69-71); BN is eval-mode affine on running stats, eps=1e-4
    (basicblock.py:69).  Layer structure is recovered from the state-dict indices
    (B.sequential flattening, basicblock.py:15-35)."""
    idxs = sorted({int(k.split(".")[1]) for k in sd if k.startswith("model.")})
    n = x
    last_conv = max(i for i in idxs if sd[f"model.{i}.weight"].dim() == 4)
    for i in idxs:
        w = sd[f"model.{i}.weight"]
        if w.dim() == 4:
            n = F.conv2d(n, w, sd[f"model.{i}.bias"], padding=1)
            nxt_is_bn = (i + 1) in idxs and sd[f"model.{i + 1}.weight"].dim() == 1
            if i != last_conv and not nxt_is_bn:
                n = F.relu(n)
        else:
            n = F.batch_norm(n, sd[f"model.{i}.running_mean"], sd[f"model.{i}.running_var"], w,
                             sd[f"model.{i}.bias"], training=False, eps=1e-4)
            n = F.relu(n)
        if taps is not None:
            taps[f"model.{i}"] = n
    return x - n
