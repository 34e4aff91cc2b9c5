"""B200-native (sm_100a) Restormer / DnCNN inference forward.

Drop-in for the model classes of leducthanhig/image-restoration-models
(``restormer.Restormer``, ``dncnn.models.network_dncnn.DnCNN``); see INTEGRATION.md.
"""
from .restormer import Restormer
from .dncnn import DnCNN

__all__ = ["Restormer", "DnCNN"]
__version__ = "0.1.0"
