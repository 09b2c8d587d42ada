"""Mirror of the reference's `light_unet.models` exports for the hot path
(reference: light_unet/models/__init__.py:6-13).  The dataset re-exports of
the reference (:18-24) belong to the NIfTI data pipeline, which is out of
scope for this path."""
from .unet3d import Lightweight3DUNet
from .losses import FocalTverskyLoss, CombinedLoss, DiceLoss, get_loss_function
from .metrics import get_connected_components

__all__ = ["Lightweight3DUNet", "FocalTverskyLoss", "CombinedLoss", "DiceLoss", "get_loss_function",
           "get_connected_components"]
