"""Mirror of the reference's `light_unet.models` exports for the hot path
(reference: light_unet/models/__init__.py:6-16).  The dataset re-exports of
the reference (:18-24) belong to the NIfTI data pipeline (light_unet.datasets
there); the device-side patch sampler of this package lives in
light_unet.datasets.device_patches."""
from .unet3d import Lightweight3DUNet
from .losses import FocalTverskyLoss, CombinedLoss, DiceLoss, get_loss_function
from .metrics import calculate_dsc, calculate_lesion_metrics, calculate_metrics, get_connected_components

__all__ = ["Lightweight3DUNet", "FocalTverskyLoss", "CombinedLoss", "DiceLoss", "get_loss_function",
           "calculate_dsc", "calculate_lesion_metrics", "calculate_metrics", "get_connected_components"]
