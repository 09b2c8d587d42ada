"""Device-side training-patch pipeline.  The reference's NIfTI datasets / DataLoader factory
(light_unet/datasets/{case_dataset,patch_dataset,loader}.py) read files through nibabel and are not part of this
package; DevicePatchSampler is the drop-in for what PatchDataset.__getitem__ computes, with the volumes resident in HBM."""
from .device_patches import DevicePatchSampler, MixedDevicePatchSampler

__all__ = ["DevicePatchSampler", "MixedDevicePatchSampler"]
