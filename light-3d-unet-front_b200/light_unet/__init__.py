"""B200-native drop-in for the volumetric hot path of Light-3D-Unet.

Same import paths as the reference package (`light_unet.models`,
`light_unet.utils`, `light_unet.core.inferencer`); the arithmetic runs in
hand-written sm_100a CUDA kernels behind the C-ABI of include/l3d.h.
"""
__version__ = "0.1.0"
