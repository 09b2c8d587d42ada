"""Minimal NIfTI-1 single-file (.nii / .nii.gz) reader and writer for the Inferencer's file boundary.

The reference goes through nibabel (inferencer.py:122-134 load the image / body mask, :164-165 save the probability map
with the input's affine and header).  nibabel is used here as well when it is importable; this module is what keeps
`Inferencer.infer_case` runnable without it.  Scope: the 348-byte NIfTI-1 header, the scalar datatypes PET / label volumes
come in, either byte order, scl_slope / scl_inter, sform / qform / pixdim affines, gzip by file name.  Arrays are returned in
nibabel's index order -- array[i, j, k] with i the fastest-varying file axis -- which is the [D, H, W] the reference feeds
to the network (inferencer.py:123).
"""
from __future__ import annotations

import gzip
import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import Tuple

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16, 768: np.uint32,
           1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).name: k for k, v in _DTYPES.items()}


@dataclass
class NiftiHeader:
    """The fields the path needs, plus the raw 348 bytes so that a written file carries the input's remaining fields
    (descrip, intent, units, ...) through unchanged -- what passing `header` to nib.Nifti1Image does."""
    shape: Tuple[int, ...]
    zooms: Tuple[float, ...]
    affine: np.ndarray
    datatype: int = 16
    raw: bytes = b""
    endian: str = "<"
    extra: dict = field(default_factory=dict)

    def get_zooms(self):
        return tuple(float(z) for z in self.zooms)

    def get_data_shape(self):
        return tuple(self.shape)


def _open(path, mode):
    path = str(path)
    return gzip.open(path, mode) if path.endswith(".gz") else open(path, mode)


def _quaternion_affine(b, c, d, qfac, pixdim, offset):
    a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    z = np.array([pixdim[1], pixdim[2], pixdim[3] * (-1.0 if qfac < 0 else 1.0)])
    A = np.eye(4)
    A[:3, :3] = R * z[None, :]
    A[:3, 3] = offset
    return A


def load(path):
    """-> (data, header): data is float64-free: the stored values as float32 (scaled by scl_slope / scl_inter when set),
    shape header.shape; header.affine is the 4x4 voxel-to-world matrix (sform, else qform, else pixdim)."""
    with _open(path, "rb") as f:
        raw = f.read(348)
        if len(raw) < 348:
            raise ValueError(f"{path}: not a NIfTI-1 file (short header)")
        endian = "<"
        if struct.unpack("<i", raw[:4])[0] != 348:
            if struct.unpack(">i", raw[:4])[0] != 348:
                raise ValueError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
            endian = ">"
        if raw[344:347] not in (b"n+1", b"ni1"):
            raise ValueError(f"{path}: bad NIfTI-1 magic {raw[344:348]!r}")
        if raw[344:347] == b"ni1":
            raise ValueError(f"{path}: two-file NIfTI (.hdr/.img) is not supported")
        dim = struct.unpack(endian + "8h", raw[40:56])
        datatype, bitpix = struct.unpack(endian + "2h", raw[70:74])
        pixdim = struct.unpack(endian + "8f", raw[76:108])
        vox_offset, slope, inter = struct.unpack(endian + "3f", raw[108:120])
        qform_code, sform_code = struct.unpack(endian + "2h", raw[252:256])
        qb, qc, qd, qx, qy, qz = struct.unpack(endian + "6f", raw[256:280])
        srow = np.array(struct.unpack(endian + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
        if datatype not in _DTYPES:
            raise ValueError(f"{path}: unsupported NIfTI datatype code {datatype}")
        ndim = dim[0]
        if not 1 <= ndim <= 7:
            raise ValueError(f"{path}: bad dim[0] = {ndim}")
        shape = tuple(int(d) for d in dim[1:1 + ndim])
        f.read(max(0, int(vox_offset) - 348))
        dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
        count = int(np.prod(shape))
        buf = f.read(count * dt.itemsize)
        if len(buf) < count * dt.itemsize:
            raise ValueError(f"{path}: truncated data ({len(buf)} of {count * dt.itemsize} bytes)")
    data = np.frombuffer(buf, dtype=dt, count=count).reshape(shape, order="F")
    if slope != 0.0 and not (slope == 1.0 and inter == 0.0) and np.isfinite(slope):
        data = data.astype(np.float64) * slope + inter
    data = np.ascontiguousarray(data, dtype=np.float32)
    if sform_code > 0:
        affine = np.vstack([srow, [0, 0, 0, 1]])
    elif qform_code > 0:
        affine = _quaternion_affine(qb, qc, qd, pixdim[0], pixdim, (qx, qy, qz))
    else:
        affine = np.diag([pixdim[1], pixdim[2], pixdim[3], 1.0])
    hdr = NiftiHeader(shape=shape, zooms=tuple(pixdim[1:1 + ndim]), affine=affine, datatype=datatype, raw=raw, endian=endian,
                      extra={"scl_slope": slope, "scl_inter": inter, "qform_code": qform_code, "sform_code": sform_code})
    return data, hdr


def save(data: np.ndarray, affine, header: NiftiHeader | None, path) -> None:
    """Write `data` (any real dtype; stored as float32, little endian, no scaling) with `affine` as the sform and the
    remaining header fields of `header` (if given) carried over.  nibabel would keep the input header's on-disk dtype and
    rescale the floats into it (the reference's prob maps of an int16 PET file are stored as scaled int16); float32 is
    used instead so that the probabilities survive the round trip exactly."""
    arr = np.asarray(data)
    if arr.ndim < 1 or arr.ndim > 7:
        raise ValueError("NIfTI supports 1 to 7 dimensions")
    arr = np.asfortranarray(arr, dtype=np.float32)
    raw = bytearray(348)
    if header is not None and len(header.raw) == 348 and header.endian == "<":
        raw[:] = header.raw
    struct.pack_into("<i", raw, 0, 348)
    dim = [arr.ndim] + list(arr.shape) + [1] * (7 - arr.ndim)
    struct.pack_into("<8h", raw, 40, *dim)
    struct.pack_into("<2h", raw, 70, _CODES["float32"], 32)
    A = np.asarray(affine if affine is not None else np.eye(4), dtype=np.float64)
    zooms = np.sqrt((A[:3, :3] ** 2).sum(axis=0))
    pixdim = list(struct.unpack("<8f", bytes(raw[76:108])))
    if pixdim[0] not in (-1.0, 1.0):
        pixdim[0] = 1.0
    for i in range(min(3, arr.ndim)):
        pixdim[1 + i] = float(zooms[i])
    struct.pack_into("<8f", raw, 76, *pixdim)
    struct.pack_into("<3f", raw, 108, 352.0, 1.0, 0.0)            # vox_offset, scl_slope, scl_inter
    struct.pack_into("<2h", raw, 252, 0, 2)                        # qform_code = 0, sform_code = 2 (aligned)
    struct.pack_into("<12f", raw, 280, *A[:3].ravel().tolist())
    raw[344:348] = b"n+1\x00"
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with _open(path, "wb") as f:
        f.write(bytes(raw))
        f.write(b"\x00\x00\x00\x00")                               # no extensions; data starts at 352
        f.write(arr.tobytes(order="F"))
