"""NIfTI-1 reader / writer used by Inferencer.infer_case when nibabel is absent (reference file
boundary: inferencer.py:122-134,164-165).  No reference fixture exists for this format (nibabel is not installed here), so the
files are built byte by byte from the NIfTI-1 specification in the test."""
import gzip
import struct

import numpy as np
import pytest

from light_unet import io_nifti


def _raw_file(arr, endian="<", code=4, slope=0.0, inter=0.0, sform=None, pixdim=(1.0, 4.0, 4.0, 4.0, 0, 0, 0, 0), vox_offset=352):
    h = bytearray(348)
    struct.pack_into(endian + "i", h, 0, 348)
    struct.pack_into(endian + "8h", h, 40, arr.ndim, *arr.shape, *([1] * (7 - arr.ndim)))
    struct.pack_into(endian + "2h", h, 70, code, arr.dtype.itemsize * 8)
    struct.pack_into(endian + "8f", h, 76, *pixdim)
    struct.pack_into(endian + "3f", h, 108, float(vox_offset), slope, inter)
    if sform is not None:
        struct.pack_into(endian + "2h", h, 252, 0, 1)
        struct.pack_into(endian + "12f", h, 280, *np.asarray(sform, dtype=np.float64)[:3].ravel())
    h[344:348] = b"n+1\x00"
    body = np.asfortranarray(arr).astype(arr.dtype.newbyteorder(endian)).tobytes(order="F")
    return bytes(h) + b"\x00" * (vox_offset - 348) + body


@pytest.mark.parametrize("endian", ["<", ">"])
def test_reads_int16_with_scaling_and_sform(tmp_path, endian):
    rng = np.random.default_rng(0)
    arr = rng.integers(-500, 3000, size=(5, 7, 9)).astype(np.int16)
    A = np.array([[-4.0, 0, 0, 100], [0, 4.0, 0, -50], [0, 0, 4.0, 7], [0, 0, 0, 1]])
    p = tmp_path / "x.nii.gz"
    with gzip.open(p, "wb") as f:
        f.write(_raw_file(arr, endian=endian, slope=0.5, inter=-3.0, sform=A, vox_offset=400))
    data, hdr = io_nifti.load(p)
    assert data.dtype == np.float32 and data.shape == (5, 7, 9) and data.flags["C_CONTIGUOUS"]
    assert np.array_equal(data, (arr.astype(np.float64) * 0.5 - 3.0).astype(np.float32))      # data[i, j, k], i fastest on disk
    assert np.allclose(hdr.affine, A) and hdr.get_zooms()[:3] == (4.0, 4.0, 4.0)


def test_round_trip_keeps_probabilities_and_geometry(tmp_path):
    rng = np.random.default_rng(1)
    prob = rng.random((6, 8, 10), dtype=np.float32)
    src = tmp_path / "img.nii"
    src.write_bytes(_raw_file(np.zeros((6, 8, 10), np.uint8), code=2, pixdim=(1.0, 2.0, 3.0, 5.0, 0, 0, 0, 0)))
    _, hdr = io_nifti.load(src)
    assert np.allclose(hdr.affine, np.diag([2.0, 3.0, 5.0, 1.0]))           # no sform / qform: pixdim
    out = tmp_path / "sub" / "p_prob.nii.gz"
    io_nifti.save(prob, hdr.affine, hdr, out)
    back, h2 = io_nifti.load(out)
    assert np.array_equal(back, prob) and np.allclose(h2.affine, hdr.affine) and h2.get_zooms() == (2.0, 3.0, 5.0)
    assert h2.datatype == 16


def test_rejects_garbage(tmp_path):
    p = tmp_path / "bad.nii"
    p.write_bytes(b"\x00" * 400)
    with pytest.raises(ValueError):
        io_nifti.load(p)
    p.write_bytes(_raw_file(np.zeros((2, 2, 2), np.int16))[:-3])
    with pytest.raises(ValueError):
        io_nifti.load(p)


def test_quaternion_affine(tmp_path):
    # identity rotation with qfac = -1: z flips
    h = bytearray(_raw_file(np.zeros((2, 3, 4), np.float32), code=16, pixdim=(-1.0, 1.0, 2.0, 3.0, 0, 0, 0, 0)))
    struct.pack_into("<2h", h, 252, 1, 0)
    struct.pack_into("<6f", h, 256, 0.0, 0.0, 0.0, 10.0, 20.0, 30.0)
    p = tmp_path / "q.nii"
    p.write_bytes(bytes(h))
    _, hdr = io_nifti.load(p)
    assert np.allclose(hdr.affine, [[1, 0, 0, 10], [0, 2, 0, 20], [0, 0, -3, 30], [0, 0, 0, 1]])
