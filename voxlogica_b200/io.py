"""Image I/O at the edge of the hot path.

VoxLogicA loads "2D and 3D images (png, bmp, nifti, ...)" (Greta.pdf p.43; the GTV spec loads
``"flair.nii.gz"``, p.48) through SimpleITK, which is not available here -- neither is nibabel.
This module is a dependency-free NIfTI-1 single-file (``.nii`` / ``.nii.gz``) reader/writer
(348-byte header, little or big endian, scl_slope/scl_inter applied, pixdim -> spacing) plus PNG /
BMP through Pillow.  Arrays are returned in the library's layout ``[nz][ny][nx]`` (x fastest),
which is exactly NIfTI's on-disk order, with the spacing as ``(sx, sy, sz)``.

Host-side only: nothing here touches the GPU; ``load_image`` feeds ``Context.upload``.
"""
from __future__ import annotations

import gzip
import struct
from typing import Tuple

import numpy as np

_NIFTI_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64,
                 256: np.int8, 512: np.uint16, 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_NIFTI_CODES = {np.dtype(v): k for k, v in _NIFTI_DTYPES.items()}


def _open(path: str, mode: str):
    return gzip.open(path, mode) if path.endswith(".gz") else open(path, mode)


def read_nifti(path: str) -> Tuple[np.ndarray, Tuple[float, float, float]]:
    """-> (f32 array [nz][ny][nx] (or [nt]... squeezed to 3D), (sx, sy, sz))"""
    data, slope, inter, sp = read_nifti_raw(path)
    arr = data.astype(np.float32)
    if slope != 1.0 or inter != 0.0:
        arr = arr * np.float32(slope) + np.float32(inter)
    return np.ascontiguousarray(arr), sp


def read_nifti_raw(path: str):
    """-> (voxels in the file's own dtype [nz][ny][nx] (native byte order), slope, inter, (sx, sy, sz)) with
    value = voxel * slope + inter (slope 1, inter 0 when the header asks for no scaling).  16-bit
    files can go to the GPU as they are: ``Context.upload_scaled(raw, slope, inter)`` halves the
    host->device bytes and gives the same f32 image as ``read_nifti`` + ``upload``."""
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError(f"{path}: shorter than a NIfTI-1 header")
    end = "<"
    if struct.unpack("<i", raw[:4])[0] != 348:
        end = ">"
        if struct.unpack(">i", raw[:4])[0] != 348:
            raise ValueError(f"{path}: sizeof_hdr is not 348 (not NIfTI-1)")
    if raw[344:347] not in (b"n+1", b"ni1"):
        raise ValueError(f"{path}: bad NIfTI magic {raw[344:348]!r}")
    if raw[344:347] == b"ni1":
        raise ValueError(f"{path}: two-file NIfTI (.hdr/.img) is not supported")
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(end + "hh", raw[70:74])
    pixdim = struct.unpack(end + "8f", raw[76:108])
    vox_offset = int(struct.unpack(end + "f", raw[108:112])[0])
    slope, inter = struct.unpack(end + "ff", raw[112:120])
    if datatype not in _NIFTI_DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype {datatype}")
    nd = dim[0]
    if not 1 <= nd <= 7:
        raise ValueError(f"{path}: bad dim[0] = {nd}")
    shape = [max(1, d) for d in dim[1:1 + nd]]
    while len(shape) > 3 and shape[-1] == 1:
        shape.pop()
    if len(shape) > 3:
        raise ValueError(f"{path}: {len(shape)}-D data; split the volumes first")
    while len(shape) < 3:
        shape.append(1)
    nx, ny, nz = shape
    dt = np.dtype(_NIFTI_DTYPES[datatype]).newbyteorder(end)
    n = nx * ny * nz
    data = np.frombuffer(raw, dtype=dt, count=n, offset=max(vox_offset, 352))
    data = np.ascontiguousarray(data.reshape(nz, ny, nx).astype(dt.newbyteorder("=")))
    if slope == 0.0:      # NIfTI-1: a zero slope means "no scaling"
        slope, inter = 1.0, 0.0
    sp = tuple(float(abs(p)) if p != 0 else 1.0 for p in pixdim[1:4])
    return data, float(slope), float(inter), sp


def write_nifti(path: str, arr: np.ndarray, spacing=(1.0, 1.0, 1.0), slope: float = 1.0, inter: float = 0.0) -> None:
    """arr: [nz][ny][nx] (or [ny][nx]); uint8 / int16 / int32 / float32 / float64.  ``slope`` / ``inter`` go to the
    header's scl_slope / scl_inter (value = voxel * slope + inter), as scanners store 12-bit data in int16 files."""
    a = np.ascontiguousarray(arr)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("write_nifti expects a 2D or 3D array")
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype not in _NIFTI_CODES:
        raise TypeError(f"unsupported dtype {a.dtype}")
    nz, ny, nx = a.shape
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into("<hh", hdr, 70, _NIFTI_CODES[a.dtype], a.dtype.itemsize * 8)
    struct.pack_into("<8f", hdr, 76, 1.0, float(spacing[0]), float(spacing[1]), float(spacing[2]), 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<ff", hdr, 112, float(slope), float(inter))
    hdr[123] = 2  # xyzt_units: millimetres
    struct.pack_into("<hh", hdr, 252, 0, 2)  # qform_code 0, sform_code 2 (aligned)
    struct.pack_into("<4f", hdr, 280, float(spacing[0]), 0.0, 0.0, 0.0)
    struct.pack_into("<4f", hdr, 296, 0.0, float(spacing[1]), 0.0, 0.0)
    struct.pack_into("<4f", hdr, 312, 0.0, 0.0, float(spacing[2]), 0.0)
    hdr[344:348] = b"n+1\0"
    with _open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(b"\0\0\0\0")
        f.write(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())


def read_png(path: str) -> np.ndarray:
    """-> uint8 array [ny][nx] (grey) or [ny][nx][3|4] (colour), as stored"""
    from PIL import Image as PILImage
    with PILImage.open(path) as im:
        if im.mode not in ("L", "RGB", "RGBA"):
            im = im.convert("RGB" if im.mode not in ("I;16", "I", "F", "1") else "L")
        return np.asarray(im).copy()


def write_png(path: str, arr: np.ndarray) -> None:
    """boolean / u8 [ny][nx] (0/1 masks are scaled to 0/255) or [ny][nx][3] RGB"""
    from PIL import Image as PILImage
    a = np.asarray(arr)
    if a.dtype == np.bool_ or (a.dtype == np.uint8 and a.ndim == 2 and a.max(initial=0) <= 1):
        a = a.astype(np.uint8) * 255
    PILImage.fromarray(a.astype(np.uint8)).save(path)


def load_image(path: str):
    """Dispatch on the extension -> dict of channel name -> (f32 array, spacing).  Single-channel
    images give {"intensity": ...}; colour images give "red" "green" "blue" (and "alpha")."""
    p = path.lower()
    if p.endswith((".nii", ".nii.gz")):
        a, sp = read_nifti(path)
        return {"intensity": (a, sp)}
    a = read_png(path)
    sp = (1.0, 1.0, 1.0)
    if a.ndim == 2:
        return {"intensity": (a.astype(np.float32)[None], sp)}
    names = ("red", "green", "blue", "alpha")
    return {names[c]: (np.ascontiguousarray(a[..., c]).astype(np.float32)[None], sp) for c in range(a.shape[2])}


def save_image(path: str, arr: np.ndarray, spacing=(1.0, 1.0, 1.0)) -> None:
    p = path.lower()
    if p.endswith((".nii", ".nii.gz")):
        write_nifti(path, arr, spacing)
    else:
        a = np.asarray(arr)
        write_png(path, a[0] if a.ndim == 3 and a.shape[0] == 1 else a)
