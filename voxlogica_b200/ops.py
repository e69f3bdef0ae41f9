"""Host-side mirror of VoxLogicA's operator table on top of the C ABI.

Reference: the ImgQL model layer registers operators by name (Greta.pdf p.43: logical
core, colour & thresholds, normalisation, texture analysis, global volume/min/max);
the F# sources are not mounted, so names follow the operators
used in the specs on p.48 and p.51-53.  Every method is a one-line call into
libvoxcuda.so -- no computation happens in Python, and there is no CPU path.

Images are handles to HBM-resident batches ``[nb][nz][ny][nx]``; numpy arrays only
appear in ``upload`` / ``download``.  Numbers that ImgQL treats as scalars (``volume``,
``min``, ``max``) are numpy vectors with one entry per volume of the batch.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi as A
from ._abi import VoxError, check  # noqa: F401  (re-exported)

_NP_OF = {A.VX_U8: np.uint8, A.VX_F32: np.float32, A.VX_U32: np.uint32}
_VX_OF = {np.dtype(np.uint8): A.VX_U8, np.dtype(np.float32): A.VX_F32, np.dtype(np.uint32): A.VX_U32}
_CMP = {"<": A.VX_LT, "<=": A.VX_LE, ">": A.VX_GT, ">=": A.VX_GE, "==": A.VX_EQ, "!=": A.VX_NE}
_ARITH = {"+": A.VX_ADD, "-": A.VX_SUB, "*": A.VX_MUL, "/": A.VX_DIV, "r-": A.VX_RSUB,
          "r/": A.VX_RDIV, "min": A.VX_MIN, "max": A.VX_MAX}


class Image:
    """Owning wrapper of one vx_img handle (released on garbage collection)."""

    __slots__ = ("ctx", "handle", "__weakref__")

    def __init__(self, ctx: "Context", handle: int):
        self.ctx = ctx
        self.handle = int(handle)

    def __del__(self):
        h, self.handle = getattr(self, "handle", 0), 0
        if h:
            try:
                A.lib().vx_release(h)
            except Exception:  # interpreter shutdown
                pass

    def release(self) -> None:
        h, self.handle = self.handle, 0
        if h:
            check(A.lib().vx_release(h))

    # -- metadata
    def info(self):
        dt, nx, ny, nz, nb = (A.i32() for _ in range(5))
        sp = (A.f32 * 3)()
        check(A.lib().vx_info(self.handle, C.byref(dt), C.byref(nx), C.byref(ny), C.byref(nz),
                              C.byref(nb), sp))
        return dt.value, nx.value, ny.value, nz.value, nb.value, tuple(sp)

    @property
    def dtype(self):
        return _NP_OF[self.info()[0]]

    @property
    def shape(self):
        _, nx, ny, nz, nb, _ = self.info()
        return (nb, nz, ny, nx)

    @property
    def spacing(self):
        return self.info()[5]

    @property
    def nb(self):
        return self.info()[4]

    def numpy(self) -> np.ndarray:
        return self.ctx.download(self)

    # -- operator sugar used by the tests and the ImgQL driver
    def __invert__(self):
        return self.ctx.not_(self)

    def __and__(self, o):
        return self.ctx.and_(self, o)

    def __or__(self, o):
        return self.ctx.or_(self, o)


class Context:
    """One libvoxcuda context = one CUDA device."""

    def __init__(self, device: int = 0):
        self.lib = A.lib()
        h = A.u64()
        check(self.lib.vx_init(device, C.byref(h)))
        self.handle = h.value
        self.device = device

    def close(self):
        if self.handle:
            check(self.lib.vx_shutdown(self.handle))
            self.handle = 0

    def sync(self):
        check(self.lib.vx_sync(self.handle))

    # ------------------------------------------------------------------ transfers
    def upload(self, arr: np.ndarray, spacing: Optional[Sequence[float]] = None,
               batched: Optional[bool] = None) -> Image:
        """numpy -> HBM.  Accepts [ny][nx], [nz][ny][nx] or (batched) [nb][nz][ny][nx]."""
        arr = np.asarray(arr)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        if arr.dtype not in _VX_OF:
            raise TypeError(f"unsupported dtype {arr.dtype}; use uint8, float32 or uint32")
        if batched is None:
            batched = arr.ndim == 4
        if arr.ndim == 2:
            arr = arr[None, None]
        elif arr.ndim == 3:
            arr = arr[None] if not batched else arr[:, None]
        elif arr.ndim != 4:
            raise ValueError("expected a 2D, 3D or batched 4D array")
        arr = np.ascontiguousarray(arr)
        nb, nz, ny, nx = arr.shape
        sp = (A.f32 * 3)(*(spacing if spacing is not None else (1.0, 1.0, 1.0)))
        out = A.u64()
        check(self.lib.vx_upload_batch(self.handle, arr.ctypes.data_as(C.c_void_p), _VX_OF[arr.dtype],
                                       nx, ny, nz, nb, sp, C.byref(out)))
        return Image(self, out.value)

    def upload_ptr(self, ptr: int, dtype, shape, spacing=None) -> Image:
        """Upload from a raw host pointer (e.g. pinned memory from host_alloc)."""
        nb, nz, ny, nx = shape
        sp = (A.f32 * 3)(*(spacing if spacing is not None else (1.0, 1.0, 1.0)))
        out = A.u64()
        check(self.lib.vx_upload_batch(self.handle, C.c_void_p(ptr), _VX_OF[np.dtype(dtype)], nx, ny,
                                       nz, nb, sp, C.byref(out)))
        return Image(self, out.value)

    def upload_scaled(self, raw: np.ndarray, slope: float = 1.0, inter: float = 0.0,
                      spacing: Optional[Sequence[float]] = None) -> Image:
        """16-bit file voxels (int16 / uint16, [nz][ny][nx] or batched 4D) -> f32 image
        ``fl(fl(raw * slope) + inter)`` formed on the device: half the host->device bytes of an f32
        upload, the same bits as ``raw.astype(f32) * f32(slope) + f32(inter)`` on the host."""
        raw = np.asarray(raw)
        if raw.dtype not in (np.int16, np.uint16):
            raise TypeError("upload_scaled takes int16 or uint16 voxels")
        if raw.ndim == 2:
            raw = raw[None, None]
        elif raw.ndim == 3:
            raw = raw[None]
        elif raw.ndim != 4:
            raise ValueError("expected a 2D, 3D or batched 4D array")
        raw = np.ascontiguousarray(raw)
        return self.upload_scaled_ptr(raw.ctypes.data, raw.dtype, raw.shape, slope, inter, spacing)

    def upload_scaled_ptr(self, ptr: int, dtype, shape, slope: float = 1.0, inter: float = 0.0, spacing=None) -> Image:
        nb, nz, ny, nx = shape
        sp = (A.f32 * 3)(*(spacing if spacing is not None else (1.0, 1.0, 1.0)))
        code = A.VX_I16 if np.dtype(dtype) == np.int16 else A.VX_U16
        out = A.u64()
        check(self.lib.vx_upload_scaled(self.handle, C.c_void_p(ptr), code, float(slope), float(inter), nx, ny, nz,
                                        nb, sp, C.byref(out)))
        return Image(self, out.value)

    def download_bits(self, img: Image) -> np.ndarray:
        """boolean image -> u8 array [nb][nz][ny][nx] through the 1-bit-per-voxel transfer"""
        dt, nx, ny, nz, nb, _ = img.info()
        n = nb * nz * ny * nx
        packed = np.empty((n + 7) // 8, np.uint8)
        check(self.lib.vx_download_bits(self.handle, img.handle, packed.ctypes.data_as(C.c_void_p), packed.nbytes))
        return np.unpackbits(packed, count=n, bitorder="little").reshape(nb, nz, ny, nx)

    def download_bits_ptr(self, img: Image, ptr: int, nbytes: int) -> None:
        check(self.lib.vx_download_bits(self.handle, img.handle, C.c_void_p(ptr), nbytes))

    def download(self, img: Image, out: Optional[np.ndarray] = None) -> np.ndarray:
        dt, nx, ny, nz, nb, _ = img.info()
        if out is None:
            out = np.empty((nb, nz, ny, nx), dtype=_NP_OF[dt])
        check(self.lib.vx_download(self.handle, img.handle, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def download_ptr(self, img: Image, ptr: int, nbytes: int) -> None:
        check(self.lib.vx_download(self.handle, img.handle, C.c_void_p(ptr), nbytes))

    def host_alloc(self, nbytes: int, write_combined: bool = False) -> int:
        p = C.c_void_p()
        fn = self.lib.vx_host_alloc_wc if write_combined else self.lib.vx_host_alloc
        check(fn(nbytes, C.byref(p)))
        return p.value

    def host_free(self, ptr: int) -> None:
        check(self.lib.vx_host_free(C.c_void_p(ptr)))

    def batch_slice(self, img: Image, first: int, count: int) -> Image:
        return self._out(lambda o: self.lib.vx_batch_slice(self.handle, img.handle, first, count, o))

    # ------------------------------------------------------------------ slab helpers / interop
    def crop_z(self, img: Image, z0: int, nz: int) -> Image:
        return self._out(lambda o: self.lib.vx_crop_z(self.handle, img.handle, int(z0), int(nz), o))

    def concat_z(self, parts: Sequence[Image]) -> Image:
        arr = (A.u64 * len(parts))(*[p.handle for p in parts])
        return self._out(lambda o: self.lib.vx_concat_z(self.handle, arr, len(parts), o))

    def transfer_from(self, img: Image) -> Image:
        """copy an image of ANOTHER context (GPU) of this process into this one (peer copy)"""
        return self._out(lambda o: self.lib.vx_transfer(self.handle, img.handle, o))

    def ipc_alloc(self, nbytes: int):
        """-> (device pointer, 64-byte IPC handle) of a zero-filled peer-visible buffer"""
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        check(self.lib.vx_ipc_alloc(self.handle, int(nbytes), C.byref(p), h))
        return p.value, h.raw

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        check(self.lib.vx_ipc_open(self.handle, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def ipc_close(self, ptr: int) -> None:
        check(self.lib.vx_ipc_close(self.handle, C.c_void_p(ptr)))

    def ipc_free(self, ptr: int) -> None:
        check(self.lib.vx_ipc_free(self.handle, C.c_void_p(ptr)))

    def copy_planes_to(self, img: Image, z0: int, nz: int, dst_ptr: int) -> None:
        check(self.lib.vx_copy_planes_to(self.handle, img.handle, int(z0), int(nz), C.c_void_p(dst_ptr)))

    def near_halo(self, a: Image, conn: int, lo_ptr: int, hi_ptr: int, erode: bool = False) -> Image:
        fn = self.lib.vx_interior_halo if erode else self.lib.vx_near_halo
        return self._out(lambda o: fn(self.handle, a.handle, conn, C.c_void_p(lo_ptr or None),
                                      C.c_void_p(hi_ptr or None), o))

    def device_ptr(self, img: Image) -> int:
        p = C.c_void_p()
        check(self.lib.vx_device_ptr(self.handle, img.handle, C.byref(p)))
        return p.value or 0

    def stream_handle(self) -> int:
        p = C.c_void_p()
        check(self.lib.vx_stream_handle(self.handle, C.byref(p)))
        return p.value or 0

    def as_tensor(self, img: Image):
        """zero-copy torch view [nb][nz][ny][nx] of an image (CUDA array interface).  The view
        is ordered after the image's producer on THIS thread's library stream; synchronise
        (``sync()``) before handing it to code running on another stream."""
        import torch

        dt, nx, ny, nz, nb, _ = img.info()
        typestr = {A.VX_U8: "|u1", A.VX_F32: "<f4", A.VX_U32: "<u4"}[dt]

        class _View:
            pass

        v = _View()
        v.__cuda_array_interface__ = {"shape": (nb, nz, ny, nx), "typestr": typestr,
                                      "data": (self.device_ptr(img), False), "version": 3, "strides": None}
        v._keepalive = img
        if dt == A.VX_U32:   # torch has no first-class uint32 arithmetic: expose the bits as int32
            v.__cuda_array_interface__["typestr"] = "<i4"
        t = torch.as_tensor(v, device=f"cuda:{self.device}")
        t._vx_keepalive = img
        return t

    def from_tensor(self, t, spacing=None) -> Image:
        """device tensor [nb][nz][ny][nx] (uint8 / float32 / int32 bits of u32) -> new image (D2D copy)"""
        import torch

        t = t.contiguous()
        dt = {torch.uint8: np.uint8, torch.float32: np.float32, torch.int32: np.uint32}[t.dtype]
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        img = self.upload_ptr(t.data_ptr(), dt, tuple(t.shape), spacing)
        self.sync()          # the source tensor may be freed by torch as soon as we return
        return img

    # ------------------------------------------------------------------ helpers
    def _out(self, call) -> Image:
        out = A.u64()
        check(call(C.byref(out)))
        return Image(self, out.value)

    def _scalars(self, call, nb: int) -> np.ndarray:
        buf = (A.f64 * nb)()
        check(call(buf))
        return np.frombuffer(buf, dtype=np.float64).copy()

    @staticmethod
    def _per_volume(v, nb: int, ctype):
        arr = np.broadcast_to(np.asarray(v, dtype=np.float64), (nb,))
        return (ctype * nb)(*[float(t) for t in arr])

    # ------------------------------------------------------------------ A1/A2 pointwise
    def tt(self, like: Image) -> Image:
        return self._out(lambda o: self.lib.vx_const_bool(self.handle, like.handle, 1, o))

    def ff(self, like: Image) -> Image:
        return self._out(lambda o: self.lib.vx_const_bool(self.handle, like.handle, 0, o))

    def constant(self, like: Image, value: float) -> Image:
        return self._out(lambda o: self.lib.vx_const_f32(self.handle, like.handle, value, o))

    def not_(self, a: Image) -> Image:
        return self._out(lambda o: self.lib.vx_not(self.handle, a.handle, o))

    def and_(self, a: Image, b: Image) -> Image:
        return self._out(lambda o: self.lib.vx_and(self.handle, a.handle, b.handle, o))

    def or_(self, a: Image, b: Image) -> Image:
        return self._out(lambda o: self.lib.vx_or(self.handle, a.handle, b.handle, o))

    def xor(self, a: Image, b: Image) -> Image:
        return self._out(lambda o: self.lib.vx_xor(self.handle, a.handle, b.handle, o))

    def cmp_scalar(self, a: Image, op: str, s) -> Image:
        """image OP scalar; ``s`` may be one number or one number per volume."""
        if np.ndim(s) == 0:
            return self._out(lambda o: self.lib.vx_cmp_scalar(self.handle, a.handle, _CMP[op], float(s), o))
        arr = self._per_volume(s, a.nb, A.f32)
        return self._out(lambda o: self.lib.vx_cmp_scalar_batch(self.handle, a.handle, _CMP[op], arr, o))

    def cmp(self, a: Image, op: str, b: Image) -> Image:
        return self._out(lambda o: self.lib.vx_cmp(self.handle, a.handle, b.handle, _CMP[op], o))

    def arith_scalar(self, a: Image, op: str, s) -> Image:
        if np.ndim(s) == 0:
            return self._out(lambda o: self.lib.vx_arith_scalar(self.handle, a.handle, _ARITH[op], float(s), o))
        arr = self._per_volume(s, a.nb, A.f32)
        return self._out(lambda o: self.lib.vx_arith_scalar_batch(self.handle, a.handle, _ARITH[op], arr, o))

    def arith(self, a: Image, op: str, b: Image) -> Image:
        return self._out(lambda o: self.lib.vx_arith(self.handle, a.handle, b.handle, _ARITH[op], o))

    def to_f32(self, a: Image) -> Image:
        return self._out(lambda o: self.lib.vx_to_f32(self.handle, a.handle, o))

    def to_bool(self, a: Image) -> Image:
        return self._out(lambda o: self.lib.vx_to_bool(self.handle, a.handle, o))

    def mask(self, a: Image, m: Image, fill: float = 0.0) -> Image:
        return self._out(lambda o: self.lib.vx_mask(self.handle, a.handle, m.handle, fill, o))

    def fused(self, program, leaves: Sequence[Image], batch_scalars=None) -> Image:
        """program: list of (opcode, dst, a, b, aux, imm) tuples (see voxcuda.h)."""
        ops = (A.vx_op * len(program))(*[A.vx_op(*p) for p in program])
        lv = (A.u64 * len(leaves))(*[l.handle for l in leaves])
        if batch_scalars is not None:
            bs = np.ascontiguousarray(batch_scalars, dtype=np.float32)
            bsp, nsc = bs.ctypes.data_as(C.POINTER(A.f32)), bs.shape[0]
        else:
            bsp, nsc = None, 0
        return self._out(lambda o: self.lib.vx_fused_pointwise(self.handle, ops, len(program), lv,
                                                               len(leaves), bsp, nsc, o))

    # ------------------------------------------------------------------ spatial operators
    def border(self, like: Image) -> Image:
        return self._out(lambda o: self.lib.vx_border(self.handle, like.handle, o))

    def near(self, a: Image, conn: int = A.VX_CONN_FULL) -> Image:
        return self._out(lambda o: self.lib.vx_near(self.handle, a.handle, conn, o))

    def interior(self, a: Image, conn: int = A.VX_CONN_FULL) -> Image:
        return self._out(lambda o: self.lib.vx_interior(self.handle, a.handle, conn, o))

    def through(self, a: Image, b: Image, conn: int = A.VX_CONN_FULL) -> Image:
        return self._out(lambda o: self.lib.vx_through(self.handle, a.handle, b.handle, conn, o))

    def components(self, a: Image, conn: int = A.VX_CONN_FULL) -> Image:
        return self._out(lambda o: self.lib.vx_components(self.handle, a.handle, conn, o))

    # kept union-find state (the stages of `through`, for slab decomposition)
    def ccl_create(self, b: Image, conn: int = A.VX_CONN_FULL) -> int:
        h = A.u64()
        check(self.lib.vx_ccl_create(self.handle, b.handle, conn, C.byref(h)))
        return h.value

    def ccl_seed(self, state: int, a: Image) -> None:
        check(self.lib.vx_ccl_seed(self.handle, state, a.handle))

    def ccl_plane(self, state: int, z: int):
        """-> (u32 label image, u8 reached image) of plane z, each [nb][1][ny][nx]"""
        l, r = A.u64(), A.u64()
        check(self.lib.vx_ccl_plane(self.handle, state, int(z), C.byref(l), C.byref(r)))
        return Image(self, l.value), Image(self, r.value)

    def ccl_flag(self, state: int, labels) -> None:
        arr = np.ascontiguousarray(labels, dtype=np.uint32)
        check(self.lib.vx_ccl_flag(self.handle, state, arr.ctypes.data_as(C.c_void_p), arr.size))

    def ccl_select(self, state: int) -> Image:
        return self._out(lambda o: self.lib.vx_ccl_select(self.handle, state, o))

    def ccl_count(self, state: int) -> int:
        """count the voxels of every local component; returns the largest count"""
        m = A.u64()
        check(self.lib.vx_ccl_count(self.handle, state, C.byref(m)))
        return int(m.value)

    def ccl_sizes(self, state: int, labels) -> np.ndarray:
        arr = np.ascontiguousarray(np.asarray(labels, dtype=np.uint32))
        out = np.zeros(arr.size, np.uint32)
        check(self.lib.vx_ccl_sizes(self.handle, state, arr.ctypes.data_as(C.c_void_p), arr.size,
                                    out.ctypes.data_as(C.c_void_p)))
        return out

    def ccl_flag_size(self, state: int, size: int) -> None:
        check(self.lib.vx_ccl_flag_size(self.handle, state, int(size)))

    def ccl_destroy(self, state: int) -> None:
        check(self.lib.vx_ccl_destroy(self.handle, state))

    def maxvol(self, a: Image, conn: int = A.VX_CONN_FULL) -> Image:
        return self._out(lambda o: self.lib.vx_maxvol(self.handle, a.handle, conn, o))

    def edt(self, a: Image) -> Image:
        return self._out(lambda o: self.lib.vx_edt(self.handle, a.handle, o))

    def edt_sq_xy(self, a: Image) -> Image:
        return self._out(lambda o: self.lib.vx_edt_sq_xy(self.handle, a.handle, o))

    def edt_z_from_sq(self, sq: Image, max_in: float) -> Image:
        return self._out(lambda o: self.lib.vx_edt_z_from_sq(self.handle, sq.handle, float(max_in), o))

    def edt_signed(self, a: Image, conn: int = A.VX_CONN_FULL) -> Image:
        return self._out(lambda o: self.lib.vx_edt_signed(self.handle, a.handle, conn, o))

    def distlt(self, r: float, a: Image) -> Image:
        return self._out(lambda o: self.lib.vx_dist_lt(self.handle, a.handle, float(r), o))

    def distleq(self, r: float, a: Image) -> Image:
        return self._out(lambda o: self.lib.vx_dist_leq(self.handle, a.handle, float(r), o))

    def distgeq(self, r: float, a: Image) -> Image:
        return self._out(lambda o: self.lib.vx_dist_geq(self.handle, a.handle, float(r), o))

    def distgt(self, r: float, a: Image) -> Image:
        return self._out(lambda o: self.lib.vx_dist_gt(self.handle, a.handle, float(r), o))

    def cross_correlation(self, rho: float, a: Image, b: Image, region: Image, m, M, k: int) -> Image:
        nb = a.nb
        mm, MM = self._per_volume(m, nb, A.f64), self._per_volume(M, nb, A.f64)
        return self._out(lambda o: self.lib.vx_cross_correlation(
            self.handle, a.handle, b.handle, region.handle, float(rho), mm, MM, int(k), o))

    def region_histogram(self, b: Image, region: Image, m, M, k: int) -> np.ndarray:
        nb = b.nb
        mm, MM = self._per_volume(m, nb, A.f64), self._per_volume(M, nb, A.f64)
        buf = (A.u64 * (nb * int(k)))()
        check(self.lib.vx_region_histogram(self.handle, b.handle, region.handle, mm, MM, int(k), buf))
        return np.frombuffer(buf, dtype=np.uint64).reshape(nb, int(k)).copy()

    def cross_correlation_h2(self, rho: float, a: Image, h2, m, M, k: int) -> Image:
        nb = a.nb
        mm, MM = self._per_volume(m, nb, A.f64), self._per_volume(M, nb, A.f64)
        h = np.ascontiguousarray(np.broadcast_to(np.asarray(h2, dtype=np.uint64), (nb, int(k))))
        return self._out(lambda o: self.lib.vx_cross_correlation_h2(
            self.handle, a.handle, h.ctypes.data_as(C.POINTER(A.u64)), float(rho), mm, MM, int(k), o))

    # ------------------------------------------------------------------ reductions
    def volume(self, a: Image) -> np.ndarray:
        return self._scalars(lambda b: self.lib.vx_volume(self.handle, a.handle, b), a.nb)

    def min(self, a: Image, mask: Optional[Image] = None) -> np.ndarray:
        if mask is None:
            return self._scalars(lambda b: self.lib.vx_min(self.handle, a.handle, b), a.nb)
        return self._scalars(lambda b: self.lib.vx_min_masked(self.handle, a.handle, mask.handle, b), a.nb)

    def max(self, a: Image, mask: Optional[Image] = None) -> np.ndarray:
        if mask is None:
            return self._scalars(lambda b: self.lib.vx_max(self.handle, a.handle, b), a.nb)
        return self._scalars(lambda b: self.lib.vx_max_masked(self.handle, a.handle, mask.handle, b), a.nb)

    def minmax(self, a: Image, mask: Optional[Image] = None):
        """(min, max) per volume from one pass over the image"""
        lo, hi = (A.f64 * a.nb)(), (A.f64 * a.nb)()
        check(self.lib.vx_minmax(self.handle, a.handle, mask.handle if mask is not None else 0, lo, hi))
        return np.array(lo[:], dtype=np.float64), np.array(hi[:], dtype=np.float64)

    def sum(self, a: Image) -> np.ndarray:
        return self._scalars(lambda b: self.lib.vx_sum(self.handle, a.handle, b), a.nb)

    # building blocks of the slab-decomposed percentiles: torch device tensors of int64 pairs
    # (sortable key << 32 | payload) owned by the caller; see voxlogica_b200/slab.py
    def _torch_ready(self, *tensors):
        import torch
        for t in tensors:
            if t is not None and t.is_cuda:
                torch.cuda.current_stream(t.device).synchronize()

    def pc_pairs(self, a: Image, mask: Image):
        """masked voxels of a single-volume image -> int64 tensor [n] of (key << 32 | voxel index)"""
        import torch
        _, nx, ny, nz, nb, _ = a.info()
        buf = torch.empty(nx * ny * nz, dtype=torch.int64, device=f"cuda:{self.device}")
        self._torch_ready(buf)
        n = A.u64()
        check(self.lib.vx_pc_pairs(self.handle, a.handle, mask.handle, C.c_void_p(buf.data_ptr()), buf.numel(), C.byref(n)))
        return buf[: int(n.value)].clone()

    def pc_sort(self, pairs):
        """stable sort of the pairs by their (unsigned) keys; returns a new tensor"""
        import torch
        out, scratch = pairs.contiguous().clone(), torch.empty_like(pairs)
        self._torch_ready(out, scratch)
        check(self.lib.vx_pc_sort(self.handle, C.c_void_p(out.data_ptr()), out.numel(), C.c_void_p(scratch.data_ptr())))
        return out

    def pc_rank(self, sorted_pairs, less_before: int, n_total: int, correction: float = 0.0):
        """f32 tensor r with r[payload] = (less_before + #keys below + correction * #equal) / n_total"""
        import torch
        dst = torch.zeros(sorted_pairs.numel(), dtype=torch.float32, device=sorted_pairs.device)
        self._torch_ready(sorted_pairs, dst)
        check(self.lib.vx_pc_rank(self.handle, C.c_void_p(sorted_pairs.data_ptr()), sorted_pairs.numel(), int(less_before),
                                  int(n_total), float(correction), C.c_void_p(dst.data_ptr())))
        return dst

    def pc_scatter(self, pairs, values, like: Image) -> Image:
        """new f32 image: out[payload(pairs[i])] = values[i], 0 elsewhere"""
        pairs, values = pairs.contiguous(), values.contiguous()
        self._torch_ready(pairs, values)
        img = self._out(lambda o: self.lib.vx_pc_scatter(self.handle, C.c_void_p(pairs.data_ptr()),
                                                         C.c_void_p(values.data_ptr()), pairs.numel(), like.handle, o))
        self.sync()          # the tensors may be freed by torch as soon as we return
        return img

    def percentiles(self, a: Image, mask: Image, correction: float = 0.0) -> Image:
        return self._out(lambda o: self.lib.vx_percentiles(self.handle, a.handle, mask.handle,
                                                           float(correction), o))

    # ------------------------------------------------------------------ measurement
    def event(self) -> int:
        e = A.u64()
        check(self.lib.vx_event_create(self.handle, C.byref(e)))
        return e.value

    def record(self, ev: int) -> None:
        check(self.lib.vx_event_record(self.handle, ev))

    def elapsed_ms(self, e0: int, e1: int) -> float:
        check(self.lib.vx_event_sync(e1))
        ms = A.f32()
        check(self.lib.vx_event_elapsed_ms(e0, e1, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = A.u64()
        check(self.lib.vx_launch_count(self.handle, C.byref(n)))
        return n.value

    def prof_enable(self, on: bool) -> None:
        check(self.lib.vx_prof_enable(self.handle, 1 if on else 0))

    def prof_reset(self) -> None:
        check(self.lib.vx_prof_reset(self.handle))

    def prof_report(self) -> dict:
        buf = C.create_string_buffer(1 << 16)
        check(self.lib.vx_prof_report(self.handle, buf, len(buf)))
        rep = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(" ", 2)
            rep[name] = (int(cnt), float(ms))
        return rep

    def flush_l2(self) -> None:
        check(self.lib.vx_flush_l2(self.handle))

    def mem_info(self):
        a, b, c = A.u64(), A.u64(), A.u64()
        check(self.lib.vx_mem_info(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return {"live_bytes": a.value, "live_handles": b.value, "pool_reserved": c.value}

    def mem_trim(self, keep_bytes: int = 0) -> None:
        check(self.lib.vx_mem_trim(self.handle, keep_bytes))

    def set_option(self, option: int, value: int) -> None:
        check(self.lib.vx_set_option(self.handle, int(option), int(value)))

    def mem_set_budget(self, nbytes: int) -> None:
        """soft cap on the device memory held by live images, scratch and the free lists (0 = none)"""
        check(self.lib.vx_mem_set_budget(self.handle, int(nbytes)))

    def mem_stats(self):
        a, b, p, c = A.u64(), A.u64(), A.u64(), A.u64()
        check(self.lib.vx_mem_stats(self.handle, C.byref(a), C.byref(b), C.byref(p), C.byref(c)))
        return {"in_use_bytes": a.value, "cached_bytes": b.value, "peak_bytes": p.value, "oom_retries": c.value}
