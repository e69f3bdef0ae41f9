"""numpy front end of the CPU oracle (oracle/vx_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of vx_oracle.c.  Imported by tests/,
``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs; never by the product
package ``voxlogica_b200``.  PARITY UNPINNED: the reference's implementation is not
mounted and SimpleITK is not installed, so this restates the documented semantics.

All functions take and return single volumes shaped [nz][ny][nx] (or [ny][nx]); use
``batched(fn, ...)`` to map over a leading batch axis.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

CMP = {"<": 0, "<=": 1, ">": 2, ">=": 3, "==": 4, "!=": 5}
ARITH = {"+": 0, "-": 1, "*": 2, "/": 3, "r-": 4, "r/": 5, "min": 6, "max": 7}


def _cpu_stamp() -> str:
    """Feature flags of this host's CPU: the library is built -march=native, so a copy built on
    another machine (the build container vs the GPU box) must be rebuilt before it is loaded."""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return " ".join(sorted(line.split(":", 1)[1].split()))
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vx_oracle.c")
    stamp_file = os.path.join(_HERE, "_build", "cpu_flags.txt")
    stamp = _cpu_stamp()
    try:
        same_cpu = open(stamp_file).read() == stamp
    except OSError:
        same_cpu = False
    if (force or not same_cpu or not os.path.exists(_SO)
            or os.path.getmtime(_SO) < os.path.getmtime(src)
            or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "Makefile"))):
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory", "-B"], stdout=subprocess.DEVNULL)
        with open(stamp_file, "w") as f:
            f.write(stamp)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for name in ("or_volume", "or_min", "or_max", "or_sum"):
            getattr(_lib, name).restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(np.asarray(a) != 0, dtype=np.uint8)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _dims(a):
    a = np.asarray(a)
    if a.ndim == 2:
        return a.shape[1], a.shape[0], 1
    if a.ndim == 3:
        return a.shape[2], a.shape[1], a.shape[0]
    raise ValueError("oracle functions take one [nz][ny][nx] or [ny][nx] volume")


def _sp(spacing):
    return (C.c_float * 3)(*(spacing if spacing is not None else (1.0, 1.0, 1.0)))


def batched(fn, *arrays, **kw):
    """Apply a single-volume oracle function to every volume of [nb][nz][ny][nx] arrays."""
    nb = arrays[0].shape[0]
    return np.stack([fn(*[a[i] for a in arrays], **kw) for i in range(nb)])


# ---- pointwise -------------------------------------------------------------------------
def not_(a):
    a = _u8(a); out = np.empty_like(a)
    lib().or_not(_p(a), _p(out), C.c_size_t(a.size)); return out


def _bool2(name, a, b):
    a, b = _u8(a), _u8(b); out = np.empty_like(a)
    getattr(lib(), name)(_p(a), _p(b), _p(out), C.c_size_t(a.size)); return out


def and_(a, b): return _bool2("or_and", a, b)
def or_(a, b): return _bool2("or_or", a, b)
def xor(a, b): return _bool2("or_xor", a, b)


def cmp_scalar(a, op, s):
    a = _f32(a); out = np.empty(a.shape, np.uint8)
    lib().or_cmp_scalar(_p(a), CMP[op], C.c_float(s), _p(out), C.c_size_t(a.size)); return out


def cmp(a, op, b):
    a, b = _f32(a), _f32(b); out = np.empty(a.shape, np.uint8)
    lib().or_cmp(_p(a), _p(b), CMP[op], _p(out), C.c_size_t(a.size)); return out


def arith_scalar(a, op, s):
    a = _f32(a); out = np.empty_like(a)
    lib().or_arith_scalar(_p(a), ARITH[op], C.c_float(s), _p(out), C.c_size_t(a.size)); return out


def arith(a, op, b):
    a, b = _f32(a), _f32(b); out = np.empty_like(a)
    lib().or_arith(_p(a), _p(b), ARITH[op], _p(out), C.c_size_t(a.size)); return out


def to_f32(a):
    a = _u8(a); out = np.empty(a.shape, np.float32)
    lib().or_to_f32(_p(a), _p(out), C.c_size_t(a.size)); return out


def to_bool(a):
    a = _f32(a); out = np.empty(a.shape, np.uint8)
    lib().or_to_bool(_p(a), _p(out), C.c_size_t(a.size)); return out


def mask(a, m, fill=0.0):
    a, m = _f32(a), _u8(m); out = np.empty_like(a)
    lib().or_mask(_p(a), _p(m), C.c_float(fill), _p(out), C.c_size_t(a.size)); return out


# ---- spatial ---------------------------------------------------------------------------
def border(like):
    nx, ny, nz = _dims(like); out = np.empty(np.asarray(like).shape, np.uint8)
    lib().or_border(_p(out), nx, ny, nz); return out


def _morph(name, a, conn):
    a = _u8(a); nx, ny, nz = _dims(a); out = np.empty_like(a)
    getattr(lib(), name)(_p(a), _p(out), nx, ny, nz, conn); return out


def near(a, conn=26): return _morph("or_near", a, conn)
def interior(a, conn=26): return _morph("or_interior", a, conn)
def maxvol(a, conn=26): return _morph("or_maxvol", a, conn)


def through(a, b, conn=26):
    a, b = _u8(a), _u8(b); nx, ny, nz = _dims(a); out = np.empty_like(a)
    lib().or_through(_p(a), _p(b), _p(out), nx, ny, nz, conn); return out


def components(a, conn=26):
    a = _u8(a); nx, ny, nz = _dims(a); out = np.empty(a.shape, np.uint32)
    lib().or_components(_p(a), _p(out), nx, ny, nz, conn); return out


def edt(a, spacing=None):
    a = _u8(a); nx, ny, nz = _dims(a); out = np.empty(a.shape, np.float32)
    lib().or_edt(_p(a), _p(out), nx, ny, nz, _sp(spacing)); return out


def edt_signed(a, spacing=None, conn=26):
    a = _u8(a); nx, ny, nz = _dims(a); out = np.empty(a.shape, np.float32)
    lib().or_edt_signed(_p(a), _p(out), nx, ny, nz, _sp(spacing), conn); return out


def dist_cmp(a, op, r, spacing=None):
    a = _u8(a); nx, ny, nz = _dims(a); out = np.empty_like(a)
    lib().or_dist_cmp(_p(a), CMP[op], C.c_float(r), _p(out), nx, ny, nz, _sp(spacing)); return out


def distlt(r, a, spacing=None): return dist_cmp(a, "<", r, spacing)
def distleq(r, a, spacing=None): return dist_cmp(a, "<=", r, spacing)
def distgeq(r, a, spacing=None): return dist_cmp(a, ">=", r, spacing)
def distgt(r, a, spacing=None): return dist_cmp(a, ">", r, spacing)


def cross_correlation(rho, a, b, region, m, M, k, spacing=None, textbook=False):
    a, b, region = _f32(a), _f32(b), _u8(region)
    nx, ny, nz = _dims(a); out = np.empty_like(a)
    lib().or_cross_correlation(_p(a), _p(b), _p(region), C.c_float(rho), C.c_double(m), C.c_double(M),
                               int(k), _p(out), nx, ny, nz, _sp(spacing), 1 if textbook else 0)
    return out


def cross_correlation_h2(rho, a, h2, m, M, k, spacing=None):
    a = _f32(a); h2 = np.ascontiguousarray(h2, dtype=np.uint64)
    nx, ny, nz = _dims(a); out = np.empty_like(a)
    lib().or_cross_correlation_h2(_p(a), _p(h2), C.c_float(rho), C.c_double(m), C.c_double(M), int(k),
                                  _p(out), nx, ny, nz, _sp(spacing))
    return out


# ---- reductions ------------------------------------------------------------------------
def volume(a):
    a = _u8(a); return lib().or_volume(_p(a), C.c_size_t(a.size))


def min_(a, mask=None):
    a = _f32(a); m = _u8(mask) if mask is not None else None
    return lib().or_min(_p(a), _p(m) if m is not None else None, C.c_size_t(a.size))


def max_(a, mask=None):
    a = _f32(a); m = _u8(mask) if mask is not None else None
    return lib().or_max(_p(a), _p(m) if m is not None else None, C.c_size_t(a.size))


def sum_(a):
    a = _f32(a); return lib().or_sum(_p(a), C.c_size_t(a.size))


def percentiles(a, mask, correction=0.0):
    a, mask = _f32(a), _u8(mask); out = np.empty_like(a)
    lib().or_percentiles(_p(a), _p(mask), C.c_double(correction), _p(out), C.c_size_t(a.size)); return out


# ---- derived operators, restated from the slide semantics --
def touch(a, b, conn=26):
    """touch(a,b) = a & through(near(b), a): the part of a connected, through a, to b
    (Greta.pdf p.51 "the dark area that touches the border"; reach with the 0<j<l side
    condition of p.38)."""
    return and_(a, through(near(b, conn), a, conn))


def grow(a, b, conn=26):
    """grow(a,b) = a | touch(b,a)  (Greta.pdf p.53 "hyperintense areas + very intense
    touching them")."""
    return or_(a, touch(b, a, conn))


def flt(r, a, spacing=None):
    """filter(r,a) = distlt(r, distgeq(r, !a)): Euclidean opening (Greta.pdf p.52 noise removal)."""
    return distlt(r, distgeq(r, not_(a), spacing), spacing)
