"""ctypes binding of libvoxcuda.so (include/voxcuda.h).

This is the Python stand-in for the F# P/Invoke table a VoxLogicA maintainer would
write: one entry per exported function, blittable types only.  The
library is built in-tree by ``__graft_entry__.build()`` (``voxlogica_b200/csrc/Makefile``)
and there is deliberately no fallback: if the shared object is missing, importing
fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VOXCUDA_LIB: another build of the library (A/B measurements); the default is the in-tree one
LIB_PATH = os.environ.get("VOXCUDA_LIB") or os.path.join(_HERE, "libvoxcuda.so")

# status codes / enums (mirror include/voxcuda.h)
VX_OK = 0
VX_E_INVALID_ARG, VX_E_SHAPE_MISMATCH, VX_E_DTYPE, VX_E_OOM = -1, -2, -3, -4
VX_E_CUDA, VX_E_NCCL, VX_E_UNSUPPORTED = -5, -6, -7
VX_U8, VX_F32, VX_U32 = 1, 2, 3
VX_OPT_QUANTISED_PATHS = 1
VX_OPT_ORDERED_UPLOADS = 2
VX_I16, VX_U16 = 4, 5   # host-side voxel types of vx_upload_scaled
VX_LT, VX_LE, VX_GT, VX_GE, VX_EQ, VX_NE = range(6)
VX_ADD, VX_SUB, VX_MUL, VX_DIV, VX_RSUB, VX_RDIV, VX_MIN, VX_MAX = range(8)
VX_CONN_FACE, VX_CONN_FULL = 6, 26
(VXP_LOAD, VXP_NOT, VXP_AND, VXP_OR, VXP_XOR, VXP_CMPS, VXP_CMP, VXP_ARITHS, VXP_ARITH,
 VXP_TOF32, VXP_TOBOOL, VXP_CONSTB, VXP_CONSTF, VXP_SELECT, VXP_CMPSB, VXP_ARITHSB) = range(16)
VXP_MAX_OPS, VXP_MAX_REGS, VXP_MAX_LEAVES = 64, 24, 8

_ERR_NAMES = {
    -1: "VX_E_INVALID_ARG", -2: "VX_E_SHAPE_MISMATCH", -3: "VX_E_DTYPE", -4: "VX_E_OOM",
    -5: "VX_E_CUDA", -6: "VX_E_NCCL", -7: "VX_E_UNSUPPORTED",
}


class VoxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class vx_op(C.Structure):
    _fields_ = [("opcode", C.c_int32), ("dst", C.c_int32), ("a", C.c_int32), ("b", C.c_int32),
                ("aux", C.c_int32), ("imm", C.c_float)]


u64, i32, f32, f64 = C.c_uint64, C.c_int32, C.c_float, C.c_double
P = C.POINTER

# name -> argtypes.  Every function returns int32 except where noted.
SIGNATURES = {
    "vx_init": [i32, P(u64)],
    "vx_shutdown": [u64],
    "vx_sync": [u64],
    "vx_upload": [u64, C.c_void_p, i32, i32, i32, i32, P(f32), P(u64)],
    "vx_upload_batch": [u64, C.c_void_p, i32, i32, i32, i32, i32, P(f32), P(u64)],
    "vx_upload_scaled": [u64, C.c_void_p, i32, f32, f32, i32, i32, i32, i32, P(f32), P(u64)],
    "vx_download": [u64, u64, C.c_void_p, C.c_size_t],
    "vx_download_bits": [u64, u64, C.c_void_p, C.c_size_t],
    "vx_info": [u64, P(i32), P(i32), P(i32), P(i32), P(i32), P(f32)],
    "vx_retain": [u64],
    "vx_release": [u64],
    "vx_batch_slice": [u64, u64, i32, i32, P(u64)],
    "vx_crop_z": [u64, u64, i32, i32, P(u64)],
    "vx_concat_z": [u64, P(u64), i32, P(u64)],
    "vx_transfer": [u64, u64, P(u64)],
    "vx_device_ptr": [u64, u64, P(C.c_void_p)],
    "vx_stream_handle": [u64, P(C.c_void_p)],
    "vx_ipc_alloc": [u64, C.c_size_t, P(C.c_void_p), C.c_void_p],
    "vx_ipc_open": [u64, C.c_void_p, P(C.c_void_p)],
    "vx_ipc_close": [u64, C.c_void_p],
    "vx_ipc_free": [u64, C.c_void_p],
    "vx_copy_planes_to": [u64, u64, i32, i32, C.c_void_p],
    "vx_host_alloc": [C.c_size_t, P(C.c_void_p)],
    "vx_host_alloc_wc": [C.c_size_t, P(C.c_void_p)],
    "vx_host_free": [C.c_void_p],
    "vx_mem_info": [u64, P(u64), P(u64), P(u64)],
    "vx_mem_trim": [u64, u64],
    "vx_set_option": [u64, i32, C.c_int64],
    "vx_mem_set_budget": [u64, u64],
    "vx_mem_stats": [u64, P(u64), P(u64), P(u64), P(u64)],
    "vx_const_bool": [u64, u64, i32, P(u64)],
    "vx_const_f32": [u64, u64, f32, P(u64)],
    "vx_not": [u64, u64, P(u64)],
    "vx_and": [u64, u64, u64, P(u64)],
    "vx_or": [u64, u64, u64, P(u64)],
    "vx_xor": [u64, u64, u64, P(u64)],
    "vx_cmp_scalar": [u64, u64, i32, f32, P(u64)],
    "vx_cmp_scalar_batch": [u64, u64, i32, P(f32), P(u64)],
    "vx_cmp": [u64, u64, u64, i32, P(u64)],
    "vx_arith_scalar": [u64, u64, i32, f32, P(u64)],
    "vx_arith_scalar_batch": [u64, u64, i32, P(f32), P(u64)],
    "vx_arith": [u64, u64, u64, i32, P(u64)],
    "vx_to_f32": [u64, u64, P(u64)],
    "vx_to_bool": [u64, u64, P(u64)],
    "vx_mask": [u64, u64, u64, f32, P(u64)],
    "vx_fused_pointwise": [u64, P(vx_op), i32, P(u64), i32, P(f32), i32, P(u64)],
    "vx_border": [u64, u64, P(u64)],
    "vx_near": [u64, u64, i32, P(u64)],
    "vx_interior": [u64, u64, i32, P(u64)],
    "vx_near_halo": [u64, u64, i32, C.c_void_p, C.c_void_p, P(u64)],
    "vx_interior_halo": [u64, u64, i32, C.c_void_p, C.c_void_p, P(u64)],
    "vx_through": [u64, u64, u64, i32, P(u64)],
    "vx_components": [u64, u64, i32, P(u64)],
    "vx_ccl_create": [u64, u64, i32, P(u64)],
    "vx_ccl_seed": [u64, u64, u64],
    "vx_ccl_plane": [u64, u64, i32, P(u64), P(u64)],
    "vx_ccl_flag": [u64, u64, C.c_void_p, u64],
    "vx_ccl_select": [u64, u64, P(u64)],
    "vx_ccl_count": [u64, u64, P(u64)],
    "vx_ccl_sizes": [u64, u64, C.c_void_p, u64, C.c_void_p],
    "vx_ccl_flag_size": [u64, u64, C.c_uint32],
    "vx_ccl_destroy": [u64, u64],
    "vx_maxvol": [u64, u64, i32, P(u64)],
    "vx_edt": [u64, u64, P(u64)],
    "vx_edt_signed": [u64, u64, i32, P(u64)],
    "vx_edt_sq_xy": [u64, u64, P(u64)],
    "vx_edt_z_from_sq": [u64, u64, f64, P(u64)],
    "vx_dist_lt": [u64, u64, f32, P(u64)],
    "vx_dist_leq": [u64, u64, f32, P(u64)],
    "vx_dist_geq": [u64, u64, f32, P(u64)],
    "vx_dist_gt": [u64, u64, f32, P(u64)],
    "vx_cross_correlation": [u64, u64, u64, u64, f32, P(f64), P(f64), i32, P(u64)],
    "vx_region_histogram": [u64, u64, u64, P(f64), P(f64), i32, P(u64)],
    "vx_cross_correlation_h2": [u64, u64, P(u64), f32, P(f64), P(f64), i32, P(u64)],
    "vx_xcorr_bin_edges": [f64, f64, i32, P(f32)],
    "vx_xcorr_region_stat": [P(u64), i32, P(f64), P(i32)],
    "vx_volume": [u64, u64, P(f64)],
    "vx_min": [u64, u64, P(f64)],
    "vx_max": [u64, u64, P(f64)],
    "vx_min_masked": [u64, u64, u64, P(f64)],
    "vx_max_masked": [u64, u64, u64, P(f64)],
    "vx_sum": [u64, u64, P(f64)],
    "vx_minmax": [u64, u64, u64, P(f64), P(f64)],
    "vx_pc_pairs": [u64, u64, u64, C.c_void_p, u64, P(u64)],
    "vx_pc_sort": [u64, C.c_void_p, u64, C.c_void_p],
    "vx_pc_rank": [u64, C.c_void_p, u64, u64, u64, f64, C.c_void_p],
    "vx_pc_scatter": [u64, C.c_void_p, C.c_void_p, u64, u64, P(u64)],
    "vx_percentiles": [u64, u64, u64, f64, P(u64)],
    "vx_event_create": [u64, P(u64)],
    "vx_event_record": [u64, u64],
    "vx_event_sync": [u64],
    "vx_event_elapsed_ms": [u64, u64, P(f32)],
    "vx_event_destroy": [u64],
    "vx_launch_count": [u64, P(u64)],
    "vx_prof_enable": [u64, i32],
    "vx_prof_reset": [u64],
    "vx_prof_report": [u64, C.c_char_p, C.c_size_t],
    "vx_flush_l2": [u64],
}


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen libvoxcuda.so and attach the signatures.  No GPU is needed for this."""
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C voxlogica_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.argtypes = argtypes
        fn.restype = i32
    lib.vx_version.argtypes = []
    lib.vx_version.restype = i32
    lib.vx_last_error.argtypes = []
    lib.vx_last_error.restype = C.c_char_p
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def check(rc: int) -> None:
    if rc != VX_OK:
        msg = lib().vx_last_error()
        raise VoxError(rc, msg.decode("utf-8", "replace") if msg else "")
