"""ctypes binding of libgempy_b200.so (the C ABI declared in include/gempy_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load, importing this module's
``lib()`` raises.  The library is built in-tree by ``gempy_b200/csrc/build.py`` (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgempy_b200.so")

GPB_KERNEL = {"cubic": 0, "exponential": 1, "matern_5_2": 2}

c_double_p = C.c_void_p   # device pointers travel as integers
c_int_p = C.c_void_p


class GpbStack(C.Structure):
    _fields_ = [
        ("n_ori", C.c_int), ("n_rest", C.c_int), ("n_surf", C.c_int), ("n_drift", C.c_int),
        ("n_faults", C.c_int), ("kernel", C.c_int),
        ("range", C.c_double), ("c_o", C.c_double), ("i_res", C.c_double), ("gi_res", C.c_double),
        ("ori_pos", C.c_void_p), ("ori_grad", C.c_void_p), ("ori_nugget", C.c_void_p),
        ("rest", C.c_void_p), ("ref", C.c_void_p), ("sp_nugget", C.c_void_p),
        ("fault_rest", C.c_void_p), ("fault_ref", C.c_void_p),
        ("surf_offsets", C.c_void_p), ("ref_unique", C.c_void_p),
    ]


class GpbRegularGrid(C.Structure):
    _fields_ = [
        ("x0", C.c_double), ("y0", C.c_double), ("z0", C.c_double),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
    ]


class GpbModelStack(C.Structure):
    _fields_ = [
        ("st", GpbStack), ("relation", C.c_int), ("fault_stacks_host", C.POINTER(C.c_int)), ("fault_stacks_dev", C.c_void_p),
        ("sp_begin", C.c_int), ("n_sp", C.c_int), ("unit_ids", C.c_void_p), ("weights", C.c_void_p),
        ("eval_table", C.c_void_p), ("isovalues", C.c_void_p),
    ]


class GpbModelDesc(C.Structure):
    _fields_ = [
        ("n_stacks", C.c_int), ("stacks", C.POINTER(GpbModelStack)), ("sp_all", C.c_void_p), ("n_sp_all", C.c_longlong),
        ("sigmoid_slope", C.c_double), ("iso_min", C.c_void_p), ("iso_max", C.c_void_p), ("fault_min", C.c_void_p),
        ("solver", C.c_int),
    ]


GPB_SEG_POINTS, GPB_SEG_REGULAR, GPB_SEG_OCTETS = 0, 1, 2


class GpbSegment(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("count", C.c_longlong), ("out_offset", C.c_longlong), ("xyz", C.c_void_p),
        ("ld_xyz", C.c_longlong), ("grid", GpbRegularGrid), ("i0", C.c_longlong), ("count_dev", C.c_void_p),
    ]


class GpbLevel(C.Structure):
    _fields_ = [
        ("ld", C.c_longlong), ("n_segments", C.c_int), ("segments", C.POINTER(GpbSegment)), ("sp_offset", C.c_longlong),
        ("Z", C.c_void_p), ("G", C.c_void_p), ("block", C.c_void_p), ("final_block", C.c_void_p),
        ("faults_block", C.c_void_p), ("squeezed", C.c_void_p), ("mask", C.c_void_p),
        ("expand_map", C.c_void_p), ("expand_src", C.c_longlong), ("expand_dst", C.c_longlong), ("expand_count", C.c_longlong),
        ("m_combine", C.c_longlong),
    ]


# name -> (restype, argtypes); every symbol include/gempy_b200.h declares
_LL = C.c_longlong
_P = C.c_void_p
SIGNATURES = {
    "gpb_last_error": (C.c_char_p, []),
    "gpb_version": (C.c_int, []),
    "gpb_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gpb_launch_count": (_LL, []),
    "gpb_bench_dfma": (C.c_int, [C.c_int, C.POINTER(C.c_double), _P]),
    "gpb_bench_dmma": (C.c_int, [C.c_int, C.POINTER(C.c_double), _P]),
    "gpb_bench_mixed": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), _P]),
    "gpb_system_size": (C.c_int, [C.POINTER(GpbStack)]),
    "gpb_assemble_cov": (C.c_int, [C.POINTER(GpbStack), _P, C.c_int, _P, _P]),
    "gpb_assemble_cov_ex": (C.c_int, [C.POINTER(GpbStack), _P, C.c_int, _P, C.c_int, _P]),
    "gpb_lu_solve": (C.c_int, [C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P]),
    "gpb_sym_solve": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P, _P]),
    "gpb_lu_set_outer_min_n": (C.c_int, [C.c_int]),
    "gpb_lu_set_outer_width": (C.c_int, [C.c_int]),
    "gpb_lu_factor": (C.c_int, [C.c_int, _P, C.c_int, _P, _P, _P]),
    "gpb_lu_apply": (C.c_int, [C.c_int, _P, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    "gpb_eval_table_doubles": (_LL, [C.POINTER(GpbStack)]),
    "gpb_pack_eval_table": (C.c_int, [C.POINTER(GpbStack), _P, _P, _P]),
    "gpb_eval_regular": (C.c_int, [C.POINTER(GpbStack), _P, C.POINTER(GpbRegularGrid), _LL, _LL, _P, _LL,
                                   _P, _P, _P, _P, _P]),
    "gpb_eval_points": (C.c_int, [C.POINTER(GpbStack), _P, _P, _LL, _LL, _P, _LL, _P, _P, _P, _P, _P]),
    "gpb_model_create": (C.c_int, [C.POINTER(GpbModelDesc), C.POINTER(C.c_void_p)]),
    "gpb_model_destroy": (None, [C.c_void_p]),
    "gpb_model_solve_stack": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GpbLevel), C.POINTER(C.c_int), _P, _LL, _P]),
    "gpb_model_workspace_bytes": (_LL, [C.c_void_p]),
    "gpb_model_eval_stack": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GpbLevel), _P]),
    "gpb_model_combine": (C.c_int, [C.c_void_p, C.POINTER(GpbLevel), _P]),
    "gpb_model_run_level": (C.c_int, [C.c_void_p, C.POINTER(GpbLevel), C.c_int, _P, _LL, _P]),
    "gpb_model_solver_path": (C.c_int, [C.c_void_p, C.c_int]),
    "gpb_corner_scratch_bytes": (_LL, [_LL]),
    "gpb_corner_unique_count": (C.c_int, [_P, _LL, _LL, C.POINTER(GpbRegularGrid), _P, _LL, C.POINTER(_LL), _P]),
    "gpb_corner_unique_emit": (C.c_int, [_P, _LL, _LL, C.c_double, C.c_double, C.c_double, _P, _LL, _P, _LL, _P, _LL, _P, _P]),
    "gpb_expand_rows": (C.c_int, [_P, _LL, _P, C.c_int, _LL, _P, _LL, _P]),
    "gpb_copy_2d": (C.c_int, [_P, _LL, _P, _LL, _LL, _LL, _P]),
    "gpb_rint": (C.c_int, [_P, _LL, _P, _P]),
    "gpb_scan_elems": (_LL, [_LL]),
    "gpb_count_marked": (C.c_int, [_P, _LL, _P, C.POINTER(_LL), _P]),
    "gpb_emit_marked": (C.c_int, [_P, _LL, _LL, _P, _P, C.c_double, C.c_double, C.c_double, _P, _LL, _P]),
    "gpb_upsample2": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "gpb_scatter_lattice": (C.c_int, [_P, _LL, _LL, C.POINTER(GpbRegularGrid), _P, C.c_int, _P, _P]),
    "gpb_dc_scratch_bytes": (_LL, [_LL]),
    "gpb_dual_contour": (C.c_int, [C.POINTER(GpbStack), _P, _P, _LL, _P, _P, _P, _LL, _LL, _P, C.POINTER(GpbRegularGrid),
                                   C.c_double, _P, _LL, _P, _P, _P, _P, _P, _P, _P]),
    "gpb_activate": (C.c_int, [_P, _LL, _P, _P, C.c_int, C.c_double, _P, _P]),
    "gpb_min": (C.c_int, [_P, _LL, _P, _P]),
    "gpb_shift": (C.c_int, [_P, _LL, _P, _P, _P]),
    "gpb_combine": (C.c_int, [_P, _P, _LL, _LL, C.c_int, C.POINTER(C.c_int), _P, _P, _P, _P, _P, _P, _P]),
    "gpb_voxel_corners": (C.c_int, [_P, _LL, _LL, C.c_double, C.c_double, C.c_double, _P, _LL, _P]),
    "gpb_mark_voxels": (C.c_int, [_P, _P, _LL, C.c_int, _P, _P]),
    "gpb_emit_children": (C.c_int, [_P, _LL, _LL, _P, C.c_double, C.c_double, C.c_double, _P, _LL,
                                    C.POINTER(_LL), _P]),
    "gpb_any8": (C.c_int, [_P, _LL, _P, _P]),
    "gpb_gravity": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _LL, _P, _P]),
    "gpb_dc_edges": (C.c_int, [_P, _LL, _P, _LL, C.c_double, _P, _P, _P, _P]),
    "gpb_dc_vertices": (C.c_int, [_P, _P, _P, _LL, C.c_double, _P, _P]),
    "gpb_mc_scratch_elems": (_LL, [_LL]),
    "gpb_mc_count": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_double, _P, _P, C.POINTER(_LL), C.POINTER(_LL), _P]),
    "gpb_mc_emit": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_double] + [C.c_double] * 6 + [_P, _P, _P, _P]),
}

_lib = None


class GpbError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the library.  Raises if it was not built -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GpbError(f"{LIB_PATH} not found: build it with `python gempy_b200/csrc/build.py` "
                           "(__graft_entry__.build()); the B200 backend has no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().gpb_last_error().decode("utf-8", "replace")
        raise GpbError(f"gempy_b200 error {rc}: {msg}")


def launch_count() -> int:
    return int(lib().gpb_launch_count())
