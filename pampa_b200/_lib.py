"""ctypes binding of include/pampa_sn.h (the C-ABI CUDA layer, libpampa_sn_b200.so).

There is no fallback: if the shared library is missing the import of this module raises, and
every compute entry point fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PAMPA_SN_LIB: another build of the same library (kernel A/B runs); never a different implementation
LIB_PATH = os.environ.get("PAMPA_SN_LIB") or os.path.join(_HERE, "lib", "libpampa_sn_b200.so")

i32, i64, f64 = C.c_int32, C.c_int64, C.c_double
p_i32, p_f64 = C.POINTER(C.c_int32), C.POINTER(C.c_double)


class Mesh(C.Structure):
    _fields_ = [("num_xy_cells", i32), ("num_layers", i32), ("has_z_faces", i32), ("max_xy_faces", i32),
                ("xy_num_faces", p_i32), ("xy_neighbor", p_i32), ("xy_face_fx", p_f64), ("xy_face_fy", p_f64),
                ("xy_face_cf", p_f64), ("xy_area", p_f64), ("xy_cx", p_f64), ("xy_cy", p_f64),
                ("xy_ij", p_i32), ("dz", p_f64), ("materials", p_i32), ("bc_minus_z", i32),
                ("bc_plus_z", i32), ("num_bcs", i32), ("bc_types", p_i32), ("xy_face_kout", p_f64),
                ("xy_face_kin", p_f64), ("face_interpolation_delta", f64)]


class XS(C.Structure):
    _fields_ = [("num_materials", i32), ("num_groups", i32), ("sigma_total", p_f64),
                ("sigma_scattering", p_f64), ("nu_sigma_fission", p_f64), ("kappa_sigma_fission", p_f64),
                ("chi_effective", p_f64), ("beta_total", p_f64)]


class Quadrature(C.Structure):
    _fields_ = [("num_directions", i32), ("directions", p_f64), ("weights", p_f64), ("reflected", p_i32)]


class LS(C.Structure):
    _fields_ = [("num_cells", i32), ("cell", p_i32), ("ptr", p_i32), ("nbr", p_i32), ("omega", p_f64),
                ("nvec", p_f64)]


class Options(C.Structure):
    _fields_ = [("device", i32), ("store_psi", i32), ("patch_cells", i32), ("tile_i", i32), ("tile_j", i32),
                ("z_chunk", i32), ("rank", i32), ("num_ranks", i32), ("shard_mode", i32), ("verbose", i32),
                ("dt_max", i32), ("generic_only", i32), ("single_stream", i32),
                ("anderson_depth", i32), ("wave_launch", i32), ("group_merge", i32), ("inline_edges", i32), ("no_graph", i32),
                ("partition_fields", i32)]


class Info(C.Structure):
    _fields_ = [("num_cells", i64), ("num_groups", i64), ("num_directions", i64), ("updates_per_sweep", i64),
                ("sweep_launches", i64), ("sweep_tasks", i64), ("num_classes", i64), ("num_chunks", i64),
                ("tile_classes", i64), ("device_bytes", i64), ("last_sweep_ms", f64), ("last_source_ms", f64),
                ("last_reduce_ms", f64), ("kernel_launches", i64), ("timed_kernel_ms", f64), ("num_tilings", i64),
                ("flow_classes", i64), ("lattice", i64), ("last_solve_ms", f64),
                ("timed_source_ms", f64), ("timed_reduce_ms", f64), ("timed_exchange_ms", f64)]


# every symbol include/pampa_sn.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pampa_sn_default_options": (None, [C.POINTER(Options)]),
    "pampa_sn_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Mesh), C.POINTER(XS), C.POINTER(Quadrature),
                                  C.POINTER(LS), C.POINTER(Options)]),
    "pampa_sn_destroy": (C.c_int, [C.c_void_p]),
    "pampa_sn_last_error": (C.c_char_p, [C.c_void_p]),
    "pampa_sn_update_xs": (C.c_int, [C.c_void_p, C.POINTER(XS)]),
    "pampa_sn_update_materials": (C.c_int, [C.c_void_p, C.POINTER(XS), p_i32]),
    "pampa_sn_source": (C.c_int, [C.c_void_p, f64]),
    "pampa_sn_sweep": (C.c_int, [C.c_void_p]),
    "pampa_sn_reduce": (C.c_int, [C.c_void_p, p_f64, p_f64, p_f64]),
    "pampa_sn_solve_keff": (C.c_int, [C.c_void_p, f64, f64, i32, f64, p_f64, p_i32]),
    "pampa_sn_iterate": (C.c_int, [C.c_void_p, i32, p_f64]),
    "pampa_sn_iterate_timed": (C.c_int, [C.c_void_p, i32, p_f64, p_f64, p_f64]),
    "pampa_sn_get": (C.c_int, [C.c_void_p, C.c_char_p, p_f64]),
    "pampa_sn_set": (C.c_int, [C.c_void_p, C.c_char_p, p_f64]),
    "pampa_sn_field_size": (i64, [C.c_void_p, C.c_char_p]),
    "pampa_sn_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, i32]),
    "pampa_sn_comm_unique_id": (C.c_int, [C.c_void_p, i32]),
    "pampa_sn_device_ptr": (C.c_void_p, [C.c_void_p, C.c_char_p, C.POINTER(i64)]),
    "pampa_sn_get_info": (C.c_int, [C.c_void_p, C.POINTER(Info)]),
    "pampa_sn_plan_check": (C.c_int, [C.POINTER(Mesh), C.POINTER(Quadrature), i32, C.POINTER(Options),
                                      C.POINTER(Info)]),
}

_lib = None


def load():
    """Load libpampa_sn_b200.so (built in-tree by pampa_b200.build / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -m pampa_b200.build` (there is no CPU fallback)"
                              % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
