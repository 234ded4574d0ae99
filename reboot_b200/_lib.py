"""ctypes loader for the C-ABI library (include/reboot_physics.h). The product path is CUDA only:
loading fails loudly when the shared object is missing -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REBOOT_B200_LIB", os.path.join(_HERE, "csrc", "libreboot_b200.so"))   # env override: A/B builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "reboot_physics.h")

RB_OK, RB_ERR_INVALID, RB_ERR_OOM, RB_ERR_CAPACITY, RB_ERR_CUDA, RB_ERR_NCCL, RB_ERR_NO_DEVICE = range(7)
RB_ACTIVE, RB_GRAVITY, RB_CONTACT = 1, 2, 4
RB_NCCL_ID_BYTES = 128
RB_IPC_HANDLE_BYTES = 80


class RbConfig(C.Structure):
    _fields_ = [("world_half", C.c_float), ("gravity", C.c_float), ("friction", C.c_float), ("pushout", C.c_float),
                ("ss_damp", C.c_float), ("margin_cap", C.c_float), ("device", C.c_int32), ("stream", C.c_void_p),
                ("max_contacts", C.c_uint64), ("slab_lo", C.c_float), ("slab_hi", C.c_float),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("rm_max_hint", C.c_float), ("halo_ghost_cap", C.c_uint32),
                ("sphere_stride", C.c_uint32), ("body_reach_hint", C.c_float), ("slab_forces", C.c_uint32)]


class RbBodyInit(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("vel", C.c_float * 3), ("flags", C.c_uint32)]


class RbContact(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("body", C.c_uint32), ("sphere", C.c_uint32), ("other", C.c_uint32),
                ("other_prim", C.c_uint32), ("depth", C.c_float), ("normal", C.c_float * 3)]


CONTACT_DTYPE = np.dtype([("kind", np.uint32), ("body", np.uint32), ("sphere", np.uint32), ("other", np.uint32),
                          ("other_prim", np.uint32), ("depth", np.float32), ("normal", np.float32, (3,))])


class RbStats(C.Structure):
    _fields_ = [("steps", C.c_uint64), ("tests_ss", C.c_uint64), ("tests_st", C.c_uint64), ("cand_ss", C.c_uint64),
                ("cand_st", C.c_uint64), ("hits_ss", C.c_uint64), ("hits_st", C.c_uint64), ("contacts_last", C.c_uint64),
                ("contacts_dropped", C.c_uint64), ("margin_overflow", C.c_uint64),
                ("n_bodies", C.c_uint32), ("n_spheres", C.c_uint32), ("n_geoms", C.c_uint32), ("n_tris", C.c_uint32),
                ("sphere_cols_x", C.c_uint32), ("sphere_cols_z", C.c_uint32), ("tri_cols_x", C.c_uint32), ("tri_cols_z", C.c_uint32),
                ("device_bytes", C.c_uint64), ("halo_sent_last", C.c_uint32), ("halo_recv_last", C.c_uint32),
                ("migrated_out_last", C.c_uint32), ("migrated_in_last", C.c_uint32),
                ("list_ticks", C.c_uint32), ("list_rebuilds", C.c_uint32), ("list_skin", C.c_float), ("list_hot_last", C.c_uint32),
                ("list_rebuild_reason", C.c_uint32), ("list_rebuilds_by_reason", C.c_uint32 * 4), ("reserved0", C.c_uint32)]


# every symbol include/reboot_physics.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "rb_config_default": (None, [C.POINTER(RbConfig)]),
    "rb_world_create": (C.c_int, [C.POINTER(RbConfig), C.POINTER(_P)]),
    "rb_world_destroy": (None, [_P]),
    "rb_last_error": (C.c_char_p, [_P]),
    "rb_add_static_geometry": (C.c_int, [_P, C.c_uint32, _P, C.POINTER(C.c_uint32)]),
    "rb_add_model": (C.c_int, [_P, C.c_uint32, _P, _P, C.POINTER(RbBodyInit), C.POINTER(C.c_uint32)]),
    "rb_add_models": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_uint32)]),
    "rb_build": (C.c_int, [_P]),
    "rb_step": (C.c_int, [_P, C.c_int]),
    "rb_integrate": (C.c_int, [_P, C.c_int]),
    "rb_collide": (C.c_int, [_P]),
    "rb_detect": (C.c_int, [_P]),
    "rb_sync": (C.c_int, [_P]),
    "rb_query_contacts": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]),
    "rb_contact_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "rb_get_state": (C.c_int, [_P, _P, _P, _P]),
    "rb_get_model_translations": (C.c_int, [_P, _P, _P]),
    "rb_get_model_matrices": (C.c_int, [_P, _P, _P]),
    "rb_export_model_matrices": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "rb_get_angular": (C.c_int, [_P, _P, _P]),
    "rb_get_world_spheres": (C.c_int, [_P, _P]),
    "rb_set_state": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P]),
    "rb_entity_set_position": (C.c_int, [_P, C.c_uint32, _P]),
    "rb_set_flags": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.c_uint32]),
    "rb_set_force": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P]),
    "rb_set_torque": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P]),
    "rb_set_angular": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P]),
    "rb_step_host": (C.c_int, [_P, C.c_int, _P, _P, C.POINTER(C.c_uint64)]),
    "rb_step_host_async": (C.c_int, [_P, C.c_int, _P, _P, C.POINTER(C.c_uint64)]),
    "rb_step_host_async3": (C.c_int, [_P, C.c_int, _P, _P, C.POINTER(C.c_uint64)]),
    "rb_host_row_capacity": (C.c_uint32, [_P]),
    "rb_get_stats": (C.c_int, [_P, C.POINTER(RbStats)]),
    "rb_profile": (C.c_int, [_P, C.c_int]),
    "rb_get_kernel_ms": (C.c_int, [_P, _P, C.POINTER(C.c_uint64)]),
    "rb_launch_count": (C.c_uint64, [_P]),
    "rb_dist_unique_id": (C.c_int, [C.c_char_p, _P]),
    "rb_dist_init": (C.c_int, [_P, C.c_char_p, _P]),
    "rb_dist_p2p_export": (C.c_int, [_P, _P]),
    "rb_dist_p2p_connect": (C.c_int, [_P, _P, _P]),
    "rb_frustum_planes": (C.c_int, [_P, _P]),
    "rb_frustum_cull_cubes": (C.c_int, [_P, _P, C.c_uint32, _P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "rb_frustum_cull_static": (C.c_int, [_P, _P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "rb_get_static_cells": (C.c_int, [_P, _P, _P]),
    "rb_triangle_cube_test": (C.c_int, [_P, C.c_uint32, _P, _P, _P]),
    "rb_ingest_trs_matrix": (C.c_int, [_P, _P]),
    "rb_ingest_mesh_triangles": (C.c_int, [C.c_uint32, _P, C.c_uint32, _P, _P, _P]),
    "rb_ingest_mesh_sphere": (C.c_int, [C.c_uint32, _P, C.c_uint32, _P, _P, _P]),
    "rb_ingest_static_mesh": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, _P, _P, C.POINTER(C.c_uint32)]),
    "rb_halo_ghost_cap_auto": (C.c_uint32, [C.c_uint32, C.c_float, C.c_float, C.c_float]),
    "rb_set_body_ids": (C.c_int, [_P, _P, C.c_uint32]),
    "rb_get_body_ids": (C.c_int, [_P, _P]),
    "rb_dist_attach_peers": (C.c_int, [_P, _P, _P]),
    "rb_step_group": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int]),
    "rb_detect_group": (C.c_int, [C.POINTER(_P), C.c_int]),
    "rb_version": (C.c_char_p, []),
}

_lib = None


def load():
    """dlopen libreboot_b200.so and bind every declared entry point. Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "reboot_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError here = a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def ptr(a):
    """numpy array (C-contiguous) or None -> void*"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
