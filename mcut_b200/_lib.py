"""ctypes declaration of the C-ABI in include/mcut_b200.h (mcut_b200/lib/libmcut_b200.so).

The library is the product; this module only binds it.  If it has not been built, or no B200 is present when a
compute entry point is called, the error is raised as is — there is no CPU path behind these calls.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmcut_b200.so")
CSRC = os.path.join(_HERE, "csrc")

c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)


class CutpathCounts(C.Structure):
    _fields_ = [("n_groups", C.c_uint64), ("n_entries", C.c_uint64), ("n_single_point_groups", C.c_uint64)]

c_u64p = C.POINTER(C.c_uint64)
c_i32p = C.POINTER(C.c_int32)
vp = C.c_void_p


class Counts(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("n_node_tests", C.c_uint64), ("n_tests", C.c_uint64), ("n_exact", C.c_uint64),
                ("n_records", C.c_uint64), ("n_cand_faces", C.c_uint64), ("status", C.c_int32), ("bad_face", C.c_uint32)]


class HostMesh(C.Structure):
    _fields_ = [("is_float", C.c_int), ("xyz", C.c_void_p), ("nv", C.c_uint32), ("face_vtx", C.c_void_p), ("face_sizes", C.c_void_p),
                ("nf", C.c_uint32)]


class BatchItem(C.Structure):
    _fields_ = [("src", HostMesh), ("cut", HostMesh), ("com", C.c_void_p), ("shift", C.c_void_p), ("perturbation", C.c_void_p),
                ("cut_eps", C.c_double), ("gp_constant", C.c_double), ("flags", C.c_uint32)]


class HostSoup(C.Structure):
    _fields_ = [("nh", C.c_uint32), ("ne", C.c_uint32), ("face_edge", C.c_void_p), ("edge_f", C.c_void_p)]


class Validation(C.Structure):
    _fields_ = [("n_components", C.c_uint32), ("n_border_edges", C.c_uint32), ("is_closed", C.c_int)]


class Record(C.Structure):
    _fields_ = [("edge", C.c_uint32), ("face", C.c_uint32), ("point", C.c_double * 3)]


class Test(C.Structure):
    _fields_ = [("edge", C.c_uint32), ("face", C.c_uint32), ("type", C.c_char), ("pip", C.c_char), ("sign_q", C.c_int8),
                ("sign_r", C.c_int8), ("exact", C.c_uint8), ("pad", C.c_uint8 * 3), ("point", C.c_double * 3)]


# every symbol include/mcut_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mcb200_device_count": (C.c_int, []),
    "mcb200_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "mcb200_ctx_destroy": (None, [vp]),
    "mcb200_last_error": (C.c_char_p, [vp]),
    "mcb200_ctx_sync": (C.c_int, [vp]),
    "mcb200_ctx_launch_count": (C.c_uint64, [vp]),
    "mcb200_ctx_set_profiling": (C.c_int, [vp, C.c_int]),
    "mcb200_ctx_profile_read": (C.c_int, [vp, C.c_char_p, C.c_size_t]),
    "mcb200_vertex_parameters": (None, [C.c_int, vp, C.c_uint32, vp, C.c_uint32, c_dp, c_dp, c_dp, c_dp]),
    "mcb200_vertex_stats": (None, [C.c_int, vp, C.c_uint32, c_dp]),
    "mcb200_vertex_parameters_from_stats": (None, [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "mcb200_cut_bbox_eps": (C.c_double, [c_dp, C.c_double, C.c_int]),
    "mcb200_soup_ids": (C.c_int, [C.c_uint32, c_u32p, c_u32p, C.c_uint32, c_u32p, c_u32p, C.c_uint32, c_u32p, c_u32p, c_u32p,
                                  c_u32p, c_u32p]),
    "mcb200_reference_edge_rank": (C.c_int, [C.c_uint32, c_u32p, c_u32p, c_u32p, C.c_uint32, C.c_uint32, c_u32p]),
    "mcb200_reference_edge_order": (C.c_int, [C.c_uint32, c_u32p, c_u32p, C.c_uint32, c_u32p, c_u32p]),
    "mcb200_mesh_create": (C.c_int, [vp, C.c_int, vp, C.c_uint32, c_u32p, c_u32p, C.c_uint32, C.POINTER(vp)]),
    "mcb200_mesh_create_trusted": (C.c_int, [vp, C.c_int, vp, C.c_uint32, c_u32p, c_u32p, C.c_uint32, C.POINTER(vp)]),
    "mcb200_mesh_adopt_device": (C.c_int, [vp, C.c_int, vp, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "mcb200_mesh_update_xyz": (C.c_int, [vp, vp, vp, C.c_uint32]),
    "mcb200_mesh_validate": (C.c_int, [vp, vp, C.POINTER(Validation)]),
    "mcb200_mesh_winding_number": (C.c_int, [vp, vp, c_dp, C.POINTER(C.c_double)]),
    "mcb200_mesh_read_components": (C.c_int, [vp, vp, c_i32p, c_i32p, c_i32p, C.c_size_t]),
    "mcb200_mesh_set_frame": (C.c_int, [vp, vp, c_dp, c_dp, c_dp]),
    "mcb200_mesh_free": (None, [vp, vp]),
    "mcb200_bvh_build": (C.c_int, [vp, vp, C.c_double]),
    "mcb200_intersection_type_without_cut": (C.c_int, [vp, vp, vp, C.POINTER(C.c_uint32)]),
    "mcb200_mesh_set_prior_face_boxes": (C.c_int, [vp, vp, C.POINTER(C.c_double), C.c_uint32]),
    "mcb200_bvh_read": (C.c_int, [vp, vp, c_dp, c_dp]),
    "mcb200_bvh_read_morton": (C.c_int, [vp, vp, c_u32p, c_u32p]),
    "mcb200_result_create": (C.c_int, [vp, C.POINTER(vp)]),
    "mcb200_result_free": (None, [vp, vp]),
    "mcb200_bvh_intersect": (C.c_int, [vp, vp, vp, vp]),
    "mcb200_result_set_shard": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "mcb200_result_set_pair_capacity": (C.c_int, [vp, vp, C.c_uint64]),
    "mcb200_soup_create": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, c_u32p, c_u32p, c_u32p, C.POINTER(vp)]),
    "mcb200_soup_number": (C.c_int, [vp, vp, vp, vp, C.POINTER(vp)]),
    "mcb200_soup_create_sized": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, c_u32p, c_u32p, c_u32p, c_u32p,
                                          C.POINTER(vp)]),
    "mcb200_soup_free": (None, [vp, vp]),
    "mcb200_soup_from_meshes": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
    "mcb200_narrowphase": (C.c_int, [vp, vp, vp, vp, vp, C.c_uint32]),
    "mcb200_intersect_stage": (C.c_int, [vp, vp, vp, C.c_double, vp, vp, C.c_uint32]),
    "mcb200_intersect_stage_host": (C.c_int, [vp, C.POINTER(HostMesh), C.POINTER(HostMesh), c_dp, c_dp, c_dp, C.c_double,
                                             C.POINTER(HostSoup), vp, C.c_uint32]),
    "mcb200_staged_soup_read": (C.c_int, [vp, c_u32p, c_u32p, c_u32p, C.c_uint32, c_u32p, c_u32p]),
    "mcb200_result_counts": (C.c_int, [vp, vp, C.POINTER(Counts)]),
    "mcb200_comm_unique_id": (C.c_int, [C.c_char_p]),
    "mcb200_comm_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_char_p, C.POINTER(vp)]),
    "mcb200_comm_destroy": (None, [vp]),
    "mcb200_comm_rank": (C.c_int, [vp]),
    "mcb200_comm_size": (C.c_int, [vp]),
    "mcb200_intersect_stage_sharded": (C.c_int, [vp, vp, vp, vp, C.c_double, vp, vp, C.c_uint32]),
    "mcb200_batch_intersect_host": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.c_uint32, C.POINTER(BatchItem), C.c_uint32,
                                              C.POINTER(Counts)]),
    "mcb200_result_queue_counts": (C.c_int, [vp, vp, C.POINTER(C.c_uint64)]),
    "mcb200_cutpath_segments": (C.c_int, [vp, vp, vp, C.POINTER(CutpathCounts)]),
    "mcb200_cutpath_read": (C.c_int, [vp, vp, C.POINTER(C.c_uint64), c_u32p, c_u32p, C.c_size_t, C.c_size_t]),
    "mcb200_result_read_pairs": (C.c_int, [vp, vp, c_u64p, C.c_size_t]),
    "mcb200_result_read_records": (C.c_int, [vp, vp, C.POINTER(Record), C.c_size_t]),
    "mcb200_result_read_tests": (C.c_int, [vp, vp, C.POINTER(Test), C.c_size_t]),
    "mcb200_result_read_planes": (C.c_int, [vp, vp, c_u32p, c_dp, c_dp, c_i32p, C.c_size_t]),
    "mcb200_result_device_ptr": (C.c_int, [vp, vp, C.c_int, C.POINTER(vp), c_u64p]),
}

_LIB = None


def build(verbose: bool = False) -> str:
    """Compile mcut_b200/lib/libmcut_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C mcut_b200/csrc` "
                               "(__graft_entry__.build()); mcut_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError here means the .so and the header disagree
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB
