"""ctypes bindings for oracle/libmcut_oracle.so (the plain-C restatement of the reference hot path)
and, when present, oracle/_ref/libref_unit.so (doorways onto the unmodified reference).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (mcut_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None
_REF: Optional[C.CDLL] = None

c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)


class Soup(C.Structure):
    _fields_ = [("nv", C.c_uint32), ("nf", C.c_uint32), ("ne", C.c_uint32), ("nh", C.c_uint32),
                ("src_nv", C.c_uint32), ("src_nf", C.c_uint32),
                ("xyz", c_dp), ("face_off", c_u32p), ("face_vtx", c_u32p), ("face_edge", c_u32p),
                ("edge_v", c_u32p), ("edge_f", c_u32p)]


class Test(C.Structure):
    _fields_ = [("edge", C.c_uint32), ("face", C.c_uint32), ("type", C.c_char), ("pip", C.c_char),
                ("sign_q", C.c_int8), ("sign_r", C.c_int8), ("exact_q", C.c_uint8), ("exact_r", C.c_uint8),
                ("pad", C.c_uint8 * 2), ("point", C.c_double * 3)]


class Record(C.Structure):
    _fields_ = [("edge", C.c_uint32), ("face", C.c_uint32), ("point", C.c_double * 3)]


class NarrowOut(C.Structure):
    _fields_ = [("status", C.c_int), ("bad_face", C.c_uint32),
                ("n_tests", C.c_size_t), ("n_records", C.c_size_t), ("n_edge_face_before_cull", C.c_size_t),
                ("n_cand_faces", C.c_size_t),
                ("tests", C.POINTER(Test)), ("records", C.POINTER(Record)),
                ("cand_faces", c_u32p), ("cand_normal", c_dp), ("cand_d", c_dp), ("cand_maxcomp", C.POINTER(C.c_int32))]


TEST_DTYPE = np.dtype([("edge", "<u4"), ("face", "<u4"), ("type", "S1"), ("pip", "S1"), ("sign_q", "i1"),
                       ("sign_r", "i1"), ("exact_q", "u1"), ("exact_r", "u1"), ("pad", "u1", (2,)),
                       ("point", "<f8", (3,))])
class CutPath(C.Structure):
    _fields_ = [("n_groups", C.c_size_t), ("n_entries", C.c_size_t), ("n_single", C.c_size_t),
                ("key", c_u64p), ("off", c_u32p), ("vtx", c_u32p)]


RECORD_DTYPE = np.dtype([("edge", "<u4"), ("face", "<u4"), ("point", "<f8", (3,))])


def build(force: bool = False) -> str:
    """Compile oracle/libmcut_oracle.so (and oracle/_ref/* when /root/reference is present)."""
    so = os.path.join(_HERE, "libmcut_oracle.so")
    src = os.path.join(_HERE, "mcut_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.mco_vertex_parameters.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, c_dp, c_dp, c_dp, c_dp]
        L.mco_vertex_parameters.restype = None
        L.mco_transform_vertices.argtypes = [C.c_int, C.c_void_p, C.c_uint32, c_dp, c_dp, c_dp, c_dp]
        L.mco_transform_vertices.restype = None
        L.mco_cut_bbox_eps.argtypes = [c_dp, C.c_double, C.c_int]
        L.mco_cut_bbox_eps.restype = C.c_double
        L.mco_face_bboxes.argtypes = [c_dp, c_u32p, c_u32p, C.c_uint32, C.c_double, c_dp, c_dp]
        L.mco_face_bboxes.restype = None
        L.mco_face_bboxes_prior.argtypes = [c_dp, c_u32p, c_u32p, C.c_uint32, C.c_double, c_dp, C.c_uint32, c_dp, c_dp]
        L.mco_face_bboxes_prior.restype = None
        L.mco_morton3D.argtypes = [C.c_float, C.c_float, C.c_float]
        L.mco_morton3D.restype = C.c_uint32
        L.mco_morton_codes.argtypes = [c_dp, C.c_uint32, c_dp, c_u32p]
        L.mco_morton_codes.restype = None
        L.mco_oibvh_size.argtypes = [C.c_int]
        L.mco_oibvh_size.restype = C.c_int
        L.mco_oibvh_pairs.argtypes = [c_dp, C.c_uint32, c_dp, C.c_uint32, C.POINTER(c_u64p), c_u64p]
        L.mco_oibvh_pairs.restype = C.c_size_t
        L.mco_grid_pairs.argtypes = [c_dp, C.c_uint32, c_dp, C.c_uint32, C.POINTER(c_u64p)]
        L.mco_grid_pairs.restype = C.c_size_t
        L.mco_free.argtypes = [C.c_void_p]
        L.mco_free.restype = None
        L.mco_soup_build.argtypes = [c_dp, C.c_uint32, c_u32p, c_u32p, C.c_uint32, c_dp, C.c_uint32, c_u32p, c_u32p,
                                     C.c_uint32, C.POINTER(Soup)]
        L.mco_soup_build.restype = C.c_int
        L.mco_soup_free.argtypes = [C.POINTER(Soup)]
        L.mco_soup_free.restype = None
        L.mco_plane_coefficients.argtypes = [c_dp, C.c_int, c_dp, c_dp]
        L.mco_plane_coefficients.restype = C.c_int
        L.mco_orient3d.argtypes = [c_dp, c_dp, c_dp, c_dp]
        L.mco_orient3d.restype = C.c_double
        L.mco_orient3d_stageA.argtypes = [c_dp, c_dp, c_dp, c_dp, C.POINTER(C.c_int)]
        L.mco_orient3d_stageA.restype = C.c_double
        L.mco_orient2d.argtypes = [c_dp, c_dp, c_dp]
        L.mco_orient2d.restype = C.c_double
        L.mco_segment_plane_type.argtypes = [c_dp, c_dp, c_dp, C.c_int, c_dp, C.c_int, c_dp, c_dp]
        L.mco_segment_plane_type.restype = C.c_char
        L.mco_segment_plane_intersection.argtypes = [c_dp, c_dp, C.c_double, c_dp, c_dp]
        L.mco_segment_plane_intersection.restype = C.c_char
        L.mco_projection_matrix.argtypes = [c_dp, C.c_int, c_dp]
        L.mco_projection_matrix.restype = None
        L.mco_point_in_polygon.argtypes = [c_dp, c_dp, C.c_int, c_dp, C.c_int]
        L.mco_point_in_polygon.restype = C.c_char
        L.mco_narrowphase.argtypes = [C.POINTER(Soup), c_u64p, C.c_size_t, c_dp, c_dp, C.c_int, C.POINTER(NarrowOut)]
        L.mco_narrowphase.restype = C.c_int
        L.mco_narrow_free.argtypes = [C.POINTER(NarrowOut)]
        L.mco_narrow_free.restype = None
        L.mco_cutpath_segments.argtypes = [c_u32p, C.c_uint32, C.POINTER(Record), C.c_size_t, C.POINTER(CutPath)]
        L.mco_cutpath_segments.restype = C.c_int
        L.mco_cutpath_free.argtypes = [C.POINTER(CutPath)]
        L.mco_cutpath_free.restype = None
        _LIB = L
    return _LIB


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_unit.so"))


def ref() -> C.CDLL:
    """Doorways onto the unmodified reference (oracle/_ref/libref_unit.so)."""
    global _REF
    if _REF is None:
        R = C.CDLL(os.path.join(_HERE, "_ref", "libref_unit.so"))
        R.ref_morton3D.argtypes = [C.c_float, C.c_float, C.c_float]
        R.ref_morton3D.restype = C.c_uint32
        R.ref_oibvh_size.argtypes = [C.c_int]
        R.ref_oibvh_size.restype = C.c_int
        R.ref_orient3d.argtypes = [c_dp, c_dp, c_dp, c_dp]
        R.ref_orient3d.restype = C.c_double
        R.ref_orient2d.argtypes = [c_dp, c_dp, c_dp]
        R.ref_orient2d.restype = C.c_double
        R.ref_plane_coefficients.argtypes = [c_dp, C.c_int, c_dp, c_dp]
        R.ref_plane_coefficients.restype = C.c_int
        R.ref_segment_plane_type.argtypes = [c_dp, c_dp, c_dp, C.c_int, c_dp, C.c_int]
        R.ref_segment_plane_type.restype = C.c_char
        R.ref_segment_plane_intersection.argtypes = [c_dp, c_dp, C.c_double, c_dp, c_dp]
        R.ref_segment_plane_intersection.restype = C.c_char
        R.ref_point_in_polygon.argtypes = [c_dp, c_dp, C.c_int, c_dp, C.c_int]
        R.ref_point_in_polygon.restype = C.c_char
        R.ref_vertex_parameters.argtypes = [C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, c_dp, c_dp, c_dp, c_dp]
        R.ref_vertex_parameters.restype = C.c_int
        R.ref_validate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        R.ref_validate.restype = C.c_int
        _REF = R
    return _REF


def ref_validate(nv: int, face_off: np.ndarray, face_vtx: np.ndarray):
    """find_connected_components + mesh_is_closed of the unmodified reference (kernel.cpp:235-364, preproc.cpp:1957-1990)."""
    R = ref()
    nf = len(face_off) - 1
    fo = np.ascontiguousarray(face_off, dtype=np.uint32)
    fv = np.ascontiguousarray(face_vtx, dtype=np.uint32)
    fcc = np.zeros(nf, dtype=np.int32)
    cv = np.zeros(nv, dtype=np.int32)
    cf = np.zeros(nv, dtype=np.int32)
    closed = C.c_int(0)
    n = R.ref_validate(nv, fo.ctypes.data, fv.ctypes.data, nf, fcc.ctypes.data, cv.ctypes.data, cf.ctypes.data, C.byref(closed))
    return n, fcc, cv[:max(n, 0)].copy(), cf[:max(n, 0)].copy(), bool(closed.value)


def winding_number(xyz: np.ndarray, face_off: np.ndarray, face_vtx: np.ndarray, query) -> float:
    """The oracle's getWindingNumber (preproc.cpp:1650-1955, sequential sum); NaN for faces with more than 4 vertices."""
    L = lib()
    L.mco_winding_number.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.mco_winding_number.restype = C.c_double
    x = np.ascontiguousarray(xyz, dtype=np.float64)
    fo = np.ascontiguousarray(face_off, dtype=np.uint32)
    fv = np.ascontiguousarray(face_vtx, dtype=np.uint32)
    q = np.ascontiguousarray(query, dtype=np.float64)
    return float(L.mco_winding_number(x.ctypes.data, fo.ctypes.data, fv.ctypes.data, len(fo) - 1, q.ctypes.data))


def solid_angle(pts, query, use_ref: bool = False) -> float:
    """calculate_signed_solid_angle of 3 or 4 points (preproc.cpp:1650-1810): the oracle's or the reference's own."""
    L = ref() if use_ref else lib()
    names = {(3, False): "mco_solid_angle_tri", (4, False): "mco_solid_angle_quad", (3, True): "ref_solid_angle_tri",
             (4, True): "ref_solid_angle_quad"}
    fn = getattr(L, names[(len(pts), use_ref)])
    fn.argtypes = [c_dp] * (len(pts) + 1)
    fn.restype = C.c_double
    arrs = [np.ascontiguousarray(p, dtype=np.float64) for p in list(pts) + [query]]
    return float(fn(*[a.ctypes.data_as(c_dp) for a in arrs]))


def validate(nv: int, face_off: np.ndarray, face_vtx: np.ndarray):
    """The oracle's restatement of the same two passes; returns (n, fccmap, cc_vertex_count, cc_face_count, border_edges)."""
    L = lib()
    nf = len(face_off) - 1
    fo = np.ascontiguousarray(face_off, dtype=np.uint32)
    fv = np.ascontiguousarray(face_vtx, dtype=np.uint32)
    fcc = np.zeros(nf, dtype=np.int32)
    cv = np.zeros(nv, dtype=np.int32)
    cf = np.zeros(nv, dtype=np.int32)
    L.mco_connected_components.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mco_connected_components.restype = C.c_int
    L.mco_border_edges.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
    L.mco_border_edges.restype = C.c_uint32
    n = L.mco_connected_components(nv, fo.ctypes.data, fv.ctypes.data, nf, fcc.ctypes.data, cv.ctypes.data, cf.ctypes.data)
    border = L.mco_border_edges(nv, fo.ctypes.data, fv.ctypes.data, nf)
    return n, fcc, cv[:n].copy(), cf[:n].copy(), int(border)


# ------------------------------------------------------------------------------------------------
# numpy-level helpers
# ------------------------------------------------------------------------------------------------
def dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_dp)


def u32p(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_u32p)


def face_offsets(faces: np.ndarray, sizes: Optional[np.ndarray]) -> np.ndarray:
    if sizes is None:
        return np.arange(0, faces.size + 1, 3, dtype=np.uint32)
    off = np.zeros(sizes.size + 1, dtype=np.uint32)
    np.cumsum(sizes, out=off[1:])
    return off


def vertex_parameters(src_xyz: np.ndarray, cut_xyz: np.ndarray):
    is_float = src_xyz.dtype == np.float32
    src = np.ascontiguousarray(src_xyz)
    cut = np.ascontiguousarray(cut_xyz)
    com = np.zeros(3)
    shift = np.zeros(3)
    sb = np.zeros(6)
    cb = np.zeros(6)
    lib().mco_vertex_parameters(int(is_float), src.ctypes.data, src.shape[0], cut.ctypes.data, cut.shape[0], dp(com),
                                dp(shift), dp(sb), dp(cb))
    return com, shift, sb, cb


def transform_vertices(xyz: np.ndarray, com: np.ndarray, shift: np.ndarray, perturbation: Optional[np.ndarray] = None):
    is_float = xyz.dtype == np.float32
    a = np.ascontiguousarray(xyz)
    out = np.zeros((a.shape[0], 3))
    pert = None if perturbation is None else dp(np.ascontiguousarray(perturbation, dtype=np.float64))
    lib().mco_transform_vertices(int(is_float), a.ctypes.data, a.shape[0], dp(com), dp(shift), pert, dp(out))
    return out


def face_bboxes(xyz: np.ndarray, off: np.ndarray, vtx: np.ndarray, eps: float, prior=None):
    """`prior` [n,6]: the boxes build_oibvh finds in the caller's face_bboxes vector (it only expands them)."""
    nf = off.size - 1
    bb = np.zeros((nf, 6))
    root = np.zeros(6)
    if prior is None or len(prior) == 0:
        lib().mco_face_bboxes(dp(xyz), u32p(off), u32p(vtx), nf, float(eps), dp(bb), dp(root))
    else:
        pb = np.ascontiguousarray(prior, dtype=np.float64)[:nf]
        lib().mco_face_bboxes_prior(dp(xyz), u32p(off), u32p(vtx), nf, float(eps), dp(pb), pb.shape[0], dp(bb), dp(root))
    return bb, root


def morton_codes(bb: np.ndarray, root: np.ndarray) -> np.ndarray:
    codes = np.zeros(bb.shape[0], dtype=np.uint32)
    lib().mco_morton_codes(dp(bb), bb.shape[0], dp(root), u32p(codes))
    return codes


def _take_pairs(n: int, ptr) -> np.ndarray:
    out = np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0, dtype=np.uint64)
    lib().mco_free(ptr)
    return out


def oibvh_pairs(src_bb: np.ndarray, cut_bb: np.ndarray) -> Tuple[np.ndarray, int]:
    ptr = c_u64p()
    nt = C.c_uint64(0)
    n = lib().mco_oibvh_pairs(dp(src_bb), src_bb.shape[0], dp(cut_bb), cut_bb.shape[0], C.byref(ptr), C.byref(nt))
    return _take_pairs(n, ptr), int(nt.value)


def grid_pairs(src_bb: np.ndarray, cut_bb: np.ndarray) -> np.ndarray:
    ptr = c_u64p()
    n = lib().mco_grid_pairs(dp(src_bb), src_bb.shape[0], dp(cut_bb), cut_bb.shape[0], C.byref(ptr))
    return np.sort(_take_pairs(n, ptr))


class SoupHandle:
    """Owns an mco_soup_t and exposes its arrays as numpy views (copied)."""

    def __init__(self, src_xyz, src_off, src_vtx, cut_xyz, cut_off, cut_vtx):
        self.c = Soup()
        rc = lib().mco_soup_build(dp(src_xyz), src_xyz.shape[0], u32p(src_off), u32p(src_vtx), src_off.size - 1,
                                  dp(cut_xyz), cut_xyz.shape[0], u32p(cut_off), u32p(cut_vtx), cut_off.size - 1,
                                  C.byref(self.c))
        if rc != 0:
            raise ValueError("soup build failed: non-manifold edge or inconsistent winding")
        c = self.c
        as_arr = np.ctypeslib.as_array
        self.xyz = as_arr(c.xyz, shape=(c.nv, 3)).copy()
        self.face_off = as_arr(c.face_off, shape=(c.nf + 1,)).copy()
        self.face_vtx = as_arr(c.face_vtx, shape=(c.nh,)).copy()
        self.face_edge = as_arr(c.face_edge, shape=(c.nh,)).copy()
        self.edge_v = as_arr(c.edge_v, shape=(c.ne, 2)).copy()
        self.edge_f = as_arr(c.edge_f, shape=(c.ne, 2)).copy()
        self.nv, self.nf, self.ne, self.nh = c.nv, c.nf, c.ne, c.nh
        self.src_nv, self.src_nf = c.src_nv, c.src_nf

    def __del__(self):
        try:
            lib().mco_soup_free(C.byref(self.c))
        except Exception:
            pass


class SoupTables:
    """An mco_soup_t over the caller's own tables (what the reference's `ps` holds): `edges` [ne,4] = source(h0), target(h0),
    face(h0), face(h1); `face_vtx` = halfedge targets around each face; `face_edge` = edge of each halfedge."""

    def __init__(self, xyz, src_nv, src_nf, edges, face_vtx, face_sizes, face_edge):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        edges = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 4)
        self.edge_v = np.ascontiguousarray(edges[:, :2])
        self.edge_f = np.ascontiguousarray(edges[:, 2:])
        self.face_vtx = np.ascontiguousarray(face_vtx, dtype=np.uint32)
        self.face_edge = np.ascontiguousarray(face_edge, dtype=np.uint32)
        self.face_off = face_offsets(self.face_vtx, np.ascontiguousarray(face_sizes, dtype=np.uint32))
        self.nv, self.nf, self.ne, self.nh = self.xyz.shape[0], self.face_off.size - 1, edges.shape[0], self.face_vtx.size
        self.src_nv, self.src_nf = int(src_nv), int(src_nf)
        self.c = Soup(self.nv, self.nf, self.ne, self.nh, self.src_nv, self.src_nf, dp(self.xyz), u32p(self.face_off),
                      u32p(self.face_vtx), u32p(self.face_edge), u32p(self.edge_v), u32p(self.edge_f))


def narrowphase(soup, pairs: np.ndarray, src_bb: np.ndarray, cut_bb: np.ndarray, stop_on_gp: bool = True
                ) -> Dict[str, object]:
    out = NarrowOut()
    pairs = np.ascontiguousarray(pairs, dtype=np.uint64)
    lib().mco_narrowphase(C.byref(soup.c), pairs.ctypes.data_as(c_u64p), pairs.size, dp(src_bb), dp(cut_bb),
                          int(stop_on_gp), C.byref(out))
    res: Dict[str, object] = {"status": int(out.status), "bad_face": int(out.bad_face),
                              "n_edge_face_before_cull": int(out.n_edge_face_before_cull)}
    nt, nr, nc = out.n_tests, out.n_records, out.n_cand_faces
    if nt:
        buf = C.string_at(out.tests, nt * C.sizeof(Test))
        res["tests"] = np.frombuffer(buf, dtype=TEST_DTYPE).copy()
    else:
        res["tests"] = np.zeros(0, dtype=TEST_DTYPE)
    if nr:
        buf = C.string_at(out.records, nr * C.sizeof(Record))
        res["records"] = np.frombuffer(buf, dtype=RECORD_DTYPE).copy()
    else:
        res["records"] = np.zeros(0, dtype=RECORD_DTYPE)
    if nc and out.status in (0, 1):
        res["cand_faces"] = np.ctypeslib.as_array(out.cand_faces, shape=(nc,)).copy()
        res["cand_normal"] = np.ctypeslib.as_array(out.cand_normal, shape=(nc, 3)).copy()
        res["cand_d"] = np.ctypeslib.as_array(out.cand_d, shape=(nc,)).copy()
        res["cand_maxcomp"] = np.ctypeslib.as_array(out.cand_maxcomp, shape=(nc,)).copy()
    else:
        res["cand_faces"] = np.zeros(0, dtype=np.uint32)
        res["cand_normal"] = np.zeros((0, 3))
        res["cand_d"] = np.zeros(0)
        res["cand_maxcomp"] = np.zeros(0, dtype=np.int32)
    lib().mco_narrow_free(C.byref(out))
    return res


def cutpath_segments(edge_f: np.ndarray, src_nf: int, records: np.ndarray) -> Dict[str, object]:
    """The cut-path segment table of a registry (mco_cutpath_segments): records[i] is registry entry i."""
    rec = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
    ef = np.ascontiguousarray(edge_f, dtype=np.uint32)
    out = CutPath()
    rc = lib().mco_cutpath_segments(u32p(ef), int(src_nf), rec.ctypes.data_as(C.POINTER(Record)), rec.size, C.byref(out))
    if rc != 0:
        raise ValueError("cutpath_segments: a record names an edge without faces")
    g, m = out.n_groups, out.n_entries
    res = {"keys": np.ctypeslib.as_array(out.key, shape=(g,)).copy() if g else np.zeros(0, dtype=np.uint64),
           "off": np.ctypeslib.as_array(out.off, shape=(g + 1,)).copy() if g else np.zeros(1, dtype=np.uint32),
           "vtx": np.ctypeslib.as_array(out.vtx, shape=(m,)).copy() if m else np.zeros(0, dtype=np.uint32),
           "n_single": int(out.n_single)}
    lib().mco_cutpath_free(C.byref(out))
    return res


def intersect_stage(src, cut, flags: int, gp_constant: float = 1e-4, perturbation=None, params=None,
                    prior_boxes=(None, None), soup_tables=None) -> Dict[str, object]:
    """The whole intersect stage of one kernel invocation on user arrays (the oracle's mcDispatch slice):
    re-centring -> face boxes -> candidate pairs -> polygon soup -> narrowphase.
    `params` = (com, shift, eps) replaces the frame derived from the arrays (zeros = the arrays already are internal
    coordinates: what a retry on a repartitioned mesh works with)."""
    from mcut_b200 import meshgen as mg  # flag constants only
    sx, sf, ss = src
    cx, cf, cs = cut
    if params is None:
        com, shift, sbb, cbb = vertex_parameters(sx, cx)
    else:
        com, shift = (np.ascontiguousarray(a, dtype=np.float64) for a in params[:2])
    sxi = transform_vertices(sx, com, shift)
    cxi0 = transform_vertices(cx, com, shift)
    cxi = cxi0 if perturbation is None else transform_vertices(cx, com, shift, perturbation)
    soff, coff = face_offsets(sf, ss), face_offsets(cf, cs)
    if params is None:
        eps = lib().mco_cut_bbox_eps(dp(cbb), gp_constant, int(bool(flags & mg.MC_DISPATCH_ENFORCE_GENERAL_POSITION_ABSOLUTE)))
    else:
        eps = float(params[2])
    sb, sroot = face_bboxes(sxi, soff, np.ascontiguousarray(sf), 0.0, prior_boxes[0])
    cb, croot = face_bboxes(cxi0, coff, np.ascontiguousarray(cf), eps, prior_boxes[1])  # from the UNperturbed cut mesh
    pairs, ntests = oibvh_pairs(sb, cb)
    if soup_tables is None:
        soup = SoupHandle(sxi, soff, np.ascontiguousarray(sf), cxi, coff, np.ascontiguousarray(cf))
    else:
        soup = SoupTables(np.concatenate([sxi, cxi]), sxi.shape[0], soff.size - 1, **soup_tables)
    nar = narrowphase(soup, pairs, sb, cb)
    nar.update({"com": com, "shift": shift, "eps": eps, "src_bboxes": sb, "cut_bboxes": cb, "src_root": sroot,
                "cut_root": croot, "pairs": pairs, "bvh_tests": ntests, "soup": soup, "src_xyz": sxi, "cut_xyz": cxi})
    return nar


INTERSECTION_TYPE_STANDARD, INTERSECTION_TYPE_INSIDE_CUTMESH, INTERSECTION_TYPE_INSIDE_SOURCEMESH, INTERSECTION_TYPE_NONE = 0, 2, 4, 8


def intersection_type_without_cut(src_xyz, src_off, src_vtx, cut_xyz, cut_off, cut_vtx, src_root, cut_root) -> int:
    """check_and_store_input_mesh_intersection_type(), preproc.cpp:1999-2122, on internal coordinates: watertightness
    (mesh_is_closed, :1957-1990), closed-interval overlap of the mesh AABBs (math.h:931-941), winding number of a first
    vertex (getWindingNumber, :1907-1955; eps 1e-7)."""
    sm_closed = validate(src_xyz.shape[0], src_off, src_vtx)[4] == 0
    cm_closed = validate(cut_xyz.shape[0], cut_off, cut_vtx)[4] == 0
    meet = all(not (src_root[j] > cut_root[3 + j] or cut_root[j] > src_root[3 + j]) for j in range(3))
    if (not sm_closed and not cm_closed) or not meet:
        return INTERSECTION_TYPE_NONE
    meshes = {"s": (src_xyz, src_off, src_vtx), "c": (cut_xyz, cut_off, cut_vtx)}

    def inside(point, mesh):
        x, off, vtx = meshes[mesh]
        return abs(1.0 - winding_number(x, off, vtx, point)) < 1e-7

    if sm_closed and cm_closed:
        def diag2(b):
            x, y, z = b[3] - b[0], b[4] - b[1], b[5] - b[2]
            return 0.0 + x * x + y * y + z * z
        sm_larger = diag2(src_root) > diag2(cut_root)
        a, b = ("s", "c") if sm_larger else ("c", "s")
        if inside(meshes[b][0][0], a):
            return INTERSECTION_TYPE_INSIDE_SOURCEMESH if sm_larger else INTERSECTION_TYPE_INSIDE_CUTMESH
        if inside(src_xyz[0], b):  # the reference takes the SOURCE mesh's first vertex here whichever mesh is "A" (:2053)
            return INTERSECTION_TYPE_INSIDE_CUTMESH if sm_larger else INTERSECTION_TYPE_INSIDE_SOURCEMESH
        return INTERSECTION_TYPE_NONE
    if sm_closed:
        return INTERSECTION_TYPE_INSIDE_SOURCEMESH if inside(cut_xyz[0], "s") else INTERSECTION_TYPE_NONE
    return INTERSECTION_TYPE_INSIDE_CUTMESH if inside(src_xyz[0], "c") else INTERSECTION_TYPE_NONE


def intersection_type(src, cut, flags: int) -> int:
    """STANDARD when the narrowphase of this invocation finds intersection points, else the verdict above."""
    r = intersect_stage(src, cut, flags)
    if r["status"] == 0 and len(r["records"]) > 0:
        return INTERSECTION_TYPE_STANDARD
    soff, coff = face_offsets(src[1], src[2]), face_offsets(cut[1], cut[2])
    return intersection_type_without_cut(r["src_xyz"], soff, np.ascontiguousarray(src[1]), r["cut_xyz"], coff,
                                         np.ascontiguousarray(cut[1]), r["src_root"], r["cut_root"])
