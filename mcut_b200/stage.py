"""Host-side mirror of the reference's intersect stage over the C-ABI (include/mcut_b200.h).

The names follow the reference: a `Context` is what mcCreateContext gives (one device, one stream), and
`intersect_stage()` walks the same steps preproc() walks between "calculate_vertex_parameters" and the end of the
kernel's "Calculate intersection points" (source/preproc.cpp:2292-2926, source/kernel.cpp:1779-3231):

    frame (com/shift)            host, sequential            mcb200_vertex_parameters
    build_oibvh(src), (cut)      device                      mcb200_bvh_build
    intersectOIBVHs              device                      mcb200_bvh_intersect
    polygon-soup ids             host, integer               mcb200_soup_from_meshes
    edge/face narrowphase        device                      mcb200_narrowphase

Everything that computes runs in the CUDA library; this module only moves arrays and raises on errors.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import Counts, Record, Test, c_dp, c_i32p, c_u32p, c_u64p

ERR_CAPACITY = -4  # include/mcut_b200.h: MCB200_ERR_CAPACITY
STAGE_SRC_RESIDENT = 2  # MCB200_STAGE_SRC_RESIDENT
STAGE_CUT_RESIDENT = 4  # MCB200_STAGE_CUT_RESIDENT

MC_DISPATCH_VERTEX_ARRAY_FLOAT = 1 << 0
MC_DISPATCH_VERTEX_ARRAY_DOUBLE = 1 << 1
MC_DISPATCH_ENFORCE_GENERAL_POSITION = 1 << 15
MC_DISPATCH_ENFORCE_GENERAL_POSITION_ABSOLUTE = 1 << 16

NARROW_LOG_TESTS = 1
NARROW_COUNT_TESTS = 8  # n_tests as the reference counts them (the side prefilter's dismissals are put through the culls too)

STATUS_SUCCESS = 0
STATUS_GENERAL_POSITION_VIOLATION = 1
STATUS_INVALID_SRC_MESH = 2
STATUS_INVALID_CUT_MESH = 3

RECORD_DTYPE = np.dtype([("edge", "<u4"), ("face", "<u4"), ("point", "<f8", (3,))])
TEST_DTYPE = np.dtype([("edge", "<u4"), ("face", "<u4"), ("type", "S1"), ("pip", "S1"), ("sign_q", "i1"), ("sign_r", "i1"),
                       ("exact", "u1"), ("pad", "u1", (3,)), ("point", "<f8", (3,))])


class Mcb200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mcut_b200 error {code}: {msg}")
        self.code = code


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_dp)


def _u32p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_u32p)


def vertex_parameters(src_xyz: np.ndarray, cut_xyz: np.ndarray):
    """com, shift, src_bbox, cut_bbox of source/preproc.cpp:2124-2290 (host; float or double arrays)."""
    L = _lib.lib()
    is_float = src_xyz.dtype == np.float32
    if cut_xyz.dtype != src_xyz.dtype:
        raise ValueError("both meshes must use the same vertex type (one MC_DISPATCH_VERTEX_ARRAY_* flag)")
    s = np.ascontiguousarray(src_xyz)
    c = np.ascontiguousarray(cut_xyz)
    com, shift, sb, cb = np.zeros(3), np.zeros(3), np.zeros(6), np.zeros(6)
    L.mcb200_vertex_parameters(int(is_float), s.ctypes.data, s.shape[0], c.ctypes.data, c.shape[0], _dp(com), _dp(shift), _dp(sb),
                               _dp(cb))
    return com, shift, sb, cb


def cut_bbox_eps(cut_bbox: np.ndarray, gp_constant: float = 1e-4, absolute: bool = False) -> float:
    return float(_lib.lib().mcb200_cut_bbox_eps(_dp(np.ascontiguousarray(cut_bbox)), gp_constant, int(absolute)))


def soup_ids(nsv: int, src_off: np.ndarray, src_vtx: np.ndarray, cut_off: np.ndarray, cut_vtx: np.ndarray):
    """Polygon-soup numbering (host).  Returns face_vtx, face_edge, edge_v[ne,2], edge_f[ne,2]."""
    L = _lib.lib()
    nh = int(src_off[-1]) + int(cut_off[-1])
    fv = np.zeros(nh, dtype=np.uint32)
    fe = np.zeros(nh, dtype=np.uint32)
    ev = np.zeros(2 * nh, dtype=np.uint32)
    ef = np.zeros(2 * nh, dtype=np.uint32)
    ne = C.c_uint32(0)
    rc = L.mcb200_soup_ids(nsv, _u32p(src_off), _u32p(src_vtx), src_off.size - 1, _u32p(cut_off), _u32p(cut_vtx), cut_off.size - 1,
                           _u32p(fv), _u32p(fe), _u32p(ev), _u32p(ef), C.byref(ne))
    if rc:
        raise Mcb200Error(rc, "soup_ids: non-manifold edge, inconsistent winding or degenerate face")
    n = ne.value
    return fv, fe, ev[:2 * n].reshape(n, 2).copy(), ef[:2 * n].reshape(n, 2).copy()


def reference_edge_rank(cand_faces: np.ndarray, face_off, face_edge: np.ndarray, ne: int, helper_threads: int = 0) -> np.ndarray:
    """rank[e] = position of polygon-soup edge e in the order in which the reference registers intersection points
    (mcb200_reference_edge_rank); the registry order is records sorted by (rank[edge], face)."""
    cand = np.ascontiguousarray(cand_faces, dtype=np.uint32)
    fe = np.ascontiguousarray(face_edge, dtype=np.uint32)
    off = None if face_off is None else np.ascontiguousarray(face_off, dtype=np.uint32)
    rank = np.zeros(max(ne, 1), dtype=np.uint32)
    rc = _lib.lib().mcb200_reference_edge_rank(cand.size, _u32p(cand), _u32p(off) if off is not None else None, _u32p(fe), ne,
                                               helper_threads, _u32p(rank))
    if rc:
        raise Mcb200Error(rc, "reference_edge_rank: invalid arguments")
    return rank[:ne]


class Context:
    """One device + one stream (what an MCUT context maps to)."""

    def __init__(self, device: int = 0, stream: int = 0):
        self.L = _lib.lib()
        self.h = C.c_void_p()
        rc = self.L.mcb200_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self.h))
        if rc:
            raise Mcb200Error(rc, self.L.mcb200_last_error(None).decode())
        self.device = device

    def check(self, rc: int):
        if rc:
            raise Mcb200Error(rc, self.L.mcb200_last_error(self.h).decode())

    def sync(self):
        self.check(self.L.mcb200_ctx_sync(self.h))

    def set_profiling(self, on: bool):
        self.check(self.L.mcb200_ctx_set_profiling(self.h, int(on)))

    def profile_read(self) -> Dict[str, Tuple[int, float]]:
        """{kernel name: (launches, total device ms)} since the last read; synchronises."""
        buf = C.create_string_buffer(1 << 16)
        n = self.L.mcb200_ctx_profile_read(self.h, buf, len(buf))
        if n < 0:
            self.check(n)
        out: Dict[str, Tuple[int, float]] = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(" ", 2)
            out[name] = (int(cnt), float(ms))
        return out

    @property
    def launches(self) -> int:
        return int(self.L.mcb200_ctx_launch_count(self.h))

    def close(self):
        if self.h:
            self.L.mcb200_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Mesh:
    def __init__(self, ctx: Context, xyz: np.ndarray, faces: np.ndarray, sizes: Optional[np.ndarray] = None):
        self.ctx = ctx
        self.h = C.c_void_p()
        if xyz.dtype not in (np.float32, np.float64):
            raise TypeError("vertices must be float32 or float64")
        self.is_float = xyz.dtype == np.float32
        xyz = np.ascontiguousarray(xyz)
        faces = np.ascontiguousarray(faces, dtype=np.uint32)
        sizes = None if sizes is None else np.ascontiguousarray(sizes, dtype=np.uint32)
        self.nv = int(xyz.shape[0])
        self.nf = int(sizes.size) if sizes is not None else int(faces.size // 3)
        ctx.check(ctx.L.mcb200_mesh_create(ctx.h, int(self.is_float), xyz.ctypes.data, self.nv, _u32p(faces), _u32p(sizes), self.nf,
                                           C.byref(self.h)))

    def set_frame(self, com=None, shift=None, perturbation=None):
        f = lambda a: None if a is None else _dp(np.ascontiguousarray(a, dtype=np.float64))
        self.ctx.check(self.ctx.L.mcb200_mesh_set_frame(self.ctx.h, self.h, f(com), f(shift), f(perturbation)))

    def build(self, eps: float = 0.0, prior_boxes=None):
        """`prior_boxes` [n,6]: what the caller's face_bboxes vector held on entry to build_oibvh (in/out there)."""
        if prior_boxes is not None and len(prior_boxes):
            pb = np.ascontiguousarray(prior_boxes, dtype=np.float64)
            self.ctx.check(self.ctx.L.mcb200_mesh_set_prior_face_boxes(self.ctx.h, self.h, _dp(pb), pb.shape[0]))
        self.ctx.check(self.ctx.L.mcb200_bvh_build(self.ctx.h, self.h, float(eps)))

    def read_bvh(self, want_boxes: bool = True):
        bb = np.zeros((self.nf, 6)) if want_boxes else None
        root = np.zeros(6)
        self.ctx.check(self.ctx.L.mcb200_bvh_read(self.ctx.h, self.h, _dp(bb), _dp(root)))
        return bb, root

    def validate(self):
        """find_connected_components + mesh_is_closed on the device.  Returns
        (n_components, fccmap[nf], cc_vertex_count[n], cc_face_count[n], n_border_edges)."""
        v = _lib.Validation()
        self.ctx.check(self.ctx.L.mcb200_mesh_validate(self.ctx.h, self.h, C.byref(v)))
        n = int(v.n_components)
        fcc = np.zeros(self.nf, dtype=np.int32)
        cv = np.zeros(max(n, 1), dtype=np.int32)
        cf = np.zeros(max(n, 1), dtype=np.int32)
        self.ctx.check(self.ctx.L.mcb200_mesh_read_components(self.ctx.h, self.h, fcc.ctypes.data_as(c_i32p), cv.ctypes.data_as(c_i32p),
                                                              cf.ctypes.data_as(c_i32p), max(n, 1)))
        return n, fcc, cv[:n], cf[:n], int(v.n_border_edges)

    def winding_number(self, query) -> float:
        """getWindingNumber of the reference (preproc.cpp:1907-1955) for a point in this mesh's internal coordinates."""
        q = np.ascontiguousarray(query, dtype=np.float64)
        out = C.c_double(0.0)
        self.ctx.check(self.ctx.L.mcb200_mesh_winding_number(self.ctx.h, self.h, q.ctypes.data_as(c_dp), C.byref(out)))
        return float(out.value)

    def read_morton(self):
        codes = np.zeros(self.nf, dtype=np.uint32)
        order = np.zeros(self.nf, dtype=np.uint32)
        self.ctx.check(self.ctx.L.mcb200_bvh_read_morton(self.ctx.h, self.h, _u32p(codes), _u32p(order)))
        return codes, order

    def free(self):
        if self.h:
            self.ctx.L.mcb200_mesh_free(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Soup:
    def __init__(self, ctx: Context, src: Mesh, cut: Mesh, tables=None):
        """`tables` = dict(edges [ne,4] = source(h0), target(h0), face(h0), face(h1); face_vtx; face_sizes; face_edge): the
        caller's own polygon soup (what the reference's `ps` holds) instead of the numbering rules applied to the arrays."""
        self.ctx = ctx
        self.h = C.c_void_p()
        if tables is None:
            ctx.check(ctx.L.mcb200_soup_from_meshes(ctx.h, src.h, cut.h, C.byref(self.h)))
            return
        u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)  # noqa: E731
        edges = u32(tables["edges"]).reshape(-1, 4)
        edge_f, fv, fe, fs = u32(edges[:, 2:]), u32(tables["face_vtx"]), u32(tables["face_edge"]), u32(tables["face_sizes"])
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))  # noqa: E731
        ctx.check(ctx.L.mcb200_soup_create_sized(ctx.h, src.nf, cut.nf, fv.size, edges.shape[0], p(fv), p(fe), p(edge_f), p(fs),
                                                 C.byref(self.h)))

    def free(self):
        if self.h:
            self.ctx.L.mcb200_soup_free(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Result:
    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.h = C.c_void_p()
        ctx.check(ctx.L.mcb200_result_create(ctx.h, C.byref(self.h)))

    def set_pair_capacity(self, max_pairs: int):
        self.ctx.check(self.ctx.L.mcb200_result_set_pair_capacity(self.ctx.h, self.h, int(max_pairs)))

    def set_shard(self, part: int, nparts: int, chunk: int = 4096):
        self.ctx.check(self.ctx.L.mcb200_result_set_shard(self.ctx.h, self.h, part, nparts, chunk))

    def counts(self) -> Counts:
        c = Counts()
        self.ctx.check(self.ctx.L.mcb200_result_counts(self.ctx.h, self.h, C.byref(c)))
        return c

    def pairs(self) -> np.ndarray:
        n = int(self.counts().n_pairs)
        out = np.zeros(n, dtype=np.uint64)
        self.ctx.check(self.ctx.L.mcb200_result_read_pairs(self.ctx.h, self.h, out.ctypes.data_as(c_u64p), n))
        return out

    def records(self) -> np.ndarray:
        n = int(self.counts().n_records)
        out = np.zeros(n, dtype=RECORD_DTYPE)
        self.ctx.check(self.ctx.L.mcb200_result_read_records(self.ctx.h, self.h, out.ctypes.data_as(C.POINTER(Record)), n))
        return out

    def cutpath(self, soup) -> Dict[str, object]:
        """Cut-path segment table of the registry (mcb200_cutpath_segments + mcb200_cutpath_read)."""
        cc = _lib.CutpathCounts()
        self.ctx.check(self.ctx.L.mcb200_cutpath_segments(self.ctx.h, soup.h, self.h, C.byref(cc)))
        g, m = int(cc.n_groups), int(cc.n_entries)
        keys, off, vtx = np.zeros(g, dtype=np.uint64), np.zeros(g + 1, dtype=np.uint32), np.zeros(m, dtype=np.uint32)
        self.ctx.check(self.ctx.L.mcb200_cutpath_read(self.ctx.h, self.h, keys.ctypes.data_as(C.POINTER(C.c_uint64)), _u32p(off),
                                                      _u32p(vtx) if m else None, g, m))
        return {"keys": keys, "off": off, "vtx": vtx, "n_single": int(cc.n_single_point_groups)}

    def tests(self) -> np.ndarray:
        # the log count is not part of mcb200_counts; ask with a generous capacity = n_tests
        n = int(self.counts().n_tests)
        out = np.zeros(n, dtype=TEST_DTYPE)
        self.ctx.check(self.ctx.L.mcb200_result_read_tests(self.ctx.h, self.h, out.ctypes.data_as(C.POINTER(Test)), n))
        return out

    def planes(self):
        n = int(self.counts().n_cand_faces)
        faces = np.zeros(n, dtype=np.uint32)
        normal = np.zeros((n, 3))
        d = np.zeros(n)
        mc = np.zeros(n, dtype=np.int32)
        self.ctx.check(self.ctx.L.mcb200_result_read_planes(self.ctx.h, self.h, _u32p(faces), _dp(normal), _dp(d),
                                                            mc.ctypes.data_as(c_i32p), n))
        return faces, normal, d, mc

    def device_ptr(self, which: int) -> Tuple[int, int]:
        p = C.c_void_p()
        n = C.c_uint64(0)
        self.ctx.check(self.ctx.L.mcb200_result_device_ptr(self.ctx.h, self.h, which, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def free(self):
        if self.h:
            self.ctx.L.mcb200_result_free(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def intersect_stage(ctx: Context, src, cut, flags: int, gp_constant: float = 1e-4, perturbation=None, log_tests: bool = False,
                    want_boxes: bool = True, params=None, prior_boxes=(None, None), soup_tables=None,
                    count_tests: bool = False, want_cutpath: bool = False) -> Dict[str, object]:
    """One kernel invocation's intersect stage on user arrays, through the C-ABI with host buffers.
    `src`/`cut` = (xyz[V,3] float32|float64, faces_flat uint32, sizes uint32|None).
    `params` = (com, shift, eps) replaces the frame derived from the arrays (zeros: the arrays are internal coordinates);
    `prior_boxes` = (src, cut) face boxes the two builds start from (see Mesh.build); `soup_tables`: see Soup.
    `count_tests`: after the run proper, the narrowphase is repeated with MCB200_NARROW_COUNT_TESTS on the same pairs and
    its n_tests (= the number of edge/face tests the reference runs) is returned as "n_tests_reference"; the records of
    that second run must equal the first run's."""
    sx, sf, ss = src
    cx, cf, cs = cut
    if params is None:
        com, shift, sbb, cbb = vertex_parameters(sx, cx)
        eps = cut_bbox_eps(cbb, gp_constant, bool(flags & MC_DISPATCH_ENFORCE_GENERAL_POSITION_ABSOLUTE))
    else:
        com, shift = (np.ascontiguousarray(a, dtype=np.float64) for a in params[:2])
        eps = float(params[2])
    ms = Mesh(ctx, sx, sf, ss)
    mc = Mesh(ctx, cx, cf, cs)
    ms.set_frame(com, shift)
    mc.set_frame(com, shift)  # boxes/BVH of the cut mesh always come from the unperturbed frame
    ms.build(0.0, prior_boxes[0])
    mc.build(eps, prior_boxes[1])
    res = Result(ctx)
    ctx.check(ctx.L.mcb200_bvh_intersect(ctx.h, ms.h, mc.h, res.h))
    soup = Soup(ctx, ms, mc, soup_tables)
    if perturbation is not None:
        mc.set_frame(com, shift, perturbation)
    ctx.check(ctx.L.mcb200_narrowphase(ctx.h, soup.h, ms.h, mc.h, res.h, NARROW_LOG_TESTS if log_tests else 0))
    c = res.counts()
    out: Dict[str, object] = {
        "com": com, "shift": shift, "eps": eps, "status": int(c.status), "bad_face": int(c.bad_face),
        "n_pairs": int(c.n_pairs), "n_node_tests": int(c.n_node_tests), "n_tests": int(c.n_tests), "n_exact": int(c.n_exact),
        "n_records": int(c.n_records), "n_cand_faces": int(c.n_cand_faces),
        "pairs": res.pairs(), "records": res.records(),
    }
    if want_boxes:
        out["src_bboxes"], out["src_root"] = ms.read_bvh()
        out["cut_bboxes"], out["cut_root"] = mc.read_bvh()
    faces, normal, d, mcmp = res.planes()
    out.update({"cand_faces": faces, "cand_normal": normal, "cand_d": d, "cand_maxcomp": mcmp})
    if want_cutpath and int(c.status) == 0:
        out["cutpath"] = res.cutpath(soup)
    if log_tests:
        out["tests"] = res.tests()
        out["n_tests_reference"] = out["n_tests"]
    elif count_tests:
        ctx.check(ctx.L.mcb200_narrowphase(ctx.h, soup.h, ms.h, mc.h, res.h, NARROW_COUNT_TESTS))
        c2 = res.counts()
        out["n_tests_reference"] = int(c2.n_tests)
        if int(c2.n_records) != out["n_records"] or int(c2.n_exact) != out["n_exact"] or int(c2.status) != out["status"] \
                or res.records().tobytes() != out["records"].tobytes():
            raise Mcb200Error(-1, "the counting run of the narrowphase differs from the plain run")
    for o in (soup, res, ms, mc):
        o.free()
    return out


INTERSECTION_TYPE_STANDARD, INTERSECTION_TYPE_INSIDE_CUTMESH, INTERSECTION_TYPE_INSIDE_SOURCEMESH, INTERSECTION_TYPE_NONE = 0, 2, 4, 8


def intersection_type(ctx: Context, src, cut, flags: int, gp_constant: float = 1e-4) -> int:
    """What MC_DISPATCH_INCLUDE_INTERSECTION_TYPE reports for one kernel invocation: STANDARD when the narrowphase finds
    intersection points (the reference decides that later, from the number of connected components: preproc.cpp:3693-3725),
    otherwise the verdict of check_and_store_input_mesh_intersection_type (mcb200_intersection_type_without_cut)."""
    sx, sf, ss = src
    cx, cf, cs = cut
    com, shift, sbb, cbb = vertex_parameters(sx, cx)
    eps = cut_bbox_eps(cbb, gp_constant, bool(flags & MC_DISPATCH_ENFORCE_GENERAL_POSITION_ABSOLUTE))
    ms, mc = Mesh(ctx, sx, sf, ss), Mesh(ctx, cx, cf, cs)
    ms.set_frame(com, shift)
    mc.set_frame(com, shift)
    ms.build(0.0)
    mc.build(eps)
    res = Result(ctx)
    ctx.check(ctx.L.mcb200_bvh_intersect(ctx.h, ms.h, mc.h, res.h))
    soup = Soup(ctx, ms, mc)
    ctx.check(ctx.L.mcb200_narrowphase(ctx.h, soup.h, ms.h, mc.h, res.h, 0))
    c = res.counts()
    if int(c.status) == STATUS_SUCCESS and int(c.n_records) > 0:
        out = INTERSECTION_TYPE_STANDARD
    else:
        t = C.c_uint32(0)
        ctx.check(ctx.L.mcb200_intersection_type_without_cut(ctx.h, ms.h, mc.h, C.byref(t)))
        out = int(t.value)
    for o in (soup, res, ms, mc):
        o.free()
    return out


def intersect_stage_host(ctx: Context, src, cut, flags: int = 0, gp_constant: float = 1e-4, perturbation=None, soup_ids_host=None,
                         log_tests: bool = False, res: "Result" = None, params=None, src_resident: bool = False,
                         cut_resident: bool = False) -> Dict[str, object]:
    """The same stage through the single pipelined C-ABI call `mcb200_intersect_stage_host`: uploads on a copy stream
    overlap the builds, the polygon-soup vertex lists are derived on the device. `soup_ids_host` = (face_edge, edge_f, ne)
    when the caller already holds `ps` (the reference does, kernel.cpp:1593-1732); None -> computed by the library.
    `params` = (com, shift, eps) to skip the host-side vertex_parameters pass (it belongs to preproc, not to this stage)."""
    from ._lib import HostMesh, HostSoup
    sx, sf, ss = src
    cx, cf, cs = cut
    if params is None:
        com, shift, sbb, cbb = vertex_parameters(sx, cx)
        eps = cut_bbox_eps(cbb, gp_constant, bool(flags & MC_DISPATCH_ENFORCE_GENERAL_POSITION_ABSOLUTE))
    else:
        com, shift, eps = params
    keep = []

    def host_mesh(x, f, sz):
        x = np.ascontiguousarray(x)
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
        f = np.ascontiguousarray(f, dtype=np.uint32)
        keep.extend([x, f])
        nf = len(f) // 3 if sz is None else len(sz)
        szp = None
        if sz is not None:
            sz = np.ascontiguousarray(sz, dtype=np.uint32)
            keep.append(sz)
            szp = sz.ctypes.data
        return HostMesh(1 if x.dtype == np.float32 else 0, x.ctypes.data, x.shape[0], f.ctypes.data, szp, nf)

    hs, hc = host_mesh(sx, sf, ss), host_mesh(cx, cf, cs)
    hsoup = None
    if soup_ids_host is not None:
        fe, ef, ne = soup_ids_host
        fe = np.ascontiguousarray(fe, dtype=np.uint32)
        ef = np.ascontiguousarray(ef, dtype=np.uint32)
        keep.extend([fe, ef])
        hsoup = C.byref(HostSoup(len(fe), int(ne), fe.ctypes.data, ef.ctypes.data))
    own = res is None
    if own:
        res = Result(ctx)
    com_a = np.ascontiguousarray(com, dtype=np.float64)
    shift_a = np.ascontiguousarray(shift, dtype=np.float64)
    pert_p = None
    if perturbation is not None:
        pert_a = np.ascontiguousarray(perturbation, dtype=np.float64)
        keep.append(pert_a)
        pert_p = pert_a.ctypes.data_as(C.POINTER(C.c_double))
    fl = (NARROW_LOG_TESTS if log_tests else 0) | (STAGE_SRC_RESIDENT if src_resident else 0) | (STAGE_CUT_RESIDENT if cut_resident else 0)
    for attempt in range(3):
        ctx.check(ctx.L.mcb200_intersect_stage_host(ctx.h, C.byref(hs), C.byref(hc), com_a.ctypes.data_as(C.POINTER(C.c_double)),
                                                    shift_a.ctypes.data_as(C.POINTER(C.c_double)), pert_p, float(eps), hsoup, res.h,
                                                    fl))
        c = Counts()
        rc = ctx.L.mcb200_result_counts(ctx.h, res.h, C.byref(c))
        if rc == ERR_CAPACITY and attempt < 2:
            continue  # the library raised the pair capacity: run again
        ctx.check(rc)
        break
    out: Dict[str, object] = {
        "com": com, "shift": shift, "eps": eps, "status": int(c.status), "bad_face": int(c.bad_face),
        "n_pairs": int(c.n_pairs), "n_tests": int(c.n_tests), "n_exact": int(c.n_exact), "n_records": int(c.n_records),
        "n_cand_faces": int(c.n_cand_faces), "pairs": res.pairs(), "records": res.records(),
    }
    faces, normal, d, mcmp = res.planes()
    out.update({"cand_faces": faces, "cand_normal": normal, "cand_d": d, "cand_maxcomp": mcmp})
    if log_tests:
        out["tests"] = res.tests()
    if own:
        res.free()
    return out


def staged_soup(ctx: Context):
    """(face_vtx, face_edge, edge_f[ne,2]) the last intersect_stage_host call of `ctx` worked with."""
    nh = C.c_uint32()
    ne = C.c_uint32()
    ctx.check(ctx.L.mcb200_staged_soup_read(ctx.h, None, None, None, 0, C.byref(nh), C.byref(ne)))
    fv = np.empty(nh.value, dtype=np.uint32)
    fe = np.empty(nh.value, dtype=np.uint32)
    ef = np.empty((ne.value, 2), dtype=np.uint32)
    ctx.check(ctx.L.mcb200_staged_soup_read(ctx.h, fv.ctypes.data_as(c_u32p), fe.ctypes.data_as(c_u32p), ef.ctypes.data_as(c_u32p),
                                            ne.value, C.byref(nh), C.byref(ne)))
    return fv, fe, ef
