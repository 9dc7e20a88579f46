#!/usr/bin/env python
"""Regenerates the committed golden vectors from the UNMODIFIED reference.

Needs oracle/_ref (built by `make -C oracle ref` from /root/reference — only possible in the build container).
Two kinds of files are written next to this script:

  unit_vectors.npz      inputs/outputs of single reference functions, called through oracle/_ref/libref_unit.so:
                        orient3d / orient2d (shewchuk.c), compute_polygon_plane_coefficients,
                        compute_segment_plane_intersection[_type], compute_point_in_polygon_test (math.cpp),
                        morton3D / get_ostensibly_implicit_bvh_size (bvh.cpp), calculate_vertex_parameters (preproc.cpp)
  stage_<case>.npz      what the reference computed inside one real mcDispatch of tests/cases.py:<case>, observed by
                        oracle/_ref/stage_harness: frame, face AABBs, mesh AABBs, candidate pairs, polygon-soup ids,
                        per-candidate-face planes, every edge/face test (type, orient3d signs, point, point-in-polygon
                        class) of every kernel invocation (with the perturbation of each retry), intersection points,
                        connected-component summary.

  corpus/bench_NNN.npz  the same record for pair NNN of the reference's own regression corpus
                        (tests/meshes/benchmarks/{src,cut}-meshNNN.off, run by tests/source/benchmark.cpp), input arrays included

Run:  python tests/golden/make_golden.py [--corpus-only | --only=case,case | --itype-only]
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from mcut_b200.mcbio import read_mcb, write_mcb  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from ref_events import SoupIndex, decode_dispatch  # noqa: E402

REFERENCE = os.environ.get("MCUT_REFERENCE", "/root/reference")
CORPUS = range(0, 61)  # benchmark.cpp runs pairs 000..060
STAGE_CASES = ["hello", "spheres_k8", "uv12", "ico_pair", "cube_cube_axis_aligned", "cube_cube_tris_offset", "patch_vs_sphere",
               "terrain_plane", "float_spheres", "coplanar_rotated", "near_coplanar_tilt", "c5_regions_small", "degenerate_edge_edge",
               "degenerate_face_vertex", "degenerate_zero_area"]


def dp(a):
    return a.ctypes.data_as(po.c_dp)


# ----------------------------------------------------------------------------------------------------------------------
def gen_orient3d(rng, n_each=600):
    pts = []
    # generic
    pts.append(rng.uniform(-50, 50, size=(n_each, 4, 3)))
    # d almost on plane(a,b,c): combination of a,b,c plus a tiny normal offset (forces stages B, C, D)
    for scale in (0.0, 1e-18, 1e-16, 1e-15, 1e-14, 1e-12):
        a = rng.uniform(1, 40, size=(n_each // 2, 3, 3))
        w = rng.uniform(-1, 2, size=(n_each // 2, 3))
        w /= w.sum(1, keepdims=True)
        d = (a * w[:, :, None]).sum(1)
        nrm = np.cross(a[:, 1] - a[:, 0], a[:, 2] - a[:, 0])
        d = d + nrm * scale * rng.uniform(-1, 1, size=(n_each // 2, 1))
        pts.append(np.concatenate([a, d[:, None, :]], 1))
    # small-integer lattice: many exact zeros
    pts.append(rng.integers(-4, 5, size=(n_each, 4, 3)).astype(np.float64))
    # lattice shifted into the positive quadrant by an irrational-ish offset (what the re-centring does)
    pts.append(rng.integers(-4, 5, size=(n_each, 4, 3)).astype(np.float64) * 1.25 + 17.123456789)
    # huge/small magnitudes
    pts.append(rng.uniform(-1, 1, size=(n_each // 2, 4, 3)) * 1e8)
    pts.append(rng.uniform(-1, 1, size=(n_each // 2, 4, 3)) * 1e-8)
    p = np.ascontiguousarray(np.concatenate(pts, 0))
    R = po.ref()
    out = np.zeros(p.shape[0])
    for i in range(p.shape[0]):
        out[i] = R.ref_orient3d(dp(p[i, 0]), dp(p[i, 1]), dp(p[i, 2]), dp(p[i, 3]))
    return p, out


def gen_orient2d(rng, n_each=500):
    pts = [rng.uniform(-50, 50, size=(n_each, 3, 2))]
    for scale in (0.0, 1e-18, 1e-16, 1e-14):
        a = rng.uniform(1, 40, size=(n_each, 2, 2))
        t = rng.uniform(-1, 2, size=(n_each, 1))
        c = a[:, 0] + t * (a[:, 1] - a[:, 0])
        nrm = np.stack([-(a[:, 1, 1] - a[:, 0, 1]), a[:, 1, 0] - a[:, 0, 0]], 1)
        c = c + nrm * scale * rng.uniform(-1, 1, size=(n_each, 1))
        pts.append(np.concatenate([a, c[:, None, :]], 1))
    pts.append(rng.integers(-4, 5, size=(n_each, 3, 2)).astype(np.float64))
    p = np.ascontiguousarray(np.concatenate(pts, 0))
    R = po.ref()
    out = np.zeros(p.shape[0])
    for i in range(p.shape[0]):
        out[i] = R.ref_orient2d(dp(p[i, 0]), dp(p[i, 1]), dp(p[i, 2]))
    return p, out


def random_polygon(rng, n, planar_noise):
    """n-gon roughly in a random plane (convex-ish), in the positive quadrant."""
    ang = np.sort(rng.uniform(0, 2 * np.pi, size=n))
    rad = rng.uniform(2.0, 6.0, size=n)
    xy = np.stack([rad * np.cos(ang), rad * np.sin(ang), planar_noise * rng.uniform(-1, 1, size=n)], 1)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    Rm = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                   [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                   [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return np.ascontiguousarray(xy @ Rm.T + rng.uniform(10, 30, size=3))


def gen_polygon_vectors(rng, count=400):
    """plane coefficients, segment/plane type, plane point and point-in-polygon on random polygons (n = 3..6)."""
    R = po.ref()
    recs = {k: [] for k in ("n", "verts", "normal", "d", "mc", "q", "r", "type", "p", "isect_ret", "pip_p", "pip_q")}
    for i in range(count):
        n = int(rng.integers(3, 7))
        if i % 25 == 0:  # degenerate: collinear / tiny
            v = np.ascontiguousarray(np.outer(np.linspace(0, 1, n), rng.uniform(1, 2, size=3)) + 5.0)
        elif i % 25 == 1:
            v = random_polygon(rng, n, 0.0) * 1e-4 + 3.0
        else:
            v = random_polygon(rng, n, 0.0 if i % 3 else 1e-3)
        normal = np.zeros(3)
        d = C.c_double(0)
        mc = R.ref_plane_coefficients(dp(v), n, dp(normal), C.byref(d))
        # a segment through / near / beside the polygon
        c = v.mean(0)
        nn = normal if np.any(normal) else np.array([0.0, 0.0, 1.0])
        mode = i % 5
        off = rng.uniform(-3, 3, size=3)
        if mode == 0:
            q, r = c + 2 * nn + 0.2 * off, c - 3 * nn + 0.2 * off
        elif mode == 1:
            q, r = c + 2 * nn + 5 * off, c - 3 * nn + 5 * off
        elif mode == 2:
            q, r = v[0].copy(), c - 3 * nn  # endpoint is a polygon vertex
        elif mode == 3:
            q, r = c + 1 * nn, c + 4 * nn + off
        else:
            q, r = (v[0] + v[1]) / 2, c + 2 * nn  # endpoint on a polygon edge (up to rounding)
        q = np.ascontiguousarray(q)
        r = np.ascontiguousarray(r)
        if np.any(normal):
            t = R.ref_segment_plane_type(dp(q), dp(r), dp(v), n, dp(normal), mc)
            p = np.zeros(3)
            ret = R.ref_segment_plane_intersection(dp(p), dp(normal), d.value, dp(q), dp(r))
            pip_p = R.ref_point_in_polygon(dp(p), dp(v), n, dp(normal), mc)
            pip_q = R.ref_point_in_polygon(dp(q), dp(v), n, dp(normal), mc)
        else:
            t, p, ret, pip_p, pip_q = b"?", np.zeros(3), b"?", b"?", b"?"
        vv = np.zeros((6, 3))
        vv[:n] = v
        for k, val in (("n", n), ("verts", vv), ("normal", normal), ("d", d.value), ("mc", mc), ("q", q), ("r", r),
                       ("type", ord(t)), ("p", p), ("isect_ret", ord(ret)), ("pip_p", ord(pip_p)), ("pip_q", ord(pip_q))):
            recs[k].append(val)
    return {f"poly_{k}": np.array(v) for k, v in recs.items()}


def gen_misc(rng):
    R = po.ref()
    xyz = np.concatenate([rng.uniform(-0.2, 1.2, size=(3000, 3)), np.array([[0.5, 0.25, 0.75], [1, 1, 1], [0, 0, 0], [np.nan, 0.5, 2.0]])]
                         ).astype(np.float32)
    codes = np.array([R.ref_morton3D(float(a), float(b), float(c)) for a, b, c in xyz], dtype=np.uint32)
    ts = np.array(list(range(1, 70)) + [100, 1000, 4095, 4096, 4097, 65535, 65536, 100000, 1000000, 1002252, 2000000, 2007372, 3998792,
                                         4000000], dtype=np.int32)
    sizes = np.array([R.ref_oibvh_size(int(t)) for t in ts], dtype=np.int32)
    out = {"morton_xyz": xyz, "morton_codes": codes, "oibvh_t": ts, "oibvh_size": sizes}
    # calculate_vertex_parameters, double and float input
    for tag, dt, flag in (("f64", np.float64, 2), ("f32", np.float32, 1)):
        vs, vc, res = [], [], []
        for k in range(6):
            ns, nc = int(rng.integers(3, 400)), int(rng.integers(3, 300))
            s = (rng.normal(size=(ns, 3)) * rng.uniform(0.5, 30) + rng.uniform(-100, 100, size=3)).astype(dt)
            c = (rng.normal(size=(nc, 3)) * rng.uniform(0.5, 30) + rng.uniform(-100, 100, size=3)).astype(dt)
            com, sh, sb, cb = np.zeros(3), np.zeros(3), np.zeros(6), np.zeros(6)
            ok = R.ref_vertex_parameters(flag, s.ctypes.data, ns, c.ctypes.data, nc, dp(com), dp(sh), dp(sb), dp(cb))
            assert ok
            vs.append(s)
            vc.append(c)
            res.append(np.concatenate([com, sh, sb, cb]))
        out[f"vp_{tag}_src"] = np.array(vs, dtype=object)
        out[f"vp_{tag}_cut"] = np.array(vc, dtype=object)
        out[f"vp_{tag}_res"] = np.array(res)
    return out


# ----------------------------------------------------------------------------------------------------------------------
STATUS_MAP = {}


def run_harness(src, cut, flags, td, tag, extra=()):
    d = {"src_xyz": src[0], "src_faces": src[1], "cut_xyz": cut[0], "cut_faces": cut[1], "flags": np.array([flags], dtype=np.uint32)}
    if src[2] is not None:
        d["src_sizes"] = src[2]
    if cut[2] is not None:
        d["cut_sizes"] = cut[2]
    ip, op = os.path.join(td, f"{tag}.in.mcb"), os.path.join(td, f"{tag}.out.mcb")
    write_mcb(ip, d)
    exe = os.path.join(ROOT, "oracle", "_ref", "stage_harness")
    subprocess.run([exe, ip, op, *extra], check=True, capture_output=True, cwd=td)
    return read_mcb(op)


def canonical_cc_hash(o):
    """Per CC: rank vertices by coordinate bit patterns, rewrite faces in ranks, rotate each face to start at its smallest
    rank, sort the faces, hash (SURVEY §8-c canonical form).  Returns uint64 per CC."""
    import hashlib
    out = []
    voff = foff = ioff = 0
    for nv, nf in zip(o["cc_nv"], o["cc_nf"]):
        v = o["cc_vertices"][voff:voff + nv]
        sizes = o["cc_face_sizes"][foff:foff + nf]
        nidx = int(sizes.sum())
        idx = o["cc_faces"][ioff:ioff + nidx]
        voff += nv
        foff += nf
        ioff += nidx
        keys = [v[i].tobytes() for i in range(nv)]
        order = sorted(range(nv), key=lambda i: keys[i])
        rank = np.zeros(nv, dtype=np.int64)
        # coincident vertices share a rank (sealed fragments may duplicate seam vertices)
        r = -1
        prev = None
        for i in order:
            if keys[i] != prev:
                r += 1
                prev = keys[i]
            rank[i] = r
        faces = []
        p = 0
        for s in sizes:
            f = [int(rank[j]) for j in idx[p:p + s]]
            p += int(s)
            m = min(range(len(f)), key=lambda t: (f[t], f[(t + 1) % len(f)]))
            faces.append(tuple(f[m:] + f[:m]))
        faces.sort()
        h = hashlib.sha256()
        h.update(b"".join(sorted(set(keys))))
        h.update(repr(faces).encode())
        out.append(int.from_bytes(h.digest()[:8], "little"))
    return np.array(out, dtype=np.uint64)


def read_off(path):
    """Minimal OFF reader (vertices as float64; faces of any size) -> (xyz, faces, sizes) as tests/cases.py returns them."""
    with open(path) as fh:
        tok = [t for line in fh for t in line.split("#")[0].split()]
    assert tok[0] == "OFF"
    nv, nf = int(tok[1]), int(tok[2])
    p = 4
    xyz = np.array([float(t) for t in tok[p:p + 3 * nv]], dtype=np.float64).reshape(nv, 3)
    p += 3 * nv
    faces, sizes = [], []
    for _ in range(nf):
        n = int(tok[p])
        faces.extend(int(t) for t in tok[p + 1:p + 1 + n])
        sizes.append(n)
        p += 1 + n
    return xyz, np.array(faces, dtype=np.uint32), np.array(sizes, dtype=np.uint32)


def stage_fixture(name, td):
    src, cut, flags = cases.ALL[name]()
    return stage_fixture_from(src, cut, flags, td, name)


def corpus_fixture(i, td):
    """Pair i of the reference's own regression corpus (tests/source/benchmark.cpp: src-meshNNN.off x cut-meshNNN.off with
    VERTEX_ARRAY_DOUBLE | ENFORCE_GENERAL_POSITION).  The input arrays are stored in the fixture: the GPU box has no
    /root/reference."""
    d = os.path.join(REFERENCE, "tests", "meshes", "benchmarks")
    src = read_off(os.path.join(d, f"src-mesh{i:03d}.off"))
    cut = read_off(os.path.join(d, f"cut-mesh{i:03d}.off"))
    flags = cases.DBL
    fx = stage_fixture_from(src, cut, flags, td, f"bench{i:03d}")
    for tag, m in (("src", src), ("cut", cut)):
        fx[f"in_{tag}_xyz"], fx[f"in_{tag}_faces"], fx[f"in_{tag}_sizes"] = m
    return fx


def stage_fixture_from(src, cut, flags, td, name):
    o = run_harness(src, cut, flags, td, name)
    fx = {"flags": np.array([flags], dtype=np.uint32), "mcDispatch_result": o["mcDispatch_result"]}
    nc2h = int(o["c2h_calls"][0])
    fx["com"] = o["c2h0_com"]
    fx["shift"] = o["c2h0_shift"]
    fx["eps"] = o["build1_eps"]
    fx["src_bboxes"], fx["cut_bboxes"] = o["build0_face_bboxes"], o["build1_face_bboxes"]
    fx["src_root"], fx["cut_root"] = o["build0_root_bbox"], o["build1_root_bbox"]
    fx["node_counts"] = np.array([o["build0_node_count"][0], o["build1_node_count"][0]], dtype=np.uint64)
    fx["src_xyz_internal"] = o["build0_xyz"]
    fx["cut_xyz_internal"] = o["build1_xyz"]
    me = o["isect0_map_entries"]
    Fs = int(o["isect0_src_face_count"][0])
    fw = me[me[:, 0] < Fs]
    fx["pairs"] = np.sort((fw[:, 0].astype(np.uint64) << np.uint64(32)) | (fw[:, 1].astype(np.uint64) - np.uint64(Fs)))
    fx["ps_edges"] = o["dispatch0_ps_edges"]
    fx["ps_face_vtx"] = o["dispatch0_ps_face_vtx"]
    fx["ps_face_sizes"] = o["dispatch0_ps_face_sizes"]
    fx["ps_face_edges"] = o["dispatch0_ps_face_edges"]
    nd = int(o["dispatch_calls"][0])
    fx["n_dispatch"] = np.array([nd], dtype=np.int32)
    # the cut-mesh conversions after the first one carry the perturbation of each retry (preproc.cpp:2650-2665); the one in
    # effect for an invocation is the latest conversion before it (a floating-polygon retry converts nothing)
    nv0 = (o["dispatch0_src_xyz"].shape[0], o["dispatch0_cut_xyz"].shape[0])
    nf0 = (o["dispatch0_src_face_sizes"].size, o["dispatch0_cut_face_sizes"].size)
    for k in range(nd):
        idx = SoupIndex(o[f"dispatch{k}_ps_xyz"], o[f"dispatch{k}_ps_face_sizes"], o[f"dispatch{k}_ps_face_vtx"], o[f"dispatch{k}_ps_edges"])
        ev = o["events"][int(o[f"dispatch{k}_event_offset"][0]):int(o[f"dispatch{k}_event_end"][0])]
        planes, tests = decode_dispatch(ev, idx)
        st = int(o[f"dispatch{k}_status"][0])
        fx[f"d{k}_status_raw"] = np.array([st], dtype=np.int32)
        last_c2h = int(o[f"dispatch{k}_c2h_calls"][0]) - 1
        has_pert = last_c2h >= 2 and int(o[f"c2h{last_c2h}_has_pert"][0]) != 0
        pert = o[f"c2h{last_c2h}_pert"] if has_pert else np.zeros(3)
        fx[f"d{k}_pert"] = pert
        fx[f"d{k}_has_pert"] = np.array([1 if has_pert else 0], dtype=np.int32)
        # a retry after the reference's floating-polygon resolution (preproc.cpp, host side, out of scope here) runs on a
        # REPARTITIONED mesh: record the meshes as that invocation got them (internal coordinates), the boxes and the
        # candidate pairs in effect, so the stage can be replayed on them with the identity frame
        nvk = (o[f"dispatch{k}_src_xyz"].shape[0], o[f"dispatch{k}_cut_xyz"].shape[0])
        nfk = (o[f"dispatch{k}_src_face_sizes"].size, o[f"dispatch{k}_cut_face_sizes"].size)
        repart = nvk != nv0 or nfk != nf0
        fx[f"d{k}_repartitioned"] = np.array([1 if repart else 0], dtype=np.int32)
        if repart:
            nb = int(o[f"dispatch{k}_build_calls"][0])
            src_b = [j for j in range(nb) if float(o[f"build{j}_eps"][0]) == 0.0 and o[f"build{j}_face_bboxes"].shape[0] == nfk[0]]
            cut_b = [j for j in range(nb) if float(o[f"build{j}_eps"][0]) > 0.0 and o[f"build{j}_face_bboxes"].shape[0] == nfk[1]]
            assert src_b and cut_b, "no build_oibvh call matches the repartitioned meshes"
            jb, jc = src_b[-1], cut_b[-1]
            assert np.array_equal(o[f"build{jb}_xyz"], o[f"dispatch{k}_src_xyz"])
            fx[f"d{k}_src_xyz"] = o[f"dispatch{k}_src_xyz"]
            fx[f"d{k}_src_faces"], fx[f"d{k}_src_sizes"] = o[f"dispatch{k}_src_face_vtx"], o[f"dispatch{k}_src_face_sizes"]
            fx[f"d{k}_cut_faces"], fx[f"d{k}_cut_sizes"] = o[f"dispatch{k}_cut_face_vtx"], o[f"dispatch{k}_cut_face_sizes"]
            fx[f"d{k}_cut_xyz_unperturbed"] = o[f"build{jc}_xyz"]
            fx[f"d{k}_src_bboxes"], fx[f"d{k}_cut_bboxes"] = o[f"build{jb}_face_bboxes"], o[f"build{jc}_face_bboxes"]
            fx[f"d{k}_eps"] = o[f"build{jc}_eps"]
            # the half-edge meshes have a history now (faces removed and added): their polygon soup is not the one the
            # numbering rules derive from flat arrays, so the reference's own tables are part of the replay input
            for name in ("ps_edges", "ps_face_vtx", "ps_face_sizes", "ps_face_edges"):
                fx[f"d{k}_{name}"] = o[f"dispatch{k}_{name}"]
            # build_oibvh's face_bboxes is in/out and preproc.cpp never clears it: the rebuild starts from the boxes of the
            # previous build of the same mesh (eps tells the two meshes apart: 0 for the source mesh)
            all_src = [j for j in range(nb) if float(o[f"build{j}_eps"][0]) == 0.0]
            all_cut = [j for j in range(nb) if float(o[f"build{j}_eps"][0]) > 0.0]
            for tag, j, side in (("src", jb, all_src), ("cut", jc, all_cut)):
                before = [i for i in side if i < j]
                fx[f"d{k}_{tag}_prior_bboxes"] = o[f"build{before[-1]}_face_bboxes"] if before else np.zeros((0, 6))
            ji = int(o[f"dispatch{k}_isect_calls"][0]) - 1
            mek = o[f"isect{ji}_map_entries"]
            Fsk = int(o[f"isect{ji}_src_face_count"][0])
            assert Fsk == nfk[0]
            fwk = mek[mek[:, 0] < Fsk]
            fx[f"d{k}_pairs"] = np.sort((fwk[:, 0].astype(np.uint64) << np.uint64(32)) | (fwk[:, 1].astype(np.uint64) - np.uint64(Fsk)))
        fx[f"d{k}_cut_xyz"] = o[f"dispatch{k}_cut_xyz"]
        faces = sorted(planes)
        fx[f"d{k}_plane_faces"] = np.array(faces, dtype=np.uint32)
        fx[f"d{k}_plane_normal"] = np.array([planes[f][0] for f in faces]).reshape(-1, 3)
        fx[f"d{k}_plane_d"] = np.array([planes[f][1] for f in faces])
        fx[f"d{k}_plane_mc"] = np.array([planes[f][2] for f in faces], dtype=np.int32)
        tests.sort(key=lambda t: (t["edge"], t["face"]))
        fx[f"d{k}_test_edge"] = np.array([t["edge"] for t in tests], dtype=np.uint32)
        fx[f"d{k}_test_face"] = np.array([t["face"] for t in tests], dtype=np.uint32)
        fx[f"d{k}_test_type"] = np.array([ord(t["type"]) for t in tests], dtype=np.uint8)
        fx[f"d{k}_test_sq"] = np.array([t["sign_q"] for t in tests], dtype=np.int8)
        fx[f"d{k}_test_sr"] = np.array([t["sign_r"] for t in tests], dtype=np.int8)
        fx[f"d{k}_test_pip"] = np.array([ord(t["pip"][-1]) if t["pip"] else 0 for t in tests], dtype=np.uint8)
        fx[f"d{k}_test_point"] = np.array([t["point"] if t["point"] is not None else np.zeros(3) for t in tests]).reshape(-1, 3)
        if f"dispatch{k}_ipoints" in o:
            ip = o[f"dispatch{k}_ipoints"]
            fx[f"d{k}_ipoints_sorted"] = ip[np.lexsort((ip[:, 2], ip[:, 1], ip[:, 0]))] if len(ip) else ip
            fx[f"d{k}_ipoints"] = ip  # in the reference's registry order (m0's intersection vertices as numbered)
    fx["cc_type"] = o["cc_type"]
    fx["cc_nv"] = o["cc_nv"]
    fx["cc_nf"] = o["cc_nf"]
    fx["cc_attrs"] = o["cc_attrs"]
    fx["cc_hash"] = canonical_cc_hash(o)
    return fx


def write_corpus(td):
    os.makedirs(os.path.join(HERE, "corpus"), exist_ok=True)
    for i in CORPUS:
        fx = corpus_fixture(i, td)
        np.savez_compressed(os.path.join(HERE, "corpus", f"bench_{i:03d}.npz"), **fx)
        print(f"corpus/bench_{i:03d}.npz: dispatches={int(fx['n_dispatch'][0])} pairs={fx['pairs'].size} "
              f"tests0={fx['d0_test_edge'].size} result={int(fx['mcDispatch_result'][0])} ccs={fx['cc_type'].size}")


def write_itype(td):
    """intersection_type.npz: what the reference reports through MC_CONTEXT_DISPATCH_INTERSECTION_TYPE for tests/itype_cases.py"""
    import itype_cases
    names, vals, rcs = [], [], []
    for name, (src, cut, flags, _) in itype_cases.CASES.items():
        o = run_harness(src, cut, flags, td, "itype", extra=["--no-events"])
        names.append(name)
        vals.append(int(o["intersection_type"][0]))
        rcs.append(int(o["mcDispatch_result"][0]))
        print(f"intersection type {name}: {vals[-1]} (mcDispatch {rcs[-1]})")
    np.savez_compressed(os.path.join(HERE, "intersection_type.npz"), names=np.array(names), types=np.array(vals, dtype=np.uint32),
                        results=np.array(rcs, dtype=np.int32))


def main():
    if "--itype-only" in sys.argv:
        with tempfile.TemporaryDirectory() as td:
            write_itype(td)
        return
    if not po.ref_available():
        raise SystemExit("oracle/_ref is missing: run `make -C oracle ref` where /root/reference exists")
    corpus_only = "--corpus-only" in sys.argv
    if corpus_only:
        with tempfile.TemporaryDirectory() as td:
            write_corpus(td)
        write_itype(td)
        return
    only = [a.split("=", 1)[1].split(",") for a in sys.argv if a.startswith("--only=")]
    if only:
        with tempfile.TemporaryDirectory() as td:
            for name in only[0]:
                np.savez_compressed(os.path.join(HERE, f"stage_{name}.npz"), **stage_fixture(name, td))
        return
    rng = np.random.default_rng(20261017)
    unit = {}
    unit["o3d_pts"], unit["o3d_out"] = gen_orient3d(rng)
    unit["o2d_pts"], unit["o2d_out"] = gen_orient2d(rng)
    unit.update(gen_polygon_vectors(rng))
    unit.update(gen_misc(rng))
    np.savez_compressed(os.path.join(HERE, "unit_vectors.npz"), **unit)
    print("unit_vectors.npz:", {k: (v.shape if hasattr(v, "shape") else None) for k, v in unit.items() if not k.startswith("poly_")})
    with tempfile.TemporaryDirectory() as td:
        for name in STAGE_CASES:
            fx = stage_fixture(name, td)
            np.savez_compressed(os.path.join(HERE, f"stage_{name}.npz"), **fx)
            print(f"stage_{name}.npz: dispatches={int(fx['n_dispatch'][0])} pairs={fx['pairs'].size} "
                  f"tests0={fx['d0_test_edge'].size} result={int(fx['mcDispatch_result'][0])} ccs={fx['cc_type'].size}")
        write_corpus(td)
        write_itype(td)


if __name__ == "__main__":
    main()
