"""Helpers shared by the golden-vector tests (CPU: oracle vs golden, GPU: CUDA vs golden)."""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STAGE_CASES = ["hello", "spheres_k8", "uv12", "ico_pair", "cube_cube_axis_aligned", "cube_cube_tris_offset", "patch_vs_sphere",
               "terrain_plane", "float_spheres", "coplanar_rotated", "near_coplanar_tilt", "c5_regions_small", "degenerate_edge_edge",
               "degenerate_face_vertex", "degenerate_zero_area"]


def beq(a, b) -> bool:
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def load_stage(name):
    return np.load(os.path.join(GOLDEN, f"stage_{name}.npz"))


def _corpus_ids():
    d = os.path.join(GOLDEN, "corpus")
    return sorted(int(f[6:9]) for f in os.listdir(d) if f.startswith("bench_") and f.endswith(".npz")) if os.path.isdir(d) else []


CORPUS_CASES = _corpus_ids()  # pairs of the reference's own regression corpus (tests/source/benchmark.cpp)


def load_corpus(i):
    """-> (fixture, src, cut, flags); the input arrays travel inside the fixture (tests/golden/make_golden.py)."""
    fx = np.load(os.path.join(GOLDEN, "corpus", f"bench_{i:03d}.npz"))
    src = (fx["in_src_xyz"], fx["in_src_faces"], fx["in_src_sizes"])
    cut = (fx["in_cut_xyz"], fx["in_cut_faces"], fx["in_cut_sizes"])
    return fx, src, cut, int(fx["flags"][0])


def replay_inputs(fx, k, src, cut):
    """keyword arguments for intersect_stage that replay kernel invocation k of a fixture: the caller's arrays, or - for a
    retry on meshes the reference repartitioned (floating-polygon resolution, host side) - the meshes that invocation got,
    which are internal coordinates already (identity frame, the recorded eps), together with the boxes build_oibvh found in
    its in/out face_bboxes argument and the reference's own polygon-soup tables (a repartitioned half-edge mesh carries the
    ids of its history; the reference hands `ps` to the narrowphase, and so does the drop-in hook)."""
    pert = fx[f"d{k}_pert"] if int(fx[f"d{k}_has_pert"][0]) else None
    if f"d{k}_repartitioned" not in fx.files or not int(fx[f"d{k}_repartitioned"][0]):
        return dict(src=src, cut=cut, perturbation=pert)
    s = (fx[f"d{k}_src_xyz"], fx[f"d{k}_src_faces"], fx[f"d{k}_src_sizes"])
    c = (fx[f"d{k}_cut_xyz_unperturbed"], fx[f"d{k}_cut_faces"], fx[f"d{k}_cut_sizes"])
    return dict(src=s, cut=c, perturbation=pert, params=(np.zeros(3), np.zeros(3), float(fx[f"d{k}_eps"][0])),
                prior_boxes=(fx[f"d{k}_src_prior_bboxes"], fx[f"d{k}_cut_prior_bboxes"]),
                soup_tables=dict(edges=fx[f"d{k}_ps_edges"], face_vtx=fx[f"d{k}_ps_face_vtx"],
                                 face_sizes=fx[f"d{k}_ps_face_sizes"], face_edge=fx[f"d{k}_ps_face_edges"]))


def load_units():
    return np.load(os.path.join(GOLDEN, "unit_vectors.npz"), allow_pickle=True)


def narrowphase_violation_expected(fx, k) -> bool:
    """General-position verdict of the narrowphase alone, from the reference's own per-test results
    (kernel.cpp:2518-2557, :2588-2597)."""
    t = fx[f"d{k}_test_type"]
    p = fx[f"d{k}_test_pip"]
    one = t == ord("1")
    touch = (t == ord("p")) | (t == ord("q")) | (t == ord("r"))
    bad1 = one & ((p == ord("e")) | (p == ord("v")))
    bad2 = touch & ((p == ord("i")) | (p == ord("e")) | (p == ord("v")))
    return bool(np.any(bad1 | bad2))


def check_tests_against_fixture(fx, k, tests, complete: bool):
    """`tests` is a structured array with edge, face, type, sign_q, sign_r, pip, point sorted by (edge, face).  When the
    reference stopped early (general-position violation), its log is a subset of ours."""
    key_mine = (tests["edge"].astype(np.uint64) << np.uint64(32)) | tests["face"].astype(np.uint64)
    key_ref = (fx[f"d{k}_test_edge"].astype(np.uint64) << np.uint64(32)) | fx[f"d{k}_test_face"].astype(np.uint64)
    pos = np.searchsorted(key_mine, key_ref)
    assert np.all(pos < key_mine.size) and np.all(key_mine[pos] == key_ref), "reference tested an (edge, face) we did not"
    if complete:
        assert key_mine.size == key_ref.size, "edge/face test key set differs"
    m = tests[pos]
    assert beq(np.frombuffer(m["type"].tobytes(), dtype=np.uint8), fx[f"d{k}_test_type"]), "segment/plane types"
    assert beq(m["sign_q"], fx[f"d{k}_test_sq"]) and beq(m["sign_r"], fx[f"d{k}_test_sr"]), "orient3d signs"
    ones = fx[f"d{k}_test_type"] == ord("1")
    assert beq(np.ascontiguousarray(m["point"][ones]), np.ascontiguousarray(fx[f"d{k}_test_point"][ones])), "plane points"
    pip_ref = fx[f"d{k}_test_pip"]
    pip_mine = np.frombuffer(m["pip"].tobytes(), dtype=np.uint8)
    has = pip_ref != 0
    assert beq(pip_mine[has], pip_ref[has]), "point-in-polygon classes"
