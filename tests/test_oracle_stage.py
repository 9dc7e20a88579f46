"""The oracle's whole intersect stage against what the unmodified reference computed inside real mcDispatch calls
(tests/golden/stage_*.npz, recorded by oracle/_ref/stage_harness through tests/golden/make_golden.py)."""
import numpy as np
import pytest

import cases
from golden_util import (CORPUS_CASES, STAGE_CASES, beq, check_tests_against_fixture, load_corpus, load_stage,
                         narrowphase_violation_expected, replay_inputs)


@pytest.mark.parametrize("case", STAGE_CASES)
def test_stage_matches_reference(oracle, case):
    fx = load_stage(case)
    src, cut, flags = cases.ALL[case]()
    assert int(fx["flags"][0]) == flags
    check_oracle_against_fixture(oracle, fx, src, cut, flags)


@pytest.mark.parametrize("pair", CORPUS_CASES)
def test_reference_regression_corpus(oracle, pair):
    """src-meshNNN.off x cut-meshNNN.off of the reference's tests/meshes/benchmarks (polygons of up to 12 vertices, open
    meshes, coincident vertices, general-position retries): every stage output of every kernel invocation."""
    fx, src, cut, flags = load_corpus(pair)
    check_oracle_against_fixture(oracle, fx, src, cut, flags)


def test_corpus_is_complete():
    assert CORPUS_CASES == list(range(61)), "tests/source/benchmark.cpp runs pairs 000..060"


def check_oracle_against_fixture(oracle, fx, src, cut, flags):
    nd = int(fx["n_dispatch"][0])
    for k in range(nd):
        kw = replay_inputs(fx, k, src, cut)
        r = oracle.intersect_stage(flags=flags, **kw)
        if "params" in kw:
            assert beq(r["src_bboxes"], fx[f"d{k}_src_bboxes"]) and beq(r["cut_bboxes"], fx[f"d{k}_cut_bboxes"]), "face AABBs"
            assert beq(r["pairs"], fx[f"d{k}_pairs"]), "candidate pair set of the retry on the repartitioned mesh"
        if k == 0:
            assert beq(r["com"], fx["com"]) and beq(r["shift"], fx["shift"]) and r["eps"] == float(fx["eps"][0])
            assert beq(r["src_xyz"], fx["src_xyz_internal"]), "re-centred source coordinates"
            assert beq(r["src_bboxes"], fx["src_bboxes"]) and beq(r["cut_bboxes"], fx["cut_bboxes"]), "face AABBs"
            assert beq(r["src_root"], fx["src_root"]) and beq(r["cut_root"], fx["cut_root"]), "mesh AABBs (bvhAABBs[0])"
            assert beq(r["pairs"], fx["pairs"]), "candidate pair set"
            assert beq(oracle.grid_pairs(r["src_bboxes"], r["cut_bboxes"]), fx["pairs"]), "pair set is tree-independent"
            nfs, nfc = r["src_bboxes"].shape[0], r["cut_bboxes"].shape[0]
            assert oracle.lib().mco_oibvh_size(nfs) == int(fx["node_counts"][0])
            assert oracle.lib().mco_oibvh_size(nfc) == int(fx["node_counts"][1])
            soup = r["soup"]
            assert beq(np.concatenate([soup.edge_v, soup.edge_f], 1), fx["ps_edges"]), "polygon-soup edge numbering"
            assert beq(soup.face_vtx, fx["ps_face_vtx"]) and beq(soup.face_edge, fx["ps_face_edges"])
        assert beq(r["cut_xyz"], fx[f"d{k}_cut_xyz"]), "cut coordinates of this kernel invocation (perturbation applied)"
        raw = int(fx[f"d{k}_status_raw"][0])
        if raw in (-1, -2):  # status_t::INVALID_SRC_MESH / INVALID_CUT_MESH (kernel.cpp:2237-2244): a degenerate candidate face
            assert r["status"] == (2 if raw == -1 else 3)
            continue
        assert beq(r["cand_faces"], fx[f"d{k}_plane_faces"])
        assert beq(r["cand_normal"], fx[f"d{k}_plane_normal"]) and beq(r["cand_d"], fx[f"d{k}_plane_d"])
        assert beq(r["cand_maxcomp"], fx[f"d{k}_plane_mc"])
        violated = narrowphase_violation_expected(fx, k)
        assert (r["status"] == 1) == violated
        check_tests_against_fixture(fx, k, r["tests"], complete=not violated)
        if not violated and f"d{k}_ipoints_sorted" in fx.files:
            pts = np.ascontiguousarray(r["records"]["point"]).reshape(-1, 3)
            pts = pts[np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0]))] if len(pts) else pts
            assert beq(pts, fx[f"d{k}_ipoints_sorted"]), "intersection point multiset (m0 vertices)"


def test_hello_world_connected_components_pinned():
    """12 connected components with these vertex/face counts."""
    fx = load_stage("hello")
    got = sorted(zip(fx["cc_type"].tolist(), fx["cc_nv"].tolist(), fx["cc_nf"].tolist()))
    want = sorted([(16, 8, 6), (16, 4, 2), (8, 10, 4), (8, 14, 10), (4, 10, 2), (4, 6, 2), (1, 10, 5), (1, 10, 5), (1, 14, 7),
                   (1, 10, 7), (1, 14, 7), (1, 10, 7)])
    assert got == want


def degenerate_case(oracle):
    """cube_cube_tris_offset with one CANDIDATE source triangle collapsed onto a line."""
    (sx, sf, ss), cut, flags = cases.cube_cube_tris_offset()
    r0 = oracle.intersect_stage((sx, sf, ss), cut, flags)
    s = int(r0["pairs"][0] >> np.uint64(32))
    a, b, c = sf[3 * s:3 * s + 3]
    sx = sx.copy()
    sx[a] = sx[b] + (sx[c] - sx[b]) * 0.5  # the collapsed face keeps overlapping the cutter's boxes
    return (sx, sf, ss), cut, flags, s


def test_degenerate_face_is_invalid_mesh(oracle):
    """kernel.cpp:2237-2312: a candidate face with |Newell normal|^2 < 1e-9 makes the mesh invalid."""
    src, cut, flags, s = degenerate_case(oracle)
    r = oracle.intersect_stage(src, cut, flags)
    assert r["status"] == 2 and r["bad_face"] <= s
