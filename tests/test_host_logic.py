"""The host-side pieces of the PRODUCT library (mcut_b200/lib/libmcut_b200.so: frame of the internal coordinates,
polygon-soup numbering) against vectors recorded from the reference.  No GPU is touched."""
import numpy as np
import pytest

import cases
from golden_util import STAGE_CASES, beq, load_stage, load_units


@pytest.fixture(scope="module")
def stage():
    from mcut_b200 import stage
    return stage


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_vertex_parameters_match_reference(stage, tag):
    U = load_units()
    for s, c, want in zip(U[f"vp_{tag}_src"], U[f"vp_{tag}_cut"], U[f"vp_{tag}_res"]):
        com, shift, sb, cb = stage.vertex_parameters(s, c)
        assert beq(np.concatenate([com, shift, sb, cb]), want)


@pytest.mark.parametrize("case", STAGE_CASES)
def test_frame_eps_and_soup_ids_match_reference(stage, oracle, case):
    fx = load_stage(case)
    (sx, sf, ss), (cx, cf, cs), flags = cases.ALL[case]()
    com, shift, sbb, cbb = stage.vertex_parameters(sx, cx)
    assert beq(com, fx["com"]) and beq(shift, fx["shift"])
    assert stage.cut_bbox_eps(cbb, 1e-4, False) == float(fx["eps"][0])
    soff, coff = oracle.face_offsets(sf, ss), oracle.face_offsets(cf, cs)
    fv, fe, ev, ef = stage.soup_ids(sx.shape[0], soff, np.ascontiguousarray(sf), coff, np.ascontiguousarray(cf))
    assert beq(np.concatenate([ev, ef], 1), fx["ps_edges"]), "edge numbering, h0 direction, incident faces"
    assert beq(fv, fx["ps_face_vtx"]) and beq(fe, fx["ps_face_edges"])


def test_frame_eps_and_soup_ids_match_reference_on_the_regression_corpus(stage, oracle):
    """The same on the 61 pairs of the reference's regression corpus (polygons of up to 32 vertices, open meshes)."""
    from golden_util import CORPUS_CASES, load_corpus
    for pair in CORPUS_CASES:
        fx, (sx, sf, ss), (cx, cf, cs), flags = load_corpus(pair)
        com, shift, sbb, cbb = stage.vertex_parameters(sx, cx)
        assert beq(com, fx["com"]) and beq(shift, fx["shift"]), pair
        assert stage.cut_bbox_eps(cbb, 1e-4, False) == float(fx["eps"][0]), pair
        soff, coff = oracle.face_offsets(sf, ss), oracle.face_offsets(cf, cs)
        fv, fe, ev, ef = stage.soup_ids(sx.shape[0], soff, np.ascontiguousarray(sf), coff, np.ascontiguousarray(cf))
        assert beq(np.concatenate([ev, ef], 1), fx["ps_edges"]), pair
        assert beq(fv, fx["ps_face_vtx"]) and beq(fe, fx["ps_face_edges"]), pair


def test_soup_ids_reject_bad_winding(stage):
    from mcut_b200.stage import Mcb200Error
    # two triangles sharing an edge in the SAME direction (inconsistent winding): hmesh.cpp:612-628 refuses the face
    src_off = np.array([0, 3, 6], dtype=np.uint32)
    src_vtx = np.array([0, 1, 2, 0, 1, 3], dtype=np.uint32)
    cut_off = np.array([0, 3], dtype=np.uint32)
    cut_vtx = np.array([0, 1, 2], dtype=np.uint32)
    with pytest.raises(Mcb200Error) as e:
        stage.soup_ids(4, src_off, src_vtx, cut_off, cut_vtx)
    assert e.value.code == -3


def test_absolute_eps(stage):
    bb = np.array([1.0, 2.0, 3.0, 4.0, 6.0, 15.0])
    assert stage.cut_bbox_eps(bb, 1e-4, True) == 1e-4
    assert stage.cut_bbox_eps(bb, 1e-4, False) == np.sqrt(0.0 + 9.0 + 16.0 + 144.0) * 1e-4


def test_reference_edge_order_and_rank_agree(stage, oracle):
    """mcb200_reference_edge_order (compact arrays: what the adapter calls inside a live dispatch) and mcb200_reference_edge_rank
    (arrays indexed by face id) replay the same hash map: order[rank[e]] == e for every edge of a candidate face, edges of no
    candidate face have no rank, and the helper-thread count changes the order only above 1024 candidate faces / edges
    (tpool.h:354-392).  The order itself is pinned by tests/test_hook_cpu.py against the unmodified reference."""
    import ctypes as C
    from mcut_b200 import _lib
    src, cut, flags = cases.ALL["spheres_k64"]()
    ref = oracle.intersect_stage(src, cut, flags)
    soup = ref["soup"]
    cand = np.ascontiguousarray(ref["cand_faces"], dtype=np.uint32)
    assert cand.size > 1024
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))  # noqa: E731
    orders = {}
    for helpers in (0, 3):
        rank = stage.reference_edge_rank(cand, soup.face_off, soup.face_edge, soup.ne, helpers)
        slots = np.ascontiguousarray(np.concatenate([soup.face_edge[soup.face_off[f]:soup.face_off[f + 1]] for f in cand]), dtype=np.uint32)
        off = np.ascontiguousarray(np.concatenate([[0], np.cumsum([soup.face_off[f + 1] - soup.face_off[f] for f in cand])]), dtype=np.uint32)
        order = np.zeros(slots.size, dtype=np.uint32)
        n = C.c_uint32(0)
        rc = _lib.lib().mcb200_reference_edge_order(cand.size, p(off), p(slots), helpers, p(order), C.byref(n))
        assert rc == 0
        order = order[:n.value]
        assert np.array_equal(np.sort(order), np.unique(slots)), "every edge of a candidate face exactly once"
        assert np.array_equal(rank[order], np.arange(order.size, dtype=np.uint32))
        no_rank = np.setdiff1d(np.arange(soup.ne, dtype=np.uint32), order)
        assert np.all(rank[no_rank] == 0xFFFFFFFF)
        orders[helpers] = order
    assert not np.array_equal(orders[0], orders[3]), "more than 1024 candidate faces: the block layout depends on the helper count"
