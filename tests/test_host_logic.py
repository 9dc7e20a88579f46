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
