"""The oracle's restatement of find_connected_components / mesh_is_closed against the unmodified reference
(oracle/_ref/libref_unit.so: ref_validate builds the half-edge mesh the reference's way and calls both)."""
import os

import numpy as np
import pytest

import validate_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_unit.so")),
                               reason="oracle/_ref is not built (needs /root/reference at build time)")


@needs_ref
@pytest.mark.parametrize("case", sorted(validate_cases.all_cases()))
def test_oracle_validation_equals_reference(oracle, case):
    nv, off, vtx = validate_cases.all_cases()[case]
    n_ref, fcc_ref, cv_ref, cf_ref, closed_ref = oracle.ref_validate(nv, off, vtx)
    n, fcc, cv, cf, border = oracle.validate(nv, off, vtx)
    assert n_ref > 0 and n == n_ref
    assert np.array_equal(fcc, fcc_ref) and np.array_equal(cv, cv_ref) and np.array_equal(cf, cf_ref)
    assert (border == 0) == closed_ref


@needs_ref
def test_oracle_validation_equals_reference_on_the_regression_corpus(oracle):
    """All 122 meshes of the reference's regression corpus (polygons of up to 32 vertices, 60 of them open)."""
    from golden_util import CORPUS_CASES, load_corpus
    for pair in CORPUS_CASES:
        _, src, cut, _ = load_corpus(pair)
        for x, f, s in (src, cut):
            off = np.concatenate([[0], np.cumsum(s)]).astype(np.uint32)
            n_ref, fcc_ref, cv_ref, cf_ref, closed_ref = oracle.ref_validate(x.shape[0], off, f)
            n, fcc, cv, cf, border = oracle.validate(x.shape[0], off, f)
            assert n == n_ref and np.array_equal(fcc, fcc_ref) and np.array_equal(cv, cv_ref) and np.array_equal(cf, cf_ref), pair
            assert (border == 0) == closed_ref, pair


def test_known_answers(oracle):
    cases = validate_cases.all_cases()
    n, fcc, cv, cf, border = oracle.validate(*cases["two_spheres_and_a_stray_vertex"])
    assert n == 3 and border == 0 and cv.tolist()[2] == 1 and cf.tolist()[2] == 0 and cv[0] == cv[1] and cf[0] == cf[1]
    n, fcc, cv, cf, border = oracle.validate(*cases["hello_open_patch"])
    assert n == 1 and border == 4  # two triangles sharing one edge
    n, fcc, cv, cf, border = oracle.validate(*cases["sphere_with_holes"])
    assert n == 1 and border > 0


@needs_ref
def test_solid_angles_equal_reference_bitwise(oracle):
    """calculate_signed_solid_angle, triangle and quad forms (preproc.cpp:1650-1810): oracle == reference, bit for bit."""
    rng = np.random.default_rng(11)
    for i in range(4000):
        pts = [rng.normal(size=3) * rng.choice([1e-3, 1.0, 50.0]) for _ in range(4)]
        q = rng.normal(size=3)
        if i % 40 == 0:
            q = pts[i % 3].copy()  # query on a vertex: the reference returns 0
        for k in (3, 4):
            a, b = oracle.solid_angle(pts[:k], q), oracle.solid_angle(pts[:k], q, use_ref=True)
            assert np.float64(a).tobytes() == np.float64(b).tobytes(), (i, k)


def test_winding_number_known_answers(oracle):
    from mcut_b200 import meshgen as mg
    x, f, s = mg.cube_sphere(12, 20.0)
    off = np.arange(0, f.size + 1, 3, dtype=np.uint32)
    assert abs(oracle.winding_number(x, off, f, [1.0, 2.0, -3.0]) - 1.0) < 1e-12  # inside (check_and_store..., preproc.cpp:2039)
    assert abs(oracle.winding_number(x, off, f, [50.0, 2.0, -3.0])) < 1e-12  # outside
    (sx, sf, ss), _, _ = mg.hello_world()  # the cube of quads
    qoff = np.concatenate([[0], np.cumsum(ss)]).astype(np.uint32)
    centre = sx.astype(np.float64).mean(axis=0)
    assert abs(abs(oracle.winding_number(sx.astype(np.float64), qoff, sf, centre)) - 1.0) < 1e-12
