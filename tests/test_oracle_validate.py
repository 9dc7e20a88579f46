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


def test_known_answers(oracle):
    cases = validate_cases.all_cases()
    n, fcc, cv, cf, border = oracle.validate(*cases["two_spheres_and_a_stray_vertex"])
    assert n == 3 and border == 0 and cv.tolist()[2] == 1 and cf.tolist()[2] == 0 and cv[0] == cv[1] and cf[0] == cf[1]
    n, fcc, cv, cf, border = oracle.validate(*cases["hello_open_patch"])
    assert n == 1 and border == 4  # two triangles sharing one edge
    n, fcc, cv, cf, border = oracle.validate(*cases["sphere_with_holes"])
    assert n == 1 and border > 0
