"""MC_DISPATCH_INCLUDE_INTERSECTION_TYPE: the oracle's restatement of check_and_store_input_mesh_intersection_type
(preproc.cpp:1999-2122) against what the unmodified reference reported (tests/golden/intersection_type.npz) on the inputs
of its own tests/source/intersectionType.cpp, and against the values those tests assert."""
import os

import numpy as np
import pytest

import itype_cases

FX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "intersection_type.npz"))
REF = dict(zip(FX["names"].tolist(), FX["types"].tolist()))


def test_fixture_covers_the_cases_and_agrees_with_the_reference_tests():
    assert sorted(REF) == sorted(itype_cases.CASES)
    assert all(int(r) == 0 for r in FX["results"])
    for name, (_, _, _, asserted) in itype_cases.CASES.items():
        if asserted is not None:
            assert REF[name] == asserted, name


@pytest.mark.parametrize("name", sorted(itype_cases.CASES))
def test_oracle_intersection_type_equals_reference(oracle, name):
    src, cut, flags, _ = itype_cases.CASES[name]
    assert oracle.intersection_type(src, cut, flags) == REF[name]
