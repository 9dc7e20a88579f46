"""CPU: the geometric comparison the drop-in tests rely on (tests/test_gpu_dropin.py: components_equivalent), exercised on
real outputs of the unmodified reference: with helper threads it numbers its intersection vertices differently, so the
component ARRAYS differ while the geometry must not."""
import numpy as np
import pytest

import cases
import test_gpu_dropin as td


@td.needs_ref
def test_reference_with_and_without_helper_threads_is_the_same_geometry(tmp_path):
    src, cut, flags = cases.ALL["spheres_k16"]()
    a = td.run_driver(str(tmp_path), "h0", src, cut, flags, [td.NODUMP])
    b = td.run_driver(str(tmp_path), "h3", src, cut, flags, [td.NODUMP], extra=["--helpers", "3"])
    assert int(a["mcDispatch_result"][0]) == int(b["mcDispatch_result"][0]) == 0
    assert td.components_equivalent(a, b, 0.0)
    assert td.canonical_components(a) == td.canonical_components(b)
    # and the comparison does notice a moved vertex or a dropped component
    c = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    c["cc_vertices"] = np.ascontiguousarray(c["cc_vertices"]).copy()
    c["cc_vertices"].reshape(-1)[7] += 1e-3
    assert not td.components_equivalent(a, c, 1e-9)
    assert td.components_equivalent(a, c, 1e-2)


@td.needs_ref
def test_parallel_contexts_of_the_reference_agree(tmp_path):
    """The driver's --contexts mode on the unmodified reference (what the GPU test compares the drop-in with)."""
    src, cut, flags = cases.ALL["hello"]()
    a = td.run_driver(str(tmp_path), "one", src, cut, flags, [td.NODUMP])
    b = td.run_driver(str(tmp_path), "six", src, cut, flags, [td.NODUMP], extra=["--contexts", "6"])
    assert int(b["contexts"][0]) == 6 and int(b["contexts_identical"][0]) == 1 and b["contexts_results"].tolist() == [0] * 6
    for k in ("cc_type", "cc_nv", "cc_nf", "cc_vertices", "cc_faces", "cc_face_sizes"):
        assert a[k].tobytes() == b[k].tobytes(), k
