"""-m gpu: the device validation passes (mcb200_mesh_validate) against the oracle, exact."""
import numpy as np
import pytest

import validate_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", sorted(validate_cases.all_cases()))
def test_device_validation_equals_oracle(oracle, gpu_ctx, case):
    from mcut_b200 import stage
    nv, off, vtx = validate_cases.all_cases()[case]
    sizes = np.diff(off).astype(np.uint32)
    xyz = np.zeros((nv, 3))
    xyz[:, 0] = np.arange(nv)
    m = stage.Mesh(gpu_ctx, xyz, vtx, None if np.all(sizes == 3) else sizes)
    n, fcc, cv, cf, border = m.validate()
    rn, rfcc, rcv, rcf, rborder = oracle.validate(nv, off, vtx)
    assert n == rn and border == rborder
    assert np.array_equal(fcc, rfcc), "face -> component map"
    assert np.array_equal(cv, rcv) and np.array_equal(cf, rcf), "per-component counts"
    m.free()


def test_device_validation_on_the_reference_corpus(oracle, gpu_ctx):
    """The 122 meshes of the reference's regression corpus (polygons of up to 32 vertices, open meshes): device == oracle,
    which equals the reference's own functions on them (tests/test_oracle_validate.py)."""
    from golden_util import CORPUS_CASES, load_corpus
    from mcut_b200 import stage
    for pair in CORPUS_CASES:
        _, src, cut, _ = load_corpus(pair)
        for x, f, s in (src, cut):
            off = np.concatenate([[0], np.cumsum(s)]).astype(np.uint32)
            m = stage.Mesh(gpu_ctx, x, f, s)
            n, fcc, cv, cf, border = m.validate()
            rn, rfcc, rcv, rcf, rborder = oracle.validate(x.shape[0], off, f)
            assert n == rn and border == rborder, pair
            assert np.array_equal(fcc, rfcc) and np.array_equal(cv, rcv) and np.array_equal(cf, rcf), pair
            m.free()


def test_device_validation_at_scale(oracle, gpu_ctx):
    """1M-triangle sphere: one closed component; the same with a cap cut off: border edges appear."""
    from mcut_b200 import meshgen as mg, stage
    x, f, s = mg.cube_sphere(289, 20.0)
    m = stage.Mesh(gpu_ctx, x, f, s)
    n, fcc, cv, cf, border = m.validate()
    assert n == 1 and border == 0 and int(cv[0]) == x.shape[0] and int(cf[0]) == f.size // 3 and not fcc.any()
    m.free()
    fo = f.reshape(-1, 3)[1000:].reshape(-1).astype(np.uint32)
    off = np.arange(0, fo.size + 1, 3, dtype=np.uint32)
    m = stage.Mesh(gpu_ctx, x, fo, None)
    got = m.validate()
    want = oracle.validate(x.shape[0], off, fo)
    assert got[0] == want[0] and got[4] == want[4] and got[4] > 0
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
    m.free()


def test_device_winding_number_equals_oracle(oracle, gpu_ctx):
    """mcb200_mesh_winding_number vs the oracle's sequential getWindingNumber: same classification (eps 1e-7 in the
    reference), values within 1e-11 (device atan2 and a different summation order), triangles and quads."""
    from mcut_b200 import meshgen as mg, stage
    rng = np.random.default_rng(5)
    x, f, s = mg.cube_sphere(40, 20.0, rotation=mg.rot_z(0.3), centre=(3.0, -2.0, 1.0))
    off = np.arange(0, f.size + 1, 3, dtype=np.uint32)
    m = stage.Mesh(gpu_ctx, x, f, s)
    for q in [np.array([3.0, -2.0, 1.0]), np.array([60.0, 0.0, 0.0]), *(rng.normal(size=3) * 15.0 for _ in range(20))]:
        want = oracle.winding_number(x, off, f, q)
        got = m.winding_number(q)
        assert abs(got - want) < 1e-11, (q, got, want)
        assert (abs(1.0 - got) < 1e-7) == (abs(1.0 - want) < 1e-7) and (abs(got) < 1e-7) == (abs(want) < 1e-7)
    m.free()
    (sx, sf, ss), _, _ = mg.hello_world()
    sx = sx.astype(np.float64)
    qoff = np.concatenate([[0], np.cumsum(ss)]).astype(np.uint32)
    m = stage.Mesh(gpu_ctx, sx, sf, ss)
    for q in [sx.mean(axis=0), sx.mean(axis=0) + 100.0, sx[0]]:
        assert abs(m.winding_number(q) - oracle.winding_number(sx, qoff, sf, q)) < 1e-12
    m.free()


def test_device_winding_number_follows_the_frame(oracle, gpu_ctx):
    """The mesh is stored in user coordinates; the query is given in the internal frame (x - com) + shift."""
    from mcut_b200 import meshgen as mg, stage
    x, f, s = mg.cube_sphere(16, 20.0)
    off = np.arange(0, f.size + 1, 3, dtype=np.uint32)
    com, shift = np.array([1.5, -2.0, 0.25]), np.array([30.0, 31.0, 29.5])
    m = stage.Mesh(gpu_ctx, x, f, s)
    m.set_frame(com, shift)
    xi = (x - com) + shift
    q = (np.array([2.0, 3.0, -4.0]) - com) + shift
    assert abs(m.winding_number(q) - oracle.winding_number(xi, off, f, q)) < 1e-11 and abs(m.winding_number(q) - 1.0) < 1e-9
    m.free()


def test_device_winding_number_refuses_big_polygons(gpu_ctx):
    from mcut_b200 import stage
    xyz = np.array([[0, 0, 0], [1, 0, 0], [1.5, 1, 0], [0.5, 2, 0], [-0.5, 1, 0], [0, 0, 1.0]])
    faces = np.array([0, 1, 2, 3, 4, 0, 1, 5], dtype=np.uint32)
    m = stage.Mesh(gpu_ctx, xyz, faces, np.array([5, 3], dtype=np.uint32))
    with pytest.raises(RuntimeError, match="four vertices"):
        m.winding_number([0.2, 0.2, 0.2])
    m.free()


def test_device_intersection_type_equals_reference(gpu_ctx):
    """mcb200_intersection_type_without_cut (+ STANDARD when the narrowphase finds points) on the inputs of the reference's
    tests/source/intersectionType.cpp: the verdicts the unmodified reference reported (tests/golden/intersection_type.npz)."""
    import itype_cases
    from mcut_b200 import stage
    from test_oracle_itype import REF
    for name, (src, cut, flags, _) in itype_cases.CASES.items():
        assert stage.intersection_type(gpu_ctx, src, cut, flags) == REF[name], name
