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
