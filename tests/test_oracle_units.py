"""The oracle's single functions against vectors recorded from the unmodified reference (tests/golden/unit_vectors.npz)."""
import ctypes as C

import numpy as np
import pytest

from golden_util import beq, load_units


@pytest.fixture(scope="module")
def U():
    return load_units()


def dp(po, a):
    return np.ascontiguousarray(a).ctypes.data_as(po.c_dp)


def test_orient3d_bit_exact(oracle, U):
    p, want = U["o3d_pts"], U["o3d_out"]
    L = oracle.lib()
    got = np.array([L.mco_orient3d(dp(oracle, p[i, 0]), dp(oracle, p[i, 1]), dp(oracle, p[i, 2]), dp(oracle, p[i, 3]))
                    for i in range(p.shape[0])])
    assert beq(got, want)
    # the vector set really exercises the adaptive stages and exact zeros
    cert = C.c_int(0)
    n_uncertain = 0
    for i in range(p.shape[0]):
        L.mco_orient3d_stageA(dp(oracle, p[i, 0]), dp(oracle, p[i, 1]), dp(oracle, p[i, 2]), dp(oracle, p[i, 3]), C.byref(cert))
        n_uncertain += 0 if cert.value else 1
    assert n_uncertain > 500 and np.count_nonzero(want == 0.0) > 20


def test_orient2d_bit_exact(oracle, U):
    p, want = U["o2d_pts"], U["o2d_out"]
    L = oracle.lib()
    got = np.array([L.mco_orient2d(dp(oracle, p[i, 0]), dp(oracle, p[i, 1]), dp(oracle, p[i, 2])) for i in range(p.shape[0])])
    assert beq(got, want)


def test_polygon_plane_segment_pip(oracle, U):
    L = oracle.lib()
    n_all = U["poly_n"]
    for i in range(n_all.size):
        n = int(n_all[i])
        v = np.ascontiguousarray(U["poly_verts"][i][:n])
        normal = np.zeros(3)
        d = C.c_double(0)
        mc = L.mco_plane_coefficients(dp(oracle, v), n, dp(oracle, normal), C.byref(d))
        assert beq(normal, U["poly_normal"][i]) and d.value == U["poly_d"][i] and mc == U["poly_mc"][i], f"plane {i}"
        if not np.any(normal):
            continue
        q, r = np.ascontiguousarray(U["poly_q"][i]), np.ascontiguousarray(U["poly_r"][i])
        t = L.mco_segment_plane_type(dp(oracle, q), dp(oracle, r), dp(oracle, v), n, dp(oracle, normal), mc, None, None)
        assert ord(t) == U["poly_type"][i], f"segment type {i}"
        p = np.zeros(3)
        ret = L.mco_segment_plane_intersection(dp(oracle, p), dp(oracle, normal), d.value, dp(oracle, q), dp(oracle, r))
        assert ord(ret) == U["poly_isect_ret"][i] and beq(p, U["poly_p"][i]), f"plane point {i}"
        assert ord(L.mco_point_in_polygon(dp(oracle, p), dp(oracle, v), n, dp(oracle, normal), mc)) == U["poly_pip_p"][i]
        assert ord(L.mco_point_in_polygon(dp(oracle, q), dp(oracle, v), n, dp(oracle, normal), mc)) == U["poly_pip_q"][i]


def test_morton_and_oibvh_size(oracle, U):
    L = oracle.lib()
    xyz = U["morton_xyz"]
    got = np.array([L.mco_morton3D(float(a), float(b), float(c)) for a, b, c in xyz], dtype=np.uint32)
    assert beq(got, U["morton_codes"])
    # SURVEY.md Appendix A
    assert L.mco_morton3D(0.5, 0.25, 0.75) == 721420288 and L.mco_morton3D(1.0, 1.0, 1.0) == 1073741823
    sizes = np.array([L.mco_oibvh_size(int(t)) for t in U["oibvh_t"]], dtype=np.int32)
    assert beq(sizes, U["oibvh_size"])
    for t, s in ((1, 1), (2, 3), (3, 6), (5, 11), (12, 24), (1000, 2001), (1000000, 2000007), (4000000, 8000007)):
        assert L.mco_oibvh_size(t) == s


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_vertex_parameters(oracle, U, tag):
    for s, c, want in zip(U[f"vp_{tag}_src"], U[f"vp_{tag}_cut"], U[f"vp_{tag}_res"]):
        com, shift, sb, cb = oracle.vertex_parameters(s, c)
        assert beq(np.concatenate([com, shift, sb, cb]), want)
