"""CPU: the side prefilter of the triangle narrowphase (mcut_b200/csrc/predicates.cuh: orient3d_side_prefilter) never claims
more than orient3d's stage A certifies.

The device dismisses an edge/face test when both endpoints are on one side of the tested plane by the bound
|det'| > 2^-43 L^3 (det' = ((T1-T0) x (T2-T0)) . (P-T0), L >= every difference).  predicates.cuh proves that the bound
implies stage A's own certificate (shewchuk.c:2367-2410) with the opposite sign of det'.  Here the same claim is tested
empirically on a few million inputs chosen to hurt: points within 1e-18..1e-9 triangle sizes of the plane, slivers, large
common translations (so that the differences round), mixed scales — with the arithmetic restated in numpy (binary64, no
fused multiply-add: what -fmad=false compiles to), stage A's restatement pinned against the oracle's C function, and the
sign checked against exact rational arithmetic on a subset."""
import ctypes as C
from fractions import Fraction

import numpy as np

ERR_A = 7.7715611723761027e-16  # o3derrboundA, shewchuk.c:420-433


def prefilter(t0, t1, t2, p):
    u, v, w = t1 - t0, t2 - t0, p - t0
    nx = u[:, 1] * v[:, 2] - u[:, 2] * v[:, 1]
    ny = u[:, 2] * v[:, 0] - u[:, 0] * v[:, 2]
    nz = u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]
    luv = np.maximum(np.abs(u).max(axis=1), np.abs(v).max(axis=1))
    l = np.maximum(luv, np.abs(w).max(axis=1))
    det = nx * w[:, 0] + ny * w[:, 1] + nz * w[:, 2]
    thr = l * l * l * 2.0 ** -43
    return np.where(det > thr, -1, np.where(det < -thr, 1, 0))  # orient3d's sign = -sign(det')


def stage_a(pa, pb, pc, pd):
    a, b, c = pa - pd, pb - pd, pc - pd
    bdxcdy, cdxbdy = b[:, 0] * c[:, 1], c[:, 0] * b[:, 1]
    cdxady, adxcdy = c[:, 0] * a[:, 1], a[:, 0] * c[:, 1]
    adxbdy, bdxady = a[:, 0] * b[:, 1], b[:, 0] * a[:, 1]
    det = a[:, 2] * (bdxcdy - cdxbdy) + b[:, 2] * (cdxady - adxcdy) + c[:, 2] * (adxbdy - bdxady)
    perm = (np.abs(bdxcdy) + np.abs(cdxbdy)) * np.abs(a[:, 2]) + (np.abs(cdxady) + np.abs(adxcdy)) * np.abs(b[:, 2]) \
        + (np.abs(adxbdy) + np.abs(bdxady)) * np.abs(c[:, 2])
    bound = ERR_A * perm
    return det, (det > bound) | (-det > bound), perm


def samples(rng, n):
    """(t0, t1, t2, p): triangles of size ~s at offset ~o, p = a point of the plane + h * normal with |h| from 1e-18 s up"""
    s = 10.0 ** rng.uniform(-3, 3, size=(n, 1))
    o = rng.normal(size=(n, 3)) * 10.0 ** rng.uniform(-2, 6, size=(n, 1)) * (rng.random((n, 1)) < 0.7)
    t0 = o + rng.normal(size=(n, 3)) * s
    e1 = rng.normal(size=(n, 3)) * s
    e2 = rng.normal(size=(n, 3)) * s
    sliver = rng.random(n) < 0.25
    e2[sliver] = e1[sliver] * rng.uniform(-2, 2, size=(sliver.sum(), 1)) + e2[sliver] * 10.0 ** rng.uniform(-9, -2, size=(sliver.sum(), 1))
    t1, t2 = t0 + e1, t0 + e2
    nrm = np.cross(e1, e2)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
    ab = rng.uniform(-1.5, 2.5, size=(n, 2))
    h = rng.choice([-1.0, 1.0], size=(n, 1)) * 10.0 ** rng.uniform(-18, 1, size=(n, 1)) * s
    h[rng.random(n) < 0.05] = 0.0
    p = t0 + ab[:, :1] * e1 + ab[:, 1:] * e2 + h * nrm
    return t0, t1, t2, p


def test_stage_a_restatement_equals_the_oracle(oracle):
    rng = np.random.default_rng(7)
    t0, t1, t2, p = samples(rng, 20000)
    det, certain, perm = stage_a(t0, t1, t2, p)
    L = oracle.lib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    for i in range(t0.shape[0]):
        c = C.c_int(0)
        a, b, cc, d = (np.ascontiguousarray(x[i]) for x in (t0, t1, t2, p))
        want = L.mco_orient3d_stageA(dp(a), dp(b), dp(cc), dp(d), C.byref(c))
        # (the oracle's function hands back the permanent instead of the determinant when the filter fails)
        assert bool(c.value) == bool(certain[i]) and want == (det[i] if certain[i] else perm[i]), i


def test_prefilter_verdicts_are_stage_a_certificates():
    rng = np.random.default_rng(11)
    decided = near = 0
    for _ in range(12):
        t0, t1, t2, p = samples(rng, 250000)
        side = prefilter(t0, t1, t2, p)
        det, certain, _ = stage_a(t0, t1, t2, p)
        d = side != 0
        assert np.all(certain[d]), "a verdict of the prefilter that stage A would not certify"
        assert np.all(np.sign(det[d]) == side[d]), "a verdict with the wrong sign"
        decided += int(d.sum())
        near += int((~certain).sum())
    # the sample really contains both regimes: most points are decided, and many are beyond stage A (the cases that matter)
    assert decided > 1500000 and near > 100000


def test_prefilter_sign_against_exact_arithmetic():
    rng = np.random.default_rng(13)
    t0, t1, t2, p = samples(rng, 6000)
    side = prefilter(t0, t1, t2, p)
    F = lambda x: [Fraction(float(v)) for v in x]  # noqa: E731
    for i in np.nonzero(side)[0]:
        a, b, c, d = F(t0[i]), F(t1[i]), F(t2[i]), F(p[i])
        r = [[a[k] - d[k] for k in range(3)], [b[k] - d[k] for k in range(3)], [c[k] - d[k] for k in range(3)]]
        exact = (r[0][0] * (r[1][1] * r[2][2] - r[1][2] * r[2][1]) - r[0][1] * (r[1][0] * r[2][2] - r[1][2] * r[2][0])
                 + r[0][2] * (r[1][0] * r[2][1] - r[1][1] * r[2][0]))
        assert exact != 0 and (1 if exact > 0 else -1) == side[i], i


# ---- the in-register exact sign (predicates.cuh: det3_sign_exact), restated in Python floats -----------------------------
def _two_sum(a, b):
    x = a + b
    bv = x - a
    av = x - bv
    return x, (a - av) + (b - bv)


def _two_product(a, b):  # Dekker / Veltkamp: the same (product, error) pair an FMA gives
    p = a * b
    def split(x):
        c = 134217729.0 * x
        hi = c - (c - x)
        return hi, x - hi
    ah, al = split(a)
    bh, bl = split(b)
    return p, al * bl - (((p - ah * bh) - al * bh) - ah * bl)


def _det3_sign_exact(r):
    """r = rows a, b, c (exact differences).  Returns (decided, sign, passes)."""
    (ax, ay, az), (bx, by, bz), (cx, cy, cz) = r
    t = []
    for x, y, z in ((bx, cy, az), (-cx, by, az), (cx, ay, bz), (-ax, cy, bz), (ax, by, cz), (-bx, ay, cz)):
        p, e = _two_product(x, y)
        lo = _two_product(e, z)
        hi = _two_product(p, z)
        t += [lo[0], lo[1], hi[0], hi[1]]
    for n_pass in range(1, 7):
        for i in range(1, 24):
            t[i], t[i - 1] = _two_sum(t[i], t[i - 1])
        rest = sum(abs(x) for x in t[:23])
        if rest == 0.0 or abs(t[23]) > 2.0 * rest:
            return True, (t[23] > 0) - (t[23] < 0), n_pass
    return False, 0, 6


def test_exact_sign_by_distillation_matches_integer_determinants():
    """rows on a lattice (integers * 2^-20), many of them nearly dependent: the determinant is tiny against its terms or
    exactly zero — the regime the device kernel k_tri_resolve is there for"""
    rng = np.random.default_rng(17)
    undecided, zeros, passes = 0, 0, [0] * 7
    for k in range(20000):
        a = rng.integers(-2 ** 30, 2 ** 30, size=3)
        b = rng.integers(-2 ** 30, 2 ** 30, size=3)
        mode = k % 4
        if mode == 0:
            c = rng.integers(-2 ** 30, 2 ** 30, size=3)
        else:  # c = i*a + j*b (+ a tiny lattice step or nothing): |det| is a few units against terms of 2^90
            i, j = rng.integers(-3, 4, size=2)
            c = i * a + j * b
            if mode != 3:
                c = c + rng.integers(-2, 3, size=3)
        rows_int = [[int(v) for v in r] for r in (a, b, c)]
        rows = [[v * 2.0 ** -20 for v in r] for r in rows_int]
        assert all(float(int(v * 2 ** 20)) == v * 2 ** 20 for r in rows for v in r)
        (ax, ay, az), (bx, by, bz), (cx, cy, cz) = rows_int
        exact = az * (bx * cy - cx * by) + bz * (cx * ay - ax * cy) + cz * (ax * by - bx * ay)
        ok, sg, n_pass = _det3_sign_exact(rows)
        if not ok:
            undecided += 1
            continue
        passes[n_pass] += 1
        zeros += exact == 0
        assert sg == (exact > 0) - (exact < 0), (k, exact)
    assert undecided == 0 and zeros > 1000 and sum(passes[2:]) > 3000  # cancellation really happened, and was resolved
