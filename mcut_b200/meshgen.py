"""Synthetic input meshes for the five BASELINE.json configs plus small test shapes.

Every generator returns ``(xyz[V,3], faces_flat[uint32], sizes[uint32] | None)``; ``sizes is None`` means a
triangle mesh (the reference's ``pFaceSizes == NULL`` convention, preproc.cpp:206).  All meshes are
single-component, edge-manifold and consistently wound (preproc.cpp:531-549) and are scaled so every
triangle passes the reference's degenerate-face rule ``|Newell normal|^2 >= 1e-9`` (math.cpp:151-174).
Plain numpy; deterministic (seeded) so the same arrays are produced here and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

Mesh = Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]

MC_DISPATCH_VERTEX_ARRAY_FLOAT = 1 << 0
MC_DISPATCH_VERTEX_ARRAY_DOUBLE = 1 << 1
MC_DISPATCH_ENFORCE_GENERAL_POSITION = 1 << 15
MC_DISPATCH_ENFORCE_GENERAL_POSITION_ABSOLUTE = 1 << 16


# ----------------------------------------------------------------------------------------------
# C1: tutorials/HelloWorld/HelloWorld.cpp:65-107 (float input, quad cube cut by two triangles)
# ----------------------------------------------------------------------------------------------
def hello_world() -> Tuple[Mesh, Mesh, int]:
    sv = np.array([[-5, -5, 5], [5, -5, 5], [5, 5, 5], [-5, 5, 5], [-5, -5, -5], [5, -5, -5], [5, 5, -5],
                   [-5, 5, -5]], dtype=np.float32)
    sf = np.array([0, 1, 2, 3, 7, 6, 5, 4, 1, 5, 6, 2, 0, 3, 7, 4, 3, 2, 6, 7, 4, 5, 1, 0], dtype=np.uint32)
    ss = np.full(6, 4, dtype=np.uint32)
    cv = np.array([[-20, -4, 0], [0, 20, 20], [20, -4, 0], [0, 20, -20]], dtype=np.float32)
    cf = np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32)
    return (sv, sf, ss), (cv, cf, None), MC_DISPATCH_VERTEX_ARRAY_FLOAT


# ----------------------------------------------------------------------------------------------
# rotations
# ----------------------------------------------------------------------------------------------
def rot_z(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def rot_x(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def rot_axis(axis, a: float) -> np.ndarray:
    u = np.asarray(axis, dtype=np.float64)
    u = u / np.linalg.norm(u)
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    return np.eye(3) + math.sin(a) * K + (1 - math.cos(a)) * (K @ K)


# ----------------------------------------------------------------------------------------------
# cube-sphere: each cube face a k x k quad grid split into 2k^2 triangles; 12k^2 tris, 6k^2+2 verts
# ----------------------------------------------------------------------------------------------
def cube_sphere_grid(k: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Unit directions, triangle indices and the per-cube-face vertex-id grids ``vid[6, k+1, k+1]`` of a cube-sphere
    with k subdivisions per cube edge.  Cell (f, i, j) holds triangles 2(f k^2 + i k + j) and +1:
    (v00, v10, v11) and (v00, v11, v01) with vAB = vid[f, i+A, j+B]."""
    n = k + 1
    # integer lattice points on the surface of the cube [0,k]^3, welded by a dict-free numpy unique
    lin = np.arange(n, dtype=np.int64)
    a, b = np.meshgrid(lin, lin, indexing="ij")
    a = a.ravel()
    b = b.ravel()
    zeros = np.zeros_like(a)
    full = np.full_like(a, k)
    # (axis order chosen so that every face is wound counter-clockwise seen from outside)
    cube_faces = [
        np.stack([full, a, b], 1),  # +x
        np.stack([zeros, b, a], 1),  # -x
        np.stack([b, full, a], 1),  # +y
        np.stack([a, zeros, b], 1),  # -y
        np.stack([a, b, full], 1),  # +z
        np.stack([b, a, zeros], 1),  # -z
    ]
    tris = []
    all_pts = np.concatenate(cube_faces, 0)
    key = (all_pts[:, 0] * n + all_pts[:, 1]) * n + all_pts[:, 2]
    uniq, first_idx, inv = np.unique(key, return_index=True, return_inverse=True)
    # keep first-appearance order for vertex ids (deterministic, locality-friendly)
    order = np.argsort(first_idx, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    vid = rank[inv]
    pts = all_pts[first_idx[order]].astype(np.float64)
    i, j = np.meshgrid(np.arange(k), np.arange(k), indexing="ij")
    i = i.ravel()
    j = j.ravel()
    for f in range(6):
        base = f * n * n
        v00 = vid[base + i * n + j]
        v10 = vid[base + (i + 1) * n + j]
        v11 = vid[base + (i + 1) * n + (j + 1)]
        v01 = vid[base + i * n + (j + 1)]
        t = np.empty((k * k, 2, 3), dtype=np.int64)
        t[:, 0, :] = np.stack([v00, v10, v11], 1)
        t[:, 1, :] = np.stack([v00, v11, v01], 1)
        tris.append(t.reshape(-1, 3))
    tris = np.concatenate(tris, 0)
    c = pts / k * 2.0 - 1.0  # cube [-1,1]^3
    # tan-warp for more uniform triangles, then normalise
    c = np.tan(c * (math.pi / 4.0))
    d = c / np.linalg.norm(c, axis=1, keepdims=True)
    return d, tris.astype(np.uint32), vid.reshape(6, n, n)


def cube_sphere_dirs(k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Unit directions + triangle indices of a cube-sphere with k subdivisions per cube edge."""
    d, tris, _ = cube_sphere_grid(k)
    return d, tris


def cube_sphere(k: int, radius: float = 20.0, rotation: Optional[np.ndarray] = None,
                centre=(0.0, 0.0, 0.0), radial: Optional[np.ndarray] = None) -> Mesh:
    d, tris = cube_sphere_dirs(k)
    if rotation is not None:
        d = d @ rotation.T
    r = radius if radial is None else radius * (1.0 + radial)
    xyz = d * (r[:, None] if isinstance(r, np.ndarray) else r) + np.asarray(centre, dtype=np.float64)
    return np.ascontiguousarray(xyz), np.ascontiguousarray(tris.ravel()), None


def c2_two_spheres(k: int = 289, radius: float = 20.0) -> Tuple[Mesh, Mesh, int]:
    """C2: CSG of two cube-spheres (k=289 -> 1,002,252 triangles each), crossing on one closed curve."""
    a = cube_sphere(k, radius)
    rot = rot_x(0.26) @ rot_z(0.37)
    b = cube_sphere(k, radius, rotation=rot, centre=(0.9 * radius, 0.13 * radius, 0.07 * radius))
    return a, b, MC_DISPATCH_VERTEX_ARRAY_DOUBLE | MC_DISPATCH_ENFORCE_GENERAL_POSITION


def _sph_field(d: np.ndarray) -> np.ndarray:
    """Smooth, sign-changing low-frequency field on the sphere (three real spherical harmonics)."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    return 0.6 * (x * y) + 0.5 * (z * (x * x - y * y)) + 0.4 * (3 * z * z - 1.0) * x


def c5_near_coplanar(k: int = 409, radius: float = 20.0, amp: float = 1e-3, noise: float = 1e-9,
                     seed: int = 77) -> Tuple[Mesh, Mesh, int]:
    """Round-1 C5 recipe (SURVEY §8-d as written): dense overlap, but NO test reaches the exact stages — the stage-A
    filter is relative to the permanent and only fails within ~1e-15 of a plane.  Kept as a dense-overlap case; the
    BASELINE config 5 workload is `c5_coplanar_regions`."""
    a = cube_sphere(k, radius)
    d, _ = cube_sphere_dirs(k)
    rot = rot_axis((1.0, 2.0, 3.0), 1e-3)
    rng = np.random.default_rng(seed)
    u = rng.uniform(-1.0, 1.0, size=d.shape[0])
    radial = amp * _sph_field(d) + noise * u
    b = cube_sphere(k, radius, rotation=rot, radial=radial)
    return a, b, MC_DISPATCH_VERTEX_ARRAY_DOUBLE | MC_DISPATCH_ENFORCE_GENERAL_POSITION


LATTICE_BITS = 47  # coordinates of the C5 meshes are multiples of 2^-47 (see c5_coplanar_regions)


def _snap(x: np.ndarray) -> np.ndarray:
    return np.round(x * float(1 << LATTICE_BITS)) / float(1 << LATTICE_BITS)


def raycast_cube_sphere(k: int, a_xyz: np.ndarray, vid: np.ndarray, d: np.ndarray):
    """Where the ray from the origin along unit direction d[i] meets the piecewise-planar surface of the cube-sphere
    (a_xyz, vid) centred at the origin: returns (points, ok, tri) with tri[i] = the three vertex ids of the hit triangle.
    The triangle is found analytically (cube face -> inverse tan-warp -> grid cell -> which of the cell's two triangles);
    ok is False where neither triangle of the cell holds the hit well inside (the ray passes close to a mesh edge)."""
    ax = np.argmax(np.abs(d), axis=1)
    sgn = np.sign(d[np.arange(d.shape[0]), ax])
    # cube-face frames exactly as cube_sphere_grid lays them out: on face f the grid coordinates (i, j) run along axes (p, q)
    face = np.where(sgn > 0, ax * 2, ax * 2 + 1)
    pq = {0: (1, 2), 1: (2, 1), 2: (2, 0), 3: (0, 2), 4: (0, 1), 5: (1, 0)}
    out = np.zeros_like(d)
    ok = np.zeros(d.shape[0], dtype=bool)
    tri = np.zeros((d.shape[0], 3), dtype=np.int64)
    for f in range(6):
        m = np.nonzero(face == f)[0]
        if m.size == 0:
            continue
        p_ax, q_ax = pq[f]
        major = np.abs(d[m, f // 2])
        s_p = np.arctan(d[m, p_ax] / major) / (math.pi / 4.0)
        s_q = np.arctan(d[m, q_ax] / major) / (math.pi / 4.0)
        ci = np.clip(np.floor((s_p + 1.0) * 0.5 * k).astype(np.int64), 0, k - 1)
        cj = np.clip(np.floor((s_q + 1.0) * 0.5 * k).astype(np.int64), 0, k - 1)
        i00, i10, i11, i01 = vid[f, ci, cj], vid[f, ci + 1, cj], vid[f, ci + 1, cj + 1], vid[f, ci, cj + 1]
        dd = d[m]
        done = np.zeros(m.size, dtype=bool)
        res = np.zeros((m.size, 3))
        rtri = np.zeros((m.size, 3), dtype=np.int64)
        for (j0, j1, j2) in ((i00, i10, i11), (i00, i11, i01)):
            t0, t1, t2 = a_xyz[j0], a_xyz[j1], a_xyz[j2]
            nrm = np.cross(t1 - t0, t2 - t0)
            t = np.einsum("ij,ij->i", nrm, t0) / np.einsum("ij,ij->i", nrm, dd)
            hit = dd * t[:, None]

            def side(p0, p1):
                return np.einsum("ij,ij->i", np.cross(p1 - p0, hit - p0), nrm)
            area2 = np.einsum("ij,ij->i", nrm, nrm)
            w0, w1, w2 = side(t1, t2) / area2, side(t2, t0) / area2, side(t0, t1) / area2
            inside = (w0 > 1e-4) & (w1 > 1e-4) & (w2 > 1e-4) & ~done
            res[inside] = hit[inside]
            rtri[inside] = np.stack([j0, j1, j2], 1)[inside]
            done |= inside
        out[m] = res
        ok[m] = done
        tri[m] = rtri
    return out, ok, tri


def _nearly_on_plane(p: np.ndarray, a0: np.ndarray, a1: np.ndarray, a2: np.ndarray, reach: int = 24) -> np.ndarray:
    """For every row: a lattice point (multiples of 2^-LATTICE_BITS) next to p that lies extremely close to — but not on —
    the plane of the lattice triangle (a0, a1, a2).  With N the (integer) plane normal, the signed distance of p + delta is
    proportional to D0 + N.delta: for each of the (2 reach + 1)^2 offsets along the two axes where |N| is larger, the offset
    along the third axis that brings the sum closest to zero is a rounding; the best non-zero sum over all of them wins.
    Exact integer arithmetic has the last word (the point must not be ON the plane)."""
    scale = 1 << LATTICE_BITS
    to_int = lambda x: np.array([[int(v) for v in row] for row in np.round(x * float(scale))], dtype=object)  # noqa: E731
    A0, A1, A2, Q0 = to_int(a0), to_int(a1), to_int(a2), to_int(p)
    U, V = A1 - A0, A2 - A0
    N = np.stack([U[:, 1] * V[:, 2] - U[:, 2] * V[:, 1], U[:, 2] * V[:, 0] - U[:, 0] * V[:, 2], U[:, 0] * V[:, 1] - U[:, 1] * V[:, 0]], 1)
    W = Q0 - A0
    D0 = N[:, 0] * W[:, 0] + N[:, 1] * W[:, 1] + N[:, 2] * W[:, 2]
    S = np.array([max(abs(n[0]), abs(n[1]), abs(n[2])) for n in N], dtype=object)
    nhat = np.array([[float(n[j]) / float(s) for j in range(3)] for n, s in zip(N, S)])
    d0 = np.array([float(d) / float(s) for d, s in zip(D0, S)])
    n_rows = p.shape[0]
    order = np.argsort(-np.abs(nhat), axis=1)  # axes by decreasing |N|: the last one is the fine adjustment
    rows = np.arange(n_rows)
    n_a, n_b, n_c = nhat[rows, order[:, 0]], nhat[rows, order[:, 1]], nhat[rows, order[:, 2]]
    r = np.arange(-reach, reach + 1, dtype=np.float64)
    ia, ib = [g.ravel() for g in np.meshgrid(r, r, indexing="ij")]
    best = np.zeros((n_rows, 3), dtype=np.int64)
    for lo in range(0, n_rows, 4096):
        hi = min(lo + 4096, n_rows)
        w = d0[lo:hi, None] + n_a[lo:hi, None] * ia[None, :] + n_b[lo:hi, None] * ib[None, :]
        nc = np.where(np.abs(n_c[lo:hi]) < 1e-6, 1e-6, n_c[lo:hi])[:, None]  # a normal (almost) along an axis: no fine adjustment
        kc = np.clip(np.round(-w / nc), -200000.0, 200000.0)
        res = np.abs(w + kc * n_c[lo:hi, None])
        res[res < 1e-9] = np.inf  # would be (or be indistinguishable from) a point ON the plane: exactly zero determinant
        pick = np.argmin(res, axis=1)
        sel = np.arange(hi - lo)
        best[np.arange(lo, hi), order[lo:hi, 0]] = ia[pick].astype(np.int64)
        best[np.arange(lo, hi), order[lo:hi, 1]] = ib[pick].astype(np.int64)
        best[np.arange(lo, hi), order[lo:hi, 2]] = kc[sel, pick].astype(np.int64)
    D = D0 + N[:, 0] * best[:, 0] + N[:, 1] * best[:, 1] + N[:, 2] * best[:, 2]
    assert all(int(v) != 0 for v in D), "a chosen lattice point lies exactly on its plane"
    q = Q0 + best.astype(object)
    return np.array([[float(int(v)) for v in row] for row in q]) / float(scale)


def c5_coplanar_regions(k: int = 409, radius: float = 20.0, amp: float = 1e-3, noise: float = 1e-9, seed: int = 77,
                        cap_cos: float = 0.93, touch: bool = False) -> Tuple[Mesh, Mesh, int]:
    """BASELINE config 5: two ~2M-triangle noisy spheres in dense overlap with NEAR-COPLANAR REGIONS that defeat the
    stage-A orient3d filter (shewchuk.c: the filter is RELATIVE, |det| <= 7.8e-16 * permanent: a point must lie within
    ~1e-15 triangle sizes of a plane, so metre-scale noise never gets there).
    A = cube-sphere.  B = A's topology rotated by 1e-3 rad about (1,2,3), radially displaced by a smooth sign-changing
    field (amplitude `amp`) + white noise: shallow crossings everywhere.  Inside six spherical caps (cos to the cap axis >
    cap_cos, ~20 % of the sphere) B is a re-triangulation of A's own piecewise-planar surface: each of its vertices is put
    next to the plane of the A triangle under it, closer than the filter can resolve but never exactly on it.
    What makes that survive the reference's re-centring x' = (x - com) + shift (preproc.cpp:124-176), whose rounding would
    otherwise push every point ~1e-14 off its plane: ALL coordinates are multiples of 2^-47.  Then both additions round
    to a grid the inputs already lie on, so within one binade of the intermediate and of the result they add the same
    constant to every coordinate — an exact translation, and determinants keep their exact values.  The near-planar
    vertex is the lattice point, among the 11^3 around the ray-cast hit, with the smallest non-zero exact determinant."""
    flags = MC_DISPATCH_VERTEX_ARRAY_DOUBLE | MC_DISPATCH_ENFORCE_GENERAL_POSITION
    # the exact-integer search takes ~25 s at k = 409: keep the arrays for later calls on the same machine (tests + bench)
    cache = None
    if k >= 128:
        import os
        import tempfile
        cache = os.path.join(tempfile.gettempdir(), f"mcut_b200_c5_{k}_{radius!r}_{amp!r}_{noise!r}_{seed}_{cap_cos!r}_{int(touch)}_v2.npz")
        if os.path.exists(cache):
            try:
                z = np.load(cache)
                return (z["a"], z["f"], None), (z["b"], z["f"], None), flags
            except Exception:
                pass
    d_a, tris, vid = cube_sphere_grid(k)
    a_xyz = _snap(np.ascontiguousarray(d_a * radius))
    rot = rot_axis((1.0, 2.0, 3.0), 1e-3)
    d_b = d_a @ rot.T
    rng = np.random.default_rng(seed)
    u = rng.uniform(-1.0, 1.0, size=d_b.shape[0])
    radial = amp * _sph_field(d_a) + noise * u
    b_xyz = _snap(d_b * (radius * (1.0 + radial))[:, None])
    axes = np.array([[1, 2, 2], [-2, 1, 2], [2, -2, 1], [-1, -2, -2], [2, -1, -2], [-2, 2, -1]], dtype=np.float64) / 3.0
    in_cap = (d_b @ axes.T).max(axis=1) > cap_cos
    idx = np.nonzero(in_cap)[0]
    hit, ok, tri = raycast_cube_sphere(k, a_xyz, vid, d_b[idx])
    idx, hit, tri = idx[ok], hit[ok], tri[ok]
    b_xyz[idx] = _nearly_on_plane(hit, a_xyz[tri[:, 0]], a_xyz[tri[:, 1]], a_xyz[tri[:, 2]])
    if touch:
        # optional: ONE vertex of B exactly on a vertex of A — an exactly zero determinant, i.e. a general-position violation
        # (status -4) on the unperturbed attempt, as any input with coincident geometry produces
        b_xyz[idx[0]] = a_xyz[tri[0, 0]]
    faces = np.ascontiguousarray(tris.ravel())
    b_xyz = np.ascontiguousarray(b_xyz)
    if cache:
        try:
            tmp = cache + f".{os.getpid()}.tmp.npz"
            np.savez(tmp, a=a_xyz, b=b_xyz, f=faces)
            os.replace(tmp, cache)
        except Exception:
            pass
    return (a_xyz, faces, None), (b_xyz, faces, None), flags


# ----------------------------------------------------------------------------------------------
# UV sphere of SURVEY.md Appendix C (the input the Appendix-A golden counts were captured on)
# ----------------------------------------------------------------------------------------------
def uv_sphere(rings: int, centre=(0.0, 0.0, 0.0), radius: float = 1.0, rho: float = 0.0) -> Mesh:
    r, s = rings, 2 * rings
    dirs = [(0.0, 0.0, 1.0)]
    for i in range(1, r):
        th = math.pi * i / r
        for j in range(s):
            ph = 2.0 * math.pi * j / s
            dirs.append((math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)))
    dirs.append((0.0, 0.0, -1.0))
    d = np.array(dirs, dtype=np.float64)
    x1 = d[:, 0] * math.cos(rho) - d[:, 1] * math.sin(rho)
    y1 = d[:, 0] * math.sin(rho) + d[:, 1] * math.cos(rho)
    y2 = y1 * math.cos(0.7 * rho) - d[:, 2] * math.sin(0.7 * rho)
    z2 = y1 * math.sin(0.7 * rho) + d[:, 2] * math.cos(0.7 * rho)
    xyz = np.asarray(centre, dtype=np.float64) + radius * np.stack([x1, y2, z2], 1)

    def vid(i, j):
        return 1 + (i - 1) * s + (j % s)

    south = 1 + (r - 1) * s
    f = []
    for j in range(s):
        f.append((0, vid(1, j), vid(1, j + 1)))
    for i in range(1, r - 1):
        for j in range(s):
            f.append((vid(i, j), vid(i + 1, j), vid(i + 1, j + 1)))
            f.append((vid(i, j), vid(i + 1, j + 1), vid(i, j + 1)))
    for j in range(s):
        f.append((south, vid(r - 1, j + 1), vid(r - 1, j)))
    return np.ascontiguousarray(xyz), np.array(f, dtype=np.uint32).ravel(), None


def two_uv_spheres(rings: int, radius: float) -> Tuple[Mesh, Mesh, int]:
    a = uv_sphere(rings, (0, 0, 0), radius, 0.0)
    b = uv_sphere(rings, (0.9 * radius, 0.13 * radius, 0.07 * radius), radius, 0.37)
    return a, b, MC_DISPATCH_VERTEX_ARRAY_DOUBLE | MC_DISPATCH_ENFORCE_GENERAL_POSITION


# ----------------------------------------------------------------------------------------------
# icosphere (C4: level 4 -> 5,120 triangles / 2,562 vertices)
# ----------------------------------------------------------------------------------------------
def icosphere(level: int, radius: float = 20.0) -> Mesh:
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(level):
        nv = v.shape[0]
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
        es = np.sort(e, 1)
        key = es[:, 0] * nv + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        mid = v[uniq // nv] + v[uniq % nv]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        m = nv + inv.reshape(3, -1)  # m[0]=ab, m[1]=bc, m[2]=ca
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        f = np.concatenate([np.stack([a, m[0], m[2]], 1), np.stack([b, m[1], m[0]], 1),
                            np.stack([c, m[2], m[1]], 1), np.stack([m[0], m[1], m[2]], 1)], 0)
        v = np.concatenate([v, mid], 0)
    return np.ascontiguousarray(v * radius), f.astype(np.uint32).ravel(), None


def random_rotation(rng: np.random.Generator) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


_ICO_CACHE = {}


def c4_pair(j: int, level: int = 4, radius: float = 20.0) -> Tuple[Mesh, Mesh, int]:
    """C4 pair j: two icospheres, B scaled U(0.8,1.2), randomly rotated, offset U(0.3R,1.2R); seed 1000+j."""
    if level not in _ICO_CACHE:
        _ICO_CACHE[level] = icosphere(level, 1.0)
    v, f, _ = _ICO_CACHE[level]
    rng = np.random.default_rng(1000 + j)
    scale = rng.uniform(0.8, 1.2)
    rot = random_rotation(rng)
    direction = rng.normal(size=3)
    direction /= np.linalg.norm(direction)
    off = rng.uniform(0.3 * radius, 1.2 * radius) * direction
    a = (np.ascontiguousarray(v * radius), f, None)
    b = (np.ascontiguousarray((v * (radius * scale)) @ rot.T + off), f, None)
    return a, b, MC_DISPATCH_VERTEX_ARRAY_DOUBLE | MC_DISPATCH_ENFORCE_GENERAL_POSITION


# ----------------------------------------------------------------------------------------------
# C3: noisy terrain heightfield + section planes
# ----------------------------------------------------------------------------------------------
def _value_noise(n: int, cells: int, rng: np.random.Generator) -> np.ndarray:
    g = rng.uniform(-1.0, 1.0, size=(cells + 1, cells + 1))
    t = np.linspace(0.0, cells, n, endpoint=True)
    i = np.minimum(t.astype(np.int64), cells - 1)
    fr = t - i
    fr = fr * fr * (3 - 2 * fr)
    gx0 = g[i][:, i]
    gx1 = g[i + 1][:, i]
    gy0 = g[i][:, i + 1]
    gy1 = g[i + 1][:, i + 1]
    fx = fr[:, None]
    fy = fr[None, :]
    return (gx0 * (1 - fx) + gx1 * fx) * (1 - fy) + (gy0 * (1 - fx) + gy1 * fx) * fy


def terrain(n: int = 1415, extent: float = 400.0, amp: float = 10.0, white: float = 0.05, seed: int = 1234) -> Mesh:
    """n x n vertex grid over [0,extent]^2 (2(n-1)^2 triangles); 3-octave value noise + white noise."""
    rng = np.random.default_rng(seed)
    h = np.zeros((n, n))
    for o, cells in enumerate((4, 8, 16)):
        h += (amp / (2 ** o)) * _value_noise(n, cells, rng)
    h += rng.uniform(-white, white, size=(n, n))
    xs = np.linspace(0.0, extent, n)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    xyz = np.stack([X.ravel(), Y.ravel(), h.ravel()], 1)
    i, j = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    i = i.ravel()
    j = j.ravel()
    v00 = i * n + j
    v10 = (i + 1) * n + j
    v11 = (i + 1) * n + j + 1
    v01 = i * n + j + 1
    t = np.empty((i.size, 2, 3), dtype=np.int64)
    t[:, 0, :] = np.stack([v00, v10, v11], 1)
    t[:, 1, :] = np.stack([v00, v11, v01], 1)
    return np.ascontiguousarray(xyz), t.reshape(-1).astype(np.uint32), None


def c3_plane(k: int, count: int = 256) -> Tuple[np.ndarray, float]:
    """Section plane k of `count`: (unit normal, sectionOffset in [0,1]) per SURVEY.md §8-d."""
    theta = 2.0 * math.pi * k / count
    phi = 0.2 + 0.6 * ((k * 0.618) % 1.0)
    nrm = np.array([math.sin(phi) * math.cos(theta), math.sin(phi) * math.sin(theta), math.cos(phi)])
    nrm /= np.linalg.norm(nrm)
    off = 0.1 + 0.8 * k / max(count - 1, 1)
    return nrm, off


def c3_supertriangle(xyz: np.ndarray, normal) -> Mesh:
    """The cut mesh mcEnqueueDispatchPlanarSection builds for a section of `xyz` (frontend.cpp:722-899,
    generate_supertriangle_from_mesh_vertices): one triangle of side ~4 bounding-box diagonals in the plane through the
    vertex mean projected along the normal (the reference computes an offset centroid and then does not use it, :802)."""
    n = np.asarray(normal, dtype=np.float64)
    n = n / np.linalg.norm(n)
    mean = xyz.mean(axis=0)
    diag = float(np.linalg.norm(xyz.max(axis=0) - xyz.min(axis=0)))
    k = int(np.argmax(np.abs(n)))
    w = np.zeros(3)
    w[k] = 1.0
    if np.array_equal(w, n):
        w[k] = 0.0
        w[(k + 1) % 3] = 1.0
    u = np.cross(n, w)
    v = np.cross(n, u)
    on_plane = mean - n * float(mean @ n)
    uv_pos = (u + v) / np.linalg.norm(u + v)
    uv_neg = (u - v) / np.linalg.norm(u - v)
    third = (uv_pos + uv_neg) / np.linalg.norm(uv_pos + uv_neg)
    tri = np.stack([on_plane + uv_pos * (2.0 * diag), on_plane + uv_neg * (2.0 * diag), on_plane - third * (2.0 * diag)])
    return np.ascontiguousarray(tri), np.array([0, 1, 2], dtype=np.uint32), None


# ----------------------------------------------------------------------------------------------
# small shapes for unit tests
# ----------------------------------------------------------------------------------------------
def quad_grid(nx: int, ny: int, origin=(0.0, 0.0, 0.0), du=(1.0, 0.0, 0.0), dv=(0.0, 1.0, 0.0), quads=False) -> Mesh:
    """Open (bordered) planar grid patch; triangles, or quads when quads=True."""
    o = np.asarray(origin, dtype=np.float64)
    du = np.asarray(du, dtype=np.float64)
    dv = np.asarray(dv, dtype=np.float64)
    i, j = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
    xyz = o + i.reshape(-1, 1) * du + j.reshape(-1, 1) * dv
    ci, cj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    ci = ci.ravel()
    cj = cj.ravel()
    n = ny + 1
    v00 = ci * n + cj
    v10 = (ci + 1) * n + cj
    v11 = (ci + 1) * n + cj + 1
    v01 = ci * n + cj + 1
    if quads:
        f = np.stack([v00, v10, v11, v01], 1).astype(np.uint32).ravel()
        return np.ascontiguousarray(xyz), f, np.full(ci.size, 4, dtype=np.uint32)
    t = np.empty((ci.size, 2, 3), dtype=np.int64)
    t[:, 0, :] = np.stack([v00, v10, v11], 1)
    t[:, 1, :] = np.stack([v00, v11, v01], 1)
    return np.ascontiguousarray(xyz), t.reshape(-1).astype(np.uint32), None


def mesh_counts(m: Mesh) -> Tuple[int, int]:
    xyz, f, s = m
    return xyz.shape[0], (s.size if s is not None else f.size // 3)
