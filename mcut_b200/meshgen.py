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
def cube_sphere_dirs(k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Unit directions + triangle indices of a cube-sphere with k subdivisions per cube edge."""
    n = k + 1
    # integer lattice points on the surface of the cube [0,k]^3, welded by a dict-free numpy unique
    faces_pts = []
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
    return d, tris.astype(np.uint32)


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
    """C5: two cube-spheres, B = A rotated by 1e-3 rad and displaced radially by a shallow sign-changing
    field + 1e-9-scale white noise: long shallow-angle crossings that stress the exact orient3d fallback."""
    a = cube_sphere(k, radius)
    d, _ = cube_sphere_dirs(k)
    rot = rot_axis((1.0, 2.0, 3.0), 1e-3)
    rng = np.random.default_rng(seed)
    u = rng.uniform(-1.0, 1.0, size=d.shape[0])
    radial = amp * _sph_field(d) + noise * u
    b = cube_sphere(k, radius, rotation=rot, radial=radial)
    return a, b, MC_DISPATCH_VERTEX_ARRAY_DOUBLE | MC_DISPATCH_ENFORCE_GENERAL_POSITION


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
