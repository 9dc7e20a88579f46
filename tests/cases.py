"""Named input cases shared by the CPU (oracle/golden) and GPU (CUDA vs oracle) tests."""
from __future__ import annotations

import numpy as np

from mcut_b200 import meshgen as mg

DBL = mg.MC_DISPATCH_VERTEX_ARRAY_DOUBLE | mg.MC_DISPATCH_ENFORCE_GENERAL_POSITION


def _cube(lo, hi, quads=True, dtype=np.float64):
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    v = np.array([[lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]],
                  [lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]]], dtype=dtype)
    q = np.array([0, 1, 2, 3, 7, 6, 5, 4, 1, 5, 6, 2, 0, 3, 7, 4, 3, 2, 6, 7, 4, 5, 1, 0], dtype=np.uint32)
    if quads:
        return v, q, np.full(6, 4, dtype=np.uint32)
    t = q.reshape(6, 4)
    tri = np.concatenate([t[:, [0, 1, 2]], t[:, [0, 2, 3]]], 1).reshape(-1).astype(np.uint32)
    return v, tri, None


def hello():
    return mg.hello_world()


def spheres_k8():
    return mg.c2_two_spheres(k=8)


def spheres_k16():
    return mg.c2_two_spheres(k=16)


def spheres_k64():
    """49,152 triangles per mesh: several radix-sort tiles, deep trees, > 32 leaf candidates per traversal step."""
    return mg.c2_two_spheres(k=64)


def uv12():
    return mg.two_uv_spheres(12, 20.0)


def ico_pair():
    return mg.c4_pair(3, level=2)


def cube_cube_axis_aligned():
    """Axis-aligned cubes with exactly coplanar/touching features: orient3d hits exact zeros -> GP violation."""
    a = _cube((-5, -5, -5), (5, 5, 5))
    b = _cube((0, 0, 0), (10, 10, 5))  # top face coplanar with a's top face
    return a, b, DBL


def cube_cube_tris_offset():
    a = _cube((-5, -5, -5), (5, 5, 5), quads=False)
    b = _cube((-1.25, -2.5, 1.75), (9.5, 8.25, 7.125), quads=False)
    return a, b, DBL


def patch_vs_sphere():
    """Open cut patch (border edges) through a closed sphere: partial-cut style topology, quads in the cutter."""
    a = mg.cube_sphere(6, 20.0)
    b = mg.quad_grid(7, 5, origin=(-31.3, -27.1, 3.37), du=(9.1, 0.3, 0.11), dv=(0.2, 10.9, 0.07), quads=True)
    return a, b, DBL


def terrain_plane():
    """C3 in miniature: heightfield cut by ONE big triangle (the planar-section supertriangle shape)."""
    a = mg.terrain(n=24, extent=100.0, amp=6.0, white=0.05, seed=7)
    tri = np.array([[-400.0, -380.0, 0.37], [620.0, -410.0, 1.91], [90.0, 700.0, -2.3]])
    return a, (tri, np.array([0, 1, 2], dtype=np.uint32), None), DBL


def near_coplanar_small():
    """C5 in miniature: shallow crossings + 1e-9 noise so the stage-A filter fails on many tests."""
    return mg.c5_near_coplanar(k=10, radius=20.0, amp=1e-3, noise=1e-9, seed=77)


def c5_regions_small():
    """BASELINE config 5 in miniature: dense shallow overlap + regions where the cutter re-triangulates the source's own
    surface to within the resolution of the stage-A orient3d filter (lattice coordinates, see meshgen.c5_coplanar_regions):
    ~800 of the tests need the exact stages, with non-zero determinants."""
    return mg.c5_coplanar_regions(k=12)


def c5_regions_touch():
    """The same with one exactly coincident vertex: general-position violation (status -4) on the unperturbed attempt."""
    return mg.c5_coplanar_regions(k=12, touch=True)


def _tilted_grids(tilt, n=12):
    """Two open triangle grids that are coplanar up to `tilt` rad, then put in a generic orientation so the
    orient3d determinant cancels: the stage-A filter fails on most tests (exact-expansion stress)."""
    R = mg.rot_axis((1.0, 2.0, 3.0), 0.83) @ mg.rot_x(0.41)
    ax, af, _ = mg.quad_grid(n, n, origin=(-20.0, -20.0, 0.0), du=(40.0 / n, 0, 0), dv=(0, 40.0 / n, 0))
    bx, bf, _ = mg.quad_grid(n + 1, n - 1, origin=(-19.3, -18.1, 0.0), du=(38.7 / (n + 1), 0.0, 0), dv=(0.0, 37.9 / (n - 1), 0))
    bx = bx.copy()
    bx[:, 2] = bx[:, 1] * tilt + bx[:, 0] * tilt * 0.37
    return (np.ascontiguousarray(ax @ R.T), af, None), (np.ascontiguousarray(bx @ R.T), bf, None), DBL


def coplanar_rotated():
    return _tilted_grids(0.0)


def near_coplanar_tilt():
    return _tilted_grids(1e-15)


def float_spheres():
    (ax, af, _), (bx, bf, _), _ = mg.c2_two_spheres(k=6)
    return (ax.astype(np.float32), af, None), (bx.astype(np.float32), bf, None), (
        mg.MC_DISPATCH_VERTEX_ARRAY_FLOAT | mg.MC_DISPATCH_ENFORCE_GENERAL_POSITION)


# ---- the reference's known-answer inputs for the general-position classification (tests/source/degenerateInput.cpp:58-144:
# float arrays, no general-position enforcement, mcDispatch must answer MC_INVALID_OPERATION) ----
_T3 = np.array([0, 1, 2], dtype=np.uint32)
_S3 = np.array([3], dtype=np.uint32)


def degenerate_edge_edge():
    """degenerateInput.cpp:58-82: one intersection point would come from two edges crossing."""
    s = np.array([[0, 0, 0], [3, 0, 0], [0, 3, 0]], dtype=np.float32)
    c = np.array([[0, 2, -1], [3, 2, -1], [0, 2, 2]], dtype=np.float32)
    return (s, _T3, _S3), (c, _T3, _S3), mg.MC_DISPATCH_VERTEX_ARRAY_FLOAT


def degenerate_face_vertex():
    """degenerateInput.cpp:87-112: a cut-mesh vertex lies on the source triangle."""
    s = np.array([[0, 0, 0], [3, 0, 0], [0, 3, 0]], dtype=np.float32)
    c = np.array([[1, 1, -3], [3, 1, -3], [1, 1, 0]], dtype=np.float32)
    return (s, _T3, _S3), (c, _T3, _S3), mg.MC_DISPATCH_VERTEX_ARRAY_FLOAT


def degenerate_zero_area():
    """degenerateInput.cpp:117-144.  The test hands DOUBLE arrays to MC_DISPATCH_VERTEX_ARRAY_FLOAT: the library reads the
    first 12 floats of each buffer, i.e. the two halves of each double; that accident is part of the known answer."""
    s = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, -1, 0]], dtype=np.float64).view(np.float32).reshape(-1, 3)[:4].copy()
    c = np.array([[-1, 0, 1], [2, 0, 1], [2, 0, -1], [-1, 0, -1]], dtype=np.float64).view(np.float32).reshape(-1, 3)[:4].copy()
    sf = np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32)
    return (s, sf, np.array([3, 3], dtype=np.uint32)), (c, np.array([0, 1, 2, 3], dtype=np.uint32), np.array([4], dtype=np.uint32)), \
        mg.MC_DISPATCH_VERTEX_ARRAY_FLOAT


ALL = {
    "hello": hello,
    "spheres_k8": spheres_k8,
    "spheres_k16": spheres_k16,
    "spheres_k64": spheres_k64,
    "uv12": uv12,
    "ico_pair": ico_pair,
    "cube_cube_axis_aligned": cube_cube_axis_aligned,
    "cube_cube_tris_offset": cube_cube_tris_offset,
    "patch_vs_sphere": patch_vs_sphere,
    "terrain_plane": terrain_plane,
    "near_coplanar_small": near_coplanar_small,
    "c5_regions_small": c5_regions_small,
    "c5_regions_touch": c5_regions_touch,
    "float_spheres": float_spheres,
    "coplanar_rotated": coplanar_rotated,
    "near_coplanar_tilt": near_coplanar_tilt,
    "degenerate_edge_edge": degenerate_edge_edge,
    "degenerate_face_vertex": degenerate_face_vertex,
    "degenerate_zero_area": degenerate_zero_area,
}
