"""The inputs of the reference's tests/source/intersectionType.cpp (:314-760): cubes of quads and single quads, with the
verdict each test asserts (McDispatchIntersectionType: 0 STANDARD, 2 INSIDE_CUTMESH, 4 INSIDE_SOURCEMESH, 8 NONE)."""
import numpy as np

from mcut_b200 import meshgen as mg

FLAGS = mg.MC_DISPATCH_VERTEX_ARRAY_DOUBLE | (1 << 17)  # MC_DISPATCH_INCLUDE_INTERSECTION_TYPE (mcut.h:440)
STANDARD, INSIDE_CUTMESH, INSIDE_SOURCEMESH, NONE = 0, 2, 4, 8


def cube(h, tx, ty, tz):
    """makeCube (intersectionType.cpp:91-200): front face first, six quads."""
    v = np.array([[-h, -h, h], [h, -h, h], [h, h, h], [-h, h, h], [-h, -h, -h], [h, -h, -h], [h, h, -h], [-h, h, -h]], dtype=np.float64)
    v += np.array([tx, ty, tz], dtype=np.float64)
    f = np.array([0, 1, 2, 3, 7, 6, 5, 4, 4, 0, 3, 7, 1, 5, 6, 2, 3, 2, 6, 7, 0, 4, 5, 1], dtype=np.uint32)
    return v, f, np.full(6, 4, dtype=np.uint32)


def quad_xz(h, tx, ty, tz):
    """makeQuad_xz (:203-256)."""
    v = np.array([[-h, 0, h], [h, 0, h], [h, 0, -h], [-h, 0, -h]], dtype=np.float64) + np.array([tx, ty, tz], dtype=np.float64)
    return v, np.array([0, 1, 2, 3], dtype=np.uint32), np.array([4], dtype=np.uint32)


def quad_xy(h, tx, ty, tz):
    """makeQuad_xy (:258-311)."""
    v = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], dtype=np.float64) + np.array([tx, ty, tz], dtype=np.float64)
    return v, np.array([0, 1, 2, 3], dtype=np.uint32), np.array([4], dtype=np.uint32)


def _sphere():
    x, f, s = mg.cube_sphere(6, 1.0)
    return x, f, None


def _corner_cube():
    v, f, s = cube(0.1, 0.9, 0.9, 0.9)  # inside the sphere's AABB, outside the sphere
    return v, f, s


# name -> (src, cut, flags, what the reference's test asserts; None: not one of its tests, the fixture holds its answer)
CASES = {
    "watertightCutMeshInsideWatertightSourceMesh": (cube(2.0, 0, 0, 0), cube(1.0, 0, 0, 0), FLAGS, INSIDE_SOURCEMESH),
    "watertightSourceMeshInsideWatertightCutMesh": (cube(1.0, 0, 0, 0), cube(2.0, 0, 0, 0), FLAGS, INSIDE_CUTMESH),
    "separatedWatertightSourceMeshAndWatertightCutMesh": (cube(1.0, -2, 0, 0), cube(1.0, 2, 0, 0), FLAGS, NONE),
    "stdIntersectionWatertightSourceMeshAndWatertightCutMesh": (cube(1.0, 0, 0, 0), cube(1.0, 0.5, 0.5, 0.5), FLAGS, STANDARD),
    "separatedWatertightSourceMeshAndOpenCutMesh": (cube(1.0, 0, 0, 0), quad_xz(2, 0, 2, 0), FLAGS, NONE),
    "openCutMeshInsideWatertightSourceMesh": (cube(10.0, 0, 0, 0), quad_xz(2, 0, 2, 0), FLAGS, INSIDE_SOURCEMESH),
    "openCutMeshIntersectsWatertightSourceMesh": (cube(1.0, 0, 0, 0), quad_xz(1, 0.5, 0, 0.5), FLAGS, STANDARD),
    "watertightCutMeshIntersectsOpenSourceMesh": (quad_xz(1, 0.5, 0, 0.5), cube(1.0, 0, 0, 0), FLAGS, STANDARD),
    "openSourceMeshInsideWatertightCutMesh": (quad_xz(1, 0, 0, 0), cube(10.0, 0, 0, 0), FLAGS, INSIDE_CUTMESH),
    "separatedOpenSourceMeshAndWatertightCutMesh": (quad_xz(1, 10, 0, 0), cube(1, -10, 0, 0), FLAGS, NONE),
    "separatedOpenSourceMeshAndOpenCutMesh": (quad_xz(1, 10, 0, 0), quad_xz(1, -10, 0, 0), FLAGS, NONE),
    "intersectingOpenSourceMeshAndOpenCutMesh": (quad_xz(1, 0, 0, 0), quad_xy(2, 0, 0, 0), FLAGS, STANDARD),
    # both watertight, AABBs overlap, surfaces apart: both winding-number questions are asked (larger mesh first)
    "extra_sphere_and_cube_in_the_corner_of_its_box": (_sphere(), _corner_cube(), FLAGS, None),
    "extra_cube_in_the_corner_of_the_box_of_a_sphere": (_corner_cube(), _sphere(), FLAGS, None),
}
