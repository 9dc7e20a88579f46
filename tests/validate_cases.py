"""Meshes for the input-validation passes: (nv, face_off, face_vtx) triples."""
import numpy as np

from mcut_b200 import meshgen as mg


def _off(faces, sizes):
    if sizes is None:
        return np.arange(0, faces.size + 1, 3, dtype=np.uint32)
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)


def all_cases():
    out = {}
    (sx, sf, ss), (cx, cf, cs), _ = mg.hello_world()
    out["hello_cube_quads"] = (sx.shape[0], _off(sf, ss), sf)
    out["hello_open_patch"] = (cx.shape[0], _off(cf, cs), cf)
    x, f, s = mg.cube_sphere(9, 20.0)
    out["sphere"] = (x.shape[0], _off(f, s), f)
    # two spheres in one mesh + a vertex nobody uses: 3 components (the reference counts the stray vertex)
    f2 = np.concatenate([f, f + x.shape[0]]).astype(np.uint32)
    out["two_spheres_and_a_stray_vertex"] = (2 * x.shape[0] + 1, _off(f2, None), f2)
    # components interleaved in the vertex numbering (ids must follow the smallest vertex of each component)
    perm = np.random.default_rng(7).permutation(2 * x.shape[0]).astype(np.uint32)
    out["two_spheres_shuffled_vertices"] = (2 * x.shape[0], _off(f2, None), perm[f2])
    # open patch: a sphere with a few faces removed
    keep = np.ones(f.size // 3, dtype=bool)
    keep[[0, 1, 17, 40]] = False
    fo = f.reshape(-1, 3)[keep].reshape(-1).astype(np.uint32)
    out["sphere_with_holes"] = (x.shape[0], _off(fo, None), fo)
    tx, tf, ts = mg.terrain(n=40)
    out["open_terrain_grid"] = (tx.shape[0], _off(tf, ts), tf)
    qx, qf, qs = mg.quad_grid(9, 7, quads=True)
    out["quad_grid"] = (qx.shape[0], _off(qf, qs), qf)
    return out
