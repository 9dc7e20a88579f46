"""-m gpu: the drop-in itself.  The UNMODIFIED reference library (oracle/_ref/libmcut_ref.so) is driven through its
public C API (mcCreateContext, mcDispatch, mcGetConnectedComponents, mcGetConnectedComponentData) by a plain client
(oracle/_ref/api_driver) twice: as is, and with mcut_b200/lib/libmcut_b200_shim.so preloaded so build_oibvh() and
intersectOIBVHs() run on the B200.  Every connected component must come back bit-identical."""
import os
import subprocess

import numpy as np
import pytest

import cases
from mcut_b200.mcbio import read_mcb, write_mcb

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "api_driver")
SHIM = os.path.join(ROOT, "mcut_b200", "lib", "libmcut_b200_shim.so")
NODUMP = os.path.join(ROOT, "oracle", "_ref", "libnodump.so")

needs_ref = pytest.mark.skipif(not (os.path.exists(DRIVER) and os.path.exists(SHIM)),
                               reason="oracle/_ref/api_driver or the shim is not built (needs /root/reference at build time)")


def run_driver(tmp, tag, src, cut, flags, preload, extra=(), driver=None):
    d = {"src_xyz": src[0], "src_faces": src[1], "cut_xyz": cut[0], "cut_faces": cut[1], "flags": np.array([flags], dtype=np.uint32)}
    if src[2] is not None:
        d["src_sizes"] = src[2]
    if cut[2] is not None:
        d["cut_sizes"] = cut[2]
    ip, op = os.path.join(tmp, f"{tag}.in.mcb"), os.path.join(tmp, f"{tag}.out.mcb")
    write_mcb(ip, d)
    env = dict(os.environ, LD_PRELOAD=":".join(preload))
    r = subprocess.run([driver or DRIVER, ip, op, *extra], capture_output=True, text=True, cwd=tmp, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    out = read_mcb(op)
    out["_stderr"] = r.stderr[-2000:]
    return out


@needs_ref
@pytest.mark.parametrize("case", ["hello", "spheres_k16", "patch_vs_sphere", "cube_cube_axis_aligned", "ico_pair", "float_spheres"])
def test_mcdispatch_with_shim_is_bit_identical(tmp_path, case):
    src, cut, flags = cases.ALL[case]()
    a = run_driver(str(tmp_path), "ref", src, cut, flags, [NODUMP])
    b = run_driver(str(tmp_path), "b200", src, cut, flags, [SHIM, NODUMP])
    assert int(a["mcDispatch_result"][0]) == int(b["mcDispatch_result"][0])
    for k in ("cc_type", "cc_attrs", "cc_nv", "cc_nf", "cc_vertices", "cc_faces", "cc_face_sizes"):
        assert a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes(), k
    assert a["cc_type"].size > 0


@needs_ref
def test_hello_world_counts_through_the_shim(tmp_path):
    """tutorials/HelloWorld/HelloWorld.cpp:65-134 -> 12 connected components."""
    src, cut, flags = cases.hello()
    b = run_driver(str(tmp_path), "hello", src, cut, flags, [SHIM, NODUMP])
    got = sorted(zip(b["cc_type"].tolist(), b["cc_nv"].tolist(), b["cc_nf"].tolist()))
    want = sorted([(16, 8, 6), (16, 4, 2), (8, 10, 4), (8, 14, 10), (4, 10, 2), (4, 6, 2), (1, 10, 5), (1, 10, 5), (1, 14, 7),
                   (1, 10, 7), (1, 14, 7), (1, 10, 7)])
    assert got == want


@needs_ref
def test_planar_section_through_the_shim(tmp_path):
    """mcEnqueueDispatchPlanarSection (config 3's call): the cut BVH is a single leaf."""
    from mcut_b200 import meshgen as mg
    ter = mg.terrain(n=40, extent=100.0, amp=6.0, seed=3)
    dummy = (np.zeros((3, 3)), np.array([0, 1, 2], dtype=np.uint32), None)
    flags = mg.MC_DISPATCH_VERTEX_ARRAY_DOUBLE | mg.MC_DISPATCH_ENFORCE_GENERAL_POSITION
    extra = ["--planar", "0.3", "0.2", "0.93", "0.55"]
    a = run_driver(str(tmp_path), "ref", ter, dummy, flags, [NODUMP], extra)
    b = run_driver(str(tmp_path), "b200", ter, dummy, flags, [SHIM, NODUMP], extra)
    assert int(a["mcDispatch_result"][0]) == int(b["mcDispatch_result"][0]) == 0
    for k in ("cc_type", "cc_nv", "cc_nf", "cc_vertices", "cc_faces"):
        assert a[k].tobytes() == b[k].tobytes(), k
    assert a["cc_type"].size > 0


# ---------------------------------------------------------------------------------------------------------------------
# The whole hot path inside a live mcDispatch: oracle/_ref/libmcut_hooked.so is the reference with the inline narrowphase
# of dispatch() (kernel.cpp:1779-3206) replaced by ONE call to mcb200_hook_narrowphase() (oracle/make_hooked_kernel.py,
# mcut_b200/csrc/shim/mcut_hook.h); broadphase AND narrowphase then run on the B200, everything else is the reference.
# The registry comes back in canonical (edge, face) order while the reference's own order depends on its thread count,
# so intersection vertices may be numbered differently: components are compared as geometry (bit-exact coordinates,
# faces as cyclic vertex sequences), not as index arrays.
# ---------------------------------------------------------------------------------------------------------------------
HOOKED = os.path.join(ROOT, "oracle", "_ref", "api_driver_hooked")
needs_hooked = pytest.mark.skipif(not (os.path.exists(DRIVER) and os.path.exists(HOOKED)),
                                  reason="oracle/_ref/api_driver_hooked is not built (needs /root/reference at build time)")


def canonical_components(out):
    comps = []
    vo = fo = so = 0
    allv = np.ascontiguousarray(out["cc_vertices"]).reshape(-1, 3)
    allf = np.ascontiguousarray(out["cc_faces"]).reshape(-1)
    alls = np.ascontiguousarray(out["cc_face_sizes"]).reshape(-1)
    attrs = np.ascontiguousarray(out["cc_attrs"]).reshape(-1, 3)
    for i in range(out["cc_type"].size):
        nv, nf = int(out["cc_nv"][i]), int(out["cc_nf"][i])
        verts = allv[vo:vo + nv]
        sizes = alls[so:so + nf]
        nidx = int(sizes.sum())
        idx = allf[fo:fo + nidx]
        vo, fo, so = vo + nv, fo + nidx, so + nf
        keys = [v.tobytes() for v in verts]
        faces, o = [], 0
        for n in sizes.tolist():
            cyc = [keys[j] for j in idx[o:o + n].tolist()]
            o += n
            k = min(range(n), key=lambda t: cyc[t:] + cyc[:t])
            faces.append(tuple(cyc[k:] + cyc[:k]))
        comps.append((int(out["cc_type"][i]), tuple(int(x) for x in attrs[i]), nv, tuple(sorted(faces))))
    return sorted(comps)


@needs_hooked
@pytest.mark.parametrize("case", ["hello", "spheres_k16", "patch_vs_sphere", "cube_cube_axis_aligned", "ico_pair", "float_spheres",
                                  "uv12", "terrain_plane"])
def test_mcdispatch_with_device_narrowphase_gives_the_same_components(tmp_path, case):
    src, cut, flags = cases.ALL[case]()
    a = run_driver(str(tmp_path), "ref", src, cut, flags, [NODUMP])
    b = run_driver(str(tmp_path), "hooked", src, cut, flags, [NODUMP], driver=HOOKED)
    assert int(a["mcDispatch_result"][0]) == int(b["mcDispatch_result"][0]), b["_stderr"]
    assert a["cc_type"].size == b["cc_type"].size and a["cc_type"].size > 0
    assert sorted(a["cc_nv"].tolist()) == sorted(b["cc_nv"].tolist()) and sorted(a["cc_nf"].tolist()) == sorted(b["cc_nf"].tolist())
    assert canonical_components(a) == canonical_components(b)
