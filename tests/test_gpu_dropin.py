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


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own regression corpus (tests/source/benchmark.cpp: pairs 000..060) through the live drop-in.  27 of the
# pairs make the reference repartition a face (floating-polygon resolution) and rebuild a BVH from a half-edge mesh with
# a history: the shim's generic path (flattened hmesh, in/out face boxes) is what runs there.
# ---------------------------------------------------------------------------------------------------------------------
from golden_util import CORPUS_CASES, load_corpus  # noqa: E402

_corpus_runs = {}


def _corpus_run(tmp_path_factory, mode, pair):
    """One driver process per pair and mode ("ref", "shim", "hooked").  A process spends a second or two creating its CUDA
    context, so the first request of a mode launches ALL its pairs eight at a time; the tests then look their result up."""
    if mode not in _corpus_runs:
        from concurrent.futures import ThreadPoolExecutor
        base = str(tmp_path_factory.mktemp(f"corpus_{mode}"))

        def one(p):
            _, src, cut, flags = load_corpus(p)
            d = os.path.join(base, f"{p:03d}")
            os.makedirs(d, exist_ok=True)
            try:
                if mode == "ref":
                    return run_driver(d, "ref", src, cut, flags, [NODUMP])
                if mode == "shim":
                    return run_driver(d, "b200", src, cut, flags, [SHIM, NODUMP])
                return run_driver(d, "hooked", src, cut, flags, [NODUMP], driver=HOOKED)
            except BaseException as e:  # surfaces in the test of that pair
                return e

        with ThreadPoolExecutor(max_workers=8) as ex:
            _corpus_runs[mode] = dict(zip(CORPUS_CASES, ex.map(one, CORPUS_CASES)))
    r = _corpus_runs[mode][pair]
    if isinstance(r, BaseException):
        raise r
    return r


@needs_ref
@pytest.mark.parametrize("pair", CORPUS_CASES)
def test_corpus_mcdispatch_with_shim_is_bit_identical(tmp_path_factory, pair):
    fx, src, cut, flags = load_corpus(pair)
    a = _corpus_run(tmp_path_factory, "ref", pair)
    b = _corpus_run(tmp_path_factory, "shim", pair)
    assert int(a["mcDispatch_result"][0]) == int(b["mcDispatch_result"][0]) == int(fx["mcDispatch_result"][0]), b["_stderr"]
    for k in ("cc_type", "cc_attrs", "cc_nv", "cc_nf", "cc_vertices", "cc_faces", "cc_face_sizes"):
        assert a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes(), k
    assert sorted(zip(a["cc_type"].tolist(), a["cc_nv"].tolist(), a["cc_nf"].tolist())) == \
        sorted(zip(fx["cc_type"].tolist(), fx["cc_nv"].tolist(), fx["cc_nf"].tolist())), "the committed fixture"


def split_components(out):
    """[(type, attrs, distinct positions [n,3], faces as tuples of indices into them)] per connected component."""
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
        uniq, inv = np.unique(verts, axis=0, return_inverse=True)  # sealed fragments may hold a seam vertex twice
        inv = inv.reshape(-1)
        faces, o = [], 0
        for n in sizes.tolist():
            faces.append(tuple(int(inv[j]) for j in idx[o:o + n]))
            o += n
        comps.append((int(out["cc_type"][i]), tuple(int(x) for x in attrs[i]), uniq, faces))
    return comps


def components_equivalent(a, b, tol):
    """Same components as geometry: every component of `a` has a partner in `b` with the same type and attributes whose
    distinct vertex positions pair up one to one within `tol` (0: bit-exact) and whose faces are the same cyclic vertex
    sequences under that pairing."""
    ca, cb = split_components(a), split_components(b)
    if len(ca) != len(cb):
        return False
    used = set()

    def canon(faces, rename):
        out = []
        for f in faces:
            g = [rename[v] for v in f]
            k = min(range(len(g)), key=lambda t: g[t:] + g[:t])
            out.append(tuple(g[k:] + g[:k]))
        return sorted(out)

    for ty, at, va, fa in ca:
        ok = False
        for j, (ty2, at2, vb, fb) in enumerate(cb):
            if j in used or ty2 != ty or at2 != at or va.shape != vb.shape or len(fa) != len(fb):
                continue
            d = np.abs(va[:, None, :] - vb[None, :, :]).max(axis=2)  # tiny meshes: all pairs
            near = d.argmin(axis=1)
            if len(set(near.tolist())) != len(near) or not np.all(d[np.arange(len(near)), near] <= tol):
                continue
            if canon(fa, {i: int(near[i]) for i in range(len(near))}) == canon(fb, {i: i for i in range(vb.shape[0])}):
                used.add(j)
                ok = True
                break
        if not ok:
            return False
    return True


@needs_hooked
@pytest.mark.parametrize("pair", CORPUS_CASES)
def test_corpus_mcdispatch_with_device_narrowphase(tmp_path_factory, pair):
    """Hooked dispatch() on the corpus: broadphase AND narrowphase on the device inside a live mcDispatch.  The hook hands
    the registry over in the reference's own order (it replays the insertion sequence of the reference's
    std::unordered_map<ed_t, ...>, kernel.cpp:1781-1852, and the block order of its parallel_for, :2415-2868), so the
    intersection vertices get the reference's numbers and EVERYTHING downstream — floating-polygon partition segments
    included (27 of the pairs) — is bit-identical: every output array of every connected component, all 61 pairs."""
    fx, src, cut, flags = load_corpus(pair)
    a = _corpus_run(tmp_path_factory, "ref", pair)
    b = _corpus_run(tmp_path_factory, "hooked", pair)
    assert int(a["mcDispatch_result"][0]) == int(b["mcDispatch_result"][0]), b["_stderr"]
    assert a["cc_type"].size == b["cc_type"].size and a["cc_type"].size > 0
    for k in ("cc_type", "cc_attrs", "cc_nv", "cc_nf", "cc_vertices", "cc_faces", "cc_face_sizes"):
        assert a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes(), k
    assert components_equivalent(a, b, 0.0)


@needs_hooked
@pytest.mark.parametrize("helpers", [0, 3])
def test_hooked_registry_order_follows_the_helper_count(tmp_path, helpers):
    """Above 1024 candidate faces the reference's parallel_for cuts its maps into blocks, so its registry order depends on
    the helper-thread count; the hook is told the count and reproduces either order bit for bit."""
    src, cut, flags = cases.ALL["spheres_k64"]()
    extra = ["--helpers", str(helpers)]
    a = run_driver(str(tmp_path), "ref", src, cut, flags, [NODUMP], extra)
    b = run_driver(str(tmp_path), "hooked", src, cut, flags, [NODUMP], extra, driver=HOOKED)
    assert int(a["mcDispatch_result"][0]) == int(b["mcDispatch_result"][0]) == 0, b["_stderr"]
    for k in ("cc_type", "cc_attrs", "cc_nv", "cc_nf", "cc_vertices", "cc_faces", "cc_face_sizes"):
        assert a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes(), k


@needs_hooked
@pytest.mark.parametrize("case", ["degenerate_edge_edge", "degenerate_face_vertex", "degenerate_zero_area"])
def test_degenerate_inputs_are_rejected_like_the_reference(tmp_path, case):
    """tests/source/degenerateInput.cpp:58-144: float input, no general-position enforcement -> MC_INVALID_OPERATION (-2),
    from the narrowphase's general-position classification (edge/edge, vertex on face) or its degenerate-face rule."""
    src, cut, flags = cases.ALL[case]()
    a = run_driver(str(tmp_path), "ref", src, cut, flags, [NODUMP])
    b = run_driver(str(tmp_path), "b200", src, cut, flags, [SHIM, NODUMP])
    c = run_driver(str(tmp_path), "hooked", src, cut, flags, [NODUMP], driver=HOOKED)
    assert int(a["mcDispatch_result"][0]) == int(b["mcDispatch_result"][0]) == int(c["mcDispatch_result"][0]) == -2, c["_stderr"]


@needs_hooked
@pytest.mark.parametrize("case", ["spheres_k16", "hello"])
def test_contexts_in_parallel_through_the_drop_in(tmp_path, case):
    """The MultipleContextsInParallel pattern (tutorials/MultipleContextsInParallel, tests/source/
    concurrentSynchronizedContexts.cpp; BASELINE config 4): six MCUT contexts dispatch the same input from six threads at the
    same time.  The adapter gives every API thread its own device context; all six must return the same components, equal
    to the reference's."""
    src, cut, flags = cases.ALL[case]()
    extra = ["--contexts", "6"]
    a = run_driver(str(tmp_path), "ref", src, cut, flags, [NODUMP])
    b = run_driver(str(tmp_path), "b200", src, cut, flags, [SHIM, NODUMP], extra)
    c = run_driver(str(tmp_path), "hooked", src, cut, flags, [NODUMP], extra, driver=HOOKED)
    for o in (b, c):
        assert int(o["contexts"][0]) == 6 and int(o["contexts_identical"][0]) == 1, o["_stderr"]
        assert o["contexts_results"].tolist() == [0] * 6
    for k in ("cc_type", "cc_attrs", "cc_nv", "cc_nf", "cc_vertices", "cc_faces", "cc_face_sizes"):
        assert a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes(), k
    assert components_equivalent(a, c, 0.0)


# ---- input validation answered from the device (check_input_mesh, mesh_is_closed, the intersection-type
# verdict are PLT-called by preproc() and interposed by the adapter) ----
ITYPE_FLAG = 1 << 17  # MC_DISPATCH_INCLUDE_INTERSECTION_TYPE (mcut.h:440)


def _two_spheres(k, r_src, r_cut, centre_cut):
    from mcut_b200 import meshgen as mg
    src = mg.cube_sphere(k, r_src)
    cut = mg.cube_sphere(k, r_cut, rotation=mg.rot_z(0.2), centre=centre_cut)
    return (src[0], src[1], None), (cut[0], cut[1], None)


@needs_ref
@pytest.mark.parametrize("name,geometry", [
    ("cut_inside_src", (8, 10.0, 3.0, (0.5, 0.2, -0.3))),  # both watertight, no contact: INSIDE_SOURCEMESH, two winding numbers asked
    ("src_inside_cut", (8, 3.0, 10.0, (0.5, 0.2, -0.3))),
    ("apart", (8, 3.0, 3.0, (20.0, 1.0, 0.0))),  # the BVHs do not overlap: preproc.cpp:2891
    ("boxes_overlap_surfaces_apart", (8, 3.0, 0.3, (2.6, 2.6, 2.6))),  # inside the source's AABB, outside the sphere
    ("crossing", (8, 3.0, 3.0, (2.0, 0.5, 0.0))),  # STANDARD
])
def test_input_checks_and_intersection_type_from_the_device(tmp_path, name, geometry):
    from mcut_b200 import meshgen as mg
    src, cut = _two_spheres(*geometry)
    flags = mg.MC_DISPATCH_VERTEX_ARRAY_DOUBLE | mg.MC_DISPATCH_ENFORCE_GENERAL_POSITION | ITYPE_FLAG
    a = run_driver(str(tmp_path), "ref", src, cut, flags, [NODUMP])
    os.environ["MCB200_SHIM_TIMING"] = "1"
    try:
        b = run_driver(str(tmp_path), "b200", src, cut, flags, [SHIM, NODUMP])
        os.environ["MCB200_SHIM_HOST_CHECKS"] = "1"
        c = run_driver(str(tmp_path), "b200_host_checks", src, cut, flags, [SHIM, NODUMP])
    finally:
        os.environ.pop("MCB200_SHIM_TIMING", None)
        os.environ.pop("MCB200_SHIM_HOST_CHECKS", None)
    for got in (b, c):
        assert int(a["mcDispatch_result"][0]) == int(got["mcDispatch_result"][0]) == 0
        assert int(a["intersection_type"][0]) == int(got["intersection_type"][0]), name
        for k in ("cc_type", "cc_attrs", "cc_nv", "cc_nf", "cc_vertices", "cc_faces", "cc_face_sizes"):
            assert a[k].shape == got[k].shape and a[k].tobytes() == got[k].tobytes(), k
    # the device really answered: the adapter's timers name the three interposed functions
    err = b["_stderr"]
    assert "check_input_mesh (device)" in err and "mesh_is_closed (device)" in err
    if name != "crossing":
        assert "intersection type (device)" in err
    assert "(device)" not in c["_stderr"]


@needs_ref
def test_two_components_are_rejected_by_the_device_check(tmp_path):
    """check_input_mesh: "Detected multiple connected components in mesh" (preproc.cpp:541-550) -> mcDispatch fails the same way
    with the component count coming from the device."""
    from mcut_b200 import meshgen as mg
    s1 = mg.cube_sphere(6, 2.0)
    s2 = mg.cube_sphere(6, 2.0, centre=(10.0, 0.0, 0.0))
    src = (np.concatenate([s1[0], s2[0]]), np.concatenate([s1[1], s2[1] + np.uint32(s1[0].shape[0])]).astype(np.uint32), None)
    c = mg.cube_sphere(6, 2.0, centre=(1.0, 0.5, 0.0))
    cut = (c[0], c[1], None)
    flags = mg.MC_DISPATCH_VERTEX_ARRAY_DOUBLE
    a = run_driver(str(tmp_path), "ref", src, cut, flags, [NODUMP])
    os.environ["MCB200_SHIM_TIMING"] = "1"
    try:
        b = run_driver(str(tmp_path), "b200", src, cut, flags, [SHIM, NODUMP])
    finally:
        os.environ.pop("MCB200_SHIM_TIMING", None)
    assert int(a["mcDispatch_result"][0]) == int(b["mcDispatch_result"][0]) != 0
    assert "check_input_mesh (device)" in b["_stderr"]
