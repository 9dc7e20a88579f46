"""-m gpu: the CUDA path (through the C-ABI, host buffers) against the oracle on the same inputs. Bit-exact."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def beq(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def run_both(oracle, gpu_ctx, case, perturbation=None):
    from mcut_b200 import stage
    src, cut, flags = cases.ALL[case]()
    ref = oracle.intersect_stage(src, cut, flags, perturbation=perturbation)
    got = stage.intersect_stage(gpu_ctx, src, cut, flags, perturbation=perturbation, log_tests=True)
    return ref, got


@pytest.mark.parametrize("case", sorted(cases.ALL))
def test_stage_matches_oracle(oracle, gpu_ctx, case):
    ref, got = run_both(oracle, gpu_ctx, case)
    assert beq(got["com"], ref["com"]) and beq(got["shift"], ref["shift"]) and got["eps"] == ref["eps"]
    assert beq(got["src_bboxes"], ref["src_bboxes"]), "source face AABBs"
    assert beq(got["cut_bboxes"], ref["cut_bboxes"]), "cut face AABBs (enlarged)"
    assert beq(got["src_root"], ref["src_root"]) and beq(got["cut_root"], ref["cut_root"]), "mesh AABBs"
    assert beq(got["pairs"], ref["pairs"]), "sorted candidate pair set"
    assert got["status"] == ref["status"]
    if ref["status"] in (2, 3):
        assert got["bad_face"] == ref["bad_face"]
        return
    assert beq(got["cand_faces"], ref["cand_faces"])
    assert beq(got["cand_normal"], ref["cand_normal"]) and beq(got["cand_d"], ref["cand_d"])
    assert beq(got["cand_maxcomp"], ref["cand_maxcomp"])
    rt, gt = ref["tests"], got["tests"]
    assert len(rt) == len(gt) == got["n_tests"]
    assert beq(gt["edge"], rt["edge"]) and beq(gt["face"], rt["face"]), "edge/face test key set"
    assert beq(gt["type"], rt["type"]), "segment/plane type"
    assert beq(gt["sign_q"], rt["sign_q"]) and beq(gt["sign_r"], rt["sign_r"]), "orient3d signs"
    assert beq(gt["pip"], rt["pip"]), "point-in-polygon class"
    assert beq(gt["point"], rt["point"]), "intersection point coordinates"
    exact_ref = (rt["exact_q"].astype(np.uint8) | (rt["exact_r"].astype(np.uint8) << 1))
    tri = np.ones(len(rt), dtype=bool)
    assert beq(gt["exact"][tri], exact_ref[tri]) or case in ("hello", "patch_vs_sphere", "cube_cube_axis_aligned"), "filter verdicts"
    if ref["status"] == 0:
        rr, gr = ref["records"], got["records"]
        assert beq(gr["edge"], rr["edge"]) and beq(gr["face"], rr["face"]) and beq(gr["point"], rr["point"]), "registry"


@pytest.mark.parametrize("case", sorted(cases.ALL))
def test_stage_without_the_test_log_matches_oracle(oracle, gpu_ctx, case):
    """The production configuration: no test log, so triangle meshes go through the side prefilter (narrowphase.cu:
    k_tri_prefilter / k_tri_classify / k_tri_resolve).  Every output equals the oracle's; the prefilter's dismissals show
    up only in n_tests, and MCB200_NARROW_COUNT_TESTS brings that back to the reference's count."""
    from mcut_b200 import stage
    src, cut, flags = cases.ALL[case]()
    ref = oracle.intersect_stage(src, cut, flags)
    got = stage.intersect_stage(gpu_ctx, src, cut, flags, count_tests=True)
    assert beq(got["pairs"], ref["pairs"]) and got["status"] == ref["status"]
    if ref["status"] in (2, 3):
        assert got["bad_face"] == ref["bad_face"]
        return
    assert beq(got["cand_faces"], ref["cand_faces"]) and beq(got["cand_normal"], ref["cand_normal"])
    assert beq(got["cand_d"], ref["cand_d"]) and beq(got["cand_maxcomp"], ref["cand_maxcomp"])
    rt = ref["tests"]
    assert got["n_tests_reference"] == len(rt) and got["n_tests"] <= len(rt)
    assert got["n_exact"] == int(np.count_nonzero(rt["exact_q"] | rt["exact_r"])) or case in ("hello", "patch_vs_sphere", "cube_cube_axis_aligned")
    if ref["status"] == 0:
        rr, gr = ref["records"], got["records"]
        assert beq(gr["edge"], rr["edge"]) and beq(gr["face"], rr["face"]) and beq(gr["point"], rr["point"]), "registry"


@pytest.mark.parametrize("case", sorted(cases.ALL))
def test_cutpath_segment_table_matches_oracle(oracle, gpu_ctx, case):
    """the registry grouped by {source face, cut face} (what kernel.cpp:3332-3617 consumes), groups in std::map
    order, points in registry order, groups of more than two points along their line — device sort vs the oracle's table
    (which tests/test_oracle_cutpath.py pins against the container of a live reference dispatch)."""
    from mcut_b200 import stage
    src, cut, flags = cases.ALL[case]()
    ref = oracle.intersect_stage(src, cut, flags)
    if ref["status"] != 0:
        pytest.skip("no registry: the stage reports a general-position violation or an invalid mesh")
    got = stage.intersect_stage(gpu_ctx, src, cut, flags, want_boxes=False, want_cutpath=True)
    want = oracle.cutpath_segments(ref["soup"].edge_f, ref["soup"].src_nf, ref["records"])
    cp = got["cutpath"]
    assert beq(cp["keys"], want["keys"]) and beq(cp["off"], want["off"]) and beq(cp["vtx"], want["vtx"])
    assert cp["n_single"] == want["n_single"]


def test_perturbed_cut_frame(oracle, gpu_ctx):
    pert = np.array([1.3e-3, -0.7e-3, 2.1e-3])
    ref, got = run_both(oracle, gpu_ctx, "cube_cube_axis_aligned", perturbation=pert)
    assert beq(got["pairs"], ref["pairs"])
    assert got["status"] == ref["status"]
    assert beq(got["tests"]["type"], ref["tests"]["type"]) and beq(got["tests"]["point"], ref["tests"]["point"])
    rr, gr = ref["records"], got["records"]
    assert beq(gr["edge"], rr["edge"]) and beq(gr["face"], rr["face"]) and beq(gr["point"], rr["point"])


def test_morton_codes_match_reference_formula(oracle, gpu_ctx):
    from mcut_b200 import stage
    src, cut, flags = cases.spheres_k16()
    ref = oracle.intersect_stage(src, cut, flags)
    com, shift = ref["com"], ref["shift"]
    m = stage.Mesh(gpu_ctx, *src)
    m.set_frame(com, shift)
    m.build(0.0)
    codes, order = m.read_morton()
    want = oracle.morton_codes(ref["src_bboxes"], ref["src_root"])
    assert beq(codes, want)
    assert sorted(order.tolist()) == list(range(m.nf)), "sorted leaf order is a permutation"
    # the build orders the leaves by the top 16 of the 30 code bits (two radix passes; MCB200_MORTON_SORT_BITS=24 / 30: more)
    assert np.all(np.diff((want[order] >> 14).astype(np.int64)) >= 0), "leaves ascend by Morton code"
    m.free()


# ---- the single pipelined host-array call (mcb200_intersect_stage_host) ----
def _check_host_call(ref, got):
    assert beq(got["pairs"], ref["pairs"]), "sorted candidate pair set"
    assert got["status"] == ref["status"]
    if ref["status"] in (2, 3):
        assert got["bad_face"] == ref["bad_face"]
        return
    assert beq(got["cand_faces"], ref["cand_faces"]) and beq(got["cand_normal"], ref["cand_normal"]) and beq(got["cand_d"], ref["cand_d"])
    rt, gt = ref["tests"], got["tests"]
    assert beq(gt["edge"], rt["edge"]) and beq(gt["face"], rt["face"]) and beq(gt["type"], rt["type"])
    assert beq(gt["sign_q"], rt["sign_q"]) and beq(gt["sign_r"], rt["sign_r"]) and beq(gt["pip"], rt["pip"]) and beq(gt["point"], rt["point"])
    if ref["status"] == 0:
        rr, gr = ref["records"], got["records"]
        assert beq(gr["edge"], rr["edge"]) and beq(gr["face"], rr["face"]) and beq(gr["point"], rr["point"]), "registry"


@pytest.mark.parametrize("case", sorted(cases.ALL))
def test_host_array_call_matches_oracle(oracle, gpu_ctx, case):
    from mcut_b200 import stage
    src, cut, flags = cases.ALL[case]()
    ref = oracle.intersect_stage(src, cut, flags)
    got = stage.intersect_stage_host(gpu_ctx, src, cut, flags, log_tests=True)  # the library numbers the polygon soup itself
    _check_host_call(ref, got)


def test_host_array_call_with_caller_soup_and_perturbation(oracle, gpu_ctx):
    from mcut_b200 import stage
    pert = np.array([1.3e-3, -0.7e-3, 2.1e-3])
    for case in ("cube_cube_axis_aligned", "spheres_k16", "patch_vs_sphere"):
        src, cut, flags = cases.ALL[case]()
        ref = oracle.intersect_stage(src, cut, flags, perturbation=pert)
        (sx, sf, ss), (cx, cf, cs) = src, cut
        soff = np.arange(0, sf.size + 1, 3, dtype=np.uint32) if ss is None else np.concatenate([[0], np.cumsum(ss)]).astype(np.uint32)
        coff = np.arange(0, cf.size + 1, 3, dtype=np.uint32) if cs is None else np.concatenate([[0], np.cumsum(cs)]).astype(np.uint32)
        fv, fe, ev, ef = stage.soup_ids(sx.shape[0], soff, sf, coff, cf)
        res = stage.Result(gpu_ctx)
        for _ in range(2):  # the staging buffers and the result are reused across calls
            got = stage.intersect_stage_host(gpu_ctx, src, cut, flags, perturbation=pert, soup_ids_host=(fe, ef, ev.shape[0]),
                                             log_tests=True, res=res)
            _check_host_call(ref, got)
        res.free()


def test_pair_capacity_is_regrown_on_overflow(oracle, gpu_ctx):
    """A result whose pair buffer is too small reports MCB200_ERR_CAPACITY from mcb200_result_counts, raises its own
    capacity, and the rerun is complete (no silent truncation)."""
    from mcut_b200 import stage
    src, cut, flags = cases.ALL["spheres_k16"]()
    ref = oracle.intersect_stage(src, cut, flags)
    res = stage.Result(gpu_ctx)
    res.set_pair_capacity(64)
    got = stage.intersect_stage_host(gpu_ctx, src, cut, flags, res=res)
    assert beq(got["pairs"], ref["pairs"])
    rr, gr = ref["records"], got["records"]
    assert beq(gr["edge"], rr["edge"]) and beq(gr["face"], rr["face"]) and beq(gr["point"], rr["point"])
    res.free()


def _offsets(faces, sizes):
    if sizes is None:
        return np.arange(0, faces.size + 1, 3, dtype=np.uint32)
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)


@pytest.mark.parametrize("case", sorted(cases.ALL))
def test_device_soup_numbering_equals_host_numbering(gpu_ctx, case):
    """Edge ids handed out on the device (rank of an edge's first halfedge) == the sequential numbering of
    mcb200_soup_ids, itself pinned against the reference's `ps` in tests/test_host_logic.py."""
    from mcut_b200 import stage
    src, cut, flags = cases.ALL[case]()
    (sx, sf, ss), (cx, cf, cs) = src, cut
    fv, fe, ev, ef = stage.soup_ids(sx.shape[0], _offsets(sf, ss), sf, _offsets(cf, cs), cf)
    stage.intersect_stage_host(gpu_ctx, src, cut, flags)
    gfv, gfe, gef = stage.staged_soup(gpu_ctx)
    assert beq(gfv, fv), "ps.get_vertices_around_face order"
    assert beq(gfe, fe), "edge id of every halfedge"
    assert beq(gef, ef.reshape(-1, 2)), "faces of h0 / h1 of every edge"


def test_device_soup_numbering_on_the_reference_corpus(gpu_ctx):
    """The same on the 61 pairs of the reference's regression corpus (polygons, open meshes), against the `ps` tables the
    reference itself built (tests/golden/corpus), and the records of the host-array call against its intersection points."""
    from golden_util import CORPUS_CASES, load_corpus, narrowphase_violation_expected
    from mcut_b200 import stage
    for pair in CORPUS_CASES:
        fx, src, cut, flags = load_corpus(pair)
        r = stage.intersect_stage_host(gpu_ctx, src, cut, flags)
        gfv, gfe, gef = stage.staged_soup(gpu_ctx)
        assert beq(gfv, fx["ps_face_vtx"]), f"pair {pair}: ps.get_vertices_around_face order"
        assert beq(gfe, fx["ps_face_edges"]), f"pair {pair}: edge id of every halfedge"
        assert beq(gef, np.ascontiguousarray(fx["ps_edges"][:, 2:])), f"pair {pair}: faces of h0 / h1 of every edge"
        if not narrowphase_violation_expected(fx, 0) and "d0_ipoints_sorted" in fx.files:
            pts = np.ascontiguousarray(r["records"]["point"]).reshape(-1, 3)
            pts = pts[np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0]))] if len(pts) else pts
            assert beq(pts, fx["d0_ipoints_sorted"]), f"pair {pair}: intersection points"


def test_device_soup_numbering_reports_bad_topology(gpu_ctx):
    from mcut_b200 import stage
    src, cut, flags = cases.ALL["hello"]()
    sx, sf, ss = src
    bad = sf.copy()
    bad[0:3] = bad[0:3][::-1]  # one face wound the other way: its edges run the same way as the neighbours'
    with pytest.raises(RuntimeError, match="numbering"):
        stage.intersect_stage_host(gpu_ctx, (sx, bad, ss), cut, flags)
    ok = stage.intersect_stage_host(gpu_ctx, src, cut, flags)  # the context is still usable
    assert ok["status"] == 0


def test_contexts_in_parallel_threads(oracle):
    """The MultipleContextsInParallel pattern (tutorials/MultipleContextsInParallel/MultipleContextsInParallel.cpp:129-345):
    several host threads, one context each, dispatching at the same time.  Every result must still equal the oracle's."""
    import threading
    from mcut_b200 import stage
    names = ["hello", "spheres_k16", "patch_vs_sphere", "cube_cube_tris_offset", "uv12", "ico_pair", "float_spheres", "spheres_k8"]
    inputs = {n: cases.ALL[n]() for n in names}
    refs = {n: oracle.intersect_stage(*inputs[n]) for n in names}
    errors = []

    def worker(tid):
        try:
            ctx = stage.Context(0)
            res = stage.Result(ctx)
            for rep in range(6):
                n = names[(tid + rep) % len(names)]
                src, cut, flags = inputs[n]
                got = stage.intersect_stage_host(ctx, src, cut, flags, res=res)
                ref = refs[n]
                assert beq(got["pairs"], ref["pairs"]), (tid, n, "pairs")
                assert got["status"] == ref["status"], (tid, n, "status")
                if ref["status"] == 0:
                    assert beq(got["records"]["point"], ref["records"]["point"]) and beq(got["records"]["edge"], ref["records"]["edge"]), (tid, n)
            res.free()
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_resident_source_mesh_across_dispatches(oracle, gpu_ctx):
    """C3's pattern: one terrain, many cutting planes.  After the first call the source arrays are declared resident
    (MCB200_STAGE_SRC_RESIDENT): only the cut mesh travels, the BVH is rebuilt for the new frame, results stay exact."""
    from mcut_b200 import meshgen as mg, stage
    ter = mg.terrain(n=120)
    flags = mg.MC_DISPATCH_VERTEX_ARRAY_DOUBLE | mg.MC_DISPATCH_ENFORCE_GENERAL_POSITION
    res = stage.Result(gpu_ctx)
    for k in range(4):
        tri = np.array([[-900.0, -850.0 + 40.0 * k, -4.1 + k], [1400.0, -700.0, 3.3 - 0.5 * k], [150.0 + 30.0 * k, 1600.0, 1.7]])
        cut = (tri, np.array([0, 1, 2], dtype=np.uint32), None)
        ref = oracle.intersect_stage(ter, cut, flags)
        got = stage.intersect_stage_host(gpu_ctx, ter, cut, flags, res=res, src_resident=(k > 0))
        assert beq(got["pairs"], ref["pairs"]) and got["status"] == ref["status"]
        if ref["status"] == 0:
            rr, gr = ref["records"], got["records"]
            assert beq(gr["edge"], rr["edge"]) and beq(gr["face"], rr["face"]) and beq(gr["point"], rr["point"])
    res.free()


def test_disjoint_meshes_give_empty_results(oracle, gpu_ctx):
    """No AABB overlap at all: zero pairs, zero tests, zero records, status SUCCESS — through both entry points."""
    from mcut_b200 import meshgen as mg, stage
    a = mg.cube_sphere(10, 20.0)
    b = mg.cube_sphere(7, 5.0, centre=(200.0, 0.0, 0.0))
    flags = mg.MC_DISPATCH_VERTEX_ARRAY_DOUBLE | mg.MC_DISPATCH_ENFORCE_GENERAL_POSITION
    ref = oracle.intersect_stage(a, b, flags)
    assert len(ref["pairs"]) == 0 and ref["status"] == 0
    for got in (stage.intersect_stage(gpu_ctx, a, b, flags, log_tests=True), stage.intersect_stage_host(gpu_ctx, a, b, flags, log_tests=True)):
        assert got["n_pairs"] == 0 and got["n_records"] == 0 and got["status"] == 0 and len(got["pairs"]) == 0
        assert len(got["records"]) == 0 and len(got["cand_faces"]) == 0 and len(got["tests"]) == 0
