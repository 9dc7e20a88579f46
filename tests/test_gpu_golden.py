"""-m gpu: the CUDA path against the golden vectors recorded from the unmodified reference (no oracle in between),
including the perturbation retries the reference went through."""
import numpy as np
import pytest

import cases
from golden_util import (CORPUS_CASES, STAGE_CASES, beq, check_tests_against_fixture, load_corpus, load_stage,
                         narrowphase_violation_expected, replay_inputs)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", STAGE_CASES)
def test_cuda_stage_matches_reference_fixture(gpu_ctx, case):
    fx = load_stage(case)
    src, cut, flags = cases.ALL[case]()
    check_cuda_against_fixture(gpu_ctx, fx, src, cut, flags)


@pytest.mark.parametrize("pair", CORPUS_CASES)
def test_cuda_stage_on_reference_regression_corpus(gpu_ctx, pair):
    """The reference's own regression corpus (tests/source/benchmark.cpp, pairs 000..060), every kernel invocation the
    reference went through: perturbation retries and the retries on meshes its floating-polygon resolution repartitioned
    (replayed with the in/out face boxes and the polygon-soup tables that invocation had)."""
    fx, src, cut, flags = load_corpus(pair)
    check_cuda_against_fixture(gpu_ctx, fx, src, cut, flags)


def check_cuda_against_fixture(gpu_ctx, fx, src, cut, flags):
    from mcut_b200 import stage
    for k in range(int(fx["n_dispatch"][0])):
        kw = replay_inputs(fx, k, src, cut)
        r = stage.intersect_stage(gpu_ctx, flags=flags, log_tests=True, **kw)
        if "params" in kw:
            assert beq(r["src_bboxes"], fx[f"d{k}_src_bboxes"]) and beq(r["cut_bboxes"], fx[f"d{k}_cut_bboxes"]), "face AABBs"
            assert beq(r["pairs"], fx[f"d{k}_pairs"]), "candidate pair set of the retry on the repartitioned mesh"
        else:
            assert beq(r["com"], fx["com"]) and beq(r["shift"], fx["shift"]) and r["eps"] == float(fx["eps"][0])
            assert beq(r["src_bboxes"], fx["src_bboxes"]) and beq(r["cut_bboxes"], fx["cut_bboxes"]), "face AABBs"
            assert beq(r["src_root"], fx["src_root"]) and beq(r["cut_root"], fx["cut_root"]), "mesh AABBs"
            assert beq(r["pairs"], fx["pairs"]), "candidate pair set"
        raw = int(fx[f"d{k}_status_raw"][0])
        if raw in (-1, -2):  # status_t::INVALID_SRC_MESH / INVALID_CUT_MESH (kernel.cpp:2237-2244): a degenerate candidate face
            assert r["status"] == (stage.STATUS_INVALID_SRC_MESH if raw == -1 else stage.STATUS_INVALID_CUT_MESH)
            continue
        assert beq(r["cand_faces"], fx[f"d{k}_plane_faces"])
        assert beq(r["cand_normal"], fx[f"d{k}_plane_normal"]) and beq(r["cand_d"], fx[f"d{k}_plane_d"])
        assert beq(r["cand_maxcomp"], fx[f"d{k}_plane_mc"])
        violated = narrowphase_violation_expected(fx, k)
        assert (r["status"] == stage.STATUS_GENERAL_POSITION_VIOLATION) == violated
        check_tests_against_fixture(fx, k, r["tests"], complete=not violated)
        if not violated and f"d{k}_ipoints_sorted" in fx.files:
            pts = np.ascontiguousarray(r["records"]["point"]).reshape(-1, 3)
            pts = pts[np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0]))] if len(pts) else pts
            assert beq(pts, fx[f"d{k}_ipoints_sorted"]), "intersection points = the reference's m0 vertices"


@pytest.mark.parametrize("pair", CORPUS_CASES)
def test_cutpath_segment_table_on_reference_regression_corpus(oracle, gpu_ctx, pair):
    """SURVEY §8-f4 on the corpus (polygon faces: face pairs with more than two intersection points occur): the device's
    segment table of the last kernel invocation of every pair equals the oracle's."""
    from mcut_b200 import stage
    fx, src, cut, flags = load_corpus(pair)
    k = int(fx["n_dispatch"][0]) - 1
    kw = replay_inputs(fx, k, src, cut)
    ref = oracle.intersect_stage(flags=flags, **kw)
    if ref["status"] != 0:
        pytest.skip("the last invocation of this pair has no registry")
    got = stage.intersect_stage(gpu_ctx, flags=flags, want_boxes=False, want_cutpath=True, **kw)
    want = oracle.cutpath_segments(ref["soup"].edge_f, ref["soup"].src_nf, ref["records"])
    cp = got["cutpath"]
    assert beq(cp["keys"], want["keys"]) and beq(cp["off"], want["off"]) and beq(cp["vtx"], want["vtx"])
    assert cp["n_single"] == want["n_single"]


def test_degenerate_candidate_face_reports_invalid_mesh(oracle, gpu_ctx):
    from mcut_b200 import stage
    from test_oracle_stage import degenerate_case
    src, cut, flags, _ = degenerate_case(oracle)
    ref = oracle.intersect_stage(src, cut, flags)
    got = stage.intersect_stage(gpu_ctx, src, cut, flags)
    assert ref["status"] == 2 and got["status"] == stage.STATUS_INVALID_SRC_MESH and got["bad_face"] == ref["bad_face"]


def test_fused_stage_call_equals_separate_calls(gpu_ctx):
    """mcb200_intersect_stage (two lanes, no host round trips) == build + intersect + narrowphase called one by one."""
    from mcut_b200 import stage
    src, cut, flags = cases.spheres_k64()
    a = stage.intersect_stage(gpu_ctx, src, cut, flags, want_boxes=False)
    ctx = gpu_ctx
    com, shift, sbb, cbb = stage.vertex_parameters(src[0], cut[0])
    eps = stage.cut_bbox_eps(cbb)
    ms, mc = stage.Mesh(ctx, *src), stage.Mesh(ctx, *cut)
    ms.set_frame(com, shift)
    mc.set_frame(com, shift)
    soup = stage.Soup(ctx, ms, mc)
    res = stage.Result(ctx)
    for _ in range(2):  # twice: buffers are reused, results must not change (idempotence)
        ctx.check(ctx.L.mcb200_intersect_stage(ctx.h, ms.h, mc.h, eps, soup.h, res.h, 0))
        assert beq(res.pairs(), a["pairs"])
        rec = res.records()
        assert beq(rec["edge"], a["records"]["edge"]) and beq(rec["face"], a["records"]["face"])
        assert beq(rec["point"], a["records"]["point"])
    for o in (soup, res, ms, mc):
        o.free()
