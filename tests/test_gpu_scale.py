"""-m gpu: BASELINE.json sizes.  The oracle port finishes these in seconds, so the comparison stays bit-exact; on top
come size-independent properties (role swap, idempotence, shard union)."""
import numpy as np
import pytest

from mcut_b200 import meshgen as mg

pytestmark = pytest.mark.gpu


def beq(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def compare(oracle, ctx, src, cut, flags, perturbation=None):
    from mcut_b200 import stage
    ref = oracle.intersect_stage(src, cut, flags, perturbation=perturbation)
    got = stage.intersect_stage(ctx, src, cut, flags, perturbation=perturbation, want_boxes=False, count_tests=True)
    assert beq(got["pairs"], ref["pairs"]), "sorted candidate pair set"
    assert got["status"] == ref["status"]
    assert got["n_tests_reference"] == len(ref["tests"]) and got["n_tests"] <= len(ref["tests"])
    assert got["n_exact"] == int(np.count_nonzero(ref["tests"]["exact_q"] | ref["tests"]["exact_r"]))
    if ref["status"] == 0:
        rr, gr = ref["records"], got["records"]
        assert beq(gr["edge"], rr["edge"]) and beq(gr["face"], rr["face"]) and beq(gr["point"], rr["point"]), "registry"
    assert beq(got["cand_faces"], ref["cand_faces"]) and beq(got["cand_normal"], ref["cand_normal"])
    return ref, got


def test_c2_two_spheres_1m(oracle, gpu_ctx):
    src, cut, flags = mg.c2_two_spheres(k=289)
    ref, got = compare(oracle, gpu_ctx, src, cut, flags)
    assert got["n_pairs"] == 34464 and got["n_records"] == 5100  # == the reference's own run (harness)


def test_c2_cutpath_segment_table(oracle, gpu_ctx):
    """f4 at BASELINE size: 5,100 registry records -> 5,100 face pairs of two points each (one closed intersection curve)."""
    from mcut_b200 import stage
    src, cut, flags = mg.c2_two_spheres(k=289)
    ref = oracle.intersect_stage(src, cut, flags)
    got = stage.intersect_stage(gpu_ctx, src, cut, flags, want_boxes=False, want_cutpath=True)
    want = oracle.cutpath_segments(ref["soup"].edge_f, ref["soup"].src_nf, ref["records"])
    cp = got["cutpath"]
    assert beq(cp["keys"], want["keys"]) and beq(cp["off"], want["off"]) and beq(cp["vtx"], want["vtx"])
    assert cp["keys"].size == 5100 and cp["n_single"] == 0 and np.all(np.diff(cp["off"]) == 2)


def test_c3_terrain_4m_one_plane(oracle, gpu_ctx):
    """One of C3's 256 dispatches: 3,998,792-triangle terrain vs a single huge triangle (a one-leaf cut BVH)."""
    ter = mg.terrain()
    tri = np.array([[-900.0, -850.0, -4.1], [1400.0, -700.0, 3.3], [150.0, 1600.0, 1.7]])
    cut = (tri, np.array([0, 1, 2], dtype=np.uint32), None)
    compare(oracle, gpu_ctx, ter, cut, mg.MC_DISPATCH_VERTEX_ARRAY_DOUBLE | mg.MC_DISPATCH_ENFORCE_GENERAL_POSITION)


def test_c4_small_pairs(oracle, gpu_ctx):
    for j in range(6):
        src, cut, flags = mg.c4_pair(j)
        compare(oracle, gpu_ctx, src, cut, flags)


def test_c5_near_coplanar_regions_2m(oracle, gpu_ctx):
    """BASELINE config 5 at full size (2 x 2,007,372 triangles): dense shallow overlap + regions where the cutter lies within
    the resolution of the stage-A orient3d filter of the source's own faces.  More than a million tests go through the
    exact-expansion kernel; pairs, every test count, every record (edge, face, point) equal the oracle bit for bit."""
    src, cut, flags = mg.c5_coplanar_regions(k=409)
    ref, got = compare(oracle, gpu_ctx, src, cut, flags)
    assert got["n_pairs"] == 21547246 and got["n_tests_reference"] == 37823783
    assert got["n_exact"] == 1037662 and got["n_exact"] >= 100000
    assert got["status"] == 0 and got["n_records"] == 749808


def test_c5_dense_overlap_without_exact_tests(oracle, gpu_ctx):
    """SURVEY's original C5 recipe at a quarter of the size: the same dense overlap, no test reaches the exact stages."""
    src, cut, flags = mg.c5_near_coplanar(k=204)
    ref, got = compare(oracle, gpu_ctx, src, cut, flags)
    assert got["n_pairs"] > 100000 and got["n_exact"] == 0


def test_role_swap_gives_transposed_pairs(gpu_ctx):
    """Swapping source and cut (eps = 0 on both sides) must give the transposed pair set: the traversal picks its query
    side by size, so this exercises both orientations."""
    from mcut_b200 import stage
    a = mg.cube_sphere(40, 20.0)
    b = mg.cube_sphere(23, 20.0, rotation=mg.rot_z(0.3), centre=(11.0, 2.0, 1.0))
    ctx = gpu_ctx

    def pairs(src, cut):
        ms, mc = stage.Mesh(ctx, *src), stage.Mesh(ctx, *cut)
        ms.build(0.0)
        mc.build(0.0)
        res = stage.Result(ctx)
        ctx.check(ctx.L.mcb200_bvh_intersect(ctx.h, ms.h, mc.h, res.h))
        p = res.pairs()
        for o in (res, ms, mc):
            o.free()
        return p

    ab, ba = pairs(a, b), pairs(b, a)
    swapped = np.sort((ba << np.uint64(32)) | (ba >> np.uint64(32)))
    assert ab.size > 0 and beq(ab, swapped)


def test_shards_partition_the_pair_set(gpu_ctx):
    """traversing the query leaves in nparts round-robin slices yields disjoint pair sets whose union is the
    unsharded set (what the NCCL all-gather reassembles)."""
    from mcut_b200 import stage
    src, cut, flags = mg.c2_two_spheres(k=64)
    ctx = gpu_ctx
    ms, mc = stage.Mesh(ctx, *src), stage.Mesh(ctx, *cut)
    com, shift, sbb, cbb = stage.vertex_parameters(src[0], cut[0])
    ms.set_frame(com, shift)
    mc.set_frame(com, shift)
    ms.build(0.0)
    mc.build(stage.cut_bbox_eps(cbb))
    res = stage.Result(ctx)
    ctx.check(ctx.L.mcb200_bvh_intersect(ctx.h, ms.h, mc.h, res.h))
    full = res.pairs()
    parts = []
    for part in range(3):
        res.set_shard(part, 3, 256)
        ctx.check(ctx.L.mcb200_bvh_intersect(ctx.h, ms.h, mc.h, res.h))
        parts.append(res.pairs())
    assert sum(p.size for p in parts) == full.size
    assert beq(np.sort(np.concatenate(parts)), full)
    for o in (res, ms, mc):
        o.free()


def test_c2_through_the_single_host_array_call(oracle, gpu_ctx):
    """BASELINE size through mcb200_intersect_stage_host: pipelined uploads, polygon soup numbered on the device, query-only
    build of the larger mesh.  Same pairs, same registry; and a second call on the same context reuses every buffer."""
    from mcut_b200 import stage
    src, cut, flags = mg.c2_two_spheres(k=289)
    ref = oracle.intersect_stage(src, cut, flags)
    res = stage.Result(gpu_ctx)
    for _ in range(2):
        got = stage.intersect_stage_host(gpu_ctx, src, cut, flags, res=res)
        assert got["n_pairs"] == 34464 and got["n_records"] == 5100
        assert beq(got["pairs"], ref["pairs"])
        rr, gr = ref["records"], got["records"]
        assert beq(gr["edge"], rr["edge"]) and beq(gr["face"], rr["face"]) and beq(gr["point"], rr["point"])
        assert beq(got["cand_faces"], ref["cand_faces"]) and beq(got["cand_normal"], ref["cand_normal"])
    res.free()


def test_query_only_mesh_is_completed_on_demand(oracle, gpu_ctx):
    """mcb200_intersect_stage builds the larger mesh query-only (groups, no node records).  Using that mesh afterwards as
    the TREE side of mcb200_bvh_intersect must still work: the library completes its build first."""
    from mcut_b200 import stage
    big = mg.cube_sphere(30, 20.0)
    small = mg.cube_sphere(12, 20.0, rotation=mg.rot_z(0.2), centre=(15.0, 1.0, 2.0))
    huge = mg.cube_sphere(41, 20.0, rotation=mg.rot_x(0.1), centre=(-14.0, 3.0, 1.0))
    ctx = gpu_ctx
    mb, ms, mh = stage.Mesh(ctx, *big), stage.Mesh(ctx, *small), stage.Mesh(ctx, *huge)
    soup = stage.Soup(ctx, mb, ms)
    res = stage.Result(ctx)
    ctx.check(ctx.L.mcb200_intersect_stage(ctx.h, mb.h, ms.h, 0.0, soup.h, res.h, 0))  # `big` is the query side here
    n1 = res.counts().n_pairs
    res2 = stage.Result(ctx)
    mh.build(0.0)
    ctx.check(ctx.L.mcb200_bvh_intersect(ctx.h, mb.h, mh.h, res2.h))  # now `big` is the tree side (huge has more faces)
    got = res2.pairs()
    # reference answer from fully built meshes
    mb2, mh2 = stage.Mesh(ctx, *big), stage.Mesh(ctx, *huge)
    mb2.build(0.0)
    mh2.build(0.0)
    res3 = stage.Result(ctx)
    ctx.check(ctx.L.mcb200_bvh_intersect(ctx.h, mb2.h, mh2.h, res3.h))
    assert n1 > 0 and got.size > 0 and beq(got, res3.pairs())
    for o in (res, res2, res3, soup, mb, ms, mh, mb2, mh2):
        o.free()
