"""-m gpu: mcb200_batch_intersect_host — many small dispatches over several context lanes of one GPU (BASELINE config 4, the
MultipleContextsInParallel pattern).  Every item's counts and status must equal the oracle's for that pair, whatever lane
it ran on and whether its stage body was issued launch by launch or replayed from a CUDA graph."""
import ctypes as C

import numpy as np
import pytest

from mcut_b200 import meshgen as mg
from mcut_b200._lib import BatchItem, Counts, HostMesh

pytestmark = pytest.mark.gpu


def run_batch(pairs, nlanes, device=0):
    from mcut_b200 import stage
    lanes = [stage.Context(device) for _ in range(nlanes)]
    res = [stage.Result(c) for c in lanes]
    L = lanes[0].L
    vp = C.c_void_p
    ctx_arr = (vp * nlanes)(*[c.h for c in lanes])
    res_arr = (vp * nlanes)(*[r.h for r in res])
    items = (BatchItem * len(pairs))()
    keep = []
    for k, (src, cut, flags) in enumerate(pairs):
        sx, sf = np.ascontiguousarray(src[0]), np.ascontiguousarray(src[1], dtype=np.uint32)
        cx, cf = np.ascontiguousarray(cut[0]), np.ascontiguousarray(cut[1], dtype=np.uint32)
        keep += [sx, sf, cx, cf]
        items[k].src = HostMesh(0, sx.ctypes.data, sx.shape[0], sf.ctypes.data, None, sf.size // 3)
        items[k].cut = HostMesh(0, cx.ctypes.data, cx.shape[0], cf.ctypes.data, None, cf.size // 3)
        items[k].com = None
        items[k].gp_constant = 1e-4
        items[k].flags = 8  # MCB200_NARROW_COUNT_TESTS: n_tests as the reference counts them
    counts = (Counts * len(pairs))()
    rc = L.mcb200_batch_intersect_host(ctx_arr, res_arr, nlanes, items, len(pairs), counts)
    assert rc == 0, L.mcb200_last_error(lanes[0].h).decode()
    out = [(int(c.n_pairs), int(c.n_tests), int(c.n_exact), int(c.n_records), int(c.status)) for c in counts]
    for r in res:
        r.free()
    for c in lanes:
        c.close()
    return out


def test_batch_of_small_pairs_matches_the_oracle(oracle):
    pairs = [mg.c4_pair(j, level=3) for j in range(14)]  # 1,280-triangle icospheres: same sizes -> the graph is replayed
    want = []
    for src, cut, flags in pairs:
        r = oracle.intersect_stage(src, cut, flags)
        t = r["tests"]
        want.append((len(r["pairs"]), len(t), int(np.count_nonzero(t["exact_q"] | t["exact_r"])), len(r["records"]) if r["status"] == 0 else None,
                     r["status"]))
    for nlanes in (1, 3):
        got = run_batch(pairs, nlanes)
        for g, w in zip(got, want):
            assert g[0] == w[0] and g[1] == w[1] and g[2] == w[2] and g[4] == w[4], (nlanes, g, w)
            if w[3] is not None:
                assert g[3] == w[3]


def test_batch_with_mixed_sizes(oracle):
    """different mesh sizes on the same lane: graphs of other signatures must not be replayed (and buffers that grow
    invalidate the captured ones)"""
    pairs = [mg.c4_pair(0, level=2), mg.c2_two_spheres(k=12), mg.c4_pair(1, level=2), mg.c2_two_spheres(k=20), mg.c4_pair(2, level=2),
             mg.c2_two_spheres(k=12), mg.c4_pair(3, level=3), mg.c4_pair(4, level=2)]
    got = run_batch(pairs, 2)
    for (src, cut, flags), g in zip(pairs, got):
        r = oracle.intersect_stage(src, cut, flags)
        assert g[0] == len(r["pairs"]) and g[1] == len(r["tests"]) and g[4] == r["status"]
        if r["status"] == 0:
            assert g[3] == len(r["records"])
