"""Last-minute GPU check (under a minute): smoke(), the deferred radix order of the registry through the host-array entry
point (more than 16384 records; plain run, graph capture, graph replay), and a small batch."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
t0 = time.time()
import __graft_entry__ as g
g.smoke()
print("smoke ok", round(time.time() - t0, 1), "s", flush=True)
from mcut_b200 import meshgen as mg, stage
from oracle import pyoracle
same = lambda a, b: np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()
ctx = stage.Context(0)
src, cut, flags = mg.c5_coplanar_regions(k=64)
ref = pyoracle.intersect_stage(src, cut, flags)
res = stage.Result(ctx)
for rep in range(4):
    got = stage.intersect_stage_host(ctx, src, cut, flags, res=res)
    assert got["n_records"] == len(ref["records"]) > 16384, (got["n_records"], len(ref["records"]))
    assert same(got["pairs"], ref["pairs"])
    assert same(got["records"]["edge"], ref["records"]["edge"]) and same(got["records"]["face"], ref["records"]["face"]) \
        and same(got["records"]["point"], ref["records"]["point"]), f"registry order, run {rep}"
    assert same(got["cand_normal"], ref["cand_normal"])
print("deferred radix order ok:", got["n_records"], "records", round(time.time() - t0, 1), "s", flush=True)
res.free()
ctx.close()
import test_gpu_batch as tb
pairs = [mg.c4_pair(j, level=3) for j in range(8)]
want = []
for s, c, f in pairs:
    r = pyoracle.intersect_stage(s, c, f)
    want.append((len(r["pairs"]), len(r["tests"]), len(r["records"]) if r["status"] == 0 else None, r["status"]))
for nl in (1, 3):
    out = tb.run_batch(pairs, nl)
    for o, w in zip(out, want):
        assert o[0] == w[0] and o[1] == w[1] and o[4] == w[3] and (w[2] is None or o[3] == w[2]), (nl, o, w)
print("batch ok", round(time.time() - t0, 1), "s", flush=True)
