"""CPU: the oracle's cut-path segment table (mco_cutpath_segments) against what a live dispatch of the
reference consumes.

oracle/_ref/api_driver_hooked_cpu runs the reference with its narrowphase region replaced by the hook (answered by the
oracle); with MCB200_HOOK_DUMP_CUTPATH the hook writes the registry in its final order and the container
cutpath_edge_creation_info exactly as "Create edges with intersection points" (kernel.cpp:3332-3617) is about to read it —
from a dispatch whose connected components equal the unmodified reference's bit for bit (tests/test_hook_cpu.py).  The
oracle's table must name the same face pairs in the same (std::map) order with the same points; groups of more than two
points must hold the same points, ordered along their line as linear_projection_sort (kernel.cpp:1496-1531) orders them,
which is restated here a second time in numpy."""
import os
import subprocess

import numpy as np
import pytest

import cases
from golden_util import CORPUS_CASES, load_corpus
from mcut_b200.mcbio import write_mcb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOOKED_CPU = os.path.join(ROOT, "oracle", "_ref", "api_driver_hooked_cpu")
NODUMP = os.path.join(ROOT, "oracle", "_ref", "libnodump.so")

pytestmark = pytest.mark.skipif(not os.path.exists(HOOKED_CPU), reason="oracle/_ref is not built (needs /root/reference at build time)")


def live_table(tmp, src, cut, flags):
    """(records in registry order, edge_f, src_nf, [(sm, cm, [v...])]) of the LAST kernel invocation of the dispatch"""
    d = {"src_xyz": src[0], "src_faces": src[1], "cut_xyz": cut[0], "cut_faces": cut[1], "flags": np.array([flags], dtype=np.uint32)}
    if src[2] is not None:
        d["src_sizes"] = src[2]
    if cut[2] is not None:
        d["cut_sizes"] = cut[2]
    ip, op, dump = os.path.join(tmp, "in.mcb"), os.path.join(tmp, "out.mcb"), os.path.join(tmp, "cutpath.txt")
    write_mcb(ip, d)
    r = subprocess.run([HOOKED_CPU, ip, op], capture_output=True, text=True, cwd=tmp,
                       env=dict(os.environ, LD_PRELOAD=NODUMP, MCB200_HOOK_DUMP_CUTPATH=dump))
    assert r.returncode == 0, r.stderr[-2000:]
    if not os.path.exists(dump):
        return None
    lines = open(dump).read().split("\n")
    tag, n, src_nf = lines[0].split()
    n, src_nf = int(n), int(src_nf)
    from oracle.pyoracle import RECORD_DTYPE
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    faces = np.zeros((n, 2), dtype=np.uint32)
    for i in range(n):
        t = lines[1 + i].split()
        rec["edge"][i], rec["face"][i] = int(t[0]), int(t[1])
        faces[i] = (int(t[2]), int(t[3]))
        rec["point"][i] = np.array([int(x, 16) for x in t[4:7]], dtype=np.uint64).view(np.float64)
    ne = int(rec["edge"].max()) + 1 if n else 1
    edge_f = np.full((ne, 2), 0xFFFFFFFF, dtype=np.uint32)
    edge_f[rec["edge"]] = faces
    g = int(lines[1 + n].split()[1])
    groups = []
    for k in range(g):
        t = [int(x) for x in lines[2 + n + k].split()]
        groups.append((t[0], t[1], t[3:3 + t[2]]))
    return rec, edge_f, src_nf, groups


def projection_order(points):
    """linear_projection_sort restated in numpy float64 (same operation order)"""
    o, d = points[0], points[1]
    v = o - d
    len2 = 0.0
    for k in range(3):
        len2 = len2 + v[k] * v[k]
    v = v / np.sqrt(len2)
    proj = []
    for p in points:
        acc = 0.0
        for k in range(3):
            acc = acc + (o[k] - p[k]) * v[k]
        proj.append(acc)
    return np.argsort(np.array(proj), kind="stable")


def check(table):
    from oracle import pyoracle
    rec, edge_f, src_nf, groups = table
    cp = pyoracle.cutpath_segments(edge_f, src_nf, rec)
    assert cp["keys"].size == len(groups), "face pairs"
    many = 0
    for g, (sm, cm, vs) in enumerate(groups):
        assert int(cp["keys"][g]) == (sm << 32 | cm), "std::map order of the face pairs"
        got = cp["vtx"][cp["off"][g]:cp["off"][g + 1]].tolist()
        if len(vs) <= 2:
            assert got == vs
        else:
            many += 1
            assert sorted(got) == sorted(vs)
            want = [vs[i] for i in projection_order(rec["point"][vs])]
            assert got == want, "order along the line"
    assert cp["n_single"] == sum(1 for _, _, vs in groups if len(vs) == 1)
    return many


@pytest.mark.parametrize("pair", CORPUS_CASES)
def test_corpus_cutpath_table_equals_the_live_dispatch(tmp_path, pair):
    _, src, cut, flags = load_corpus(pair)
    table = live_table(str(tmp_path), src, cut, flags)
    assert table is not None
    check(table)


@pytest.mark.parametrize("case", ["hello", "spheres_k16", "patch_vs_sphere", "ico_pair", "spheres_k64"])
def test_cutpath_table_equals_the_live_dispatch(tmp_path, case):
    src, cut, flags = cases.ALL[case]()
    table = live_table(str(tmp_path), src, cut, flags)
    assert table is not None and len(table[3]) > 0
    check(table)


def test_the_corpus_has_groups_of_more_than_two_points(tmp_path_factory):
    """the >2 branch (kernel.cpp:3487) is really exercised: count such groups over the corpus"""
    total = 0
    for pair in CORPUS_CASES[:40]:
        _, src, cut, flags = load_corpus(pair)
        table = live_table(str(tmp_path_factory.mktemp("cp")), src, cut, flags)
        total += sum(1 for _, _, vs in table[3] if len(vs) > 2)
    assert total > 0
