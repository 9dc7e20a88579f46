"""CPU: the HOST half of the narrowphase hook against the unmodified reference.

oracle/_ref/api_driver_hooked_cpu is the reference with the inline narrowphase of dispatch() (kernel.cpp:1779-3206) replaced
by one call to mcb200_hook_narrowphase() — the same patched kernel the device drop-in uses — but answered by the oracle
(oracle/hook_oracle.cpp) instead of the B200.  What is under test is the code both hooks share
(mcut_b200/csrc/shim/hook_fill.h + mcb200_reference_edge_rank in host_logic.cpp): how flat records become the reference's
containers and in which order the registry is handed over.  Every output array of every connected component must equal the
unmodified reference's bit for bit — including the 27 corpus pairs whose floating-polygon resolution depends on how the
intersection vertices are numbered."""
import os
import subprocess

import numpy as np
import pytest

import cases
from golden_util import CORPUS_CASES, load_corpus
from mcut_b200.mcbio import read_mcb, write_mcb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "api_driver")
HOOKED_CPU = os.path.join(ROOT, "oracle", "_ref", "api_driver_hooked_cpu")
NODUMP = os.path.join(ROOT, "oracle", "_ref", "libnodump.so")
KEYS = ("cc_type", "cc_attrs", "cc_nv", "cc_nf", "cc_vertices", "cc_faces", "cc_face_sizes")

pytestmark = pytest.mark.skipif(not (os.path.exists(DRIVER) and os.path.exists(HOOKED_CPU)),
                                reason="oracle/_ref is not built (needs /root/reference at build time)")


def run(tmp, tag, driver, src, cut, flags, extra=()):
    d = {"src_xyz": src[0], "src_faces": src[1], "cut_xyz": cut[0], "cut_faces": cut[1], "flags": np.array([flags], dtype=np.uint32)}
    if src[2] is not None:
        d["src_sizes"] = src[2]
    if cut[2] is not None:
        d["cut_sizes"] = cut[2]
    ip, op = os.path.join(tmp, f"{tag}.in.mcb"), os.path.join(tmp, f"{tag}.out.mcb")
    write_mcb(ip, d)
    r = subprocess.run([driver, ip, op, *extra], capture_output=True, text=True, cwd=tmp, env=dict(os.environ, LD_PRELOAD=NODUMP))
    assert r.returncode == 0, r.stderr[-2000:]
    out = read_mcb(op)
    out["_stderr"] = r.stderr[-2000:]
    return out


def same(a, b):
    assert int(a["mcDispatch_result"][0]) == int(b["mcDispatch_result"][0]), b["_stderr"]
    for k in KEYS:
        assert a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes(), k


@pytest.mark.parametrize("pair", CORPUS_CASES)
def test_corpus_through_the_cpu_hook_is_bit_identical(tmp_path, pair):
    _, src, cut, flags = load_corpus(pair)
    a = run(str(tmp_path), "ref", DRIVER, src, cut, flags)
    b = run(str(tmp_path), "hooked_cpu", HOOKED_CPU, src, cut, flags)
    same(a, b)
    assert a["cc_type"].size > 0


@pytest.mark.parametrize("helpers", [0, 1, 3])
def test_registry_order_follows_the_helper_count(tmp_path, helpers):
    """More than 1024 candidate faces: the reference's parallel_for cuts its maps into blocks (one per thread), so its
    registry order depends on the helper-thread count; the replay is told the count and follows."""
    src, cut, flags = cases.ALL["spheres_k64"]()
    extra = ["--helpers", str(helpers)]
    a = run(str(tmp_path), "ref", DRIVER, src, cut, flags, extra)
    b = run(str(tmp_path), "hooked_cpu", HOOKED_CPU, src, cut, flags, extra)
    same(a, b)
    assert a["cc_type"].size > 0


@pytest.mark.parametrize("case", ["hello", "patch_vs_sphere", "cube_cube_axis_aligned", "float_spheres", "terrain_plane"])
def test_cases_through_the_cpu_hook(tmp_path, case):
    src, cut, flags = cases.ALL[case]()
    a = run(str(tmp_path), "ref", DRIVER, src, cut, flags)
    b = run(str(tmp_path), "hooked_cpu", HOOKED_CPU, src, cut, flags)
    same(a, b)
