#!/usr/bin/env python
"""Debug aid: run pairs of the regression corpus through the reference and through the hooked drop-in, print what differs.
usage: python tools/debug_hooked.py 7 11 20"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["MCB200_HOOK_DEBUG"] = "1"
os.environ["MCB200_SHIM_TIMING"] = "1"
import test_gpu_dropin as td  # noqa: E402
from golden_util import load_corpus  # noqa: E402

for a in sys.argv[1:]:
    pair = int(a)
    fx, src, cut, flags = load_corpus(pair)
    tmp = tempfile.mkdtemp()
    ref = td.run_driver(tmp, "ref", src, cut, flags, [td.NODUMP])
    hk = td.run_driver(tmp, "hooked", src, cut, flags, [td.NODUMP], driver=td.HOOKED)
    print(f"== pair {pair}: reference rc={int(ref['mcDispatch_result'][0])} ccs={ref['cc_type'].size}  "
          f"hooked rc={int(hk['mcDispatch_result'][0])} ccs={hk['cc_type'].size}")
    print("   fixture records per invocation:", [len(fx[f"d{k}_ipoints_sorted"]) if f"d{k}_ipoints_sorted" in fx.files else None
                                                for k in range(int(fx["n_dispatch"][0]))])
    print(hk["_stderr"])
    inv = lambda o: sorted(zip(o["cc_type"].tolist(), map(tuple, o["cc_attrs"].reshape(-1, 3).tolist())))  # noqa: E731
    print("   SUMMARY pair", pair, "inventory", inv(ref) == inv(hk), "exact", td.components_equivalent(ref, hk, 0.0), "tol1e-9",
          td.components_equivalent(ref, hk, 1e-9), "tol1e-6", td.components_equivalent(ref, hk, 1e-6))
    if os.environ.get("BRIEF"):
        continue
    ca, cb = td.canonical_components(ref), td.canonical_components(hk)
    for x, y in zip(ca, cb):
        if x != y:
            print("   component differs: ref", x[:3], "faces", len(x[3]), " hooked", y[:3], "faces", len(y[3]))
            fa, fb = set(x[3]), set(y[3])
            va = {k for f in x[3] for k in f}
            vb = {k for f in y[3] for k in f}
            import numpy as np
            print("     vertices only in ref:", [np.frombuffer(k, dtype=np.float64).tolist() for k in sorted(va - vb)][:6])
            print("     vertices only in hooked:", [np.frombuffer(k, dtype=np.float64).tolist() for k in sorted(vb - va)][:6])
            print("     faces only in ref:", len(fa - fb), " only in hooked:", len(fb - fa))
