"""Whole mcDispatch through the public C API: unmodified reference vs the hooked library (broadphase + narrowphase on the
B200), with the adapter's own wall times (MCB200_SHIM_TIMING).  (run on the GPU box)"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from mcut_b200 import meshgen as mg
from mcut_b200.mcbio import write_mcb, read_mcb

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 100
repeat = sys.argv[2] if len(sys.argv) > 2 else "1"
src, cut, flags = mg.c2_two_spheres(k=k)
with tempfile.TemporaryDirectory() as td:
    ip = os.path.join(td, "in.mcb")
    write_mcb(ip, {"src_xyz": src[0], "src_faces": src[1], "cut_xyz": cut[0], "cut_faces": cut[1], "flags": np.array([flags], dtype=np.uint32)})
    for tag, drv in (("reference", "api_driver"), ("hooked", "api_driver_hooked")):
        env = dict(os.environ, LD_PRELOAD=os.path.join(ROOT, "oracle", "_ref", "libnodump.so"), MCB200_SHIM_TIMING="1")
        for rep in range(1):
            t0 = time.perf_counter()
            r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", drv), ip, os.path.join(td, tag + ".out.mcb"), "--repeat", repeat], capture_output=True, text=True, cwd=td, env=env)
            dt = time.perf_counter() - t0
        out = read_mcb(os.path.join(td, tag + ".out.mcb"))
        print(f"{tag:10s} k={k} ({12*k*k} tris/mesh): process wall {dt:.2f} s, mcDispatch={int(out['mcDispatch_result'][0])}, components={out['cc_type'].size}")
        for line in r.stderr.splitlines():
            if "shim" in line: print("    ", line)
