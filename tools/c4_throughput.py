"""C4-style throughput on ONE GPU: T host threads, one context each, every thread pushes small independent dispatches
(two 5,120-triangle icospheres) through mcb200_intersect_stage_host and reads the counts.  (run on the GPU box)"""
import ctypes, os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from mcut_b200 import stage, meshgen

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
pairs = [meshgen.c4_pair(j) for j in range(8)]


def worker(tid, n, out):
    ctx = stage.Context(0)
    res = stage.Result(ctx)
    done = 0
    for i in range(n):
        src, cut, flags = pairs[(tid + i) % len(pairs)]
        r = stage.intersect_stage_host(ctx, src, cut, flags, res=res)
        done += 1 if r["status"] in (0, 1) else 0
    res.free()
    ctx.close()
    out[tid] = done


for T in (1, 2, 4, 8, 16):
    out = [0] * T
    th = [threading.Thread(target=worker, args=(t, N // T, out)) for t in range(T)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"threads={T:2d} dispatches={sum(out)} in {dt:.3f} s -> {sum(out)/dt:8.0f} dispatches/s ({1e3*dt/sum(out):.3f} ms each)")
