"""What the adapter's call sequence costs per call, after the device has been idle for a while (a live mcDispatch spends
seconds on the host between its device calls).  (run on the GPU box)   usage: idle_probe.py [k] [idle seconds ...]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from mcut_b200 import meshgen as mg, stage, _lib

k = int(sys.argv[1]) if len(sys.argv) > 1 else 289
idles = [float(a) for a in sys.argv[2:]] or [0.0, 0.0, 2.0, 10.0]
src, cut, flags = mg.c2_two_spheres(k=k)
ctx = stage.Context(0)
L = ctx.L
com, shift, sbb, cbb = stage.vertex_parameters(src[0], cut[0])
eps = stage.cut_bbox_eps(cbb)


class T:
    def __init__(self, what):
        self.what = what
    def __enter__(self):
        self.t = time.perf_counter()
    def __exit__(self, *a):
        print(f"    {self.what:34s} {1e3 * (time.perf_counter() - self.t):9.3f} ms", flush=True)


for idle in idles:
    print(f"idle {idle} s", flush=True)
    time.sleep(idle)
    with T("mesh_create(src)"):
        ms = stage.Mesh(ctx, src[0], src[1])
    with T("set_frame + bvh_build(src)"):
        ms.set_frame(com, shift); ms.build(0.0)
    with T("bvh_read(src) root only"):
        ms.read_bvh(want_boxes=False)
    with T("mesh_create(cut)"):
        mc = stage.Mesh(ctx, cut[0], cut[1])
    with T("set_frame + bvh_build(cut)"):
        mc.set_frame(com, shift); mc.build(eps)
    with T("bvh_read(cut) root only"):
        mc.read_bvh(want_boxes=False)
    res = stage.Result(ctx)
    with T("bvh_intersect + counts"):
        ctx.check(L.mcb200_bvh_intersect(ctx.h, ms.h, mc.h, res.h)); n = res.counts()
    time.sleep(idle)
    soup = C.c_void_p()
    with T("soup_number"):
        ctx.check(L.mcb200_soup_number(ctx.h, ms.h, mc.h, res.h, C.byref(soup)))
    with T("narrowphase + counts"):
        ctx.check(L.mcb200_narrowphase(ctx.h, soup, ms.h, mc.h, res.h, 0)); n = res.counts()
    with T("read planes + records"):
        res.planes(); res.records()
    print("    pairs", n.n_pairs, "records", n.n_records)
    L.mcb200_soup_free(ctx.h, soup)
    res.free(); ms.free(); mc.free()
