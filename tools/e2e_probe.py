"""Ad-hoc: where does the host-array call spend its time?  (run on the GPU box)"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from mcut_b200 import stage, meshgen
from mcut_b200._lib import HostMesh, HostSoup

(src, cut, flags) = meshgen.c2_two_spheres(289)
(sx, sf, ss), (cx, cf, cs) = src, cut
dev = torch.device("cuda:0")
ts = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ts)
ctx = stage.Context(0, ts.cuda_stream)
com, shift, sbb, cbb = stage.vertex_parameters(sx, cx)
eps = stage.cut_bbox_eps(cbb, 1e-4, False)
keep = []
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); keep.append(t); return t.numpy()
hsx, hsf, hcx, hcf = pin(sx), pin(sf), pin(cx), pin(cf)
hm_s = HostMesh(0, hsx.ctypes.data, sx.shape[0], hsf.ctypes.data, None, sf.size // 3)
hm_c = HostMesh(0, hcx.ctypes.data, cx.shape[0], hcf.ctypes.data, None, cf.size // 3)
res = stage.Result(ctx)
L = ctx.L
def call():
    ctx.check(L.mcb200_intersect_stage_host(ctx.h, ctypes.byref(hm_s), ctypes.byref(hm_c), com.ctypes.data_as(stage.c_dp),
                                            shift.ctypes.data_as(stage.c_dp), None, eps, None, res.h, 0))
for _ in range(3):
    call(); res.counts()
torch.cuda.synchronize()
# host time of the call itself vs until completion
for _ in range(3):
    t0 = time.perf_counter(); call(); t1 = time.perf_counter(); c = res.counts(); t2 = time.perf_counter()
    print(f"call returns after {1e3*(t1-t0):.3f} ms, counts after {1e3*(t2-t0):.3f} ms")
ctx.set_profiling(True)
for _ in range(5):
    call(); res.counts()
prof = ctx.profile_read()
ctx.set_profiling(False)
for k, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:40s} n={cnt/5:4.1f} us/step={ms/5*1e3:8.1f}")
# raw copy speed
a = torch.empty(48 * 1024 * 1024, dtype=torch.uint8).pin_memory(); b = torch.empty_like(a, device=dev)
for _ in range(2): b.copy_(a, non_blocking=True)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); b.copy_(a, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("48 MiB H2D pinned:", e0.elapsed_time(e1), "ms")
