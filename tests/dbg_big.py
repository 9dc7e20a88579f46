import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from mcut_b200 import meshgen as mg, stage
k = int(sys.argv[1]) if len(sys.argv) > 1 else 289
src, cut, flags = mg.c2_two_spheres(k=k)
ctx = stage.Context(0)
com, shift, sbb, cbb = stage.vertex_parameters(src[0], cut[0])
eps = stage.cut_bbox_eps(cbb)
ms = stage.Mesh(ctx, *src); mc = stage.Mesh(ctx, *cut)
ms.set_frame(com, shift); mc.set_frame(com, shift)
print('build src'); ms.build(0.0); ctx.sync()
print('build cut'); mc.build(eps); ctx.sync()
res = stage.Result(ctx)
print('intersect'); ctx.check(ctx.L.mcb200_bvh_intersect(ctx.h, ms.h, mc.h, res.h)); ctx.sync()
c = res.counts(); print('pairs', c.n_pairs, 'node tests', c.n_node_tests)
soup = stage.Soup(ctx, ms, mc)
print('narrow'); ctx.check(ctx.L.mcb200_narrowphase(ctx.h, soup.h, ms.h, mc.h, res.h, 0)); ctx.sync()
c = res.counts(); print('tests', c.n_tests, 'exact', c.n_exact, 'records', c.n_records, 'status', c.status)
print('stage call'); 
for i in range(3):
    ctx.check(ctx.L.mcb200_intersect_stage(ctx.h, ms.h, mc.h, eps, soup.h, res.h, 0)); ctx.sync()
    c = res.counts(); print(i, 'pairs', c.n_pairs, 'tests', c.n_tests, 'records', c.n_records)
