"""One rank of tests/test_gpu_sharded.py (launched by torch.distributed.run): a dispatch split over all ranks through
mcb200_intersect_stage_sharded must leave, on EVERY rank, exactly the pairs / records / counts one GPU produces."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
from mcut_b200 import meshgen as mg  # noqa: E402
from mcut_b200 import stage  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    ctx = stage.Context(local)
    L = ctx.L
    idbuf = ctypes.create_string_buffer(128)
    if rank == 0:
        ctx.check(L.mcb200_comm_unique_id(idbuf))
    t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
    dist.broadcast(t, 0)
    comm = ctypes.c_void_p()
    ctx.check(L.mcb200_comm_create(ctx.h, world, rank, bytes(t.cpu().numpy().tobytes()), ctypes.byref(comm)))
    inputs = {"spheres_k64": cases.ALL["spheres_k64"](), "c5_k96": mg.c5_coplanar_regions(k=96), "hello_tris": cases.ALL["cube_cube_tris_offset"]()}
    for name, (src, cut, flags) in inputs.items():
        ms, mc = stage.Mesh(ctx, *src), stage.Mesh(ctx, *cut)
        com, shift, sbb, cbb = stage.vertex_parameters(src[0], cut[0])
        ms.set_frame(com, shift)
        mc.set_frame(com, shift)
        eps = stage.cut_bbox_eps(cbb)
        soup = stage.Soup(ctx, ms, mc)
        one, many = stage.Result(ctx), stage.Result(ctx)
        for _ in range(5):
            try:
                ctx.check(L.mcb200_intersect_stage(ctx.h, ms.h, mc.h, eps, soup.h, one.h, 0))
                c1 = one.counts()
                break
            except stage.Mcb200Error as e:
                if e.code != stage.ERR_CAPACITY:
                    raise
        for _ in range(5):
            rc = L.mcb200_intersect_stage_sharded(ctx.h, comm, ms.h, mc.h, eps, soup.h, many.h, 0)
            if rc != stage.ERR_CAPACITY:
                ctx.check(rc)
                break
        cn = many.counts()
        for k in ("n_pairs", "n_tests", "n_exact", "n_records", "n_cand_faces", "status"):
            assert getattr(c1, k) == getattr(cn, k), (name, rank, k, getattr(c1, k), getattr(cn, k))
        assert one.pairs().tobytes() == many.pairs().tobytes(), (name, rank, "pairs")
        assert one.records().tobytes() == many.records().tobytes(), (name, rank, "records")
        p1, pn = one.planes(), many.planes()
        for a, b in zip(p1, pn):
            assert np.asarray(a).tobytes() == np.asarray(b).tobytes(), (name, rank, "plane rows")
        if rank == 0:
            print(f"sharded {name}: {cn.n_pairs} pairs, {cn.n_records} records, {cn.n_exact} exact tests identical on {world} ranks", flush=True)
        for o in (one, many, soup, ms, mc):
            o.free()
    L.mcb200_comm_destroy(comm)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
