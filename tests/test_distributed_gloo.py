"""world_size-2 gloo run of the host-side multi-GPU logic (sharded single dispatch): the ranks' partial
pair / record buffers are all-gathered and merged into exactly the single-rank output."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pairs_np, rec_np, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcut_b200 import distributed as D
    # deal the candidate pairs the way the traversal shards its query leaves: by chunks of the source order
    src = (pairs_np >> np.uint64(32)).astype(np.int64)
    mine = np.array([D.shard_of_leaf(int(s), world, chunk=64) == rank for s in src], dtype=bool)
    local = torch.from_numpy(pairs_np[mine].view(np.int64).copy())
    merged = D.merge_pairs(local)
    ok_pairs = np.array_equal(merged.numpy().view(np.uint64), np.sort(pairs_np))
    redge = rec_np.view(np.uint32).reshape(-1, 8)[:, 0]
    rmine = (redge % world) == rank
    lrec = torch.from_numpy(rec_np[rmine].copy())
    mrec = D.merge_records(lrec)
    ok_rec = np.array_equal(mrec.numpy().tobytes(), rec_np.tobytes())
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as fp:
        fp.write(f"{int(ok_pairs)} {int(ok_rec)} {local.numel()} {lrec.shape[0]}")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_merge_equals_single_rank(oracle, tmp_path):
    src, cut, flags = cases.spheres_k16()
    r = oracle.intersect_stage(src, cut, flags)
    pairs = r["pairs"]
    rec = r["records"]
    rec_rows = np.zeros((len(rec), 4))
    rec_rows.view(np.uint32).reshape(-1, 8)[:, 0] = rec["edge"]
    rec_rows.view(np.uint32).reshape(-1, 8)[:, 1] = rec["face"]
    rec_rows[:, 1:4] = rec["point"]
    port = _free_port()
    mp.spawn(_worker, args=(2, port, pairs, rec_rows, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        ok_pairs, ok_rec, npairs, nrec = open(tmp_path / f"rank{rank}.txt").read().split()
        assert ok_pairs == "1" and ok_rec == "1"
        assert 0 < int(npairs) < pairs.size, "both ranks got a share of the pairs"
