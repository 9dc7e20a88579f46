"""-m gpu (needs two GPUs on the box): one dispatch split over several ranks inside the C-ABI (mcb200_intersect_stage_sharded:
NCCL all-gather of counts, grouped broadcast of pairs / records, all-reduce of per-face counts and candidate flags) gives every
rank the single-GPU result byte for byte."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(gpu_count() < 2, reason="needs at least two GPUs")
def test_sharded_dispatch_equals_single_gpu():
    n = min(gpu_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "tests", "sharded_worker.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "identical on" in r.stdout
