"""GPU (-m gpu), needs >= 2 devices (skipped otherwise): the slab-decomposed step over NCCL must reproduce the
single-GPU result BITWISE (same neighbour order, same arithmetic, faces cut by the slab boundary evaluated
identically on both ranks) and keep ownership a partition while particles migrate."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case,steps", [("kh", 4), ("sedov", 3), ("fb", 3)])
def test_ranks_bitwise_equal_single_gpu(case, steps, world):
    """world = 4 adds what 2 ranks cannot show: interior slabs with two distinct neighbours"""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + world), os.path.join(ROOT, "tests", "mgpu_worker.py"), case, str(steps)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "bitwise_equal=True" in r.stdout
