"""2-GPU run of the row-block shard + single ncclAllGather (needs >= 2 GPUs: `gpurun --gpus 2`)."""
import ctypes as C
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_block_shard_allgather_bitwise(built):
    n = C.c_int32(0)
    built.lib().gvt_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29671", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-5000:]
    assert p.stdout.count(" ok") == world
