"""One hashgraph on several GPUs of one box (sw_peer_connect, M > 64): needs >= 2 CUDA devices; the single-GPU run of
the suite skips it.  `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs it."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_round_steps_over_nvlink():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = 2 if n < 4 else 4 if n < 8 else 8
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "_shard_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("SHARD_GPU_OK") == 4, out.stdout[-2000:]
