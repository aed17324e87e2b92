"""GPU: row-sharded fit over NCCL with one rank per GPU (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port, env_extra=None):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (covered on CPU by test_distributed_gloo.py)" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("DIST_GPU_OK") >= 5
    return r.stdout


# world 2: the two blocks are "opposite" (checkerboard split of the off-diagonal block);
# world 3: odd ring, every off-diagonal block has one owner; world 4: both kinds at once
@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_fit_peer_symmetric_gram(world):
    out = _run(world, 29640 + world)
    print(out)


def test_two_gpu_sharded_fit_without_peer_access():
    """FXB200_NO_PEER=1: every rank computes its whole row block (the path taken when CUDA IPC /
    NVLink peer access is unavailable); same results."""
    out = _run(2, 29649, {"FXB200_NO_PEER": "1"})
    assert "peer=1" not in out
