"""N > 1 host logic on CPU: two gloo ranks run the row-sharded algorithm with the collectives
placed where the CUDA path places its NCCL calls (tests/dist_worker.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# world 2 splits the opposite blocks in a checkerboard; world 3 has one owner per off-diagonal block
@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharded_solver_and_symmetric_gram(world):
    env = dict(os.environ, OMP_NUM_THREADS="2", OPENBLAS_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29629 + world),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIST_OK" in r.stdout
