"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): NCCL halo + all-reduce path vs
the N-rank oracle on the same decomposition.  Launched through torchrun like bench.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_solvers(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True,
                       text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-4000:]
    assert p.stdout.count("MULTI-GPU-OK") == 9
