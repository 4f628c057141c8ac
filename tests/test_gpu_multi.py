"""Multi-GPU paths on real devices (skipped with fewer than two GPUs; the driver's 1-GPU run skips them, the builder's
`gpurun --gpus 2` run records them under profiles/): the graph-partitioned training step with its halo exchanges
through torch.distributed (DistExchange) and through the library's own NCCL transport (mgn_dp_*, mgn_halo_exchange)."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu, pytest.mark.timeout(900)]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, *args, nproc=2):
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    return r.returncode


@pytest.mark.parametrize("transport", ["torch", "abi"])
def test_two_gpu_halo_step_matches_unpartitioned(transport):
    assert _torchrun("dist_halo_worker.py", transport) == 0
