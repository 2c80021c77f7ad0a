"""Multi-rank paths: world_size-2 gloo on CPU for the host logic; NCCL with one GPU per rank on the GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _launch(mode, nproc, port, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), WORKER, mode]
    e = dict(os.environ, **(env or {}))
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=e)


def test_partition_and_halo_logic_world2_gloo():
    from ikarus_b200 import build

    build.build()
    r = _launch("gloo", 2, 29611)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_partition_and_halo_logic_world3_gloo():
    r = _launch("gloo", 3, 29612)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def _gpu_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("eas", [0, 9])
def test_partitioned_assembly_and_distributed_newton_nccl(eas):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _launch("nccl", 2, 29613 + eas, {"IKB_TEST_EAS": str(eas)})
    assert r.returncode == 0 and "NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
