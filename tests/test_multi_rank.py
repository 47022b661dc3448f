"""N>1 path: element partition, particle migration and node halo sum must reproduce the single-rank result.

CPU: world_size 2 and 3 over gloo with the oracle as the per-rank engine (host/transport logic of piclas_b200.multi).
GPU: world_size 2 over NCCL with libpiclas_gpu.so (needs 2 GPUs: `gpurun --gpus 2`).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multi_worker.py")


def _run(world, engine, port, depo="cvwm", case="synthetic"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, "--engine", engine, "--depo", depo, "--case", case]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("world", [2, 3])
def test_multi_rank_gloo_oracle(world):
    _run(world, "oracle", 29611 + world)


@pytest.mark.parametrize("world,depo", [(2, "sf"), (3, "cc")])
def test_multi_rank_gloo_oracle_shape_function(world, depo):
    """DOF halo of the shape-function deposition over gloo."""
    _run(world, "oracle", 29617 + world, depo=depo)


@pytest.mark.parametrize("world,case", [(2, "periodic_ref"), (5, "periodic_ref"), (2, "ansa_ref")])
def test_multi_rank_gloo_reproduces_the_references_state_files(world, case):
    """The reference runs NIG_tracking_DSMC/periodic with MPI = 1,2,5,10 and ANSA_box with MPI = 1,2 against one committed
    PartInt; so do we (element partition of loaddistribution.f90:362-369, migration after tracking)."""
    _run(world, "oracle", 29640 + world + (10 if case == "ansa_ref" else 0), case=case)


@pytest.mark.gpu
@pytest.mark.parametrize("depo", ["cvwm", "sf", "cc"])
def test_multi_rank_gpu_nccl(depo):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "gpu", 29631 + ["cvwm", "sf", "cc"].index(depo), depo=depo)
