"""One rank per GPU over real NVLink: spawns `torchrun tests/mp_gpu_check.py` at every N in {2, 4, 8} the box offers (VMM slab,
NVLS multicast from 4 ranks, peer-to-peer below).  Skipped on single-GPU boxes - there the loop-back tests of this
directory exercise the same kernels with N ranks sharing the GPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("n", [2, 4, 8])
def test_one_rank_per_gpu(n):
    if _ngpus() < n:
        pytest.skip("needs %d GPUs, %d visible" % (n, _ngpus()))
    env = dict(os.environ)
    for k in ("CUDA_MODULE_LOADING", "MLSL_BACKEND", "MLSL_HEAP_SIZE_GB", "MLSL_WATCHDOG_SEC"):
        env.pop(k, None)
    env["MLSL_BACKEND"] = "cuda"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + n), os.path.join(ROOT, "tests", "mp_gpu_check.py")]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    tail = res.stdout[-6000:]
    assert res.returncode == 0 and "ALL PASSED" in res.stdout and "FAILED" not in res.stdout.replace("0 FAILED", ""), tail


def test_message_priority_one_rank_per_gpu():
    """MLSL_MSG_PRIORITY=1 on real GPUs: the progress threads launch queued gradient all-reduces newest first, in the
    same order on every rank (tests/mp_priority_check.py)."""
    n = 2 if _ngpus() < 4 else 4
    if _ngpus() < n:
        pytest.skip("needs %d GPUs, %d visible" % (n, _ngpus()))
    env = dict(os.environ)
    for k in ("CUDA_MODULE_LOADING", "MLSL_BACKEND", "MLSL_HEAP_SIZE_GB", "MLSL_WATCHDOG_SEC", "MLSL_STREAM_MODE"):
        env.pop(k, None)
    env["MLSL_MSG_PRIORITY"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29633", os.path.join(ROOT, "tests", "mp_priority_check.py")]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0 and "ALL PASSED" in res.stdout, res.stdout[-4000:]


@pytest.mark.parametrize("groups,dist_update", [(1, 1), (2, 0)])
def test_reference_functional_scenario_one_rank_per_gpu(groups, dist_update):
    """The reference's functional test (two OT_CC layers, closed-form values; tests/examples/mlsl_test/) through the Python
    object model on real GPUs: data parallel with distributed update, and model parallel."""
    n = 2
    if _ngpus() < n:
        pytest.skip("needs %d GPUs, %d visible" % (n, _ngpus()))
    env = dict(os.environ)
    for k in ("CUDA_MODULE_LOADING", "MLSL_HEAP_SIZE_GB", "MLSL_WATCHDOG_SEC", "MLSL_STREAM_MODE"):
        env.pop(k, None)
    env["MLSL_BACKEND"] = "cuda"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(29650 + groups), os.path.join(ROOT, "examples", "mlsl_test.py"), str(groups), str(dist_update)]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0 and "0 FAILED" in res.stdout and "Run FAILED" not in res.stdout, res.stdout[-4000:]
