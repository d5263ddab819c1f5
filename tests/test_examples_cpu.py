"""The shipped examples must keep working: each is run as a user would (mlslrun / --inproc) on the host backend."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

MLSLRUN = os.path.join(ROOT, "bin", "mlslrun")


def _run(cmd, timeout=240):
    env = dict(os.environ, MLSL_BACKEND="host", MLSL_HEAP_SIZE_GB="0.25", MLSL_WATCHDOG_SEC="60",
               MLSL_JOB_ID="ex%d" % os.getpid())
    p = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-3000:]
    return p.stdout


@pytest.mark.parametrize("cfg", [["1"], ["4"], ["2", "1"], ["2", "1", "1", "1"]])
def test_python_functional_test_inproc(cfg):
    out = _run([sys.executable, "examples/mlsl_test.py", *cfg, "--inproc", "4"])
    assert "Run PASSED." in out and ": FAILED" not in out
    assert out.count("summary:") == 4


def test_python_functional_test_multiprocess():
    out = _run([MLSLRUN, "-n", "4", sys.executable, "examples/mlsl_test.py", "2", "1"])
    assert out.count("0 FAILED") == 4 and ": FAILED" not in out


def test_workflow_example():
    out = _run([sys.executable, "examples/mlsl_example.py", "--inproc", "3"])
    assert out.count("PASSED") == 3


def test_allreduce_sample():
    out = _run([MLSLRUN, "-n", "4", sys.executable, "examples/allreduce_sample.py"])
    assert out.count("PASSED") == 4


@pytest.mark.parametrize("mode", [["--mode", "fused"], ["--mode", "allreduce", "--optimizer", "sgd"],
                                  ["--mode", "allreduce", "--compress"]])
def test_data_parallel_training_example(mode):
    out = _run([MLSLRUN, "-n", "2", sys.executable, "examples/train_data_parallel.py", "--steps", "6", *mode])
    assert out.count("replicas identical: True") == 2, out[-3000:]


def test_tensor_parallel_training_example():
    out = _run([MLSLRUN, "-n", "4", sys.executable, "examples/train_tensor_parallel.py", "--model-parts", "2", "--steps", "6"])
    assert out.count("PASSED") == 4


def test_torch_ddp_example():
    out = _run([MLSLRUN, "-n", "4", sys.executable, "examples/torch_ddp.py"])
    assert out.count("PASSED") == 4 and "FAILED" not in out


def test_pipeline_parallel_training_example():
    out = _run([MLSLRUN, "-n", "4", sys.executable, "examples/train_pipeline_parallel.py", "--stages", "2"])
    assert out.count("PASSED") == 4 and "FAILED" not in out


def test_parallel_transformer_training_example():
    out = _run([MLSLRUN, "-n", "4", sys.executable, "examples/train_parallel_transformer.py", "--model-parts", "2"])
    assert out.count("PASSED") == 4 and "FAILED" not in out
