#!/bin/bash
# hang diagnosis of the device fuzz test: launch trace + Python stacks of all rank threads while the kernels spin
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/diag_gpus.txt
MLSL_TRACE_LAUNCH=1 MLSL_TEST_DUMP_AFTER=9 timeout 300 python -m pytest tests/test_zz_fuzz_gpu.py -x -q -m gpu -k "4-2" > gpurun_out/diag_fuzz42.log 2>&1
echo "fuzz42 rc=$?" >> gpurun_out/diag_fuzz42.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/diag_pytest_all.log 2>&1
echo "all rc=$?" >> gpurun_out/diag_pytest_all.log
tail -5 gpurun_out/diag_fuzz42.log; tail -15 gpurun_out/diag_pytest_all.log
