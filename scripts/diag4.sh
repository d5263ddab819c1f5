#!/bin/bash
mkdir -p gpurun_out
MLSL_BACKEND=cuda CUDA_MODULE_LOADING=EAGER MLSL_WATCHDOG_SEC=15 timeout 300 python examples/mlsl_test.py 2 1 --inproc 4 > gpurun_out/d4_mlsl_test_py.log 2>&1; echo "rc=$?" >> gpurun_out/d4_mlsl_test_py.log
MLSL_TEST_STRATEGIES_GPU=1 timeout 400 python -m pytest tests/test_zz_strategies_gpu.py -q -m gpu > gpurun_out/d4_strategies.log 2>&1
tail -4 gpurun_out/d4_mlsl_test_py.log; tail -3 gpurun_out/d4_strategies.log
