#!/bin/bash
# 1-GPU validation: the whole `-m gpu` suite, then the opt-in kernels one pytest process each (a hang or a sticky CUDA
# error in one must not take the others down).  Logs: gpurun_out/t_*.log
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/t_all.log
MLSL_TEST_2CTA=1 timeout 300 python -m pytest tests/test_gemm_rs_gpu.py -q -m gpu -k two_cta > gpurun_out/t_2cta.log 2>&1; echo "rc=$?" >> gpurun_out/t_2cta.log
MLSL_TEST_AGGEMM=1 timeout 300 python -m pytest tests/test_tensor_parallel.py -q -m gpu -k fused_gpu > gpurun_out/t_aggemm.log 2>&1; echo "rc=$?" >> gpurun_out/t_aggemm.log
MLSL_TEST_STRATEGIES_GPU=1 timeout 400 python -m pytest tests/test_zz_strategies_gpu.py -q -m gpu > gpurun_out/t_strategies.log 2>&1; echo "rc=$?" >> gpurun_out/t_strategies.log
MLSL_BACKEND=cuda CUDA_MODULE_LOADING=EAGER timeout 300 python examples/mlsl_test.py 2 1 --inproc 4 > gpurun_out/t_mlsl_test_py.log 2>&1; echo "rc=$?" >> gpurun_out/t_mlsl_test_py.log
timeout 120 bin/mlslrun -n 2 bin/mlsl_example_cuda > gpurun_out/t_example_cuda.log 2>&1; echo "rc=$?" >> gpurun_out/t_example_cuda.log
for f in all 2cta aggemm strategies mlsl_test_py example_cuda; do echo "== $f"; tail -4 gpurun_out/t_$f.log; done
