#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d3_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/d3_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/d3_bench1.json 2> gpurun_out/d3_bench1.err; echo "bench rc=$?" >> gpurun_out/d3_bench1.err
MLSL_BACKEND=cuda CUDA_MODULE_LOADING=EAGER timeout 300 python examples/mlsl_test.py 2 1 --inproc 4 > gpurun_out/d3_mlsl_test_py.log 2>&1; echo "rc=$?" >> gpurun_out/d3_mlsl_test_py.log
MLSL_TEST_STRATEGIES_GPU=1 MLSL_TEST_DUMP_AFTER=12 timeout 200 python -m pytest tests/test_zz_strategies_gpu.py -x -q -m gpu -k pipeline > gpurun_out/d3_pipe.log 2>&1
MLSL_TEST_STRATEGIES_GPU=1 MLSL_TEST_DUMP_AFTER=12 timeout 200 python -m pytest tests/test_zz_strategies_gpu.py -x -q -m gpu -k transformer > gpurun_out/d3_tfm.log 2>&1
tail -3 gpurun_out/d3_smoke.log; tail -2 gpurun_out/d3_bench1.err; cut -c1-1500 gpurun_out/d3_bench1.json; tail -3 gpurun_out/d3_mlsl_test_py.log; tail -2 gpurun_out/d3_pipe.log gpurun_out/d3_tfm.log
