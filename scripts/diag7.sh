#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MLSL_BACKEND=cuda timeout 100 $TR --master-port 29751 examples/mlsl_test.py 1 1 > gpurun_out/d7_a.log 2>&1; echo "rc=$?" >> gpurun_out/d7_a.log
MLSL_BACKEND=cuda timeout 100 $TR --master-port 29752 examples/mlsl_test.py 2 0 > gpurun_out/d7_b.log 2>&1; echo "rc=$?" >> gpurun_out/d7_b.log
grep -E "summary|rc=|Error" gpurun_out/d7_a.log | head -5; grep -E "summary|rc=|Error" gpurun_out/d7_b.log | head -5
