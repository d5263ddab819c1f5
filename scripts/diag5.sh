#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tests/mp_gpu_check.py > gpurun_out/d5_mp.log 2>&1
CUDA_VISIBLE_DEVICES=0 MLSL_WATCHDOG_SEC=8 timeout 240 python -m pytest tests/test_collectives_gpu.py -x -q -m gpu -k "split or heap_pool or quant" > gpurun_out/d5_coll.log 2>&1
grep -v PASSED gpurun_out/d5_mp.log | grep -v "^\s*$" | grep -v "^\*\|OMP" | head -12; grep -c PASSED gpurun_out/d5_mp.log
tail -3 gpurun_out/d5_coll.log; grep -E "^E  " gpurun_out/d5_coll.log | head -5
