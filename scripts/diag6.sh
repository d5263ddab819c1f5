#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29701 tests/mp_gpu_check.py > gpurun_out/d6_mp.log 2>&1; echo "rc=$?" >> gpurun_out/d6_mp.log
timeout 120 $TR --master-port 29741 examples/torch_ddp.py --device cuda > gpurun_out/d6_ddp.log 2>&1; echo "rc=$?" >> gpurun_out/d6_ddp.log
grep -v PASSED gpurun_out/d6_mp.log | grep -v "^\s*$" | grep -v "^\*\|OMP" | head -12; grep -c PASSED gpurun_out/d6_mp.log; grep "gemm" gpurun_out/d6_mp.log
grep -v "^\*\|OMP\|^\s*$" gpurun_out/d6_ddp.log | tail -6
