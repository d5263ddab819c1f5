#!/bin/bash
# progress-engine session at N GPUs: correctness through the server thread, priority order, latency with / without a server
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29701 tests/mp_gpu_check.py > gpurun_out/mp_check_srv_$N.log 2>&1; echo "mp_check(server) rc=$?" | tee -a gpurun_out/mp_check_srv_$N.log
MLSL_MSG_PRIORITY=1 timeout 300 $TR --master-port 29711 tests/mp_priority_check.py > gpurun_out/prio_$N.log 2>&1; echo "prio rc=$?" | tee -a gpurun_out/prio_$N.log
MLSL_MSG_PRIORITY=0 timeout 300 $TR --master-port 29712 tests/mp_priority_check.py > gpurun_out/prio_off_$N.log 2>&1
for srv in 0 1; do
  MLSL_STREAM_MODE=comm MLSL_NUM_SERVERS=$srv MLSL_BENCH_GRAPH=0 timeout 300 $TR --master-port 2972$srv bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-nccl --sweep-sizes 1024,65536,1048576,16777216,268435456 > gpurun_out/bench_${N}_srv$srv.json 2> gpurun_out/bench_${N}_srv$srv.err
  python scripts/show_bench.py gpurun_out/bench_${N}_srv$srv.json | grep -v "e2e\|clocks"
done
grep -c PASSED gpurun_out/mp_check_srv_$N.log; grep FAILED gpurun_out/mp_check_srv_$N.log | head -5; tail -2 gpurun_out/mp_check_srv_$N.log
cat gpurun_out/prio_$N.log | grep -v "^\*\|OMP" | tail -8; grep "iteration 2" -A1 gpurun_out/prio_off_$N.log
