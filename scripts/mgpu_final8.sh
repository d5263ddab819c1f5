#!/bin/bash
# final N-GPU session of the round: correctness, bench (ours + NCCL arm inside), fp8 vs fp32 at 64 MiB, priorities, bcast / RS
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 tests/mp_gpu_check.py > gpurun_out/f_mp_check_$N.log 2>&1; echo "mp_check rc=$?" >> gpurun_out/f_mp_check_$N.log
timeout 300 $TR --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/f_bench_$N.json 2> gpurun_out/f_bench_$N.err; echo "bench rc=$?" >> gpurun_out/f_bench_$N.err
timeout 120 $TR --master-port 29707 bench.py --gpus $N --steps 20 --warmup 5 --bytes 67108864 --no-sweep --no-e2e --no-nccl > gpurun_out/f_bench_${N}_fp32_64mb.json 2> /dev/null
timeout 120 $TR --master-port 29708 bench.py --gpus $N --steps 20 --warmup 5 --bytes 67108864 --no-sweep --no-e2e --no-nccl --compress > gpurun_out/f_bench_${N}_fp8_64mb.json 2> /dev/null
MLSL_MSG_PRIORITY=1 timeout 120 $TR --master-port 29711 tests/mp_priority_check.py > gpurun_out/f_prio_$N.log 2>&1
timeout 200 $TR --master-port 29705 bench/collectives_bench.py --max-mb 64 > gpurun_out/f_coll_$N.jsonl 2> /dev/null
grep -c PASSED gpurun_out/f_mp_check_$N.log; grep FAILED gpurun_out/f_mp_check_$N.log | head -5; tail -2 gpurun_out/f_mp_check_$N.log
python scripts/show_bench.py gpurun_out/f_bench_$N.json; tail -2 gpurun_out/f_bench_$N.err
for f in fp32 fp8; do python scripts/show_bench.py gpurun_out/f_bench_${N}_${f}_64mb.json | head -1; done
grep -v "^\*\|OMP" gpurun_out/f_prio_$N.log | tail -4
grep -E "broadcast|reduce_scatter" gpurun_out/f_coll_$N.jsonl | cut -c1-190
