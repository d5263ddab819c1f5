#!/bin/bash
# second N-GPU session: hybrid all-reduce sweep, fp8 vs fp32 at 64 MiB, bcast/RS recheck, bench.py
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29706 bench/tune_allreduce.py --mb 1024,256 --channels 64,148 --unroll 2 --hybrid 10:25,15:25,20:33,25:33,30:40,20:50 > gpurun_out/tune_hybrid_$N.jsonl 2> gpurun_out/tune_hybrid_$N.err
timeout 300 $TR --master-port 29707 bench.py --gpus $N --steps 20 --warmup 5 --bytes 67108864 --no-sweep --no-e2e --no-nccl > gpurun_out/bench_${N}_fp32_64mb.json 2> gpurun_out/bench_${N}_fp32_64mb.err
timeout 300 $TR --master-port 29708 bench.py --gpus $N --steps 20 --warmup 5 --bytes 67108864 --no-sweep --no-e2e --no-nccl --compress > gpurun_out/bench_${N}_fp8_64mb.json 2> gpurun_out/bench_${N}_fp8_64mb.err
timeout 300 $TR --master-port 29705 bench/collectives_bench.py --max-mb 64 > gpurun_out/coll2_$N.jsonl 2> gpurun_out/coll2_$N.err
timeout 600 $TR --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench rc=$?" >> gpurun_out/bench_$N.err
grep -v '"unroll"' gpurun_out/tune_hybrid_$N.jsonl; grep BEST gpurun_out/tune_hybrid_$N.jsonl
for f in fp32 fp8; do python scripts/show_bench.py gpurun_out/bench_${N}_${f}_64mb.json | head -1; tail -1 gpurun_out/bench_${N}_${f}_64mb.err; done
grep -E "broadcast|reduce_scatter" gpurun_out/coll2_$N.jsonl | cut -c1-200
python scripts/show_bench.py gpurun_out/bench_$N.json; tail -2 gpurun_out/bench_$N.err
