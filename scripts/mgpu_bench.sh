#!/bin/bash
# bench.py (ours, with the NCCL arm) at N GPUs; extra env through EXTRA="K=V K=V"
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
env $EXTRA timeout 900 $TR --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}_$TAG.json 2> gpurun_out/bench_${N}_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_${N}_$TAG.err
python scripts/show_bench.py gpurun_out/bench_${N}_$TAG.json 2>/dev/null || cut -c1-600 gpurun_out/bench_${N}_$TAG.json
tail -2 gpurun_out/bench_${N}_$TAG.err
