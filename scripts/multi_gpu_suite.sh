#!/bin/bash
# usage: multi_gpu_suite.sh N   - correctness + all benchmark configurations on N GPUs of one box; results -> gpurun_out/
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
port=29600
run() { port=$((port + 1)); timeout ${T:-300} $TR --master-port $port "$@"; }
run tests/mp_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | grep -v ": PASSED" | tail -8
run bench.py --gpus $N --steps 10 --warmup 3 --nccl > gpurun_out/bench$N.json 2> gpurun_out/bench$N.err
python scripts/show_bench.py gpurun_out/bench$N.json
run bench.py --gpus $N --steps 20 --warmup 3 --bytes 67108864 --compress --no-sweep --no-e2e > gpurun_out/bench${N}_fp8_64mb.json 2>> gpurun_out/bench$N.err
run bench.py --gpus $N --steps 20 --warmup 3 --bytes 67108864 --no-sweep --no-e2e > gpurun_out/bench${N}_fp32_64mb.json 2>> gpurun_out/bench$N.err
python scripts/show_bench.py gpurun_out/bench${N}_fp8_64mb.json gpurun_out/bench${N}_fp32_64mb.json
run bench/gemm_rs_bench.py 2>&1 | grep "^{" | tee gpurun_out/gemm_rs_$N.jsonl | cut -c1-400
for cfg in "mlsl --mode fused" "mlsl --mode allreduce" "ddp"; do
  run bench/train_bench.py --model resnet50 --steps 15 --warmup 5 --impl $cfg 2>&1 | grep "^{" | tee -a gpurun_out/train_$N.jsonl
done
if [ "${BERT:-1}" = "1" ]; then
  for cfg in "mlsl --mode fused" "ddp"; do
    T=400 run bench/train_bench.py --model bert-large --steps 8 --warmup 3 --impl $cfg 2>&1 | grep "^{" | tee -a gpurun_out/train_$N.jsonl
  done
fi
