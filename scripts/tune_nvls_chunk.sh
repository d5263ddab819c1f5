#!/bin/bash
# usage: tune_nvls_chunk.sh [N=8]   - 1 GiB all-reduce on N GPUs for several launch granularities of giant messages
# (MLSL_NVLS_CHUNK_MB; 0 = one launch) and with NVLS off (peer-to-peer two-shot): which one to make the default.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
port=29700
for cfg in "MLSL_NVLS_CHUNK_MB=0" "MLSL_NVLS_CHUNK_MB=128" "MLSL_NVLS_CHUNK_MB=256" "MLSL_NVLS_CHUNK_MB=512" "MLSL_NVLS=0"; do
  port=$((port + 1))
  env $cfg timeout 300 $TR --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-sweep --no-e2e > gpurun_out/chunk_$N.$cfg.json 2> gpurun_out/chunk_$N.err
  python - "$cfg" gpurun_out/chunk_$N.$cfg.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-24s busbw %8.1f GB/s  %.3f ms" % (sys.argv[1], d["value"], d["ms_per_step"]))
except Exception as e:
    print("%-24s failed: %s" % (sys.argv[1], e))
PY
done
