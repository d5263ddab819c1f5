#!/bin/bash
# usage: tune_ar.sh N  -> prints busbw at 64MB/256MB/1GiB for a few (channels, unroll) settings
N=$1
for cfg in "96 0" "128 0" "128 8" "64 8" "128 4"; do
  set -- $cfg
  MLSL_NUM_CHANNELS=$1 MLSL_AR_UNROLL=$2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((20 + RANDOM % 70)) bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --sweep-sizes 67108864,268435456 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('channels=$1 unroll=$2  1GiB=%.1f  ' % d['value'] + '  '.join('%dMB=%.1f' % (s['bytes'] >> 20, s['busbw_GBps']) for s in d['sweep']))
"
done
