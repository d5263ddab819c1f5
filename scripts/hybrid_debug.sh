#!/bin/bash
# Repeat the hybrid session-graph GPU test with launch tracing until it fails (loop-back flake hunt).
mkdir -p gpurun_out
for i in 1 2 3 4; do
  MLSL_TRACE_LAUNCH=1 MLSL_DEBUG_ERRORS=1 timeout 120 python -m pytest tests/test_collectives_gpu.py -q -x > gpurun_out/hybrid_$i.log 2>&1
  rc=$?
  echo "iteration $i rc=$rc $(tail -1 gpurun_out/hybrid_$i.log)"
  if [ $rc -ne 0 ]; then break; fi
done
