#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/probe_blocking_torch.py > gpurun_out/probe_blocking_torch.txt 2>&1
MLSL_TRACE_LAUNCH=1 timeout 300 python -m pytest tests/test_zz_fuzz_gpu.py -x -q -m gpu -k "4-2" 2>&1 | grep -v "^  File\|^    " > gpurun_out/diag2_fuzz42.log
cat gpurun_out/probe_blocking_torch.txt
grep "^\[" gpurun_out/diag2_fuzz42.log | head -60
