#!/bin/bash
# One-GPU check run on the B200 box: probe, GPU tests, smoke, short bench.  Output -> gpurun_out/
mkdir -p gpurun_out
python scripts/probe_gpu.py > gpurun_out/probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
tail -3 gpurun_out/smoke.txt
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
