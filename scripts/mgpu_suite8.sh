#!/bin/bash
# the N-GPU session (N = 4 or 8): correctness, tuning sweep of the large all-reduce, PCIe ceiling, bench (both arms), other collectives
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 500 $TR --master-port 29701 tests/mp_gpu_check.py > gpurun_out/mp_check_$N.log 2>&1; echo "mp_check rc=$?" | tee -a gpurun_out/mp_check_$N.log
timeout 300 $TR --master-port 29706 bench/tune_allreduce.py > gpurun_out/tune_ar_$N.jsonl 2> gpurun_out/tune_ar_$N.err
timeout 120 $TR --master-port 29702 bench/pcie_probe.py > gpurun_out/pcie_$N.json 2> gpurun_out/pcie_$N.err
timeout 600 $TR --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench rc=$?" >> gpurun_out/bench_$N.err
timeout 400 $TR --master-port 29705 bench/collectives_bench.py --max-mb 256 > gpurun_out/coll_$N.jsonl 2> gpurun_out/coll_$N.err
timeout 500 python bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_ref_$N.json 2> gpurun_out/bench_ref_$N.err
grep -c PASSED gpurun_out/mp_check_$N.log; grep FAILED gpurun_out/mp_check_$N.log | head; tail -2 gpurun_out/mp_check_$N.log
grep BEST gpurun_out/tune_ar_$N.jsonl; cat gpurun_out/pcie_$N.json
python scripts/show_bench.py gpurun_out/bench_$N.json; cut -c1-300 gpurun_out/bench_ref_$N.json; tail -3 gpurun_out/bench_$N.err
