#!/bin/bash
# multi-GPU session: correctness (one rank per GPU), the PCIe ceiling, bench.py (both arms), the other collectives vs NCCL
# usage: scripts/mgpu_suite.sh N [quick]
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 600 $TR --master-port 29701 tests/mp_gpu_check.py > gpurun_out/mp_check_$N.log 2>&1; echo "mp_check rc=$?" | tee -a gpurun_out/mp_check_$N.log
timeout 200 $TR --master-port 29702 bench/pcie_probe.py > gpurun_out/pcie_$N.json 2> gpurun_out/pcie_$N.err
timeout 200 $TR --master-port 29703 bench/pcie_probe.py --no-bind > gpurun_out/pcie_nobind_$N.json 2>> gpurun_out/pcie_$N.err
timeout 900 $TR --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench rc=$?" >> gpurun_out/bench_$N.err
timeout 600 python bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_ref_$N.json 2> gpurun_out/bench_ref_$N.err
timeout 600 $TR --master-port 29705 bench/collectives_bench.py --max-mb 256 > gpurun_out/coll_$N.jsonl 2> gpurun_out/coll_$N.err
tail -3 gpurun_out/mp_check_$N.log; cat gpurun_out/pcie_$N.json gpurun_out/pcie_nobind_$N.json; cut -c1-400 gpurun_out/bench_$N.json; cut -c1-300 gpurun_out/bench_ref_$N.json; tail -3 gpurun_out/bench_$N.err
