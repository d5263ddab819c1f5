TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 200 $TR --master-port 29611 tests/mp_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | grep -v ": PASSED" | tail -8
timeout 300 $TR --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --nccl > gpurun_out/bench2.json 2> gpurun_out/bench2.err
python scripts/show_bench.py gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err
# torch.distributed "mlsl" backend with CUDA tensors (collectives, member-made sub-groups need >= 4 ranks, DDP, quantised hook)
timeout 300 $TR --master-port 29613 tests/torch_backend_worker.py env cuda 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6
timeout 200 $TR --master-port 29614 examples/torch_ddp.py --device cuda 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -4
