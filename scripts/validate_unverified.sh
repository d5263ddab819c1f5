#!/bin/bash
# First GPU call of the next round: everything docs/NEXT_STEPS.md lists as "never run on hardware", in order of risk, each
# step under its own timeout and with its own log in gpurun_out/ so one hang or failure does not hide the rest.
#   gpurun --timeout 1500 -- 'bash scripts/validate_unverified.sh one'          (1 GPU,  ~12 min)
#   gpurun --gpus 4 --timeout 900 -- 'bash scripts/validate_unverified.sh multi 4'
mkdir -p gpurun_out
export CUDA_MODULE_LOADING=EAGER
step() {   # step <name> <timeout> <command...>
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/validate_summary.txt
  timeout "$t" "$@" > "gpurun_out/v_$name.txt" 2>&1
  local rc=$?
  echo "$name rc=$rc : $(grep -v '^$' gpurun_out/v_$name.txt | tail -1 | cut -c1-200)" | tee -a gpurun_out/validate_summary.txt
}
if [ "${1:-one}" = "one" ]; then
  # 1. the default path as the driver runs it
  step pytest_gpu 900 python -m pytest tests -m gpu -q
  step smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
  step bench1 600 python bench.py --gpus 1 --steps 10 --warmup 3
  grep -o '"value": [0-9.]*\|"e2e": {"value": [0-9.]*\|"correct": [a-z]*' gpurun_out/v_bench1.txt | tr '\n' ' ' | tee -a gpurun_out/validate_summary.txt; echo
  # 2. device paths that exist only behind examples
  step python_twin_cuda 300 env MLSL_BACKEND=cuda python examples/mlsl_test.py 2 1 --inproc 4
  # 3. opt-in tests (new strategies on device tensors, CTA-pair GEMM on several ranks, all-gather + GEMM)
  step strategies 300 env MLSL_TEST_STRATEGIES_GPU=1 python -m pytest tests/test_zz_strategies_gpu.py -m gpu -q
  step gemm_2cta 300 env MLSL_TEST_2CTA=1 python -m pytest tests/test_gemm_rs_gpu.py -m gpu -q -k two_cta
  step ag_gemm 300 env MLSL_TEST_AGGEMM=1 python -m pytest tests/test_tensor_parallel.py -m gpu -q -k fused_gpu
  # 4. one ncu capture of the kernel behind the N = 1 headline
  step ncu_scale_copy 300 ncu --set full --clock-control none --import-source on -k regex:k_scale_copy -c 1 \
       -o gpurun_out/prof_scale_copy python bench.py --gpus 1 --steps 1 --warmup 1 --no-sweep --no-e2e
else
  N=${2:-2}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  step mp_gpu_check 300 $TR --master-port 29611 tests/mp_gpu_check.py
  step bench$N 400 $TR --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --nccl
  step torch_backend 300 $TR --master-port 29613 tests/torch_backend_worker.py env cuda
  step torch_ddp 200 $TR --master-port 29614 examples/torch_ddp.py --device cuda
  step torch_bench_mlsl 300 $TR --master-port 29615 bench/torch_backend_bench.py --backend mlsl --device cuda --max-mb 256
  step torch_bench_nccl 300 $TR --master-port 29616 bench/torch_backend_bench.py --backend nccl --device cuda --max-mb 256
  step pipeline 200 $TR --master-port 29617 examples/train_pipeline_parallel.py --stages 2
  step collectives_vs_nccl 400 $TR --master-port 29618 bench/collectives_bench.py --max-mb 256
  step example_cuda 120 bin/mlslrun -n $N bin/mlsl_example_cuda
  if [ "$N" -ge 4 ]; then   # two "nodes" of N/2 GPUs each on one box: two-level collectives with NCCL between them
    H=$((N / 2))
    (CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((H - 1))) timeout 300 python -m torch.distributed.run --nnodes 2 --node-rank 0 --nproc-per-node $H \
       --master-addr 127.0.0.1 --master-port 29700 tests/multinode_worker.py cuda > gpurun_out/v_multinode0.txt 2>&1 &)
    CUDA_VISIBLE_DEVICES=$(seq -s, $H $((N - 1))) step multinode1 300 python -m torch.distributed.run --nnodes 2 --node-rank 1 \
       --nproc-per-node $H --master-addr 127.0.0.1 --master-port 29700 tests/multinode_worker.py cuda
  fi
fi
echo; cat gpurun_out/validate_summary.txt
