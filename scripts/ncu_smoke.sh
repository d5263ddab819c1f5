#!/bin/bash
# profiler evidence on ONE GPU: every kernel smoke() launches with its device time, and full captures of the headline
# (single-GPU) kernel and of the solo-mode collective kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1
echo "ncu_rc=$?" >> gpurun_out/ncu_smoke.log
ncu --set full --clock-control none --import-source on -k regex:"k_scale_copy|k_allreduce_quant|k_pull_copy|k_gemm_rs" -c 6 -o gpurun_out/prof_r2 python -c "
import os
os.environ['MLSL_FORCE_KERNEL_SOLO']='1'; os.environ['MLSL_BACKEND']='cuda'; os.environ['MLSL_HEAP_SIZE_GB']='1.5'
import torch, mlsl_b200 as mlsl
from mlsl_b200.ops.gemm_rs import gemm_reduce_scatter
torch.cuda.set_device(0); s=torch.cuda.Stream(); torch.cuda.set_stream(s)
mlsl.init()
n=(256<<20)//4
x=mlsl.alloc_tensor(n, torch.float32, zero=False); y=mlsl.alloc_tensor(n, torch.float32, zero=False); x.fill_(1.0)
os.environ.pop('MLSL_FORCE_KERNEL_SOLO')
mlsl.allreduce(x, out=y, scale=0.5)                       # k_scale_copy: the N=1 headline kernel
q=(16<<20)//4
mlsl.allreduce(x[:q], out=y[:q], compress=True)            # k_allreduce_quant (solo)
mlsl.allgather(x[:q], out=y[:q])                           # k_pull_copy(_bulk) (solo)
a=(torch.randn(4096,4096,device='cuda')*0.1).bfloat16(); w=(torch.randn(4096,4096,device='cuda')*0.1).bfloat16()
o=gemm_reduce_scatter(a,w,group='data')                    # k_gemm_rs / k_gemm_rs2
torch.cuda.synchronize(); mlsl.finalize()
" > gpurun_out/ncu_full.log 2>&1
echo "ncu_full_rc=$?" >> gpurun_out/ncu_full.log
tail -3 gpurun_out/ncu_smoke.log; tail -3 gpurun_out/ncu_full.log; grep -c "k_" gpurun_out/launches_smoke.csv
