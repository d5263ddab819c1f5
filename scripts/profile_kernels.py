"""Single-GPU driver for ncu: runs the collective kernels with a 1-rank group (MLSL_FORCE_KERNEL_SOLO=1), so each
kernel completes on its own (ncu serialises kernels; multi-rank kernels wait for each other and cannot be captured).
usage: python scripts/profile_kernels.py [bytes]"""
import os
import sys

os.environ["MLSL_FORCE_KERNEL_SOLO"] = "1"
os.environ.setdefault("MLSL_BACKEND", "cuda")
os.environ.setdefault("MLSL_HEAP_SIZE_GB", "3")
os.environ.setdefault("MLSL_STREAM_MODE", "inline")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mlsl_b200 as mlsl  # noqa: E402

nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 256 << 20
torch.cuda.set_device(0)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
mlsl.init()
n = nbytes // 4
x = mlsl.alloc_tensor(n, torch.float32)
y = mlsl.alloc_tensor(n, torch.float32)
x.fill_(1.0)
for _ in range(3):
    mlsl.allreduce(x, out=y, scale=0.5, group="global")                    # k_allreduce<float, OpSum, 4>
xb = mlsl.alloc_tensor(n, torch.bfloat16)
yb = mlsl.alloc_tensor(n, torch.bfloat16)
for _ in range(2):
    mlsl.allreduce(xb, out=yb, group="global")                             # k_allreduce<bf16>
for _ in range(2):
    mlsl.allreduce(x, out=y, compress=True, group="global")                # k_allreduce_quant
sh = mlsl.reduce_scatter(x, group="global")                                # k_reduce_pull
full = mlsl.allgather(sh, group="global")                                  # k_pull_copy
torch.cuda.synchronize()
print("ok", float(y[0]), float(full[0]))
mlsl.finalize()
