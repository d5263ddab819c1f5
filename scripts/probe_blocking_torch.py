"""Which torch-level operations block the host while another stream runs a long kernel?  (Loop-back ranks dead-lock on
any of them, see csrc/tools/probe_blocking.cu for the raw CUDA calls.)  Usage: python scripts/probe_blocking_torch.py"""
import os
import sys
import time

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.cuda.set_device(0)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
SPIN = int(1.0 * 1.9e9)          # ~1 s of cycles
x = torch.zeros(1 << 20, device="cuda")
torch.cuda.synchronize()
host = torch.ones(1 << 18)
pinned = torch.ones(1 << 18).pin_memory()
pre_ev = torch.cuda.Event()
pre_ev.record()
torch.cuda.synchronize()
keep = []


def probe(name, fn):
    with torch.cuda.stream(sa):
        torch.cuda._sleep(SPIN)
    t0 = time.perf_counter()
    with torch.cuda.stream(sb):
        r = fn()
    dt = (time.perf_counter() - t0) * 1e3
    keep.append(r)
    torch.cuda.synchronize()
    print("%-56s %10.3f ms %s" % (name, dt, "  <-- BLOCKS" if dt > 400 else ""), flush=True)


probe("Event() + record (first new event)", lambda: (lambda e: (e.record(), e)[1])(torch.cuda.Event()))
probe("Event() + record x100", lambda: [(lambda e: (e.record(), e)[1])(torch.cuda.Event()) for _ in range(100)])
probe("Event(enable_timing) + record", lambda: (lambda e: (e.record(), e)[1])(torch.cuda.Event(enable_timing=True)))
probe("existing event record", lambda: pre_ev.record())
probe("event query", lambda: pre_ev.query())
probe("zeros(1000) on device", lambda: torch.zeros(1000, device="cuda"))
probe("x.clone()", lambda: x.clone())
probe("pageable.to(cuda) 1 MB", lambda: host.to("cuda"))
probe("pinned.to(cuda, non_blocking)", lambda: pinned.to("cuda", non_blocking=True))
probe("x[:1000].cpu() (sync own stream)", lambda: x[:1000].cpu())
probe("empty(256 MB) (cudaMalloc)", lambda: torch.empty(64 << 20, device="cuda"))
probe("x.sum().item()", lambda: x.sum().item())
probe("torch.cuda.Stream()", lambda: torch.cuda.Stream())
probe("x.double() (new dtype kernel)", lambda: x.double())
probe("torch.randn (first RNG kernel)", lambda: torch.randn(1000, device="cuda"))
probe("matmul bf16 512^3 (first cuBLAS call)", lambda: torch.randn(512, 512, device="cuda", dtype=torch.bfloat16) @ torch.randn(512, 512, device="cuda", dtype=torch.bfloat16))
probe("matmul again", lambda: torch.randn(512, 512, device="cuda", dtype=torch.bfloat16) @ torch.randn(512, 512, device="cuda", dtype=torch.bfloat16))
probe("del big tensor + empty_cache (cudaFree)", lambda: (keep.clear(), torch.cuda.empty_cache()))

# the library's own calls
os.environ["MLSL_BACKEND"] = "cuda"
os.environ["MLSL_HEAP_SIZE_GB"] = "0.5"
import mlsl_b200 as mlsl

with torch.cuda.stream(sb):
    env = mlsl.init()
    t = mlsl.alloc_tensor(4096, torch.float32)
torch.cuda.synchronize()


def lib_alloc_free():
    a = mlsl.alloc_tensor(1 << 16, torch.float32)
    del a
    return mlsl.alloc_tensor(1 << 10, torch.float32)


probe("mlsl.alloc_tensor + drop + alloc", lib_alloc_free)
probe("mlsl.allreduce (1 rank, out of place)", lambda: mlsl.allreduce(t, out=mlsl.alloc_tensor(4096, torch.float32)))
probe("mlsl.allreduce foreign tensor", lambda: mlsl.allreduce(x[:4096].clone()))
probe("mlsl.allgather", lambda: mlsl.allgather(t))
mlsl.finalize()
