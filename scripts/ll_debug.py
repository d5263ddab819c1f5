"""Loop-back diagnosis of the LL all-reduce with many in-process ranks on one GPU: per call wall time + correctness.

usage: python scripts/ll_debug.py [world] [inline|default]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from conftest import run_ranks  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "default"


def body(r, mlsl):
    log = []
    for it in range(6):
        x = torch.full((4,), float(r + 1), device="cuda")
        y = torch.empty_like(x)
        t0 = time.time()
        mlsl.allreduce(x, out=y)
        torch.cuda.current_stream().synchronize()
        t1 = time.time()
        h = mlsl.alloc_tensor(4, torch.float32)
        h.copy_(x)
        mlsl.allreduce(h)
        torch.cuda.current_stream().synchronize()
        t2 = time.time()
        log.append((it, round(t1 - t0, 3), y[0].item(), round(t2 - t1, 3), h[0].item()))
    return log


env = {"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "5", "MLSL_DEBUG_ERRORS": "1"}
if mode == "inline":
    env["MLSL_STREAM_MODE"] = "inline"
try:
    outs = run_ranks(world, body, backend="cuda", env=env)
    want = world * (world + 1) / 2
    for r, o in enumerate(outs):
        print("rank", r, "want", want, o)
except BaseException as e:  # noqa: BLE001
    print("FAILED:", repr(e)[:500])
