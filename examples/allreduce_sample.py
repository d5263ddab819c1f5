#!/usr/bin/env python
"""Smallest possible program: all-reduce 128 floats, each rank contributes its index, every element must become
(P - 1) * P / 2 (the reference's migration sample, mlsl_to_oneccl/mlsl_sample.cpp, in Python; the C++ version is
csrc/tests/mlsl_sample.cpp).        bin/mlslrun -n 4 python examples/allreduce_sample.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsl_b200 as mlsl  # noqa: E402

mlsl.init()
rank, world = mlsl.rank(), mlsl.world_size()
buf = mlsl.alloc_tensor(128, torch.float32)
buf.fill_(float(rank))
mlsl.allreduce(buf)                      # in place, SUM over the world, returns when the result may be used
if mlsl.is_device():
    torch.cuda.current_stream().synchronize()
expected = (world - 1) * world / 2
ok = bool((buf == expected).all())
print("[%d] %s" % (rank, "PASSED" if ok else "FAILED: got %s, expected %s" % (float(buf[0]), expected)), flush=True)
del buf
mlsl.finalize()
sys.exit(0 if ok else 1)
