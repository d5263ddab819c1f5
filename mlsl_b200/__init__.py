"""mlsl_b200: a Blackwell-native deep-learning collective library with the capabilities of Intel MLSL.

    import mlsl_b200 as mlsl
    mlsl.init()                                   # RANK / WORLD_SIZE / LOCAL_RANK from the environment (torchrun)
    g = mlsl.alloc_tensor(n, torch.float32)       # lives in the symmetric, peer-mapped heap
    mlsl.allreduce(g, scale=1.0 / mlsl.world_size())
    opt = mlsl.DistributedOptimizer(model.parameters(), lr=0.1)

The object model of the reference (Environment / Session / Distribution / Operation / ...) is in `mlsl_b200.api`.
"""
import os as _os

# Several ranks may share one GPU (in-process loopback used by the tests and smoke()): their kernels spin on each
# other, so every stream needs its own hardware queue - the default of 8 connections makes unrelated streams
# serialise behind a waiting kernel.  Must be set before the CUDA context exists; a user setting wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib, api  # noqa: E402
from ._lib import MLSLError  # noqa: E402
from .api import (CompressionType, DataType, GroupType, InprocWorld, MLSL, OperationType, OptimizerType, PhaseType,
                  ReductionType, cuda_available)
from .comm import (Work, alloc_tensor, allgather, allgatherv, allreduce, alltoall, alltoallv, barrier, bcast, bind_thread_state, env, finalize,
                   free_tensor, gather, heap_pool, init, is_device, is_initialized, rank, reduce, reduce_scatter, ring_shift, scatter, tensor_from_address,
                   world_distribution, world_size)

__version__ = "2026.1"


def __getattr__(name):
    # heavier modules are imported on demand
    if name == "DistributedOptimizer":
        from .optim import DistributedOptimizer
        return DistributedOptimizer
    if name == "DistributedDataParallel":
        from .parallel.data_parallel import DistributedDataParallel
        return DistributedDataParallel
    raise AttributeError(name)
