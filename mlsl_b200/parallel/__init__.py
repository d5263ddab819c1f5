"""Parallelism helpers built on Distribution / Session: data parallel (gradient all-reduce or fused distributed
update), the (data x model) hybrid groups of the reference, and tensor-parallel linear layers whose partial sums are
reduced by the fused GEMM + reduce-scatter kernel; sequence_parallel, pipeline_parallel and multinode are imported on
demand (`from mlsl_b200.parallel import pipeline_parallel`)."""
from .data_parallel import DistributedDataParallel, broadcast_parameters

__all__ = ["DistributedDataParallel", "broadcast_parameters"]
