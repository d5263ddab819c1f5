"""Fused compute + collective ops (hand-written sm_100a kernels, see csrc/cuda)."""
from .ag_gemm import allgather_gemm
from .gemm_rs import gemm_reduce_scatter

__all__ = ["gemm_reduce_scatter", "allgather_gemm"]
