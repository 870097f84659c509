"""memilio_b200 -- B200-native backend for MEmilio's Graph-ODE ensemble hot path (host-side Python mirror).

The compute path is the CUDA library `memilio_b200/csrc/libmio_cuda.so` behind the C-ABI of include/mio_cuda.h.
There is no CPU fallback: every compute call raises if the library or a GPU is missing.
"""
from . import cdefs, flat, problem, study  # noqa: F401
from .problem import GraphProblem  # noqa: F401

__all__ = ["cdefs", "flat", "problem", "study", "GraphProblem"]
