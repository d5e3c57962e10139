"""hvi_b200 -- B200-native negative-ELBO + gradient path of HybridVariationalInference.jl.

Host-side mirror (Python, since Julia is absent from this image) of the reference's interface
for that one path, on top of the C-ABI CUDA library `csrc/libhvi_b200.so`
(include/hvi_b200.h).  There is no CPU fallback: importing works without the library, calling
the path without it raises.
"""
from . import _abi
from .problem import (HybridProblem, DoubleMMCase, construct_problem_test_HybridProblem,
                      gen_hybridproblem_synthetic, draw_ξ)

__all__ = ["_abi", "HybridProblem", "DoubleMMCase", "construct_problem_test_HybridProblem",
           "gen_hybridproblem_synthetic", "draw_ξ"]
