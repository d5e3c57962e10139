"""hvi_b200 -- B200-native negative-ELBO + gradient path of HybridVariationalInference.jl.

Host-side mirror (Python, since Julia is absent from this image) of the reference's interface
for that one path, on top of the C-ABI CUDA library `csrc/libhvi_b200.so`
(include/hvi_b200.h).  There is no CPU fallback: importing works without the library, calling
the path without it raises.
"""
from . import _abi
from .problem import (HybridProblem, DoubleMMCase, construct_problem_test_HybridProblem,
                      gen_hybridproblem_synthetic, draw_ξ)
from .elbo import (NegElboPlan, NegElboComponents, HVIError, neg_elbo_gtf, neg_elbo_gtf_components,
                   get_loss_elbo, pbm_compile_check, vec2utri_pos, utri2vec_pos, get_cor_count, COMPONENT_NAMES)

__all__ = ["_abi", "HybridProblem", "DoubleMMCase", "construct_problem_test_HybridProblem",
           "gen_hybridproblem_synthetic", "draw_ξ", "NegElboPlan", "NegElboComponents", "HVIError",
           "neg_elbo_gtf", "neg_elbo_gtf_components", "get_loss_elbo", "pbm_compile_check", "vec2utri_pos", "utri2vec_pos",
           "get_cor_count", "COMPONENT_NAMES"]
