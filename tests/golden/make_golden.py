"""Generates tests/golden/*.npz: small input/output vectors of the neg-ELBO path.

The reference (Julia) cannot run in this image, so these vectors come from the float64 oracle
(oracle/hvi_oracle.py, a restatement pinned component-wise against the reference's known-answer
tests -- tests/test_oracle_pins.py).  They freeze the oracle: a later change of the oracle or of
the CUDA path that moves these numbers is caught.  julia/dump_golden.jl writes the same file
layout from the real package on a host that has Julia; drop its output here to replace them.

Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from helpers import O, hvi_b200, oracle_spec  # noqa: E402

CASES = {
    "c1_test_hybridproblem": (lambda: hvi_b200.construct_problem_test_HybridProblem(dtype="f64"), 20, 3, 0.0),
    "doublemm_default": (lambda: hvi_b200.DoubleMMCase(dtype="f64"), 20, 3, 0.0),
    "doublemm_drivernan": (lambda: hvi_b200.DoubleMMCase(dtype="f64"), 24, 4, 0.25),
    "doublemm_covark2": (lambda: hvi_b200.DoubleMMCase(("covarK2",), dtype="f64"), 12, 3, 0.0),
}


def main():
    for name, (mk, nb, M, nanf) in CASES.items():
        prob = mk()
        data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, driver_nan_frac=nanf)
        zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
        ϕ = prob.init_ϕ(perturbed=True)
        nL, comps, g = O.value_and_grad(oracle_spec(prob), ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), phi=ϕ, xM=data["xM"], xP=data["xP"], y_o=data["y_o"],
                            y_unc=data["y_unc"], zP=zP, zMs=zMs, nL=nL, comps=comps, grad=g)
        print(name, nL)


if __name__ == "__main__":
    main()
