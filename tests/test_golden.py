"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the
float64 oracle): the oracle and the C restatement must reproduce them on the CPU, the CUDA path
on the GPU (Float64 within rel 1e-10, Float32 within 1e-4)."""
import glob
import os

import numpy as np
import pytest

from helpers import O, c_oracle_value_and_grad, hvi_b200, oracle_spec

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
PROBLEMS = {
    "c1_test_hybridproblem": lambda dt: hvi_b200.construct_problem_test_HybridProblem(dtype=dt),
    "doublemm_default": lambda dt: hvi_b200.DoubleMMCase(dtype=dt),
    "doublemm_drivernan": lambda dt: hvi_b200.DoubleMMCase(dtype=dt),
    "doublemm_covark2": lambda dt: hvi_b200.DoubleMMCase(("covarK2",), dtype=dt),
}


def load(path):
    name = os.path.splitext(os.path.basename(path))[0]
    return name, dict(np.load(path))


def test_golden_files_present():
    assert len(FILES) == len(PROBLEMS)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracles_reproduce_golden(path):
    name, G = load(path)
    prob = PROBLEMS[name]("f64")
    nL, comps, g = O.value_and_grad(oracle_spec(prob), G["phi"], G["xM"], G["xP"], G["y_o"], G["y_unc"], G["zP"], G["zMs"])
    assert nL == pytest.approx(float(G["nL"]), rel=1e-13)
    np.testing.assert_allclose(g, G["grad"], rtol=1e-11, atol=1e-11 * np.abs(G["grad"]).max())
    nLc, compsc, gc = c_oracle_value_and_grad(prob, G["phi"], G, G["zP"], G["zMs"])
    assert nLc == pytest.approx(float(G["nL"]), rel=1e-12)
    np.testing.assert_allclose(compsc, G["comps"], rtol=1e-12, atol=1e-12 * np.abs(G["comps"]).max())
    np.testing.assert_allclose(gc, G["grad"], rtol=1e-10, atol=1e-11 * np.abs(G["grad"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_reproduces_golden(path, dtype):
    name, G = load(path)
    prob = PROBLEMS[name](dtype)
    tol = {"f64": 1e-10, "f32": 1e-4}[dtype]
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(G["xM"], G["xP"], G["y_o"], G["y_unc"])
    nL, comps, g = plan.neg_elbo_withgradient(G["phi"].astype(prob.np_dtype), G["zP"], G["zMs"])
    if dtype == "f32":   # the golden inputs are float64: compare against the oracle on the rounded inputs
        f = lambda a: np.asarray(a, dtype=np.float32)
        ref = O.value_and_grad(oracle_spec(prob), f(G["phi"]), f(G["xM"]), f(G["xP"]), f(G["y_o"]), f(G["y_unc"]),
                               f(G["zP"]), f(G["zMs"]), "f32")
    else:
        ref = (float(G["nL"]), G["comps"], G["grad"])
    assert abs(nL - ref[0]) <= tol * abs(ref[0])
    for nme, r in prob.positions().items():
        if len(r):
            assert np.max(np.abs(g[r] - ref[2][r])) <= tol * np.max(np.abs(ref[2][r])), nme
