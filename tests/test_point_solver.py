"""Point-solver loss loss_gf (src/gf.jl:227-295): oracle restatement checked by finite differences on the CPU, the
CUDA path (hvi_loss_gf_grad) against the oracle on the GPU -- Float64 rel 1e-10, Float32 rel 1e-4."""
import numpy as np
import pytest

from helpers import O, hvi_b200, oracle_spec

CASES = {
    "C1": lambda dt: hvi_b200.construct_problem_test_HybridProblem(dtype=dt),
    "default": lambda dt: hvi_b200.DoubleMMCase(dtype=dt),
    "covarK2": lambda dt: hvi_b200.DoubleMMCase(("covarK2",), dtype=dt),
    "omit_r0": lambda dt: hvi_b200.DoubleMMCase(("omit_r0",), dtype=dt),
    "no_globals": lambda dt: hvi_b200.DoubleMMCase(("no_globals",), dtype=dt),
    "transIdent": lambda dt: hvi_b200.DoubleMMCase(("transIdent",), dtype=dt),
    "rangescaling": lambda dt: hvi_b200.DoubleMMCase(("use_rangescaling",), dtype=dt),
}


def phi_gf_of(prob, scale=1.0):
    ϕ = prob.init_ϕ(perturbed=True)
    pos = prob.positions()
    ϕP = ϕ[pos["μP"]] + 0.05 * scale
    return np.concatenate([ϕ[pos["ϕg"]], ϕP]).astype(np.float64)


def test_oracle_loss_gf_gradient_by_finite_differences():
    prob = CASES["covarK2"]("f64")
    spec = oracle_spec(prob)
    d = hvi_b200.gen_hybridproblem_synthetic(prob, 40, seed=2)
    ϕ = phi_gf_of(prob)
    nL, comps, g = O.loss_gf_value_and_grad(spec, ϕ, d["xM"], d["xP"], d["y_o"], d["y_unc"])
    assert comps[0] == pytest.approx(comps[1] + comps[2], rel=1e-14)
    rng = np.random.default_rng(0)
    for i in list(rng.choice(prob.n_phig, 12, replace=False)) + list(range(prob.n_phig, ϕ.size)):
        h = 1e-6 * max(1.0, abs(ϕ[i]))
        e = np.zeros_like(ϕ); e[i] = h
        fp = O.loss_gf_value_and_grad(spec, ϕ + e, d["xM"], d["xP"], d["y_o"], d["y_unc"])[0]
        fm = O.loss_gf_value_and_grad(spec, ϕ - e, d["xM"], d["xP"], d["y_o"], d["y_unc"])[0]
        assert (fp - fm) / (2 * h) == pytest.approx(g[i], rel=2e-5, abs=1e-6 * np.max(np.abs(g)))


def test_oracle_loss_gf_is_the_elbo_likelihood_at_the_mean():
    """with ξ = 0 the draws of the ELBO path collapse onto the point prediction: nLy and the prior terms agree
    (θ far from the clamp)"""
    prob = CASES["default"]("f64")
    spec = oracle_spec(prob)
    d = hvi_b200.gen_hybridproblem_synthetic(prob, 30, seed=4)
    ϕ = prob.init_ϕ(perturbed=True)
    pos = prob.positions()
    ϕ_gf = np.concatenate([ϕ[pos["ϕg"]], ϕ[pos["μP"]]])
    zP, zMs = np.zeros((2, prob.n_P)), np.zeros((2, prob.n_M * 30))
    _, c_elbo, _ = O.value_and_grad(spec, ϕ, d["xM"], d["xP"], d["y_o"], d["y_unc"], zP, zMs)
    _, c_pt, _ = O.loss_gf_value_and_grad(spec, ϕ_gf, d["xM"], d["xP"], d["y_o"], d["y_unc"])
    assert c_pt[1] == pytest.approx(c_elbo[3], rel=1e-13)
    assert c_pt[2] == pytest.approx(c_elbo[4], rel=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 1e-4)])
@pytest.mark.parametrize("case", sorted(CASES))
def test_loss_gf_cuda_matches_oracle(case, dtype, tol):
    prob = CASES[case](dtype)
    spec = oracle_spec(prob)
    nb = 333
    d = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=9, driver_nan_frac=0.1 if case == "default" else 0.0)
    ϕ = phi_gf_of(prob).astype(prob.np_dtype)
    plan = hvi_b200.NegElboPlan(prob, device=0)
    plan.set_data(d["xM"], d["xP"], d["y_o"], d["y_unc"])
    nL, comps, g = plan.loss_gf_withgradient(ϕ)
    nL_o, comps_o, g_o = O.loss_gf_value_and_grad(spec, ϕ.astype(np.float64), d["xM"], d["xP"], d["y_o"], d["y_unc"])
    assert abs(nL - nL_o) <= tol * abs(nL_o)
    assert np.max(np.abs(comps - comps_o)) <= tol * np.max(np.abs(comps_o))
    ng = prob.n_phig
    for name, sl in (("ϕg", slice(0, ng)), ("ϕP", slice(ng, ng + prob.n_P))):
        if g_o[sl].size:
            assert np.max(np.abs(g[sl] - g_o[sl])) <= tol * np.max(np.abs(g_o[sl])), name
    nL2, _, none = plan.loss_gf_withgradient(ϕ, want_grad=False)
    assert nL2 == nL and none is None
