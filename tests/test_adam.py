"""Device-resident Adam iterations (hvi_adam_*): the same trajectory as the host loop
`(nL, g) = withgradient(loss_elbo, ϕ); ϕ = Optimisers.Adam update` of src/HybridSolver.jl:256-260."""
import numpy as np
import pytest

import hvi_b200

pytestmark = pytest.mark.gpu


def host_adam(plan, ϕ, n_iter, n_MC, seed0, eta, b1=0.9, b2=0.999, eps=1e-8):
    ϕ = ϕ.astype(np.float64).copy()
    m, v = np.zeros_like(ϕ), np.zeros_like(ϕ)
    trace = []
    for it in range(n_iter):
        nL, comps, g = plan.neg_elbo_withgradient(ϕ.astype(plan.dt), n_MC=n_MC, seed=seed0 + it)
        g = g.astype(np.float64)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        ϕ = (ϕ.astype(plan.dt) - eta * (m / (1 - b1 ** (it + 1))) / (np.sqrt(v / (1 - b2 ** (it + 1))) + eps)).astype(plan.dt).astype(np.float64)
        trace.append(nL)
    return ϕ, np.array(trace)


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 2e-4)])
def test_adam_run_matches_host_loop(dtype, tol):
    prob = hvi_b200.construct_problem_test_HybridProblem(dtype=dtype)
    data = hvi_b200.gen_hybridproblem_synthetic(prob, 200, seed=3)
    plan = hvi_b200.NegElboPlan(prob, device=0)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    ϕ0 = prob.init_ϕ()
    want_ϕ, want_trace = host_adam(plan, ϕ0, 8, 3, 11, 0.02)
    plan.adam_init(ϕ0, eta=0.02)
    trace = plan.adam_run(8, 3, seed0=11)
    got = plan.adam_phi().astype(np.float64)
    assert np.all(np.isfinite(trace))
    assert np.max(np.abs(trace - want_trace)) <= tol * np.max(np.abs(want_trace))
    assert np.max(np.abs(got - want_ϕ)) <= tol * max(1.0, np.max(np.abs(want_ϕ)))


def test_adam_lowers_the_loss_and_needs_init():
    prob = hvi_b200.DoubleMMCase(dtype="f32")
    data = hvi_b200.gen_hybridproblem_synthetic(prob, 800)
    plan = hvi_b200.NegElboPlan(prob, device=0)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    with pytest.raises(Exception):
        plan.adam_run(1, 3)
    plan.adam_init(prob.init_ϕ(), eta=0.02)
    tr = plan.adam_run(300, 12, seed0=1)
    assert np.all(np.isfinite(tr))
    assert np.mean(tr[-20:]) < 0.5 * np.mean(tr[:20])
    assert plan.adam_run(5, 12, seed0=400, trace=False) is None     # nothing synchronised
    assert np.all(np.isfinite(plan.adam_phi()))
