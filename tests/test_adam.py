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


def test_sharded_adam_steps_equal_unsharded():
    """two site shards (two plans on one device stand in for two ranks; their result vectors are added where NCCL
    would all-reduce them) walk the same trajectory as one plan over all sites"""
    import torch
    from hvi_b200.shard import adam_step_sharded, site_range
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    nb, M, steps = 301, 4, 6
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=8)
    ϕ0 = prob.init_ϕ(perturbed=True)
    whole = hvi_b200.NegElboPlan(prob, device=0)
    whole.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    whole.adam_init(ϕ0, eta=0.01)
    want_trace = whole.adam_run(steps, M, seed0=21)
    want = whole.adam_phi()
    dev = torch.device("cuda", 0)
    plans, offs, outs = [], [], []
    for r in range(2):
        lo, hi = site_range(nb, r, 2)
        p = hvi_b200.NegElboPlan(prob, device=0, count_globals=(r == 0))
        p.set_data(*(data[k][:, lo:hi] for k in ("xM", "xP", "y_o", "y_unc")))
        p.adam_init(ϕ0, eta=0.01)
        plans.append(p); offs.append(lo)
        outs.append(torch.zeros(prob.n_phi + 8, dtype=torch.float64, device=dev))

    trace = []
    for it in range(steps):
        for p, lo, o in zip(plans, offs, outs):
            p.neg_elbo_async(p.adam_phi_dev(), None, None, o, M, 0, seed=21 + it, site_offset=lo)
        torch.cuda.synchronize()
        total = outs[0] + outs[1]
        for p in plans:
            p.adam_apply_dev(total)
        torch.cuda.synchronize()
        trace.append(float(total[prob.n_phi + 6]))
    for p in plans:
        got = p.adam_phi()
        assert np.max(np.abs(got - want)) <= 1e-10 * max(1.0, np.max(np.abs(want)))
    assert np.max(np.abs(np.array(trace) - want_trace)) <= 1e-10 * np.max(np.abs(want_trace))
    assert callable(adam_step_sharded)
