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


CASES_FUSED = {
    "DoubleMM_default": lambda dt: hvi_b200.DoubleMMCase(dtype=dt),
    "C1_test_HybridProblem": lambda dt: hvi_b200.construct_problem_test_HybridProblem(dtype=dt),
    "omit_r0": lambda dt: hvi_b200.DoubleMMCase(("omit_r0",), dtype=dt),
    "no_globals": lambda dt: hvi_b200.DoubleMMCase(("no_globals",), dtype=dt),
    "K1global": lambda dt: hvi_b200.DoubleMMCase(("omit_r0", "K1global"), dtype=dt),
    "transIdent_rangescaling": lambda dt: hvi_b200.DoubleMMCase(("transIdent", "use_rangescaling"), dtype=dt),
}


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-11), ("f32", 3e-5)])
@pytest.mark.parametrize("case", list(CASES_FUSED))
@pytest.mark.parametrize("nb,M", [(20, 3), (1, 1), (333, 12), (1000, 16)])
def test_single_launch_evaluation_equals_the_multi_kernel_path(case, nb, M, dtype, tol):
    """the reference's operating point (20 sites, n_MC = 3; src/DoubleMM/f_doubleMM.jl:315-325) runs as ONE launch
    (k_small_iter): same value, components and gradient as the eight-kernel path on the same device draws, up to the
    summation order of the partial sums"""
    prob = CASES_FUSED[case](dtype)
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=12, driver_nan_frac=0.2 if nb > 20 else 0.0)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    if nb <= 48:      # the default applies the break-even: the reference's batch size takes the single launch by itself
        n0 = plan.launch_count
        plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=77)
        assert plan.launch_count - n0 == 1
    else:             # a batch of hundreds of sites is faster as eight launches of the whole GPU
        n0 = plan.launch_count
        plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=77)
        assert plan.launch_count - n0 >= 6
    plan.use_fused(2)   # force the single launch wherever it is supported
    n0 = plan.launch_count
    nL, comps, g = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=77)
    assert plan.launch_count - n0 == 1
    plan.use_fused(False)
    n0 = plan.launch_count
    nL2, comps2, g2 = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=77)
    assert plan.launch_count - n0 >= 6
    assert abs(nL - nL2) <= tol * abs(nL2)
    assert np.max(np.abs(np.array(comps) - np.array(comps2))) <= tol * np.max(np.abs(comps2))
    for name, r in prob.positions().items():
        if len(r):
            assert np.max(np.abs(g[r].astype(np.float64) - g2[r])) <= tol * np.max(np.abs(g2[r])), name
    nL3, _, g3 = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=77, want_grad=False)
    assert g3 is None and nL3 == nL2


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 3e-4)])
def test_single_launch_optimiser_loop_equals_the_multi_kernel_loop(dtype, tol):
    prob = hvi_b200.DoubleMMCase(dtype=dtype)
    data = hvi_b200.gen_hybridproblem_synthetic(prob, 800, seed=3)
    ϕ0 = prob.init_ϕ()
    res = []
    for fused in (True, False):
        plan = hvi_b200.NegElboPlan(prob)
        plan.use_fused(fused)
        plan.set_dataset(data["xM"], data["xP"], data["y_o"], data["y_unc"])
        plan.adam_init(ϕ0, eta=0.02)
        n0 = plan.launch_count
        tr = plan.adam_run_minibatches(45, 3, 20, seed0=5)           # 40 minibatches per epoch: wraps around
        per_iter = (plan.launch_count - n0) / 45
        plan.select_range(0, 20)
        tr_b = plan.adam_run(10, 12, seed0=500)                        # the loop on one registered batch
        res.append((tr, tr_b, plan.adam_phi().astype(np.float64), per_iter))
    (tr, tr_b, ϕa, la), (tr2, tr2_b, ϕb, lb) = res
    assert la < 2 and lb >= 8                                          # one launch (+ the counter reset) vs >= 8 per iteration
    assert np.all(np.isfinite(tr)) and np.all(np.isfinite(tr_b))
    assert np.max(np.abs(tr - tr2)) <= tol * np.max(np.abs(tr2))
    assert np.max(np.abs(tr_b - tr2_b)) <= tol * np.max(np.abs(tr2_b))
    assert np.max(np.abs(ϕa - ϕb)) <= tol * max(1.0, np.max(np.abs(ϕb)))
