"""Double-buffered batch registration (hvi_plan_prefetch_data / hvi_plan_commit_data): the prefetched batch must
not disturb evaluations of the current one, and after the commit results are bit-identical to set_data."""
import numpy as np
import pytest

import hvi_b200

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_prefetch_commit_equals_set_data(dtype):
    prob = hvi_b200.DoubleMMCase(dtype=dtype)
    M = 8
    ϕ = prob.init_ϕ(perturbed=True)
    batches = [hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=100 + nb) for nb in (300, 517, 517, 64)]
    ref = hvi_b200.NegElboPlan(prob, device=0)
    want = []
    for b in batches:
        ref.set_data(b["xM"], b["xP"], b["y_o"], b["y_unc"])
        want.append(ref.neg_elbo_withgradient(ϕ, n_MC=M, seed=7))
    plan = hvi_b200.NegElboPlan(prob, device=0)
    plan.set_data(*(batches[0][k] for k in ("xM", "xP", "y_o", "y_unc")))
    for i in range(len(batches)):
        if i + 1 < len(batches):
            b = batches[i + 1]
            plan.prefetch_data(b["xM"], b["xP"], b["y_o"], b["y_unc"])
        nL, comps, g = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=7)   # still batch i
        assert nL == want[i][0]
        assert np.array_equal(g, want[i][2])
        if i + 1 < len(batches):
            plan.commit_data()
    # a plain set_data drops a pending prefetch
    plan.prefetch_data(*(batches[1][k] for k in ("xM", "xP", "y_o", "y_unc")))
    plan.set_data(*(batches[0][k] for k in ("xM", "xP", "y_o", "y_unc")))
    nL, comps, g = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=7)
    assert nL == want[0][0]
    with pytest.raises(Exception):
        plan.commit_data()


def test_prefetch_commit_with_site_indices():
    """MeanVarSep batches carry their global site indices: commit_data(i_sites) must behave like set_data(i_sites)"""
    prob = hvi_b200.DoubleMMCase(("sepvar", "few_sites"), dtype="f64")
    nb, M = 20, 4
    ϕ = prob.init_ϕ(perturbed=True)
    b1 = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=1)
    b2 = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=2)
    i1, i2 = np.arange(0, 20), np.arange(60, 80)[::-1].copy()
    ref = hvi_b200.NegElboPlan(prob, device=0)
    ref.set_data(b2["xM"], b2["xP"], b2["y_o"], b2["y_unc"], i_sites=i2)
    want = ref.neg_elbo_withgradient(ϕ, n_MC=M, seed=3)
    plan = hvi_b200.NegElboPlan(prob, device=0)
    plan.set_data(b1["xM"], b1["xP"], b1["y_o"], b1["y_unc"], i_sites=i1)
    plan.prefetch_data(b2["xM"], b2["xP"], b2["y_o"], b2["y_unc"])
    plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=3)
    plan.commit_data(i_sites=i2)
    got = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=3)
    assert got[0] == want[0] and np.array_equal(got[2], want[2])


def test_prefetch_on_a_fresh_plan():
    prob = hvi_b200.DoubleMMCase(dtype="f32")
    b = hvi_b200.gen_hybridproblem_synthetic(prob, 97, seed=5)
    ϕ = prob.init_ϕ(perturbed=True)
    ref = hvi_b200.NegElboPlan(prob, device=0)
    ref.set_data(b["xM"], b["xP"], b["y_o"], b["y_unc"])
    want = ref.neg_elbo_withgradient(ϕ, n_MC=8, seed=2)
    plan = hvi_b200.NegElboPlan(prob, device=0)
    plan.prefetch_data(b["xM"], b["xP"], b["y_o"], b["y_unc"])
    plan.commit_data()
    got = plan.neg_elbo_withgradient(ϕ, n_MC=8, seed=2)
    assert got[0] == want[0] and np.array_equal(got[2], want[2])
    # prediction-only batch through the same route
    plan.prefetch_data(b["xM"], b["xP"])
    plan.commit_data()
    y, θP, θMs, ent = plan.predict_hvi(ϕ, n_sample_pred=8, seed=3)
    ref.set_data(b["xM"], b["xP"])
    y2, _, θMs2, ent2 = ref.predict_hvi(ϕ, n_sample_pred=8, seed=3)
    assert np.array_equal(y, y2) and np.array_equal(θMs, θMs2) and ent == ent2


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_broadcast_registration_equals_full_arrays(dtype):
    """drivers and observation uncertainty shared by all sites given as one column (hvi_plan_set_data_bcast): identical
    results to the full (rows x n_site) arrays, for the value + gradient and for the prediction path"""
    prob = hvi_b200.DoubleMMCase(dtype=dtype)
    nb, M = 1234, 8
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=6)         # DoubleMM: same drivers for every site
    assert np.all(data["xP"] == data["xP"][:, :1]) and np.all(data["y_unc"] == data["y_unc"][:, :1])
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    full = hvi_b200.NegElboPlan(prob)
    full.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    ref = full.neg_elbo_withgradient(ϕ, zP, zMs)
    for xs, us in ((True, True), (True, False), (False, True)):
        plan = hvi_b200.NegElboPlan(prob)
        plan.set_data_bcast(data["xM"], data["xP"][:, :1] if xs else data["xP"], data["y_o"],
                            data["y_unc"][:, 0] if us else data["y_unc"])
        got = plan.neg_elbo_withgradient(ϕ, zP, zMs)
        assert got[0] == ref[0] and np.array_equal(got[2], ref[2])
        assert np.array_equal(plan.predict_hvi(ϕ, zP, zMs)[0], full.predict_hvi(ϕ, zP, zMs)[0])
