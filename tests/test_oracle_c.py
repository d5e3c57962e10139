"""The C restatement with the hand-derived reverse pass (oracle/hvi_oracle.c; also the timed CPU
baseline) against the torch-autograd oracle: two independent derivations of the same gradient.
CPU only."""
import numpy as np
import pytest

from helpers import O, c_oracle_value_and_grad, hvi_b200, oracle_spec

CASES = {
    "C1": lambda dt: hvi_b200.construct_problem_test_HybridProblem(dtype=dt),
    "C1_covarK2": lambda dt: hvi_b200.construct_problem_test_HybridProblem(("covarK2",), dtype=dt),
    "default": lambda dt: hvi_b200.DoubleMMCase(dtype=dt),
    "covarK2": lambda dt: hvi_b200.DoubleMMCase(("covarK2",), dtype=dt),
    "omit_r0": lambda dt: hvi_b200.DoubleMMCase(("omit_r0",), dtype=dt),
    "K1global_covar": lambda dt: hvi_b200.DoubleMMCase(("omit_r0", "K1global", "covarK2"), dtype=dt),
    "no_globals": lambda dt: hvi_b200.DoubleMMCase(("no_globals",), dtype=dt),
    "stackedMS": lambda dt: hvi_b200.DoubleMMCase(("stackedMS",), dtype=dt),
    "transIdent": lambda dt: hvi_b200.DoubleMMCase(("transIdent",), dtype=dt),
    "neglect_cor": lambda dt: hvi_b200.DoubleMMCase(("neglect_cor",), dtype=dt),
    "rangescaling": lambda dt: hvi_b200.DoubleMMCase(("use_rangescaling",), dtype=dt),
}


@pytest.mark.parametrize("case", list(CASES))
def test_c_restatement_matches_autograd_f64(case):
    prob = CASES[case]("f64")
    nb, M = 20, 3
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, driver_nan_frac=0.2)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    nL, comps, g = O.value_and_grad(oracle_spec(prob), ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs)
    nLc, compsc, gc = c_oracle_value_and_grad(prob, ϕ, data, zP, zMs)
    assert nLc == pytest.approx(nL, rel=1e-12)
    assert np.allclose(compsc, comps, rtol=1e-12, atol=1e-12 * np.abs(comps).max())
    for name, r in prob.positions().items():
        if len(r):
            assert np.max(np.abs(gc[r] - g[r])) <= 1e-11 * np.max(np.abs(g[r])), name


def test_c_restatement_f32_and_threads():
    prob = hvi_b200.DoubleMMCase(dtype="f32")
    nb, M = 300, 8
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    nL, comps, g = O.value_and_grad(oracle_spec(prob), ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs, "f32")
    a = c_oracle_value_and_grad(prob, ϕ, data, zP, zMs, nthreads=1)
    b = c_oracle_value_and_grad(prob, ϕ, data, zP, zMs, nthreads=4)
    assert a[0] == pytest.approx(nL, rel=1e-4) and b[0] == pytest.approx(nL, rel=1e-4)
    for name, r in prob.positions().items():
        if len(r):
            assert np.max(np.abs(a[2][r] - g[r])) <= 2e-4 * np.max(np.abs(g[r])), name
            assert np.max(np.abs(b[2][r] - g[r])) <= 2e-4 * np.max(np.abs(g[r])), name


def test_site_shards_sum_to_the_whole():
    """with the global-block terms counted on one shard only, the sum over site shards
    of [∇ϕ ; components] equals the unsharded evaluation (what the all-reduce relies on)."""
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    nb, M = 24, 4
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    full = c_oracle_value_and_grad(prob, ϕ, data, zP, zMs)
    tot_g, tot_c = np.zeros(prob.n_phi), np.zeros(6)
    for r, (lo, hi) in enumerate([(0, 10), (10, 24)]):
        d = {k: data[k][:, lo:hi] for k in ("xM", "xP", "y_o", "y_unc")}
        z = zMs[:, lo * prob.n_M:hi * prob.n_M]
        nL, c, g = c_oracle_value_and_grad(prob, ϕ, d, zP, z, count_globals=(r == 0))
        tot_g += g
        tot_c += c
    assert np.allclose(tot_c, full[1], rtol=1e-12)
    assert np.allclose(tot_g, full[2], rtol=1e-10, atol=1e-10 * np.abs(full[2]).max())


def test_meanvarsep_c_restatement_matches_autograd():
    """MeanVarSepHVIApproximation (src/elbo_sepvec.jl:3-48): per-site log-variances gathered by i_sites."""
    prob = hvi_b200.DoubleMMCase(("sepvar", "few_sites"), dtype="f64")
    assert prob.n_phi == prob.n_phig + 2 + 2 * 100 + 1 + 1 + 2     # init_hybrid_params.jl:147-152
    nb, M = 20, 3
    i_sites = np.arange(40, 60)          # the third minibatch of a DataLoader with batchsize 20
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    nL, comps, g = O.value_and_grad(oracle_spec(prob), ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs,
                                    i_sites=i_sites)
    nLc, compsc, gc = c_oracle_value_and_grad(prob, ϕ, data, zP, zMs, i_sites=i_sites)
    assert nLc == pytest.approx(nL, rel=1e-12)
    np.testing.assert_allclose(gc, g, rtol=1e-10, atol=1e-11 * np.abs(g).max())
    r = prob.positions()["logσ2_ζMs"]
    touched = np.zeros(len(r), bool)
    touched[(i_sites[:, None] * 2 + np.arange(2)).ravel()] = True
    assert np.all(g[r][~touched] == 0) and np.all(g[r][touched] != 0)


def test_threads_agree():
    """the OpenMP site loop must give the single-thread result (loop counters private, per-thread accumulators)"""
    for scen in ((), ("covarK2",)):
        prob = hvi_b200.DoubleMMCase(scen, dtype="f64")
        n, M = 1500, 8
        data = hvi_b200.gen_hybridproblem_synthetic(prob, n, seed=5)
        zP, zMs = hvi_b200.draw_ξ(prob, n, M)
        ϕ = prob.init_ϕ(perturbed=True)
        nL1, c1, g1 = c_oracle_value_and_grad(prob, ϕ, data, zP, zMs, nthreads=1)
        nL4, c4, g4 = c_oracle_value_and_grad(prob, ϕ, data, zP, zMs, nthreads=4)
        assert abs(nL1 - nL4) <= 1e-12 * abs(nL1)
        assert np.max(np.abs(g1 - g4)) <= 1e-12 * np.max(np.abs(g1))
