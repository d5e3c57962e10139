"""GPU parity: the CUDA path (through the C-ABI) against the float64 oracle on identical inputs
and identical standard-normal draws ξ.

Tolerances (BASELINE.json north_star): value and gradient within rel 1e-10 (Float64) and 1e-4
(Float32).  "rel" for the gradient is max|g − g_ref| / max|g_ref| per ϕ component block
(ϕg, logσ2_ζP, coef_logσ2_ζMs, ρsP, ρsM, μP); for the scalar it is |v − v_ref| / |v_ref| and
for the six components |c − c_ref| / max|c_ref|.
"""
import numpy as np
import pytest

from helpers import O, hvi_b200, oracle_spec

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-10, "f32": 1e-4}

CASES = {
    "C1_test_HybridProblem": lambda dt: hvi_b200.construct_problem_test_HybridProblem(dtype=dt),
    "DoubleMM_default": lambda dt: hvi_b200.DoubleMMCase(dtype=dt),
    "DoubleMM_omit_r0": lambda dt: hvi_b200.DoubleMMCase(("omit_r0",), dtype=dt),
    "DoubleMM_K1global": lambda dt: hvi_b200.DoubleMMCase(("omit_r0", "K1global"), dtype=dt),
    "DoubleMM_no_globals": lambda dt: hvi_b200.DoubleMMCase(("no_globals",), dtype=dt),
    "DoubleMM_stackedMS": lambda dt: hvi_b200.DoubleMMCase(("stackedMS",), dtype=dt),
    "DoubleMM_transIdent": lambda dt: hvi_b200.DoubleMMCase(("transIdent",), dtype=dt),
    "DoubleMM_neglect_cor": lambda dt: hvi_b200.DoubleMMCase(("neglect_cor",), dtype=dt),
    "DoubleMM_rangescaling": lambda dt: hvi_b200.DoubleMMCase(("use_rangescaling",), dtype=dt),
    # pbm covariates: g([xM; ζP_m[idx]]) per draw (src/elbo.jl:491-518; docs/src/tutorials/corr_site_global.md)
    "DoubleMM_covarK2": lambda dt: hvi_b200.DoubleMMCase(("covarK2",), dtype=dt),
    "C1_covarK2": lambda dt: hvi_b200.construct_problem_test_HybridProblem(("covarK2",), dtype=dt),
    "DoubleMM_K1global_covarK1": lambda dt: hvi_b200.DoubleMMCase(("omit_r0", "K1global", "covarK2"), dtype=dt),
}


def check(prob, nb, M, nan_frac=0.0, perturbed=True, seed=111):
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=seed, driver_nan_frac=nan_frac)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=perturbed)
    plan = hvi_b200.NegElboPlan(prob)
    assert {k: (r.start, r.stop) for k, r in plan.positions().items()} == \
           {k: (r.start, r.stop) for k, r in prob.positions().items()}
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    nL, comps, g = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    nL_o, comps_o, g_o = O.value_and_grad(oracle_spec(prob), ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"],
                                          zP, zMs, prob.dtype)
    tol = TOL[prob.dtype]
    assert abs(nL - nL_o) <= tol * abs(nL_o), (nL, nL_o)
    assert np.max(np.abs(np.array(comps) - comps_o)) <= tol * np.max(np.abs(comps_o)), (comps, comps_o)
    for name, r in prob.positions().items():
        if len(r) == 0:
            continue
        err = np.max(np.abs(g[r].astype(np.float64) - g_o[r]))
        scale = np.max(np.abs(g_o[r]))
        assert err <= tol * scale, (name, err, scale, g[r][:4], g_o[r][:4])
    # value-only call gives the same scalar
    nL2, _, g2 = plan.neg_elbo_withgradient(ϕ, zP, zMs, want_grad=False)
    assert g2 is None and nL2 == nL
    plan.close()
    return nL, g


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("case", list(CASES))
def test_config1_sizes(case, dtype):
    """BASELINE config 1 sizes: 20 sites, n_MC = 3 (test/test_HybridProblem.jl:252)."""
    check(CASES[case](dtype), nb=20, M=3)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("nb,M", [(1, 4), (31, 8), (33, 32), (257, 64), (1000, 12), (64, 36)])
def test_ragged_sizes_staged_and_direct(nb, M, dtype):
    """site counts around the 32-site tile, draw counts on both the cp.async path (M % 4 == 0)
    and the direct path, partial last stage (M = 36, 12)."""
    check(hvi_b200.DoubleMMCase(dtype=dtype), nb=nb, M=M)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_missing_drivers(dtype):
    """:driverNAN generalised (src/DoubleMM/f_doubleMM.jl:332-340): NaN drivers with NaN
    observations are skipped in value and gradient (test/test_missingdriver.jl:84-137)."""
    nL, g = check(hvi_b200.DoubleMMCase(dtype=dtype), nb=200, M=8, nan_frac=0.3)
    assert np.isfinite(nL) and np.all(np.isfinite(g))


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_initial_parameters(dtype):
    """ϕ as initialised by the reference (logσ2 = −10, ρ = 0)."""
    check(hvi_b200.DoubleMMCase(dtype=dtype), nb=100, M=16, perturbed=False)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("nb,M", [(33, 8), (300, 12), (129, 5)])
def test_pbm_covars_ragged(nb, M, dtype):
    check(hvi_b200.DoubleMMCase(("covarK2",), dtype=dtype), nb=nb, M=M, nan_frac=0.2)


def test_generate_zeta_with_pbm_covars():
    import torch
    prob = hvi_b200.DoubleMMCase(("covarK2",), dtype="f64")
    nb, M = 40, 6
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    ζP, ζMs, σ = plan.generate_ζ(ϕ, zP, zMs)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    ζP_o, ζMs_o, σ_o = O.generate_zeta(oracle_spec(prob), t(ϕ), t(data["xM"]), t(zP), t(zMs))
    np.testing.assert_allclose(ζP, ζP_o.numpy(), rtol=1e-12)
    np.testing.assert_allclose(ζMs, ζMs_o.numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(σ, σ_o.numpy(), rtol=1e-12)


def test_run_to_run_bit_reproducible():
    prob = hvi_b200.DoubleMMCase(dtype="f32")
    nb, M = 5000, 32
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    a = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    b = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    assert a[0] == b[0] and np.array_equal(a[2], b[2])


def test_anomaly_flag_like_reference_nan():
    """driver NaN but observation finite: the reference's loss is NaN;
    the library reports HVI_ENONFINITE instead of silently masking."""
    prob = hvi_b200.DoubleMMCase(dtype="f32")
    nb, M = 40, 4
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    data["xP"][3, 5] = np.nan
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    with pytest.raises(hvi_b200.HVIError) as ei:
        plan.neg_elbo_withgradient(prob.init_ϕ(), zP, zMs)
    assert ei.value.code == hvi_b200._abi.HVI_ENONFINITE


def test_generate_zeta_matches_oracle():
    import torch
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    nb, M = 50, 7
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    ζP, ζMs, σ = plan.generate_ζ(ϕ, zP, zMs)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    ζP_o, ζMs_o, σ_o = O.generate_zeta(oracle_spec(prob), t(ϕ), t(data["xM"]), t(zP), t(zMs))
    assert ζMs.shape == (nb, prob.n_M, M) and ζP.shape == (prob.n_P, M)   # test/test_elbo.jl:98-127
    np.testing.assert_allclose(ζP, ζP_o.numpy(), rtol=1e-12)
    np.testing.assert_allclose(ζMs, ζMs_o.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(σ, σ_o.numpy(), rtol=1e-12)


def test_device_rng_mode():
    """device-side draws (CUDA.randn stand-in): the generator's moments, seed reproducibility,
    sharding invariance of the columns, and the ELBO evaluated with in-kernel draws equals the ELBO
    evaluated on the same draws passed in explicitly."""
    prob = hvi_b200.DoubleMMCase(dtype="f32")
    nb, M = 2000, 32
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    z = plan.randn(M, prob.n_M * nb, seed=7).astype(np.float64)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.01
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3) < 0.06
    assert abs(np.corrcoef(z[0], z[1])[0, 1]) < 0.05 and abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.5
    np.testing.assert_array_equal(plan.randn(M, 100, seed=7, col_offset=300), z[:, 300:400].astype(np.float32))
    a = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=7)
    b = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=7)
    c = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=8)
    assert a[0] == b[0] and np.array_equal(a[2], b[2]) and a[0] != c[0]
    # same draws passed in: zP are the columns 2^62 … of the generator
    zP = plan.randn(M, prob.n_P, seed=7, col_offset=1 << 62)
    d = plan.neg_elbo_withgradient(ϕ, zP, z.astype(np.float32))
    assert abs(a[0] - d[0]) <= 1e-5 * abs(d[0])
    np.testing.assert_allclose(a[2], d[2], rtol=1e-3, atol=1e-4 * np.abs(d[2]).max())


@pytest.mark.parametrize("case", ["DoubleMM_default", "DoubleMM_covarK2"])
def test_site_shards_on_one_gpu_sum_to_the_whole(case):
    """SURVEY 8(e) on the CUDA path: two plans own contiguous site ranges, only the first counts
    the global-block terms; the sum of their [∇ϕ ; components] equals the unsharded evaluation
    (what the NCCL all-reduce computes across ranks)."""
    from hvi_b200.shard import shard_batch
    prob = CASES[case]("f64")
    nb, M = 90, 8
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    full = hvi_b200.NegElboPlan(prob)
    full.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    nL, comps, g = full.neg_elbo_withgradient(ϕ, zP, zMs)
    tot_g, tot_c = np.zeros_like(g, dtype=np.float64), np.zeros(6)
    for r in range(2):
        d, z, lo = shard_batch(data, zMs, prob.n_M, r, 2)
        plan = hvi_b200.NegElboPlan(prob, count_globals=(r == 0))
        plan.set_data(d["xM"], d["xP"], d["y_o"], d["y_unc"])
        _, c, gg = plan.neg_elbo_withgradient(ϕ, zP, z)
        tot_g += gg
        tot_c += np.array(c)
    np.testing.assert_allclose(tot_c, np.array(comps), rtol=1e-11)
    np.testing.assert_allclose(tot_g, g, rtol=1e-9, atol=1e-10 * np.abs(g).max())


def test_device_rng_is_sharding_invariant():
    """the device generator is keyed by the global site index: a shard with site_offset draws the
    same ξ as the corresponding sites of the unsharded evaluation."""
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    nb, M = 64, 8
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    ϕ = prob.init_ϕ(perturbed=True)
    full = hvi_b200.NegElboPlan(prob)
    full.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    nL, comps, g = full.neg_elbo_withgradient(ϕ, n_MC=M, seed=5)
    tot = np.zeros(6)
    for r, (lo, hi) in enumerate([(0, 24), (24, 64)]):
        plan = hvi_b200.NegElboPlan(prob, count_globals=(r == 0))
        plan.set_data(*(data[k][:, lo:hi] for k in ("xM", "xP", "y_o", "y_unc")))
        tot += np.array(plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=5, site_offset=lo)[1])
    np.testing.assert_allclose(tot, np.array(comps), rtol=1e-11)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("case", ["DoubleMM_default", "C1_test_HybridProblem", "DoubleMM_covarK2", "DoubleMM_no_globals"])
def test_predict_hvi_matches_oracle(case, dtype):
    """predict_hvi / sample_posterior forward path (src/elbo.jl:320-458): shapes as documented
    (:305-313), values against the oracle, NaN predictions exactly where a driver is missing,
    prediction-only batch (no observations)."""
    prob = CASES[case](dtype)
    nb, ns = 45, 9
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, driver_nan_frac=0.2)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, ns)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(data["xM"], data["xP"])          # no y_o / y_unc
    y, θP, θMs, ent = plan.predict_hvi(ϕ, zP, zMs)
    assert y.shape == (prob.n_obs, nb, ns) and θP.shape == (prob.n_P, ns) and θMs.shape == (nb, prob.n_M, ns)
    f = (lambda a: np.asarray(a, dtype=np.float32)) if dtype == "f32" else (lambda a: a)
    y_o, θP_o, θMs_o, ent_o = O.predict_hvi(oracle_spec(prob), f(ϕ), f(data["xM"]), f(data["xP"]), f(zP), f(zMs))
    tol = {"f64": 1e-10, "f32": 2e-5}[dtype]
    assert np.array_equal(np.isnan(y), np.isnan(y_o)) and np.isnan(y).any()
    np.testing.assert_allclose(θP, θP_o, rtol=tol)
    np.testing.assert_allclose(θMs, θMs_o, rtol=tol, atol=tol)
    np.testing.assert_allclose(y, y_o, rtol=tol, atol=tol, equal_nan=True)
    assert abs(ent - ent_o) <= 1e-6 * abs(ent_o) if dtype == "f32" else abs(ent - ent_o) <= 1e-11 * abs(ent_o)
    # the device generator gives finite predictions of the same shape
    y2, _, _, ent2 = plan.predict_hvi(ϕ, n_sample_pred=12, seed=3)
    assert y2.shape == (prob.n_obs, nb, 12) and np.isfinite(y2[~np.isnan(y2)]).all() and abs(ent2 - ent) <= 1e-9 * abs(ent)
    # without observations the ELBO of the batch has no likelihood term but still evaluates
    nL, comps, g = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    assert comps.nLy == 0.0 and np.isfinite(nL)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("scen", [("sepvar", "few_sites"), ("sepvar", "few_sites", "covarK2")])
def test_meanvarsep_approximation(scen, dtype):
    """MeanVarSepHVIApproximation (src/elbo_sepvec.jl:3-48): per-site log-variances read at the
    batch's i_sites; their gradient is owner-local, zero for sites outside the batch."""
    prob = hvi_b200.DoubleMMCase(scen, dtype=dtype)
    nb, M = 20, 4
    i_sites = np.arange(60, 80)[::-1].copy()       # any order
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob)
    assert {k: (r.start, r.stop) for k, r in plan.positions().items()} == \
           {k: (r.start, r.stop) for k, r in prob.positions().items()}
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"], i_sites=i_sites)
    nL, comps, g = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    nL_o, comps_o, g_o = O.value_and_grad(oracle_spec(prob), ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"],
                                          zP, zMs, dtype, i_sites=i_sites)
    tol = TOL[dtype]
    assert abs(nL - nL_o) <= tol * abs(nL_o)
    assert np.max(np.abs(np.array(comps) - comps_o)) <= tol * np.max(np.abs(comps_o))
    for name, r in prob.positions().items():
        if len(r):
            assert np.max(np.abs(g[r] - g_o[r])) <= tol * np.max(np.abs(g_o[r])), name
    r = prob.positions()["logσ2_ζMs"]
    outside = np.ones(len(r), bool)
    outside[(i_sites[:, None] * 2 + np.arange(2)).ravel()] = False
    assert np.all(g[r][outside] == 0)
    y, θP, θMs, ent = plan.predict_hvi(ϕ, zP, zMs)
    y_o, θP_o, θMs_o, ent_o = O.predict_hvi(oracle_spec(prob), ϕ, data["xM"], data["xP"], zP, zMs, i_sites=i_sites)
    np.testing.assert_allclose(θMs, θMs_o, rtol=max(tol, 2e-5), atol=tol)
    assert abs(ent - ent_o) <= 1e-6 * abs(ent_o)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_many_draws_global_pd_path(dtype):
    """n_MC = 700: the per-draw records of the global block no longer fit the 16 KB shared-memory
    budget and are read from global memory (n_sample_pred-sized evaluations)."""
    check(hvi_b200.DoubleMMCase(dtype=dtype), nb=40, M=700)


@pytest.mark.parametrize("M", [5, 6, 64])
def test_device_rng_equals_explicit_draws_any_n_mc(M):
    """in-kernel draws, including n_MC not a multiple of 4 (tail draws) and Float64 (chunks of 2)."""
    for dtype in ("f32", "f64"):
        prob = hvi_b200.DoubleMMCase(dtype=dtype)
        nb = 70
        data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
        ϕ = prob.init_ϕ(perturbed=True)
        plan = hvi_b200.NegElboPlan(prob)
        plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
        a = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=11)
        z = plan.randn(M, prob.n_M * nb, seed=11)
        zP = plan.randn(M, prob.n_P, seed=11, col_offset=1 << 62)
        d = plan.neg_elbo_withgradient(ϕ, zP, z)
        tol = 1e-5 if dtype == "f32" else 1e-12
        assert abs(a[0] - d[0]) <= tol * abs(d[0]), (dtype, M)
        np.testing.assert_allclose(a[2], d[2], rtol=100 * tol, atol=10 * tol * np.abs(d[2]).max())
