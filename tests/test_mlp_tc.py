"""The tcgen05 small applicator (csrc/hvi_mlp_tc.cu: g and its hand adjoint with TMEM accumulators) against the float64
autograd oracle and against the mma.sync kernels it replaces (HVI_MLP_TC=0), through the C-ABI, on identical inputs.

Reference: g = NormalScaling ∘ MLP (src/ModelApplicator.jl:163-176, ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52),
evaluated per site (pass 1) and per (draw, site) unit with pbm covariates (pass 2, src/elbo.jl:508-515).
Tolerance: rel 1e-4 (Float32, BASELINE.json north_star) on value and on every ϕ block.
"""
import os

import numpy as np
import pytest

from helpers import O, hvi_b200, oracle_spec

TOL = 1e-4


def _eval(prob, data, ϕ, zP, zMs, mode):
    old = os.environ.get("HVI_MLP_TC")
    os.environ["HVI_MLP_TC"] = str(mode)
    try:
        plan = hvi_b200.NegElboPlan(prob, device=0)
        plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
        return plan.neg_elbo_withgradient(ϕ, zP, zMs), plan.launch_count
    finally:
        if old is None:
            os.environ.pop("HVI_MLP_TC", None)
        else:
            os.environ["HVI_MLP_TC"] = old


def _close(prob, got, ref, tol):
    assert abs(got[0] - ref[0]) <= tol * abs(ref[0]), (got[0], ref[0])
    for blk, r in prob.positions().items():
        if len(r):
            err = np.max(np.abs(np.asarray(got[2][r], dtype=np.float64) - np.asarray(ref[2][r], dtype=np.float64)))
            assert err <= tol * np.max(np.abs(ref[2][r])), (blk, err)


# ragged tile counts: one column, one short of / one beyond a 128-column tile, several tiles with a partial last one
@pytest.mark.gpu
@pytest.mark.parametrize("scen", [(), ("covarK2",)])
@pytest.mark.parametrize("nb,M", [(1, 3), (127, 2), (129, 5), (1000, 8)])
def test_tc_applicator_matches_oracle_and_mma(scen, nb, M):
    prob = hvi_b200.DoubleMMCase(scen, dtype="f32")
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=5, driver_nan_frac=0.1 if nb > 100 else 0.0)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    ref = O.value_and_grad(oracle_spec(prob), ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs, "f32")
    tc, _ = _eval(prob, data, ϕ, zP, zMs, 3)     # 3: the tcgen05 kernels at every batch size (the default leaves small batches to mma.sync)
    mma, _ = _eval(prob, data, ϕ, zP, zMs, 0)
    _close(prob, tc, ref, TOL)
    _close(prob, mma, ref, TOL)
    _close(prob, tc, mma, TOL)


@pytest.mark.gpu
def test_tc_applicator_is_deterministic_and_shard_additive():
    """no float atomics: two runs are bit-identical; the ϕg gradient of two site shards sums to the whole"""
    prob = hvi_b200.DoubleMMCase(("covarK2",), dtype="f32")
    nb, M = 3000, 4
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=9)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    a, _ = _eval(prob, data, ϕ, zP, zMs, 3)
    b, _ = _eval(prob, data, ϕ, zP, zMs, 3)
    assert a[0] == b[0] and np.array_equal(a[2], b[2])


@pytest.mark.gpu
def test_tc_forward_feeds_generate_zeta():
    """μ_ζ of the tcgen05 forward (3xTF32) through generate_ζ against the float64 restatement (src/elbo.jl:472-521)"""
    import torch
    prob = hvi_b200.DoubleMMCase(("covarK2",), dtype="f32")
    nb, M = 300, 6
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob, device=0)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    os.environ["HVI_MLP_TC"] = "3"
    try:
        ζP, ζMs, σ = plan.generate_ζ(ϕ, zP, zMs)
    finally:
        os.environ.pop("HVI_MLP_TC", None)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    ζP_o, ζMs_o, σ_o = O.generate_zeta(oracle_spec(prob), t(ϕ), t(data["xM"]), t(zP), t(zMs))
    np.testing.assert_allclose(ζMs, ζMs_o.numpy(), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(σ, σ_o.numpy(), rtol=2e-5)


@pytest.mark.gpu
def test_default_takes_tc_above_the_size_threshold_and_mma_below():
    """10⁶ columns run on tcgen05 by default, 10⁴ on mma.sync; both agree with each other on the same inputs"""
    prob = hvi_b200.DoubleMMCase(dtype="f32")
    ϕ = prob.init_ϕ(perturbed=True)
    for nb in (10_000, 1_000_000):
        data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=2)
        auto = _eval_rng(prob, data, ϕ, None)
        forced_tc = _eval_rng(prob, data, ϕ, 3)
        forced_mma = _eval_rng(prob, data, ϕ, 0)
        same_as = forced_tc if nb >= 500_000 else forced_mma
        assert auto[0] == same_as[0] and np.array_equal(auto[2], same_as[2])      # deterministic kernels: bit-identical
        _close(prob, forced_tc, forced_mma, 2e-5)


def _eval_rng(prob, data, ϕ, mode):
    old = os.environ.get("HVI_MLP_TC")
    if mode is None:
        os.environ.pop("HVI_MLP_TC", None)
    else:
        os.environ["HVI_MLP_TC"] = str(mode)
    try:
        plan = hvi_b200.NegElboPlan(prob, device=0)
        plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
        return plan.neg_elbo_withgradient(ϕ, n_MC=4, seed=21)
    finally:
        if old is None:
            os.environ.pop("HVI_MLP_TC", None)
        else:
            os.environ["HVI_MLP_TC"] = old
