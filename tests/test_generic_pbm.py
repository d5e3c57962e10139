"""The generic process-based-model seam: the reference's `f` is any AbstractPBMApplicator callable
(src/PBMApplicator.jl:26-32); here a model is CUDA C++ source compiled at run time (hvi_plan_set_pbm, hvi_pbm.cuh) whose
Jacobian comes from forward-mode dual numbers.  Checked: (1) f_doubleMM through the seam against the hand-written DoubleMM
kernel, (2) a second model and (3) a user model with exp/sqrt and a floss_penalty (src/elbo.jl:153,282-286) against the
float64 torch-autograd oracle evaluating the same formula."""
import numpy as np
import pytest
import torch

from helpers import O, hvi_b200, oracle_spec

TOL = {"f64": 1e-10, "f32": 1e-4}

USER_SRC = r"""
struct UserPbm {
  static constexpr bool HAS_PENALTY = true;
  // y_o = fix[0] + θP[0] · exp(−θM[0] · x_o) + 0.1 · sqrt(θM[1]) · x_{n_obs + o}
  template <class S, class T>
  __device__ static void f(const S* thP, const S* thM, const T* fix, const T* x, S* y) {
    for (int o = 0; o < HVI_NO; o++) y[o] = fix[0] + thP[0] * hexp(-thM[0] * x[o]) + T(0.1) * hsqrt(thM[1]) * x[HVI_NO + o];
  }
  // ½ (θM[0] − 1)² + 0.01 Σ_o y_o per site
  template <class S, class T>
  __device__ static S penalty(const S* thP, const S* thM, const T* fix, const T* x, const S* y) {
    S p = T(0.5) * (thM[0] - T(1)) * (thM[0] - T(1));
    for (int o = 0; o < HVI_NO; o++) p = p + T(0.01) * y[o];
    return p;
  }
};
"""


def user_f(fix0):
    def f(thP, thM, xP):
        nO = xP.shape[0] // 2
        return fix0 + thP[0] * torch.exp(-thM[:, 0] * xP[:nO]) + 0.1 * torch.sqrt(thM[:, 1]) * xP[nO:]
    return f


def user_penalty(thP, thM, xP, y):
    return (0.5 * (thM[:, 0] - 1.0) ** 2).sum() + 0.01 * y.sum()


def mm1_f(thP, thM, xP):
    return thP[0] + thM[:, 0] * xP / (thM[:, 1] + xP)


def _compare(prob, got, ref, tol):
    nL, comps, g = got
    nL_o, comps_o, g_o = ref
    assert abs(nL - nL_o) <= tol * abs(nL_o), (nL, nL_o)
    assert np.max(np.abs(np.array(comps) - comps_o)) <= tol * np.max(np.abs(comps_o)), (comps, comps_o)
    for blk, r in prob.positions().items():
        if len(r):
            assert np.max(np.abs(np.asarray(g[r], dtype=np.float64) - g_o[r])) <= tol * np.max(np.abs(g_o[r])), blk


def test_sources_compile_without_a_device():
    """NVRTC needs no GPU: the seam's header and the built-in / user models build for sm_100a on the CPU box"""
    p32, p64 = hvi_b200.DoubleMMCase(dtype="f32"), hvi_b200.DoubleMMCase(("covarK2", "omit_r0"), dtype="f64")
    for prob in (p32, p64):
        ok, log = hvi_b200.pbm_compile_check("builtin:doubleMM", prob)
        assert ok, log
        ok, log = hvi_b200.pbm_compile_check(USER_SRC, prob)
        assert ok, log
    ok, log = hvi_b200.pbm_compile_check("builtin:mm1", p64, n_x=8)
    assert ok, log
    ok, log = hvi_b200.pbm_compile_check("struct UserPbm { int nothing; };", p32)
    assert not ok and "no member" in log
    ok, log = hvi_b200.pbm_compile_check("builtin:nope", p32)
    assert not ok and "unknown built-in" in log
    ok, log = hvi_b200.pbm_compile_check("builtin:mm1", hvi_b200.DoubleMMCase(("no_globals",), dtype="f32"), n_x=8)
    assert not ok


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("scen", [(), ("covarK2",), ("omit_r0",), ("no_globals",), ("sepvar", "few_sites"), ("stackedMS",)])
def test_doubleMM_through_the_seam_equals_the_handwritten_kernel(scen, dtype):
    prob = hvi_b200.DoubleMMCase(scen, dtype=dtype)
    nb, M = (100, 5) if "sepvar" in scen else (173, 9)
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=4, driver_nan_frac=0.2)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    fast = hvi_b200.NegElboPlan(prob)
    fast.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    ref = fast.neg_elbo_withgradient(ϕ, zP, zMs)
    gen = hvi_b200.NegElboPlan(prob)
    gen.set_pbm("builtin:doubleMM")
    gen.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    got = gen.neg_elbo_withgradient(ϕ, zP, zMs)
    _compare(prob, got, (ref[0], np.array(ref[1]), ref[2].astype(np.float64)), 2e-5 if dtype == "f32" else 1e-11)
    # device draws: both kernels use the same generator
    a, b = fast.neg_elbo_withgradient(ϕ, n_MC=M, seed=5), gen.neg_elbo_withgradient(ϕ, n_MC=M, seed=5)
    assert abs(a[0] - b[0]) <= (2e-5 if dtype == "f32" else 1e-11) * abs(a[0])
    # back to the hand-written kernel
    gen.set_pbm("doubleMM")
    gen.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    again = gen.neg_elbo_withgradient(ϕ, zP, zMs)
    assert again[0] == ref[0] and np.array_equal(again[2], ref[2])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_second_builtin_model_against_the_oracle(dtype):
    """builtin:mm1 on the (:omit_r0) parameterisation: θP = (K2,), θM = (r1, K1), n_x = n_obs drivers"""
    prob = hvi_b200.DoubleMMCase(("omit_r0",), dtype=dtype)
    nb, M = 150, 7
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=9)
    xP = np.asfortranarray(data["xP"][:8])
    y_o = data["y_o"].copy()
    y_o[3, ::7] = np.nan                                   # missing observations
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_pbm("builtin:mm1", n_x=8)
    plan.set_data(data["xM"], xP, y_o, data["y_unc"])
    got = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    spec = oracle_spec(prob)
    spec.pbm = mm1_f
    ref = O.value_and_grad(spec, ϕ, data["xM"], xP, y_o, data["y_unc"], zP, zMs, dtype)
    _compare(prob, got, ref, TOL[dtype])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("scen", [("omit_r0",), ("omit_r0", "covarK2")])
def test_user_model_with_penalty_against_the_oracle(scen, dtype):
    prob = hvi_b200.DoubleMMCase(scen, dtype=dtype)
    nb, M = 97, 6
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=10)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_pbm(USER_SRC, n_x=16, fix=[0.25])
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    got = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    spec = oracle_spec(prob)
    spec.pbm, spec.penalty = user_f(0.25), user_penalty
    ref = O.value_and_grad(spec, ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs, dtype)
    assert ref[1][2] != 0.0                               # the loss_penalty component is live
    _compare(prob, got, ref, TOL[dtype])
    # the device optimiser loop (graph replay) runs on the generic kernel too
    plan.adam_init(prob.init_ϕ(), eta=0.01)
    tr = plan.adam_run(12, M, seed0=3)
    assert np.all(np.isfinite(tr))


@pytest.mark.gpu
def test_generic_model_errors():
    prob = hvi_b200.DoubleMMCase(dtype="f32")
    plan = hvi_b200.NegElboPlan(prob)
    with pytest.raises(hvi_b200.HVIError) as e:
        plan.set_pbm("struct UserPbm { int nothing; };")
    assert "compilation failed" in str(e.value)
    plan.set_pbm("builtin:doubleMM")
    data = hvi_b200.gen_hybridproblem_synthetic(prob, 40, seed=1)
    with pytest.raises(hvi_b200.HVIError):                  # no batch after set_pbm
        plan.neg_elbo_withgradient(prob.init_ϕ(), n_MC=3, seed=1)
    xP = data["xP"].copy()
    xP[0, 5] = np.nan                                       # driver missing, observation present: NaN as in the reference
    plan.set_data(data["xM"], xP, data["y_o"], data["y_unc"])
    with pytest.raises(hvi_b200.HVIError) as e:
        plan.neg_elbo_withgradient(prob.init_ϕ(), n_MC=3, seed=1)
    assert e.value.code == hvi_b200._abi.HVI_ENONFINITE
    with pytest.raises(hvi_b200.HVIError):
        plan.predict_hvi(prob.init_ϕ(), n_sample_pred=4, seed=1)
