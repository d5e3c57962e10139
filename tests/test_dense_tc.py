"""The tcgen05 dense layer (3xTF32, TMEM accumulators) against a float64 matmul: the building
block of the wide applicator (BASELINE config 5).  Tolerance: 3e-5 relative to max|C| for K up to 512 (the
tensor core's fp32 accumulation over 3·K/8 instructions dominates; plain TF32 gives ~5e-4)."""
import ctypes as C

import numpy as np
import pytest

from helpers import hvi_b200

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Mc,N,K,act", [(128, 128, 32, 0), (128, 128, 512, 0), (300, 256, 64, 1), (1000, 512, 512, 1),
                                        (77, 128, 96, 2)])
def test_dense_tc_matches_float64(Mc, N, K, act):
    lib = hvi_b200._abi.load_library()
    rng = np.random.default_rng(Mc + N + K)
    A = rng.standard_normal((Mc, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    out = np.zeros((Mc, N), dtype=np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.hvi_dense_tc(0, Mc, N, K, p(A), p(W), p(b), act, p(out))
    assert rc == 0, lib.hvi_last_error(None)
    z = A.astype(np.float64) @ W.astype(np.float64).T + b
    ref = np.tanh(z) if act == 1 else 1 / (1 + np.exp(-z)) if act == 2 else z
    err = np.max(np.abs(out - ref)) / np.max(np.abs(ref))
    assert err < 3e-5, err


@pytest.mark.parametrize("dtype,hidden,scen", [("f32", (512, 512, 512, 512), ()), ("f32", (128, 256), ()),
                                               ("f32", (512, 512), ("covarK2",)), ("f64", (96, 160), ())])
def test_wide_applicator_forward(dtype, hidden, scen):
    """g for networks wider than HVI_MAX_WIDTH (BASELINE config 5 shape 5-512-512-512-512-2): hidden
    layers on the tensor cores (Float32) -- forward paths and the ELBO gradient against the oracle."""
    import torch
    from helpers import O, oracle_spec
    prob = hvi_b200.DoubleMMCase(scen, dtype=dtype, hidden=hidden)
    nb, ns = 333, 5
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, ns)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    spec = oracle_spec(prob)
    tol = 1e-10 if dtype == "f64" else 2e-4
    if not scen:
        mu = plan.apply_g(ϕ)
        mu_o = O.apply_g(spec, t(data["xM"]), t(ϕ[prob.positions()["ϕg"]])).numpy()
        assert np.max(np.abs(mu - mu_o)) <= tol * np.max(np.abs(mu_o))
    y, θP, θMs, ent = plan.predict_hvi(ϕ, zP, zMs)
    y_o, θP_o, θMs_o, ent_o = O.predict_hvi(spec, ϕ, data["xM"], data["xP"], zP, zMs)
    np.testing.assert_allclose(θMs, θMs_o, rtol=5 * tol, atol=tol)
    np.testing.assert_allclose(y, y_o, rtol=5 * tol, atol=tol)
    # the ELBO and its gradient through the wide applicator: tensor-core forward, data-gradient and
    # weight-gradient GEMMs (Float32), CUDA-core fallbacks (Float64)
    nL, comps, g = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    nL_o, comps_o, g_o = O.value_and_grad(spec, ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs, dtype)
    gtol = 1e-9 if dtype == "f64" else 1e-4
    assert abs(nL - nL_o) <= gtol * abs(nL_o), (nL, nL_o)
    for name, r in prob.positions().items():
        if len(r):
            err = np.max(np.abs(g[r].astype(np.float64) - g_o[r]))
            assert err <= gtol * np.max(np.abs(g_o[r])), (name, err, np.max(np.abs(g_o[r])))
