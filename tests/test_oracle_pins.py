"""Pins the oracle (oracle/hvi_oracle.py) against every known-answer / property test the
reference holds for the pieces of the neg-ELBO path.  The whole-path value is
unpinned in the reference (test/test_elbo.jl:364,375 only assert types); it is checked here
against central finite differences and an mpmath evaluation instead.

CPU only.  file:line citations are relative to /root/reference.
"""
import math

import numpy as np
import pytest
import torch
from scipy import stats

from helpers import O, hvi_b200, oracle_spec


# ---- test/test_cholesky_structure.jl ---------------------------------------------------------
def test_invsumn():  # :55-64
    ns = [1, 2, 3, 6]
    assert [O.invsumn(sum(range(1, n + 1))) for n in ns] == ns
    with pytest.raises(Exception):
        O.invsumn(5)


def test_get_cor_count_table():  # :66-77
    assert O.get_cor_count_total([]) == 0
    assert O.get_cor_count_total([1]) == 0
    assert O.get_cor_count_total([2]) == 1
    assert O.get_cor_count_total([3]) == 3
    assert O.get_cor_count_total([4]) == 6
    assert O.get_cor_count(4) == 6
    assert O.get_cor_count_total([1, 4]) == 0 + 3
    assert O.get_cor_count_total([2, 4]) == 1 + 1
    assert O.get_cor_count_total([3, 4]) == 3 + 0
    assert O.get_cor_count_total([2, 5]) == 1 + 3


def test_vec2uutri_roundtrip():  # :92-107
    v = torch.arange(1.0, 7.0, dtype=torch.float64)
    U = O.vec2uutri(v)
    assert U.shape == (4, 4)
    assert torch.equal(torch.diag(U), torch.ones(4, dtype=torch.float64))
    assert torch.equal(torch.tril(U, -1), torch.zeros(4, 4, dtype=torch.float64))
    # column-wise packing: v = [U12, U13, U23, U14, U24, U34]
    assert [float(U[0, 1]), float(U[0, 2]), float(U[1, 2]), float(U[0, 3]), float(U[1, 3]), float(U[2, 3])] == \
        [1, 2, 3, 4, 5, 6]
    assert torch.equal(O.uutri2vec(U), v)


def test_utri2vec_pos_known_answers():  # :109-119
    assert O.utri2vec_pos(1, 1) == 1
    assert O.utri2vec_pos(1, 2) == 2
    assert O.utri2vec_pos(2, 2) == 3
    assert O.utri2vec_pos(1, 3) == 4
    assert O.utri2vec_pos(1, 4) == 7
    assert O.utri2vec_pos(5, 5) == 15
    with pytest.raises(AssertionError):
        O.utri2vec_pos(2, 1)
    for s in range(1, 200):  # vec2utri_pos is the inverse (src/cholesky.jl:30-37)
        assert O.utri2vec_pos(*O.vec2utri_pos(s)) == s


def test_transformU_cholesky1_unit_diag():  # :154-168
    v = torch.arange(1.0, 7.0, dtype=torch.float64)
    ch = O.transformU_cholesky1(v)
    assert torch.allclose(torch.diag(ch.T @ ch), torch.ones(4, dtype=torch.float64))
    assert torch.equal(torch.tril(ch, -1), torch.zeros(4, 4, dtype=torch.float64))


def test_transformU_block_cholesky1_blocks():  # :170-243
    v = torch.tensor([1.0, 2.0, 3.0, 5.0], dtype=torch.float64)
    cor_ends = [3, 4]   # get_ca_ends of (b1 = 1:3, b2 = [5])
    rho = torch.arange(1.0, 1.0 + O.get_cor_count_total(cor_ends), dtype=torch.float64)
    U = O.transformU_block_cholesky1(rho, cor_ends)
    assert torch.allclose(torch.diag(U.T @ U), torch.ones(4, dtype=torch.float64))
    assert torch.equal(U[0:3, 3:4], torch.zeros(3, 1, dtype=torch.float64))
    # degenerate: one parameter, no correlation
    U1 = O.transformU_block_cholesky1(torch.zeros(0, dtype=torch.float64), [1])
    assert U1.shape == (1, 1) and float(U1[0, 0]) == 1.0
    # no parameters at all
    U0 = O.transformU_block_cholesky1(torch.zeros(0, dtype=torch.float64), [0])
    assert U0.shape == (0, 0)
    del v


def test_cholcor_coefficient_single():  # test/test_elbo.jl:151-152
    rho_true = -0.6
    a = math.copysign(math.sqrt(rho_true ** 2 / (1 - rho_true ** 2)), rho_true)  # src/cholesky.jl:192-195
    assert a == pytest.approx(-0.75)
    U = O.transformU_cholesky1(torch.tensor([a], dtype=torch.float64))
    assert float((U.T @ U)[0, 1]) == pytest.approx(rho_true)


# ---- test/test_logden_normal.jl ----------------------------------------------------------------
def test_neg_logden_indep_normal_vs_normal_logpdf():  # :9-31
    mu, sigma, y = np.array([1.0, 1.0]), np.array([1.1, 2.0]), np.array([1.2, 1.1])
    ls2 = np.log(sigma ** 2)
    true = stats.norm.logpdf(y, mu, sigma)
    tmp = (ls2 + (y - mu) ** 2 * np.exp(-ls2)) / 2
    assert np.allclose(tmp - tmp[0], -(true - true[0]))
    t = lambda a: torch.as_tensor(a)
    assert float(O.neg_logden_indep_normal(t(y), t(mu), t(ls2))) == pytest.approx(tmp.sum())
    # NaN observations are dropped (src/logden_normal.jl:34-39)
    y2 = np.array([1.2, np.nan])
    assert float(O.neg_logden_indep_normal(t(y2), t(mu), t(ls2))) == pytest.approx(tmp[0])


def test_entropy_MvNormal():  # :33-37
    rng = np.random.default_rng(0)
    S = np.diag([4.0, 5.0]) + rng.random((2, 2))
    S2 = S @ S.T
    ref = stats.multivariate_normal(mean=np.zeros(2), cov=S2).entropy()
    assert O.entropy_MvNormal(2, np.linalg.slogdet(S2)[1]) == pytest.approx(ref)


# ---- test/test_bijectors_utils.jl ---------------------------------------------------------------
def test_exp_bijector_value_logjac_and_gradient():  # :41-52
    x = torch.tensor([[0.1, -0.3], [1.2, 0.7]], dtype=torch.float64, requires_grad=True)
    th, lj = O._transform([O.TR_EXP, O.TR_EXP], x)
    assert torch.allclose(th, torch.exp(x)) and float(lj) == pytest.approx(float(x.sum()))
    (g,) = torch.autograd.grad(th.sum() + lj, x)
    assert torch.allclose(g, torch.exp(x.detach()) + 1.0)


def test_logistic_bijector_logjac():  # :54-69
    x = torch.tensor([[0.3], [-2.0], [4.0]], dtype=torch.float64)
    th, lj = O._transform([O.TR_LOGISTIC], x)
    s = torch.sigmoid(x)
    assert torch.allclose(th, s)
    assert float(lj) == pytest.approx(float(torch.log(s * (1 - s)).sum()))


# ---- test/test_ModelApplicator.jl:28-38 : Φ(Φ⁻¹(r)) ≈ r ----------------------------------------
def test_norminvcdf_roundtrip():
    r = torch.tensor([1e-6, 0.01, 0.3, 0.5, 0.77, 0.999], dtype=torch.float64)
    q = torch.special.ndtri(r)
    assert np.allclose(stats.norm.cdf(q.numpy()), r.numpy(), rtol=1e-12)


# ---- test/test_doubleMM.jl:32-37 : priors -------------------------------------------------------
def test_prior_mode_and_upper_quantile():
    mu, sigma = O.fit_lognormal_mode_q95(2.0, 6.0)
    assert math.exp(mu - sigma ** 2) == pytest.approx(2.0)                 # mode(priors[:K2]) == θall.K2
    assert stats.lognorm(s=sigma, scale=math.exp(mu)).ppf(0.95) == pytest.approx(6.0)
    mu, sigma = O.fit_lognormal_mean_q95(2.0, 6.0)                          # test/test_HybridProblem.jl:84-85
    d = stats.lognorm(s=sigma, scale=math.exp(mu))
    assert d.mean() == pytest.approx(2.0) and d.ppf(0.95) == pytest.approx(6.0)
    mu, sigma = O.fit_normal_mode_q975(0.5, 1.5)                            # :86
    assert stats.norm(mu, sigma).ppf(0.975) == pytest.approx(1.5)
    # the package's own host-side fits are the same numbers
    assert hvi_b200.problem.fit_lognormal_mode_upper(2.0, 6.0)[1:] == pytest.approx(O.fit_lognormal_mode_q95(2.0, 6.0))
    assert hvi_b200.problem.fit_lognormal_mean_upper(2.0, 6.0)[1:] == pytest.approx(O.fit_lognormal_mean_q95(2.0, 6.0))


# ---- test/test_doubleMM.jl:60-106 / test/test_PBMApplicator.jl:41-44 ------------------------------
def test_f_doubleMM_sites_formula_and_nan_driver_gradients():
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    spec = oracle_spec(prob)
    nb = 4
    thP = torch.tensor([0.3, 2.0], dtype=torch.float64, requires_grad=True)
    thM = torch.tensor([[0.5, 0.2], [0.6, 0.25], [0.4, 0.15], [0.55, 0.3]], dtype=torch.float64, requires_grad=True)
    S1 = np.array([0.5, 0.5, 0.5, 0.5, 0.4, 0.3, 0.2, 0.1])
    S2 = np.array([1.0, 3.0, 4.0, 5.0, 5.0, 5.0, 5.0, 5.0])
    xP = np.repeat(np.concatenate([S1, S2]).reshape(-1, 1), nb, axis=1)
    y = O.f_doubleMM_sites(spec, thP, thM, torch.as_tensor(xP))
    i = 2
    r0, K2, r1, K1 = 0.3, 2.0, 0.4, 0.15
    assert np.allclose(y[:, i].detach().numpy(), r0 + r1 * S1 / (K1 + S1) * S2 / (K2 + S2))
    # NaN drivers: prediction NaN there, gradients finite everywhere (strong-zero masking)
    xP[6:8, 1] = np.nan
    y = O.f_doubleMM_sites(spec, thP, thM, torch.as_tensor(xP))
    assert torch.isnan(y[6:8, 1]).all() and torch.isfinite(y[:6, 1]).all()
    yo = torch.zeros_like(y)
    yo[6:8, 1] = float("nan")
    loss = O.neg_logden_indep_normal(yo, y, torch.zeros_like(y))
    gP, gM = torch.autograd.grad(loss, (thP, thM))
    assert torch.isfinite(gP).all() and torch.isfinite(gM).all()


# ---- layout of ϕ (Appendix B) ----------------------------------------------------------------------
def test_phi_layout_known_lengths():
    # docs/src/tutorials/blocks_corr.md:86-89 : nP = 1, nM = 2 → length(ϕq) = 7
    p = hvi_b200.DoubleMMCase(("omit_r0",))
    assert p.n_phi - p.n_phig == 7
    # default DoubleMM: |ϕq| = 10, |ϕg| = 580 (5-20-20-2), |ϕ| = 590
    p = hvi_b200.DoubleMMCase()
    assert (p.n_phig, p.n_phi) == (580, 590)
    pos = p.positions()
    assert list(pos) == ["ϕg", "logσ2_ζP", "coef_logσ2_ζMs", "ρsP", "ρsM", "μP"]
    assert (pos["logσ2_ζP"], pos["coef_logσ2_ζMs"], pos["ρsP"], pos["ρsM"], pos["μP"]) == \
        (range(580, 582), range(582, 586), range(586, 587), range(587, 588), range(588, 590))
    # C1 (test/test_HybridProblem.jl): cor_ends P = 1:2 → no ρsP, |ϕq| = 9
    p = hvi_b200.construct_problem_test_HybridProblem()
    assert p.n_phi - p.n_phig == 9 and len(p.positions()["ρsP"]) == 0
    # with one pbm covariate the network is 6-24-24-2 = 816 parameters
    assert hvi_b200.DoubleMMCase(("covarK2",)).n_phig == 816
    assert oracle_spec(p).offsets()["_len"] == p.n_phi


# ---- test/test_elbo.jl:98-127, 129-242 : shapes and 10 000-draw statistics -----------------------
def test_generate_zeta_shapes_and_statistics():
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    spec = oracle_spec(prob)
    nb, M = 6, 10_000
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    ϕ = prob.init_ϕ()
    pos = prob.positions()
    sd_P, sd_Ma, b = [0.2, 20.0], [0.1, 2.0], [-0.3, 0.2]
    ϕ[pos["logσ2_ζP"]] = np.log(np.square(sd_P))
    ϕ[pos["coef_logσ2_ζMs"]] = [math.log(sd_Ma[0] ** 2), b[0], math.log(sd_Ma[1] ** 2), b[1]]
    cc = lambda r: math.copysign(math.sqrt(r * r / (1 - r * r)), r)
    ϕ[pos["ρsP"]] = cc(0.8)
    ϕ[pos["ρsM"]] = cc(-0.6)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    zsP, zsMs_tr, sigma = O.generate_zeta(spec, t(ϕ), t(data["xM"]), t(zP), t(zMs))
    assert zsP.shape == (2, M) and zsMs_tr.shape == (nb, 2, M) and sigma.shape == (2 + 2 * nb,)
    assert np.allclose(zsP.std(dim=1).numpy(), sd_P, rtol=0.05)
    mu = O.apply_g(spec, t(data["xM"]), t(ϕ[pos["ϕg"]])).numpy()         # (nM × nb)
    sd_M_expected = np.exp((np.log(np.square(sd_Ma)).reshape(-1, 1) + np.array(b).reshape(-1, 1) * mu) / 2)
    assert np.allclose(zsMs_tr.std(dim=2).numpy(), sd_M_expected.T, rtol=0.2)
    assert np.corrcoef(zsP.numpy())[0, 1] == pytest.approx(0.8, abs=0.02)
    for i in range(nb):
        assert np.corrcoef(zsMs_tr[i].numpy())[0, 1] == pytest.approx(-0.6, abs=0.03)
    # different sites are uncorrelated
    assert abs(np.corrcoef(zsMs_tr[0, 0].numpy(), zsMs_tr[1, 0].numpy())[0, 1]) < 0.1
    assert np.allclose(sigma[:2].numpy(), sd_P)


# ---- whole path: finite differences and mpmath ----------------------------------------------------
@pytest.mark.parametrize("case", ["default", "C1", "covarK2"])
def test_autograd_gradient_vs_central_differences(case):
    prob = {"default": hvi_b200.DoubleMMCase(dtype="f64"),
            "C1": hvi_b200.construct_problem_test_HybridProblem(dtype="f64"),
            "covarK2": hvi_b200.DoubleMMCase(("covarK2",), dtype="f64")}[case]
    spec = oracle_spec(prob)
    nb, M = 5, 3
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, driver_nan_frac=0.3)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True).astype(np.float64)
    nL, comps, g = O.value_and_grad(spec, ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs)
    assert comps[0] - comps[1] + comps[2] == pytest.approx(nL)
    assert comps[3] + comps[4] + comps[5] == pytest.approx(comps[0])
    rng = np.random.default_rng(0)
    idx = np.concatenate([rng.choice(prob.n_phig, 12, replace=False), np.arange(prob.n_phig, prob.n_phi)])
    f = lambda p: O.value_and_grad(spec, p, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs)[0]
    for i in idx:
        h = 1e-6 * max(1.0, abs(ϕ[i]))
        e = np.zeros_like(ϕ)
        e[i] = h
        fd = (f(ϕ + e) - f(ϕ - e)) / (2 * h)
        assert fd == pytest.approx(g[i], rel=2e-5, abs=2e-5 * np.max(np.abs(g))), i


def test_scalar_vs_mpmath_40_digits():
    """Bounds the rounding of the float64 oracle itself (needed to claim 1e-10 parity)."""
    import mpmath as mp
    mp.mp.dps = 40
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    spec = oracle_spec(prob)
    nb, M = 3, 2
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True).astype(np.float64)
    nL = O.value_and_grad(spec, ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs)[0]
    pos = prob.positions()
    F = lambda a: mp.mpf(float(a))
    ϕg = [F(v) for v in ϕ[pos["ϕg"]]]
    # network
    def g(x):
        h, p = [F(v) for v in x], 0
        for (ni, no, act, hb) in prob.layers:
            z = []
            for j in range(no):
                s = sum(ϕg[p + i * no + j] * h[i] for i in range(ni))
                if hb:
                    s += ϕg[p + ni * no + j]
                z.append(mp.tanh(s) if act == 1 else 1 / (1 + mp.exp(-s)))
            p += ni * no + (no if hb else 0)
            h = z
        return [F(prob.scale_mu[k]) + F(prob.scale_sigma[k]) * mp.sqrt(2) * mp.erfinv(2 * h[k] - 1) for k in range(2)]
    ls2P = [F(v) for v in ϕ[pos["logσ2_ζP"]]]
    cf = [F(v) for v in ϕ[pos["coef_logσ2_ζMs"]]]
    rP, rM = F(ϕ[pos["ρsP"]][0]), F(ϕ[pos["ρsM"]][0])
    muP = [F(v) for v in ϕ[pos["μP"]]]
    U = lambda r: [[1, r / mp.sqrt(1 + r * r)], [0, 1 / mp.sqrt(1 + r * r)]]
    UP, UM = U(rP), U(rM)
    sigP = [mp.exp(l / 2) for l in ls2P]
    prior = lambda name, th: (lambda a, b: mp.log(th) + mp.log(F(b)) + mp.log(2 * mp.pi) / 2 + (mp.log(th) - F(a)) ** 2 / (2 * F(b) ** 2))(*prob.priors[name][1:])
    tot_y = tot_p = tot_j = mp.mpf(0)
    logsig = sum(l / 2 for l in ls2P)
    mus = [g(data["xM"][:, i]) for i in range(nb)]
    for i in range(nb):
        for k in range(2):
            logsig += (cf[2 * k] + cf[2 * k + 1] * mus[i][k]) / 2
    for m in range(M):
        z = [F(zP[m, 0]), F(zP[m, 1])]
        ζP = [muP[0] + sigP[0] * UP[0][0] * z[0], muP[1] + sigP[1] * (UP[0][1] * z[0] + UP[1][1] * z[1])]
        r0, K2 = mp.exp(ζP[0]), mp.exp(ζP[1])
        tot_j += ζP[0] + ζP[1]
        tot_p += prior("r0", r0) + prior("K2", K2)
        for i in range(nb):
            zz = [F(zMs[m, 2 * i]), F(zMs[m, 2 * i + 1])]
            sg = [mp.exp((cf[2 * k] + cf[2 * k + 1] * mus[i][k]) / 2) for k in range(2)]
            ζ = [mus[i][0] + sg[0] * zz[0], mus[i][1] + sg[1] * (UM[0][1] * zz[0] + UM[1][1] * zz[1])]
            r1, K1 = mp.exp(ζ[0]), mp.exp(ζ[1])
            tot_j += ζ[0] + ζ[1]
            tot_p += prior("r1", r1) + prior("K1", K1)
            for o in range(8):
                S1, S2 = F(data["xP"][o, i]), F(data["xP"][8 + o, i])
                y = r0 + r1 * S1 / (K1 + S1) * S2 / (K2 + S2)
                u = F(data["y_unc"][o, i])
                tot_y += (u + (F(data["y_o"][o, i]) - y) ** 2 * mp.exp(-u)) / 2
    K = 2 + 2 * nb
    ent = (K * mp.log(2 * mp.pi * mp.e) + 2 * logsig) / 2
    ref = (tot_y + tot_p - tot_j) / M - ent
    assert abs(mp.mpf(nL) - ref) / abs(ref) < mp.mpf("1e-13")
