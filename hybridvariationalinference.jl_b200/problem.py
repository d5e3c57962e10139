"""Host-side mirror of the reference's problem description for the neg-ELBO path.

The reference queries a problem through ``get_hybridproblem_*`` accessors
(src/AbstractHybridProblem.jl:30-311) and captures the results as keyword arguments of
``neg_elbo_gtf`` (src/elbo.jl:46-66, src/HybridSolver.jl:193-245).  `HybridProblem` here holds
exactly those pieces -- names, bijector codes, priors, correlation blocks, the MLP shape and
output scaling, the PBM slot map -- and flattens them into the C `hvi_desc`.

`DoubleMMCase` mirrors src/DoubleMM/f_doubleMM.jl:161-417 (scenario flags select the
variants); `construct_problem_test_HybridProblem` mirrors test/test_HybridProblem.jl:32-111.
All file:line citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _abi as A

Z95 = 1.6448536269514722
Z975 = 1.959963984540054


# ---- DistributionFits stand-ins (inputs to the boundary, computed on the host) -----------
def fit_lognormal_mode_upper(mode: float, q95: float) -> Tuple[int, float, float]:
    """fit(LogNormal, mode, QuantilePoint(q95, 0.95), Val(:mode)) -- src/DoubleMM/f_doubleMM.jl:188"""
    L = math.log(q95 / mode)
    s = (-Z95 + math.sqrt(Z95 * Z95 + 4 * L)) / 2
    return (A.HVI_PR_LOGNORMAL, math.log(mode) + s * s, s)


def fit_lognormal_mean_upper(mean: float, q95: float) -> Tuple[int, float, float]:
    """fit(LogNormal, mean, QuantilePoint(q95, 0.95)) (default Val(:mean)) --
    test/test_HybridProblem.jl:84-85"""
    L = math.log(q95 / mean)
    s = Z95 - math.sqrt(Z95 * Z95 - 2 * L)
    return (A.HVI_PR_LOGNORMAL, math.log(mean) - s * s / 2, s)


def fit_normal_mode_uu(mode: float, q975: float) -> Tuple[int, float, float]:
    """fit(Normal, mode, qp_uu(q), Val(:mode)) -- test/test_HybridProblem.jl:86"""
    return (A.HVI_PR_NORMAL, mode, (q975 - mode) / Z975)


def prior_quantile(prior, z):
    kind, a, b = prior
    return a + b * z if kind == A.HVI_PR_NORMAL else math.exp(a + b * z)


def inverse_transform(code: int, theta: float) -> float:
    if code == A.HVI_TR_EXP:
        return math.log(theta)
    if code == A.HVI_TR_LOGISTIC:
        return math.log(theta / (1 - theta))
    return theta


def normal_scaling(priorsM, transM):
    """get_quantile_transformed (src/HybridProblem.jl:291-295) then fit(Normal, @qp_l, @qp_u)
    per output (src/ModelApplicator.jl:150-161)."""
    mu, sg = [], []
    for pr, tr in zip(priorsM, transM):
        lo = inverse_transform(tr, prior_quantile(pr, -Z95))
        hi = inverse_transform(tr, prior_quantile(pr, Z95))
        mu.append((lo + hi) / 2)
        sg.append((hi - lo) / (2 * Z95))
    return mu, sg


def get_cor_counts(cor_ends: Sequence[int]) -> List[int]:
    """src/cholesky.jl:236-244"""
    if len(cor_ends) == 0:
        return [0]
    return [((e if i == 0 else e - cor_ends[i - 1]) - 1) * (e if i == 0 else e - cor_ends[i - 1]) // 2
            for i, e in enumerate(cor_ends)]


@dataclass
class HybridProblem:
    """The pieces of a reference `HybridProblem` (src/HybridProblem.jl:29-72) that the
    neg-ELBO path reads."""
    θP: Dict[str, float]                  # par_templates.θP (names + template values)
    θM: Dict[str, float]
    θFix: Dict[str, float]
    transP: List[int]
    transM: List[int]
    priors: Dict[str, Tuple[int, float, float]]
    cor_ends: Dict[str, List[int]]        # (P=…, M=…)
    layers: List[Tuple[int, int, int, bool]]
    scale_mu: List[float]
    scale_sigma: List[float]
    pbm_covars: Tuple[str, ...] = ()
    n_obs: int = 8
    n_covar: int = 5
    dtype: str = "f32"
    is_omit_priors: bool = False
    scale_kind: int = A.HVI_SCALE_NORMAL
    n_site: int = 800
    n_batch: int = 20
    approx: str = "MeanHVIApproximationMat"   # or "MeanVarSepHVIApproximation" (src/HVIApproximation.jl:7-24)

    # ---- sizes and the ϕ layout (src/init_hybrid_params.jl:25-33,119-124; HybridProblem.jl:101-104)
    @property
    def n_P(self): return len(self.θP)
    @property
    def n_M(self): return len(self.θM)
    @property
    def n_phig(self): return sum(i * o + (o if b else 0) for i, o, _, b in self.layers)
    @property
    def n_rhoP(self): return sum(get_cor_counts(self.cor_ends["P"])) if self.n_P else 0
    @property
    def n_rhoM(self): return sum(get_cor_counts(self.cor_ends["M"]))
    @property
    def np_dtype(self): return np.float32 if self.dtype == "f32" else np.float64
    @property
    def is_sepvar(self): return self.approx == "MeanVarSepHVIApproximation"

    def positions(self) -> Dict[str, range]:
        """0-based ranges of the components of ϕ -- `get_positions` of the interpreters
        `int_ϕg_ϕq` / `int_ϕq` (src/ComponentArrayInterpreter.jl:281-288)."""
        out, p = {}, 0
        # MeanVarSep: logσ2_ζMs (n_θM × n_site) replaces the coefficients (src/init_hybrid_params.jl:147-152)
        third = ("logσ2_ζMs", self.n_M * self.n_site) if self.is_sepvar else ("coef_logσ2_ζMs", 2 * self.n_M)
        for name, ln in (("ϕg", self.n_phig), ("logσ2_ζP", self.n_P), third,
                         ("ρsP", self.n_rhoP), ("ρsM", self.n_rhoM), ("μP", self.n_P)):
            out[name] = range(p, p + ln)
            p += ln
        return out

    @property
    def n_phi(self): return self.positions()["μP"].stop

    def pbm_covar_indices(self) -> List[int]:
        """get_pbm_covar_indices (src/elbo.jl:549-557), 0-based."""
        names = list(self.θP)
        return [names.index(k) for k in self.pbm_covars]

    def slots(self):
        """(r0, r1, K1, K2) addressed by name in hcat(θP, θMs, θFix)
        (src/PBMApplicator.jl:234-239,283-287)."""
        out = []
        for name in ("r0", "r1", "K1", "K2"):
            if name in self.θP:
                out.append((A.HVI_SRC_P, list(self.θP).index(name), 0.0))
            elif name in self.θM:
                out.append((A.HVI_SRC_M, list(self.θM).index(name), 0.0))
            else:
                out.append((A.HVI_SRC_FIX, 0, float(self.θFix[name])))
        return out

    def priorsP(self): return [self.priors[k] for k in self.θP]
    def priorsM(self): return [self.priors[k] for k in self.θM]

    def to_desc(self, device: int = 0, count_globals: bool = True) -> A.hvi_desc:
        d = A.hvi_desc()
        import ctypes
        d.struct_size = ctypes.sizeof(A.hvi_desc)
        d.dtype = A.HVI_F32 if self.dtype == "f32" else A.HVI_F64
        d.device = device
        d.flags = A.HVI_FLAG_OMIT_PRIORS if self.is_omit_priors else 0
        d.n_P, d.n_M, d.n_obs, d.n_covar = self.n_P, self.n_M, self.n_obs, self.n_covar
        d.n_layers = len(self.layers)
        for l, (ni, no, act, b) in enumerate(self.layers):
            d.layers[l].n_in, d.layers[l].n_out, d.layers[l].act, d.layers[l].has_bias = ni, no, act, int(b)
        d.scale_kind = self.scale_kind
        for k in range(self.n_M):
            d.scale_a[k], d.scale_b[k] = self.scale_mu[k], self.scale_sigma[k]
            d.trans_M[k] = self.transM[k]
            pr = self.priorsM()[k]
            d.prior_M[k].kind, d.prior_M[k].a, d.prior_M[k].b = pr
        for j in range(self.n_P):
            d.trans_P[j] = self.transP[j]
            pr = self.priorsP()[j]
            d.prior_P[j].kind, d.prior_P[j].a, d.prior_P[j].b = pr
        d.n_cor_ends_P = len(self.cor_ends["P"])
        for i, e in enumerate(self.cor_ends["P"]):
            d.cor_ends_P[i] = e
        d.n_cor_ends_M = len(self.cor_ends["M"])
        for i, e in enumerate(self.cor_ends["M"]):
            d.cor_ends_M[i] = e
        for s, (src, idx, fix) in enumerate(self.slots()):
            d.slot_src[s], d.slot_idx[s], d.slot_fix[s] = src, idx, fix
        idx = self.pbm_covar_indices()
        d.n_pbm_covar = len(idx)
        for i, v in enumerate(idx):
            d.pbm_covar_idx[i] = v
        d.count_globals = int(count_globals)
        d.approx = A.HVI_APPROX_MEANVAR_SEP if self.is_sepvar else A.HVI_APPROX_MEAN
        d.n_site_total = self.n_site if self.is_sepvar else 0
        return d

    def as_plain_dict(self) -> dict:
        """Plain-data view (used by the tests to build the oracle's own Spec)."""
        return dict(nP=self.n_P, nM=self.n_M, nO=self.n_obs, nC=self.n_covar, layers=list(self.layers),
                    scale_mu=list(self.scale_mu), scale_sigma=list(self.scale_sigma),
                    transP=list(self.transP), transM=list(self.transM),
                    priorsP=self.priorsP(), priorsM=self.priorsM(),
                    cor_ends_P=list(self.cor_ends["P"]), cor_ends_M=list(self.cor_ends["M"]),
                    slots=[(s, (i if s != A.HVI_SRC_FIX else f)) for s, i, f in self.slots()],
                    pbm_covar_idx=self.pbm_covar_indices(), omit_priors=self.is_omit_priors,
                    scale_kind="range" if self.scale_kind == A.HVI_SCALE_RANGE else "normal",
                    approx="sepvar" if self.is_sepvar else "mean", n_site_total=self.n_site if self.is_sepvar else 0)

    # ---- initial parameters -------------------------------------------------------------
    def init_ϕq(self, ρ0: float = 0.0) -> np.ndarray:
        """init_hybrid_ϕunc + update_μP_by_θP (src/init_hybrid_params.jl:105-125,
        src/HybridProblem.jl:101-104): logσ2_ζP=-10, coef=[-10,0] per column, ρ=ρ0, μP=T⁻¹(θP)."""
        q = [-10.0] * self.n_P
        if self.is_sepvar:
            # init_hybrid_ϕunc(::AbstractMeanVarSepHVIApproximation) (src/init_hybrid_params.jl:127-153):
            # σ = relative error 0.01 of the template, per transform (compute_σ_unconstrained :155-168)
            rel = 0.01
            ls = [2 * math.log(math.sqrt(math.log(rel * rel + 1))) if t == A.HVI_TR_EXP else 2 * math.log(rel * v)
                  for t, v in zip(self.transM, self.θM.values())]
            q += ls * self.n_site
        else:
            for _ in range(self.n_M):
                q += [-10.0, 0.0]
        q += [ρ0] * (self.n_rhoP + self.n_rhoM)
        q += [inverse_transform(t, v) for t, v in zip(self.transP, self.θP.values())]
        return np.asarray(q, dtype=self.np_dtype)

    def init_ϕg(self, seed: int = 111) -> np.ndarray:
        """Glorot-uniform weights, zero biases, laid out like SimpleChains' TurboDense params:
        per layer vec(W[out×in]) column-major then b (ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52).
        (The reference draws from StableRNG(111); the stream is not reproducible outside
        Julia and ϕg is an input of the boundary.)"""
        rng = np.random.Generator(np.random.PCG64(seed))
        parts = []
        for ni, no, _, b in self.layers:
            lim = math.sqrt(6.0 / (ni + no))
            parts.append(rng.uniform(-lim, lim, size=ni * no))
            if b:
                parts.append(np.zeros(no))
        return np.concatenate(parts).astype(self.np_dtype)

    def init_ϕ(self, seed: int = 111, perturbed: bool = False) -> np.ndarray:
        """ϕ = (ϕg, ϕq) (src/init_hybrid_params.jl:25-33).  `perturbed` uses the non-trivial
        uncertainty parameters of test/test_elbo.jl:133-149 so that every gradient is exercised."""
        ϕ = np.concatenate([self.init_ϕg(seed), self.init_ϕq()]).astype(self.np_dtype)
        if perturbed:
            pos = self.positions()
            ϕ[pos["ρsM"]] = np.linspace(-0.75, 0.6, len(pos["ρsM"])) if len(pos["ρsM"]) else 0
            ϕ[pos["ρsP"]] = np.linspace(1.33, -0.4, len(pos["ρsP"])) if len(pos["ρsP"]) else 0
            cf = np.array([[math.log(0.1 ** 2), -0.3], [math.log(2.0 ** 2) - 8.0, 0.2],
                           [-6.0, 0.1], [-7.0, -0.15]])
            if self.is_sepvar:
                r = pos["logσ2_ζMs"]
                ϕ[r] = -7.0 + np.random.Generator(np.random.PCG64(seed + 1)).standard_normal(len(r))
            else:
                for k in range(self.n_M):
                    ϕ[pos["coef_logσ2_ζMs"][2 * k]] = cf[k % 4, 0] - 2.0
                    ϕ[pos["coef_logσ2_ζMs"][2 * k + 1]] = cf[k % 4, 1]
            ϕ[pos["logσ2_ζP"]] = np.linspace(-6.0, -5.0, self.n_P) if self.n_P else 0
        return ϕ


# ---- the reference's example case --------------------------------------------------------
θP_DEFAULT = dict(r0=0.3, K2=2.0)      # src/DoubleMM/f_doubleMM.jl:3
θM_DEFAULT = dict(r1=0.5, K1=0.2)      # :4
xP_S1 = np.array([0.5, 0.5, 0.5, 0.5, 0.4, 0.3, 0.2, 0.1])  # :12
xP_S2 = np.array([1.0, 3.0, 4.0, 5.0, 5.0, 5.0, 5.0, 5.0])  # :13


def _three_layer(n_input, n_out, hidden=None):
    """construct_3layer_MLApplicator: n_in → 4n_in tanh → 4n_in tanh → n_out logistic, no
    bias on the last layer (ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52).  `hidden`
    overrides the hidden widths (BASELINE config 5: 4×512)."""
    widths = [n_input * 4, n_input * 4] if hidden is None else list(hidden)
    layers, prev = [], n_input
    for w in widths:
        layers.append((prev, w, A.HVI_ACT_TANH, True))
        prev = w
    layers.append((prev, n_out, A.HVI_ACT_LOGISTIC, False))
    return layers


def DoubleMMCase(scenario: Sequence[str] = (), dtype: str = "f32", hidden=None) -> HybridProblem:
    """The accessors of `DoubleMMCase` for a scenario tuple (src/DoubleMM/f_doubleMM.jl:161-417)."""
    scen = set(scenario)
    θall = {**θP_DEFAULT, **θM_DEFAULT}
    if "no_globals" in scen:                         # :163-165
        θP, θM = {}, dict(θM_DEFAULT)
    elif "omit_r0" in scen and "K1global" in scen:   # :166-171
        θP, θM = dict(K1=0.2), dict(r1=0.5, K2=2.0)
    elif "omit_r0" in scen:
        θP, θM = dict(K2=2.0), dict(θM_DEFAULT)
    else:
        θP, θM = dict(θP_DEFAULT), dict(θM_DEFAULT)
    θFix = {k: v for k, v in θall.items() if k not in θP and k not in θM}          # :273-274
    priors = {k: fit_lognormal_mode_upper(v, 3 * v) for k, v in θall.items()}       # :187-189
    if "stackedMS" in scen:                                                          # :232-234
        transP = [A.HVI_TR_EXP] * len(θP)
        transM = [A.HVI_TR_IDENTITY] + [A.HVI_TR_EXP] * (len(θM) - 1)
    elif "transIdent" in scen:                                                       # :235-238
        transP, transM = [A.HVI_TR_IDENTITY] * len(θP), [A.HVI_TR_IDENTITY] * len(θM)
    else:                                                                            # :239-240
        transP, transM = [A.HVI_TR_EXP] * len(θP), [A.HVI_TR_EXP] * len(θM)
    if "neglect_cor" in scen:                                                        # :398-401
        cor_ends = dict(P=list(range(1, len(θP) + 1)), M=list(range(1, len(θM) + 1)))
    else:                                                                            # :403
        cor_ends = dict(P=[len(θP)], M=[len(θM)])
    pbm_covars = ()
    if "covarK2" in scen:                                                            # :219-227
        pbm_covars = ("K1",) if "K1global" in scen else ("K2",)
    n_covar = 5                                                                      # :313
    priorsM = [priors[k] for k in θM]
    mu, sg = normal_scaling(priorsM, transM)                                         # :203-213
    scale_kind = A.HVI_SCALE_NORMAL
    if "use_rangescaling" in scen:  # RangeScalingModelApplicator(lowers, uppers) :209-210
        lo = [m - Z95 * s for m, s in zip(mu, sg)]
        mu, sg = lo, [2 * Z95 * s for s in sg]
        scale_kind = A.HVI_SCALE_RANGE
    n_site = 100 if "few_sites" in scen else 20 if "sites20" in scen else 800        # :315-325
    approx = "MeanVarSepHVIApproximation" if "sepvar" in scen else "MeanHVIApproximationMat"   # :408
    return HybridProblem(approx=approx, θP=θP, θM=θM, θFix=θFix, transP=transP, transM=transM, priors=priors,
                         cor_ends=cor_ends, layers=_three_layer(n_covar + len(pbm_covars), len(θM), hidden),
                         scale_mu=mu, scale_sigma=sg, pbm_covars=pbm_covars, n_obs=8, n_covar=n_covar,
                         dtype=dtype, scale_kind=scale_kind, n_site=n_site, n_batch=20)


def construct_problem_test_HybridProblem(scenario: Sequence[str] = (), dtype: str = "f32") -> HybridProblem:
    """test/test_HybridProblem.jl:32-111: transM = (identity, Exp), cor_ends = (P=1:2, M=[2]),
    LogNormal priors fitted by mean except r1 ~ Normal."""
    scen = set(scenario)
    θP, θM = dict(θP_DEFAULT), dict(θM_DEFAULT)
    θall = {**θP, **θM}
    priors = {k: fit_lognormal_mean_upper(v, 3 * v) for k, v in θall.items()}
    priors["r1"] = fit_normal_mode_uu(θall["r1"], 3 * θall["r1"])
    transP = [A.HVI_TR_EXP, A.HVI_TR_EXP]
    transM = [A.HVI_TR_IDENTITY, A.HVI_TR_EXP]
    pbm_covars = ("K2",) if "covarK2" in scen else ()
    mu, sg = normal_scaling([priors[k] for k in θM], transM)
    return HybridProblem(θP=θP, θM=θM, θFix={}, transP=transP, transM=transM, priors=priors,
                         cor_ends=dict(P=[1, 2], M=[2]), layers=_three_layer(5 + len(pbm_covars), 2),
                         scale_mu=mu, scale_sigma=sg, pbm_covars=pbm_covars, dtype=dtype,
                         n_site=800, n_batch=20)


# ---- synthetic data (inputs of the boundary) -------------------------------------------
def gen_hybridproblem_synthetic(prob: HybridProblem, n_site: int, seed: int = 111, driver_nan_frac: float = 0.0):
    """gen_hybridproblem_synthetic (src/DoubleMM/f_doubleMM.jl:358-394) with gen_cov_pred /
    compute_correlated_covars / scale_centered_at (src/gencovar.jl:9-64): correlated covariates
    xM (5×n), site parameters θMs_true scattered 10 % around the template, identical drivers
    for every site, y_o = f(θ_true) + N(0, 0.01²), y_unc = log(0.01²).  numpy PCG64 replaces
    StableRNG (streams are not reproducible outside Julia; the data are inputs).
    `driver_nan_frac` generalises the :driverNAN scenario (:332-340): for that fraction of sites
    the last two S1 drivers and observations are NaN."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ft = prob.np_dtype
    n_pc, n_covar, nM, nO = 2, prob.n_covar, prob.n_M, prob.n_obs
    x_pc = rng.random((n_pc, n_site))
    A1 = rng.random((n_covar, n_pc))
    rhos = np.concatenate([[1.0], np.exp(-np.arange(1, n_covar) / 8.0)]).reshape(-1, 1)
    noise = rng.standard_normal((n_covar, n_site)) * 0.2
    xM = rhos * (A1 @ x_pc) + (1 - rhos) * noise
    x_all = np.vstack([x_pc, x_pc[0:1] * x_pc[1:2]])
    A2 = rng.random((nM, x_all.shape[0]))
    th0 = A2 @ x_all
    m = np.array(list(prob.θM.values())).reshape(-1, 1)
    th0 = (th0 - th0.mean(axis=1, keepdims=True)) / th0.std(axis=1, ddof=1, keepdims=True)
    θMs_true = m + th0 * (0.1 * m)                       # (nM × n_site)
    S1 = np.repeat(xP_S1[:nO].reshape(-1, 1), n_site, axis=1)
    S2 = np.repeat(xP_S2[:nO].reshape(-1, 1), n_site, axis=1)
    par = {}
    for name in ("r0", "r1", "K1", "K2"):
        if name in prob.θP:
            par[name] = np.full(n_site, prob.θP[name])
        elif name in prob.θM:
            par[name] = θMs_true[list(prob.θM).index(name)]
        else:
            par[name] = np.full(n_site, prob.θFix[name])
    y_true = par["r0"] + par["r1"] * S1 / (par["K1"] + S1) * S2 / (par["K2"] + S2)
    σ_o = 0.01
    y_o = y_true + rng.standard_normal(y_true.shape) * σ_o
    y_unc = np.full(y_o.shape, 2 * math.log(σ_o))
    xP = np.vstack([S1, S2])
    if driver_nan_frac > 0:
        bad = rng.random(n_site) < driver_nan_frac
        xP[nO - 2:nO, bad] = np.nan
        y_o[nO - 2:, bad] = np.nan
    f = lambda a: np.asfortranarray(a.astype(ft))
    return dict(xM=f(xM), xP=f(xP), y_o=f(y_o), y_unc=f(y_unc), θMs_true=θMs_true, y_true=y_true)


def draw_ξ(prob: HybridProblem, n_site: int, n_MC: int, seed: int = 211):
    """The standard-normal draws of sample_ζresid_norm (src/elbo.jl:605-606): first
    zP (n_MC × n_θP), then zMs (n_MC × n_θM·n_site), column-major."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ft = prob.np_dtype
    zP = np.asfortranarray(rng.standard_normal((n_MC, prob.n_P)).astype(ft))
    zMs = np.asfortranarray(rng.standard_normal((n_MC, prob.n_M * n_site), dtype=np.float32).astype(ft)
                            if ft == np.float32 else rng.standard_normal((n_MC, prob.n_M * n_site)))
    return zP, zMs
