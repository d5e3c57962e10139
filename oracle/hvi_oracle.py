"""CPU ORACLE (test infrastructure, NOT the product) -- float64 restatement of the
reference's negative-ELBO path, function by function, in torch so that reverse-mode
autograd gives a gradient that is independent of the hand-written CUDA adjoint.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package never does.

PARITY STATUS: the reference is Julia and cannot run in this image (no julia binary), and
its tests never pin the whole-path ELBO value/gradient (test/test_elbo.jl:364,375 are
type assertions only).  The *components* are pinned against the reference's own
known-answer tests (tests/test_oracle_pins.py: cholesky index maps and diag(U'U)==1,
logden vs Normal logpdf, entropy vs MvNormal entropy, Exp/Logistic log-Jacobians,
prior mode/quantile fits, NaN-driver finiteness, 10 000-draw covariance statistics).
The whole-path scalar and gradient are "parity unpinned": they are checked against central
finite differences, an mpmath 40-digit evaluation and the independent C restatement
(oracle/hvi_oracle.c), not against a run of the Julia package.  julia/dump_golden.jl is kept
ready to produce real golden vectors on a host that has Julia.

All ``file:line`` citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

F64 = torch.float64

# transform codes (src/bijectors_utils.jl:10-23, 38-52; identity via Stacked)
TR_IDENTITY, TR_EXP, TR_LOGISTIC = 0, 1, 2
# prior codes (Distributions.logpdf call sites src/elbo.jl:217-235)
PR_NONE, PR_NORMAL, PR_LOGNORMAL = 0, 1, 2
# PBM slot sources (src/PBMApplicator.jl:234-239, hcat(θP, θMs, θFix) at :283-287)
SRC_P, SRC_M, SRC_FIX = 0, 1, 2
ACT_IDENTITY, ACT_TANH, ACT_LOGISTIC = 0, 1, 2


# --------------------------------------------------------------------------------------
# src/cholesky.jl
# --------------------------------------------------------------------------------------
def sumn(n: int) -> int:
    """src/cholesky.jl:1"""
    return n * (n + 1) // 2


def invsumn(s: int) -> int:
    """src/cholesky.jl:8 -- errors (InexactError) if s is not a triangular number."""
    n = -0.5 + math.sqrt(0.25 + 2 * s)
    if abs(n - round(n)) > 1e-9:
        raise ValueError("InexactError: %r is not n(n+1)/2" % (s,))
    return int(round(n))


def vec2utri_pos(s: int) -> Tuple[int, int]:
    """src/cholesky.jl:30-37 -- one-based (row, col) of one-based packed position s."""
    col = math.ceil(-0.5 + math.sqrt(0.25 + 2 * s))
    cl = col - 1
    s1 = cl * (cl + 1) // 2
    return (s - s1, col)


def utri2vec_pos(row: int, col: int) -> int:
    """src/cholesky.jl:42-50"""
    assert row <= col
    return (col - 1) * col // 2 + row


def get_cor_count(n_par: int) -> int:
    """src/cholesky.jl:245-247"""
    return sumn(n_par - 1)


def get_cor_counts(cor_ends: Sequence[int]) -> List[int]:
    """src/cholesky.jl:236-244"""
    if len(cor_ends) == 0:
        return [0]
    out = []
    for i, e in enumerate(cor_ends):
        nblk = e if i == 0 else e - cor_ends[i - 1]
        out.append(get_cor_count(nblk))
    return out


def get_cor_count_total(cor_ends: Sequence[int]) -> int:
    return sum(get_cor_counts(cor_ends))


def vec2uutri(v: torch.Tensor, n: Optional[int] = None) -> torch.Tensor:
    """src/cholesky.jl:52-73 -- unit upper triangular, packed column-wise above the diagonal
    (the comprehension iterates i fastest, so k advances down each column)."""
    if n is None:
        n = invsumn(v.numel()) + 1
    m = torch.zeros((n, n), dtype=v.dtype)
    rows, cols = [], []
    for j in range(n):
        for i in range(j):
            rows.append(i)
            cols.append(j)
    m = m + torch.eye(n, dtype=v.dtype)
    if rows:
        m = m.index_put((torch.tensor(rows), torch.tensor(cols)), v)
    return m


def uutri2vec(X: torch.Tensor) -> torch.Tensor:
    """src/cholesky.jl:109-126"""
    n = X.shape[0]
    out = [X[i, j] for j in range(n) for i in range(j)]
    return torch.stack(out) if out else X.new_zeros((0,))


def transformU_cholesky1(v: torch.Tensor, n: Optional[int] = None) -> torch.Tensor:
    """src/cholesky.jl:156-174 -- columns of the unit-upper matrix scaled to unit norm."""
    Us = vec2uutri(v, n)
    return Us / torch.sqrt((Us * Us).sum(dim=0, keepdim=True))


def transformU_block_cholesky1(v: torch.Tensor, cor_ends: Sequence[int]) -> torch.Tensor:
    """src/cholesky.jl:262-313 -- block diagonal of per-block factors.

    The reference's ``ranges`` (:274-280) treat per-block counts as cumulative ends, which is
    only right for one block, all-singleton blocks or one block followed by singletons
    (SURVEY Appendix C-6).  This restatement uses the intended cumulative offsets, which
    coincides with the reference on every configuration where the reference is correct.
    """
    cor_ends = list(cor_ends)
    if len(cor_ends) == 0:
        return torch.ones((1, 1), dtype=v.dtype)
    if cor_ends == [0]:
        return torch.zeros((0, 0), dtype=v.dtype)
    n = cor_ends[-1]
    U = torch.zeros((n, n), dtype=v.dtype)
    counts = get_cor_counts(cor_ends)
    off, start = 0, 0
    for cnt, end in zip(counts, cor_ends):
        nb = end - start
        blk = transformU_cholesky1(v[off:off + cnt], nb)
        pad = torch.zeros((n, n), dtype=v.dtype)
        pad[start:end, start:end] = blk
        U = U + pad
        off += cnt
        start = end
    return U


# --------------------------------------------------------------------------------------
# src/logden_normal.jl
# --------------------------------------------------------------------------------------
def neg_logden_indep_normal(obs, mu, logsigma2, sigfac=1.0):
    """src/logden_normal.jl:16-43 -- NaN observations are dropped (boolean compaction)."""
    fin = ~torch.isnan(obs)
    o, m, l2 = obs[fin], mu[fin], logsigma2[fin]
    return (sigfac * l2 + (o - m) ** 2 * torch.exp(-l2)).sum() / 2


def entropy_MvNormal(K, logdetSigma):
    """src/logden_normal.jl:74"""
    return (K * math.log(2 * math.pi * math.e) + logdetSigma) / 2


# --------------------------------------------------------------------------------------
# problem description crossing the boundary (SURVEY 8(b) hvi_desc)
# --------------------------------------------------------------------------------------
@dataclass
class Spec:
    nP: int
    nM: int
    nO: int
    nC: int
    layers: List[Tuple[int, int, int, bool]]  # (n_in, n_out, act, has_bias)
    scale_mu: Sequence[float]   # NormalScalingModelApplicator.μ (src/ModelApplicator.jl:132-161)
    scale_sigma: Sequence[float]
    transP: Sequence[int]
    transM: Sequence[int]
    priorsP: Sequence[Tuple[int, float, float]]
    priorsM: Sequence[Tuple[int, float, float]]
    cor_ends_P: Sequence[int]
    cor_ends_M: Sequence[int]
    # (r0, r1, K1, K2) -> (source, index-or-value)
    slots: Sequence[Tuple[int, float]]
    pbm_covar_idx: Sequence[int] = field(default_factory=list)  # 0-based positions in θP
    omit_priors: bool = False
    scale_kind: str = "normal"  # or "range" (offset=scale_mu, width=scale_sigma)
    # "mean": MeanHVIApproximationMat (src/elbo.jl:633-678); "sepvar": MeanVarSepHVIApproximation
    # (src/elbo_sepvec.jl:3-48), ϕq then holds logσ2_ζMs (nM × n_site_total) instead of the coefficients
    approx: str = "mean"
    n_site_total: int = 0
    # generic process-based model f(θP, θMs_tr, xP) -> y (n_obs × nb) (AbstractPBMApplicator, src/PBMApplicator.jl:26-32);
    # None: f_doubleMM_sites.  penalty(θP, θMs_tr, xP, y_pred) -> scalar: floss_penalty (src/elbo.jl:153,282-286)
    pbm: Optional[Callable] = None
    penalty: Optional[Callable] = None

    @property
    def n_phig(self):
        return sum(i * o + (o if b else 0) for i, o, _, b in self.layers)

    @property
    def n_rhoP(self):
        return get_cor_count_total(self.cor_ends_P) if self.nP > 0 else 0

    @property
    def n_rhoM(self):
        return get_cor_count_total(self.cor_ends_M)

    def offsets(self):
        """ϕ = [ϕg ; logσ2_ζP ; coef_logσ2_ζMs (2×nM col-major) ; ρsP ; ρsM ; μP]
        (src/init_hybrid_params.jl:26,119-124; src/HybridProblem.jl:101-104)."""
        o = {}
        p = 0
        n_coef = self.nM * self.n_site_total if self.approx == "sepvar" else 2 * self.nM
        for name, ln in (("phig", self.n_phig), ("logsigma2_P", self.nP), ("coef", n_coef),
                         ("rhoP", self.n_rhoP), ("rhoM", self.n_rhoM), ("muP", self.nP)):
            o[name] = (p, ln)
            p += ln
        o["_len"] = p
        return o

    @property
    def n_phi(self):
        return self.offsets()["_len"]


# --------------------------------------------------------------------------------------
# g: MLP + NormalScaling (src/ModelApplicator.jl:163-176; ext/…SimpleChainsExt.jl:43-52)
# --------------------------------------------------------------------------------------
def mlp(spec: Spec, x: torch.Tensor, phig: torch.Tensor) -> torch.Tensor:
    """x (n_in × n_col); per layer params vec(W[out×in]) column-major, then b."""
    h = x
    p = 0
    for (n_in, n_out, act, has_b) in spec.layers:
        W = phig[p:p + n_in * n_out].reshape(n_in, n_out).T  # col-major (out×in)
        p += n_in * n_out
        z = W @ h
        if has_b:
            z = z + phig[p:p + n_out].reshape(-1, 1)
            p += n_out
        if act == ACT_TANH:
            h = torch.tanh(z)
        elif act == ACT_LOGISTIC:
            h = torch.sigmoid(z)
        else:
            h = z
    return h


def apply_g(spec: Spec, x: torch.Tensor, phig: torch.Tensor) -> torch.Tensor:
    """norminvcdf.(μ, σ, MLP(x)) = μ + σ·Φ⁻¹(p)  (src/ModelApplicator.jl:168)."""
    p = mlp(spec, x, phig)
    mu = torch.tensor(spec.scale_mu, dtype=F64).reshape(-1, 1)
    sg = torch.tensor(spec.scale_sigma, dtype=F64).reshape(-1, 1)
    if spec.scale_kind == "range":  # src/ModelApplicator.jl:191-194
        return p * sg + mu
    return mu + sg * torch.special.ndtri(p)


# --------------------------------------------------------------------------------------
# src/elbo.jl
# --------------------------------------------------------------------------------------
def _append_each_covars(xM, zP_covar):
    """src/elbo.jl:534-547"""
    if zP_covar.numel() == 0:
        return xM
    rep = zP_covar.reshape(-1, 1).expand(-1, xM.shape[1])
    return torch.cat([xM, rep], dim=0)


def sample_zeta_resid_norm(spec: Spec, zP, zMs, mu_zMs, phiq, i_sites=None):
    """MeanHVIApproximationMat method, src/elbo.jl:633-678 with _compute_choleskyfactor
    :721-743.  zP (n_MC×nP), zMs (n_MC × nM·nb) column index (i·nM + k);
    mu_zMs (nM × nb).  Returns ζP_resids (nP×M), ζMs_parfirst_resids (nM×nb×M), σ."""
    o = _phiq_offsets(spec)
    nM, nb = mu_zMs.shape
    M = zP.shape[0]
    rhoP = phiq[o["rhoP"][0]:o["rhoP"][0] + o["rhoP"][1]]
    rhoM = phiq[o["rhoM"][0]:o["rhoM"][0] + o["rhoM"][1]]
    UP = transformU_block_cholesky1(rhoP, spec.cor_ends_P) if spec.nP > 0 else torch.zeros((0, 0), dtype=F64)
    UM = transformU_block_cholesky1(rhoM, spec.cor_ends_M)
    if spec.approx == "sepvar":   # src/elbo_sepvec.jl:23: logσ2_ζMs[:, i_sites]
        blk = phiq[o["coef"][0]:o["coef"][0] + o["coef"][1]].reshape(spec.n_site_total, nM).T
        idx = torch.arange(nb) if i_sites is None else torch.as_tensor(np.asarray(i_sites), dtype=torch.long)
        logs2M = blk[:, idx]
    else:
        cf = phiq[o["coef"][0]:o["coef"][0] + 2 * nM].reshape(nM, 2).T  # (2 × nM) col-major
        logs2M = cf[0, :].reshape(-1, 1) + cf[1, :].reshape(-1, 1) * mu_zMs
    logs2P = phiq[o["logsigma2_P"][0]:o["logsigma2_P"][0] + spec.nP]
    sigM = torch.exp(logs2M / 2)
    sigP = torch.exp(logs2P / 2)
    # Uσ' * z' per block:  r[k] = σ[k] Σ_j U[j,k] z[j]
    rP = (zP @ UP).T * sigP.reshape(-1, 1) if spec.nP > 0 else torch.zeros((0, M), dtype=F64)
    z3 = zMs.reshape(M, nb, nM)                     # [m, i, j]
    t = torch.einsum("mij,jk->kim", z3, UM)         # [k, i, m]
    rM = t * sigM.unsqueeze(-1)
    sigma = torch.cat([sigP, sigM.T.reshape(-1)])   # vcat(σP, vec(σMs)) col-major
    return rP, rM, sigma


def _phiq_offsets(spec: Spec):
    o = spec.offsets()
    base = o["logsigma2_P"][0]
    return {k: (v[0] - base, v[1]) for k, v in o.items() if k not in ("phig", "_len")}


def generate_zeta(spec: Spec, phi, xM, zP, zMs, i_sites=None):
    """src/elbo.jl:472-521.  Returns ζsP (nP×M), ζsMs_tr (nb×nM×M), σ."""
    o = spec.offsets()
    phig = phi[o["phig"][0]:o["phig"][0] + o["phig"][1]]
    phiq = phi[o["logsigma2_P"][0]:]
    mu_zP = phi[o["muP"][0]:o["muP"][0] + spec.nP]
    idx = list(spec.pbm_covar_idx)
    xMP0 = _append_each_covars(xM, mu_zP[idx] if idx else mu_zP[:0])
    mu_zMs0 = apply_g(spec, xMP0, phig)
    rP, rM, sigma = sample_zeta_resid_norm(spec, zP, zMs, mu_zMs0, phiq, i_sites)
    zsP = mu_zP.reshape(-1, 1) + rP if spec.nP > 0 else rP
    M = zP.shape[0]
    if not idx:
        zsMs_tr = (mu_zMs0.unsqueeze(-1) + rM).permute(1, 0, 2)
    else:
        cols = []
        for m in range(M):
            xMP = _append_each_covars(xM, zsP[idx, m])
            mu_m = apply_g(spec, xMP, phig)
            cols.append((mu_m + rM[:, :, m]).T)
        zsMs_tr = torch.stack(cols, dim=2)
    return zsP, zsMs_tr, sigma


def _clamp_like_reference(theta, dtype_name):
    """src/elbo.jl:797-798, 803-804: θ ← max(√floatmin, min(√floatmax, θ)) for the
    *working* float type of the reference run (Float32 or Float64)."""
    fi = np.finfo(np.float32 if dtype_name == "f32" else np.float64)
    lo, hi = math.sqrt(float(fi.tiny)), math.sqrt(float(fi.max))
    return torch.clamp(theta, min=lo, max=hi)


def _transform(codes, zeta_cols):
    """Stacked/StackedArray of per-column bijectors (src/bijectors_utils.jl:65-114).
    zeta_cols: (..., n_par) last dim = parameter.  Returns θ and logabsdetjac."""
    outs, lj = [], zeta_cols.new_zeros(())
    for k, c in enumerate(codes):
        z = zeta_cols[..., k]
        if c == TR_EXP:
            outs.append(torch.exp(z))
            lj = lj + z.sum()
        elif c == TR_LOGISTIC:
            outs.append(torch.sigmoid(z))
            lj = lj + (torch.nn.functional.logsigmoid(z) + torch.nn.functional.logsigmoid(-z)).sum()
        else:
            outs.append(z)
    th = torch.stack(outs, dim=-1) if outs else zeta_cols
    return th, lj


def transform_and_logjac_zeta(spec: Spec, zP, zMs_tr, dtype_name):
    """src/elbo.jl:783-808"""
    thM, ljM = _transform(spec.transM, zMs_tr)
    thM = _clamp_like_reference(thM, dtype_name)
    if spec.nP == 0:
        return zP, thM, ljM
    thP, ljP = _transform(spec.transP, zP)
    thP = _clamp_like_reference(thP, dtype_name)
    return thP, thM, ljP + ljM


def f_doubleMM_sites(spec: Spec, thP, thMs_tr, xP):
    """src/DoubleMM/f_doubleMM.jl:101-139 through PBMPopulationApplicator
    (src/PBMApplicator.jl:263-292).  xP (2nO × nb) rows S1[1:nO]; S2[1:nO]."""
    nO = spec.nO
    S1, S2 = xP[:nO, :], xP[nO:, :]
    valid = torch.isfinite(S1) & torch.isfinite(S2)
    nb = thMs_tr.shape[0]
    pars = []
    for (src, v) in spec.slots:
        if src == SRC_P:
            p = thP[int(v)].expand(nb)
        elif src == SRC_M:
            p = thMs_tr[:, int(v)]
        else:
            p = torch.full((nb,), float(v), dtype=F64)
        # is_valid .* p' with Julia's strong zero false*NaN == 0
        pars.append(torch.where(valid, p.reshape(1, -1).expand(nO, -1), torch.zeros((), dtype=F64)))
    r0, r1, K1, K2 = pars
    return r0 + r1 * S1 / (K1 + S1) * S2 / (K2 + S2)


def _neg_logpdf(kind, a, b, th):
    """Distributions.logpdf for Normal(μ,σ) / LogNormal(μ,σ) (call sites src/elbo.jl:217-235)."""
    if kind == PR_NORMAL:
        return 0.5 * ((th - a) / b) ** 2 + math.log(b) + 0.5 * math.log(2 * math.pi)
    if kind == PR_LOGNORMAL:
        lt = torch.log(th)
        return lt + math.log(b) + 0.5 * math.log(2 * math.pi) + (lt - a) ** 2 / (2 * b * b)
    return torch.zeros_like(th)


def compute_priors_logdensity(spec: Spec, thP, thMs_tr):
    """src/elbo.jl:203-260"""
    if spec.omit_priors:
        return torch.zeros((), dtype=F64)
    nl = torch.zeros((), dtype=F64)
    for j, (kind, a, b) in enumerate(spec.priorsP):
        nl = nl + _neg_logpdf(kind, a, b, thP[j])
    for k, (kind, a, b) in enumerate(spec.priorsM):
        nl = nl + _neg_logpdf(kind, a, b, thMs_tr[:, k]).sum()
    return nl


def neg_elbo_zeta_tf(spec: Spec, zsP, zsMs_tr, sigma, xP, y_ob, y_unc, dtype_name):
    """src/elbo.jl:132-201"""
    M = zsMs_tr.shape[2]
    nLy = nlp = nlj = pen = torch.zeros((), dtype=F64)
    for m in range(M):
        zP_m = zsP[:, m]
        thP, thM, lj = transform_and_logjac_zeta(spec, zP_m, zsMs_tr[:, :, m], dtype_name)
        y_pred = spec.pbm(thP, thM, xP) if spec.pbm is not None else f_doubleMM_sites(spec, thP, thM, xP)
        nLy = nLy + neg_logden_indep_normal(y_ob, y_pred, y_unc)
        if spec.penalty is not None:
            pen = pen + spec.penalty(thP, thM, xP, y_pred)
        nlp = nlp + compute_priors_logdensity(spec, thP, thM)
        nlj = nlj - lj
    nLy, nlp, nlj = nLy / M, nlp / M, nlj / M
    logdetS = 2 * torch.log(sigma).sum()
    n_theta = spec.nP + zsMs_tr.shape[0] * zsMs_tr.shape[1]
    assert sigma.numel() == n_theta
    ent = entropy_MvNormal(n_theta, logdetS)
    return dict(nLjoint=nLy + nlp + nlj, entropy_zeta=ent, loss_penalty=pen / M,
                nLy=nLy, neg_log_prior=nlp, neg_log_jac=nlj)


COMPONENTS = ("nLjoint", "entropy_zeta", "loss_penalty", "nLy", "neg_log_prior", "neg_log_jac")


def neg_elbo_gtf_components(spec: Spec, phi, xM, xP, y_ob, y_unc, zP, zMs, dtype_name="f64", i_sites=None):
    """src/elbo.jl:46-88 (n_MC_cap == n_MC; priors_θ_mean empty)."""
    zsP, zsMs_tr, sigma = generate_zeta(spec, phi, xM, zP, zMs, i_sites)
    return neg_elbo_zeta_tf(spec, zsP, zsMs_tr, sigma, xP, y_ob, y_unc, dtype_name)


def neg_elbo_gtf(spec: Spec, phi, xM, xP, y_ob, y_unc, zP, zMs, dtype_name="f64"):
    """src/elbo.jl:30-44: nL = nLjoint − entropy_ζ + loss_penalty."""
    c = neg_elbo_gtf_components(spec, phi, xM, xP, y_ob, y_unc, zP, zMs, dtype_name)
    return c["nLjoint"] - c["entropy_zeta"] + c["loss_penalty"]


def value_and_grad(spec: Spec, phi, xM, xP, y_ob, y_unc, zP, zMs, dtype_name="f64", i_sites=None):
    """Scalar, six components and ∇ϕ by torch reverse-mode autograd in float64 (stands in
    for Zygote.withgradient, src/HybridSolver.jl:256-258).  Inputs: numpy arrays in the
    reference's shapes; they are converted (exactly) to float64."""
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    ph = t(phi).clone().requires_grad_(True)
    comps = neg_elbo_gtf_components(spec, ph, t(xM), t(xP), t(y_ob), t(y_unc), t(zP), t(zMs), dtype_name, i_sites)
    nL = comps["nLjoint"] - comps["entropy_zeta"] + comps["loss_penalty"]
    (g,) = torch.autograd.grad(nL, ph)
    cv = np.array([float(comps[k].detach()) for k in COMPONENTS])
    return float(nL.detach()), cv, g.numpy().copy()


# --------------------------------------------------------------------------------------
# src/gf.jl: the point-solver loss
# --------------------------------------------------------------------------------------
def loss_gf_components(spec: Spec, phi_gf, xM, xP, y_ob, y_unc):
    """loss_gf (src/gf.jl:227-295) with gf (:153-177) and gtrans (:183-199): ϕ = [ϕg; ϕP];
    θMs = transMs(g([xM; ϕP[covars]], ϕg)'), θP = transP(ϕP) -- no clamp on this path; nLy = py(y_o, y_pred, y_unc)
    (src/logden_normal.jl:16-43); neg_log_prior (src/elbo.jl:203-260); floss_penalty = zero."""
    n_g = spec.n_phig
    phig, zP = phi_gf[:n_g], phi_gf[n_g:n_g + spec.nP]
    idx = list(spec.pbm_covar_idx)
    xMP = _append_each_covars(xM, zP[idx] if idx else zP[:0])
    zMs_tr = apply_g(spec, xMP, phig).T
    thM, _ = _transform(spec.transM, zMs_tr)
    thP, _ = _transform(spec.transP, zP) if spec.nP > 0 else (zP, None)
    y_pred = f_doubleMM_sites(spec, thP, thM, xP)
    nLy = neg_logden_indep_normal(y_ob, y_pred, y_unc)
    nlp = compute_priors_logdensity(spec, thP, thM)
    return dict(nLjoint_pen=nLy + nlp, nLy=nLy, neg_log_prior=nlp)


def loss_gf_value_and_grad(spec: Spec, phi_gf, xM, xP, y_ob, y_unc):
    """value, (nLjoint_pen, nLy, neg_log_prior) and ∇ϕ by float64 autograd (stands in for the AutoZygote gradient of
    src/HybridSolver.jl:95-100)."""
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    ph = t(phi_gf).clone().requires_grad_(True)
    c = loss_gf_components(spec, ph, t(xM), t(xP), t(y_ob), t(y_unc))
    (g,) = torch.autograd.grad(c["nLjoint_pen"], ph)
    return float(c["nLjoint_pen"].detach()), np.array([float(c[k].detach()) for k in ("nLjoint_pen", "nLy", "neg_log_prior")]), \
        g.numpy().copy()


# --------------------------------------------------------------------------------------
# prior / scaling fits done by DistributionFits in the reference (inputs to the boundary)
# --------------------------------------------------------------------------------------
Z95 = 1.6448536269514722   # quantile(Normal(), 0.95)
Z975 = 1.959963984540054


def fit_lognormal_mode_q95(mode, q95):
    """fit(LogNormal, mode, QuantilePoint(q95, 0.95), Val(:mode)) (src/DoubleMM/f_doubleMM.jl:188):
    mode = exp(μ−σ²), q = exp(μ+zσ)  ⇒  σ² + zσ − ln(q/mode) = 0."""
    L = math.log(q95 / mode)
    sigma = (-Z95 + math.sqrt(Z95 * Z95 + 4 * L)) / 2
    return (math.log(mode) + sigma * sigma, sigma)


def fit_lognormal_mean_q95(mean, q95):
    """fit(LogNormal, mean, QuantilePoint(q95,0.95)) -- DistributionFits default Val(:mean)
    (test/test_HybridProblem.jl:84-85): mean = exp(μ+σ²/2)  ⇒  σ²/2 − zσ + ln(q/mean) = 0,
    smaller root."""
    L = math.log(q95 / mean)
    sigma = Z95 - math.sqrt(Z95 * Z95 - 2 * L)
    return (math.log(mean) - sigma * sigma / 2, sigma)


def fit_normal_mode_q975(mode, q975):
    """fit(Normal, mode, qp_uu(q)) (test/test_HybridProblem.jl:86)."""
    return (mode, (q975 - mode) / Z975)


def quantile_prior(kind, a, b, z):
    return a + b * z if kind == PR_NORMAL else math.exp(a + b * z)


def normal_scaling_from_priors(priorsM, transM):
    """get_quantile_transformed (src/HybridProblem.jl:291-295) + fit(Normal, @qp_l, @qp_u)
    (src/ModelApplicator.jl:150-161): 5 %/95 % prior quantiles pulled back through transM,
    μ=(l+u)/2, σ=(u−l)/(2·z95)."""
    mus, sgs = [], []
    for (kind, a, b), tr in zip(priorsM, transM):
        lo, hi = quantile_prior(kind, a, b, -Z95), quantile_prior(kind, a, b, Z95)
        if tr == TR_EXP:
            lo, hi = math.log(lo), math.log(hi)
        elif tr == TR_LOGISTIC:
            lo, hi = math.log(lo / (1 - lo)), math.log(hi / (1 - hi))
        mus.append((lo + hi) / 2)
        sgs.append((hi - lo) / (2 * Z95))
    return mus, sgs


# --------------------------------------------------------------------------------------
# prediction path (src/elbo.jl:320-458, 815-835; src/PBMApplicator.jl:51-57)
# --------------------------------------------------------------------------------------
def transform_zetas(spec: Spec, zsP, zsMs_tr):
    """transform_ζs (src/elbo.jl:815-835): bijectors only, no clamp, no log-Jacobian."""
    thP = _transform(spec.transP, zsP.T)[0].T if spec.nP > 0 else zsP
    thMs = _transform(spec.transM, zsMs_tr.permute(0, 2, 1))[0].permute(0, 2, 1)
    return thP, thMs


def predict_hvi(spec: Spec, phi, xM, xP, zP, zMs, i_sites=None):
    """predict_hvi (src/elbo.jl:320-352) with the draws passed in: y (nO × nb × n_sample),
    θsP (nP × n_sample), θsMs_tr (nb × nM × n_sample), entropy_ζ."""
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    zsP, zsMs_tr, sigma = generate_zeta(spec, t(phi), t(xM), t(zP), t(zMs), i_sites)
    ent = entropy_MvNormal(sigma.numel(), 2 * torch.log(sigma).sum())
    thP, thMs = transform_zetas(spec, zsP, zsMs_tr)
    ys = [f_doubleMM_sites(spec, thP[:, m] if spec.nP > 0 else thP[:0, 0], thMs[:, :, m], t(xP)) for m in range(thMs.shape[2])]
    return torch.stack(ys, dim=2).numpy(), thP.numpy(), thMs.numpy(), float(ent)
