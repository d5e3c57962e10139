/* CPU ORACLE, C restatement (test infrastructure and timed CPU baseline -- NOT the product).
 *
 * Restates the reference's negative-ELBO evaluation (src/elbo.jl:30-88, 132-260, 472-521,
 * 633-678, 783-808; src/cholesky.jl:52-73,156-174,262-313; src/logden_normal.jl:16-43,74;
 * src/DoubleMM/f_doubleMM.jl:101-139; src/ModelApplicator.jl:163-176 -- all relative to
 * /root/reference) with a hand-derived reverse pass, as plain loops
 * over sites and Monte-Carlo draws.  It is checked against the torch-autograd restatement
 * (oracle/hvi_oracle.py) in tests/test_oracle_c.py, and is the "port" CPU baseline that
 * bench.py times.  Parity status: see the header of oracle/hvi_oracle.py ("unpinned" for
 * the whole-path value; components pinned against the reference's known-answer tests).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 *
 * Build: gcc -O2 -fopenmp -fPIC -shared -DREAL=double -DSUFFIX=_f64 ...   (oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/hvi_b200.h"

#ifndef REAL
#define REAL double
#endif
#ifndef SUFFIX
#define SUFFIX _f64
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

typedef REAL real;

#define MAXW 1024 /* widest layer handled by the oracle */

/* ---- Φ⁻¹ (StatsFuns.norminvcdf; Wichura AS241 PPND16, |rel err| < 1e-16) -------------- */
static double ndtri(double p) {
  double q = p - 0.5, r, v;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    v = q * (((((((2.5090809287301226727e3 * r + 3.3430575583588128105e4) * r + 6.7265770927008700853e4) * r +
                 4.5921953931549871457e4) * r + 1.3731693765509461125e4) * r + 1.9715909503065514427e3) * r +
               1.3314166789178437745e2) * r + 3.3871328727963666080e0) /
        (((((((5.2264952788528545610e3 * r + 2.8729085735721942674e4) * r + 3.9307895800092710610e4) * r +
             2.1213794301586595867e4) * r + 5.3941960214247511077e3) * r + 6.8718700749205790830e2) * r +
           4.2313330701600911252e1) * r + 1.0);
    return v;
  }
  r = q < 0 ? p : 1.0 - p;
  if (r <= 0) return q < 0 ? -INFINITY : INFINITY;
  r = sqrt(-log(r));
  if (r <= 5.0) {
    r -= 1.6;
    v = (((((((7.74545014278341407640e-4 * r + 2.27238449892691845833e-2) * r + 2.41780725177450611770e-1) * r +
             1.27045825245236838258e0) * r + 3.64784832476320460504e0) * r + 5.76949722146069140550e0) * r +
          4.63033784615654529590e0) * r + 1.42343711074968357734e0) /
        (((((((1.05075007164441684324e-9 * r + 5.47593808499534494600e-4) * r + 1.51986665636164571966e-2) * r +
             1.48103976427480074590e-1) * r + 6.89767334985100004550e-1) * r + 1.67638483018380384940e0) * r +
           2.05319162663775882187e0) * r + 1.0);
  } else {
    r -= 5.0;
    v = (((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 1.24266094738807843860e-3) * r +
             2.65321895265761230930e-2) * r + 2.96560571828504891230e-1) * r + 1.78482653991729133580e0) * r +
          5.46378491116411436990e0) * r + 6.65790464350110377720e0) /
        (((((((2.04426310338993978564e-15 * r + 1.42151175831644588870e-7) * r + 1.84631831751005468180e-5) * r +
             7.86869131145613259100e-4) * r + 1.48753612908506148525e-2) * r + 1.36929880922735805310e-1) * r +
           5.99832206555887937690e-1) * r + 1.0);
  }
  return q < 0 ? -v : v;
}

typedef struct {
  int nP, nM, nO, nC, nL;
  int woff[HVI_MAX_LAYERS], boff[HVI_MAX_LAYERS];
  int nphig, n_in;
  int o_ls2P, o_coef, o_rhoP, o_rhoM, o_muP, nphi;
  int blk_start_P[HVI_MAX_PAR], rho_off_P[HVI_MAX_PAR];
  int blk_start_M[HVI_MAX_PAR], rho_off_M[HVI_MAX_PAR];
  int slot_code[4];
  double clamp_lo, clamp_hi;
  int sepvar;   /* MeanVarSepHVIApproximation: o_coef is the offset of logσ2_ζMs (nM x n_site_total) */
} layout_t;

static int cor_count_n(int n) { return (n - 1) * n / 2; }

/* per column k: first column of its block and offset of its packed coefficients
 * (intended cumulative offsets) */
static int block_maps(const int32_t* ends, int n_ends, int n, int* blk_start, int* rho_off) {
  int total = 0, start = 0, b, k;
  if (n == 0) return 0;
  if (n_ends == 0) { /* single parameter, no correlation (src/cholesky.jl:265-269) */
    for (k = 0; k < n; k++) { blk_start[k] = k; rho_off[k] = 0; }
    return 0;
  }
  for (b = 0; b < n_ends; b++) {
    int end = ends[b];
    for (k = start; k < end; k++) {
      int kk = k - start;
      blk_start[k] = start;
      rho_off[k] = total + kk * (kk - 1) / 2;
    }
    total += cor_count_n(end - start);
    start = end;
  }
  return total;
}

static void make_layout(const hvi_desc* d, layout_t* L) {
  int l, p = 0, s;
  memset(L, 0, sizeof(*L));
  L->nP = d->n_P; L->nM = d->n_M; L->nO = d->n_obs; L->nC = d->n_covar; L->nL = d->n_layers;
  for (l = 0; l < d->n_layers; l++) {
    L->woff[l] = p; p += d->layers[l].n_in * d->layers[l].n_out;
    L->boff[l] = p; if (d->layers[l].has_bias) p += d->layers[l].n_out;
  }
  L->nphig = p; L->n_in = d->layers[0].n_in;
  L->o_ls2P = p; p += L->nP;
  L->sepvar = d->approx == HVI_APPROX_MEANVAR_SEP;
  L->o_coef = p; p += L->sepvar ? L->nM * (int)d->n_site_total : 2 * L->nM;
  L->o_rhoP = p; p += block_maps(d->cor_ends_P, d->n_cor_ends_P, L->nP, L->blk_start_P, L->rho_off_P);
  L->o_rhoM = p; p += block_maps(d->cor_ends_M, d->n_cor_ends_M, L->nM, L->blk_start_M, L->rho_off_M);
  L->o_muP = p; p += L->nP;
  L->nphi = p;
  for (s = 0; s < 4; s++)
    L->slot_code[s] = d->slot_src[s] == HVI_SRC_P ? d->slot_idx[s]
                    : d->slot_src[s] == HVI_SRC_M ? L->nP + d->slot_idx[s] : -1;
  if (d->dtype == HVI_F32) { L->clamp_lo = sqrt(1.17549435082228750797e-38); L->clamp_hi = sqrt(3.40282346638528859812e+38); }
  else { L->clamp_lo = sqrt(2.2250738585072014e-308); L->clamp_hi = sqrt(1.7976931348623157e+308); }
}

/* U = unit-upper(v) with unit-norm columns, per block (src/cholesky.jl:156-174,262-313) */
static void build_U(int n, const int* blk_start, const int* rho_off, const real* rho, real U[HVI_MAX_PAR][HVI_MAX_PAR],
                    real* colnorm) {
  int i, k;
  for (k = 0; k < n; k++) {
    real ss = 1;
    for (i = 0; i < n; i++) U[i][k] = 0;
    for (i = blk_start[k]; i < k; i++) { real v = rho[rho_off[k] + (i - blk_start[k])]; U[i][k] = v; ss += v * v; }
    U[k][k] = 1;
    colnorm[k] = (real)sqrt((double)ss);
    for (i = blk_start[k]; i <= k; i++) U[i][k] /= colnorm[k];
  }
}

/* pullback of build_U: rho_bar += strictly-upper part of ((Ubar - U (U.Ubar)) / norm) */
static void build_U_adj(int n, const int* blk_start, const int* rho_off, real U[HVI_MAX_PAR][HVI_MAX_PAR],
                        const real* colnorm, double Ubar[HVI_MAX_PAR][HVI_MAX_PAR], double* rho_bar) {
  int i, k;
  for (k = 0; k < n; k++) {
    double dot = 0;
    for (i = blk_start[k]; i <= k; i++) dot += (double)U[i][k] * Ubar[i][k];
    for (i = blk_start[k]; i < k; i++) rho_bar[rho_off[k] + (i - blk_start[k])] += (Ubar[i][k] - (double)U[i][k] * dot) / (double)colnorm[k];
  }
}

/* θ = T(ζ) with the reference's clamp; returns θ, dθ/dζ (0 when clamped), the log-Jacobian
 * term and its derivative (src/elbo.jl:783-808, src/bijectors_utils.jl:10-23,38-52) */
static void transform1(int code, real zeta, double lo, double hi, real* th, real* dth, real* lj, real* dlj) {
  real t, d;
  if (code == HVI_TR_EXP) { t = (real)exp((double)zeta); d = t; *lj = zeta; *dlj = 1; }
  else if (code == HVI_TR_LOGISTIC) {
    t = (real)(1.0 / (1.0 + exp(-(double)zeta))); d = t * (1 - t);
    *lj = (real)(-log1p(exp(-(double)zeta)) - log1p(exp((double)zeta))); *dlj = 1 - 2 * t;
  } else { t = zeta; d = 1; *lj = 0; *dlj = 0; }
  if (t > (real)hi) { t = (real)hi; d = 0; }
  if (t < (real)lo) { t = (real)lo; d = 0; }
  *th = t; *dth = d;
}

/* −logpdf(prior, θ) and its derivative w.r.t. θ */
static void prior1(const hvi_prior* pr, real th, real* nl, real* dnl) {
  if (pr->kind == HVI_PR_NORMAL) {
    real u = (th - (real)pr->a) / (real)pr->b;
    *nl = (real)0.5 * u * u + (real)(log(pr->b) + 0.9189385332046727418);
    *dnl = u / (real)pr->b;
  } else if (pr->kind == HVI_PR_LOGNORMAL) {
    real lt = (real)log((double)th), u = (lt - (real)pr->a) / (real)pr->b;
    *nl = lt + (real)0.5 * u * u + (real)(log(pr->b) + 0.9189385332046727418);
    *dnl = (1 + u / (real)pr->b) / th;
  } else { *nl = 0; *dnl = 0; }
}

static real act_f(int act, real z) {
  if (act == HVI_ACT_TANH) return (real)tanh((double)z);
  if (act == HVI_ACT_LOGISTIC) return (real)(1.0 / (1.0 + exp(-(double)z)));
  return z;
}
static real act_d(int act, real a) { /* derivative expressed through the activation value */
  if (act == HVI_ACT_TANH) return 1 - a * a;
  if (act == HVI_ACT_LOGISTIC) return a * (1 - a);
  return 1;
}

/* g forward for one column: acts[l] (l = 0 input … nL = output p), returns μ_ζ[k], q[k] */
static void g_forward(const hvi_desc* d, const layout_t* L, const real* phig, const real* x, real** acts, real* mu_zeta,
                      real* q) {
  int l, i, j, k;
  for (i = 0; i < L->n_in; i++) acts[0][i] = x[i];
  for (l = 0; l < L->nL; l++) {
    int ni = d->layers[l].n_in, no = d->layers[l].n_out;
    const real* W = phig + L->woff[l];
    for (j = 0; j < no; j++) {
      real z = d->layers[l].has_bias ? phig[L->boff[l] + j] : 0;
      for (i = 0; i < ni; i++) z += W[i * no + j] * acts[l][i];
      acts[l + 1][j] = act_f(d->layers[l].act, z);
    }
  }
  for (k = 0; k < L->nM; k++) {
    real p = acts[L->nL][k];
    if (d->scale_kind == HVI_SCALE_RANGE) { q[k] = p; mu_zeta[k] = p * (real)d->scale_b[k] + (real)d->scale_a[k]; }
    else { q[k] = (real)ndtri((double)p); mu_zeta[k] = (real)d->scale_a[k] + (real)d->scale_b[k] * q[k]; }
  }
}

/* g backward for one column: adds to gphig, returns cotangent of the input x in xbar */
static void g_backward(const hvi_desc* d, const layout_t* L, const real* phig, real** acts, const real* q,
                       const real* mu_bar, double* gphig, real* xbar, real* da, real* db) {
  int l, i, j, k, nout = d->layers[L->nL - 1].n_out;
  for (k = 0; k < nout; k++) {
    real pb = 0;
    if (k < L->nM) {
      if (d->scale_kind == HVI_SCALE_RANGE) pb = mu_bar[k] * (real)d->scale_b[k];
      else pb = mu_bar[k] * (real)d->scale_b[k] * (real)(2.50662827463100050242 * exp(0.5 * (double)q[k] * (double)q[k]));
    }
    da[k] = pb;
  }
  for (l = L->nL - 1; l >= 0; l--) {
    int ni = d->layers[l].n_in, no = d->layers[l].n_out;
    const real* W = phig + L->woff[l];
    for (j = 0; j < no; j++) da[j] *= act_d(d->layers[l].act, acts[l + 1][j]); /* δ_l */
    for (i = 0; i < ni; i++) db[i] = 0;
    for (j = 0; j < no; j++) {
      real dl = da[j];
      if (d->layers[l].has_bias) gphig[L->boff[l] + j] += (double)dl;
      for (i = 0; i < ni; i++) {
        gphig[L->woff[l] + i * no + j] += (double)(dl * acts[l][i]);
        db[i] += W[i * no + j] * dl;
      }
    }
    for (i = 0; i < ni; i++) da[i] = db[i];
  }
  for (i = 0; i < L->n_in; i++) xbar[i] = da[i];
}

/* one (site, draw): everything after ζ is known.  Returns contributions (unscaled by 1/M). */
typedef struct { double nly, nlp, lj; } draw_sums;

static void draw_eval(const hvi_desc* d, const layout_t* L, const real* thP, const real* zetaM, const real* S1,
                      const real* S2, const real* yo, const real* yu, draw_sums* s, real* zetaM_bar,
                      real* thP_bar /* += */) {
  int k, o, sl, nP = L->nP, nM = L->nM;
  real thM[HVI_MAX_PAR], dth[HVI_MAX_PAR], dlj[HVI_MAX_PAR], dnl[HVI_MAX_PAR], par[4], pbar[4] = {0, 0, 0, 0};
  real thbar[HVI_MAX_PAR];
  for (k = 0; k < nM; k++) {
    real lj, nl = 0;
    transform1(d->trans_M[k], zetaM[k], L->clamp_lo, L->clamp_hi, &thM[k], &dth[k], &lj, &dlj[k]);
    s->lj += (double)lj;
    dnl[k] = 0;
    if (!(d->flags & HVI_FLAG_OMIT_PRIORS)) { prior1(&d->prior_M[k], thM[k], &nl, &dnl[k]); s->nlp += (double)nl; }
    thbar[k] = 0;
  }
  for (sl = 0; sl < 4; sl++) {
    int c = L->slot_code[sl];
    par[sl] = c < 0 ? (real)d->slot_fix[sl] : c < nP ? thP[c] : thM[c - nP];
  }
  for (o = 0; o < L->nO; o++) {
    real s1 = S1[o], s2 = S2[o], y, res, dy, e, A, B, d1, d2;
    /* is_valid .* p and NaN-observation compaction (f_doubleMM.jl:111-137; logden_normal.jl:34-39):
     * an entry contributes iff the observation is not NaN; an invalid driver with a finite
     * observation yields NaN exactly like the reference */
    if (isnan((double)yo[o])) continue;
    if (!(isfinite((double)s1) && isfinite((double)s2))) { s->nly += NAN; continue; }
    d1 = par[2] + s1; d2 = par[3] + s2;
    A = s1 / d1; B = s2 / d2;
    y = par[0] + par[1] * s1 / d1 * s2 / d2;
    res = y - yo[o];
    e = (real)exp(-(double)yu[o]);
    s->nly += 0.5 * ((double)yu[o] + (double)(res * res * e));
    dy = res * e;
    pbar[0] += dy;
    pbar[1] += dy * A * B;
    pbar[2] -= dy * par[1] * A * B / d1;
    pbar[3] -= dy * par[1] * A * B / d2;
  }
  for (sl = 0; sl < 4; sl++) {
    int c = L->slot_code[sl];
    if (c < 0) continue;
    if (c < nP) thP_bar[c] += pbar[sl]; else thbar[c - nP] += pbar[sl];
  }
  for (k = 0; k < nM; k++) zetaM_bar[k] = (thbar[k] + dnl[k]) * dth[k] - dlj[k];
}

/* The whole evaluation.  phi (nphi), xM (nC×nb), xP (2nO×nb), y_o,y_unc (nO×nb),
 * zP (M×nP), zMs (M×nM·nb), all column-major `real`.  grad may be NULL. */
int FN(hvi_oracle_neg_elbo_grad)(const hvi_desc* d, int64_t nb, int M, const real* phi, const real* xM, const real* xP,
                                 const real* y_o, const real* y_unc, const real* zP, const real* zMs, double comps[6],
                                 double* nL_out, real* grad, int nthreads, const int64_t* i_sites) {
  layout_t L;
  real UP[HVI_MAX_PAR][HVI_MAX_PAR], UM[HVI_MAX_PAR][HVI_MAX_PAR], nrmP[HVI_MAX_PAR], nrmM[HVI_MAX_PAR];
  real sigP[HVI_MAX_PAR];
  int nP, nM, nO, nC, j, jj, k, m, npc = d->n_pbm_covar;
  double w;
  double tot_nly = 0, tot_nlp = 0, tot_lj = 0, tot_logsig = 0;
  double* G; /* gradient accumulator, double */
  double UMbar[HVI_MAX_PAR][HVI_MAX_PAR], UPbar[HVI_MAX_PAR][HVI_MAX_PAR];
  real *rP, *zetaP, *thP, *gP;   /* [nP][M] */
  double *zetaP_bar;             /* [nP][M] */
  make_layout(d, &L);
  nP = L.nP; nM = L.nM; nO = L.nO; nC = L.nC; w = 1.0 / M;
  if (L.n_in != nC + npc) return HVI_EINVAL;
  G = (double*)calloc((size_t)L.nphi, sizeof(double));
  rP = (real*)calloc((size_t)(4 * (nP + 1) * M), sizeof(real));
  zetaP = rP + (nP + 1) * M; thP = zetaP + (nP + 1) * M; gP = thP + (nP + 1) * M;
  zetaP_bar = (double*)calloc((size_t)((nP + 1) * M), sizeof(double));
  memset(UMbar, 0, sizeof(UMbar)); memset(UPbar, 0, sizeof(UPbar));
  build_U(nP, L.blk_start_P, L.rho_off_P, phi + L.o_rhoP, UP, nrmP);
  build_U(nM, L.blk_start_M, L.rho_off_M, phi + L.o_rhoM, UM, nrmM);
  /* global block: σP, residuals, ζP, θP per draw (src/elbo.jl:657,664-670,496) */
  for (j = 0; j < nP; j++) sigP[j] = (real)exp((double)phi[L.o_ls2P + j] / 2);
  for (m = 0; m < M; m++)
    for (j = 0; j < nP; j++) {
      real t = 0, lj, dlj, nl = 0, dnl = 0;
      for (jj = L.blk_start_P[j]; jj <= j; jj++) t += UP[jj][j] * zP[m + (int64_t)M * jj];
      rP[j * M + m] = sigP[j] * t;
      zetaP[j * M + m] = phi[L.o_muP + j] + rP[j * M + m];
      transform1(d->trans_P[j], zetaP[j * M + m], L.clamp_lo, L.clamp_hi, &thP[j * M + m], &gP[j * M + m], &lj, &dlj);
      if (d->count_globals) {
        tot_lj += (double)lj;
        if (!(d->flags & HVI_FLAG_OMIT_PRIORS)) { prior1(&d->prior_P[j], thP[j * M + m], &nl, &dnl); tot_nlp += (double)nl; }
        zetaP_bar[j * M + m] += w * ((double)(dnl * gP[j * M + m]) - (double)dlj);
      }
    }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel private(m, k, j, jj)
  {
    double* Gt = (double*)calloc((size_t)L.nphi, sizeof(double));
    double* zPb = (double*)calloc((size_t)((nP + 1) * M), sizeof(double));
    double UMb[HVI_MAX_PAR][HVI_MAX_PAR];
    double t_nly = 0, t_nlp = 0, t_lj = 0, t_ls = 0;
    real* actbuf = (real*)malloc(sizeof(real) * (size_t)(HVI_MAX_LAYERS + 1) * MAXW);
    real* acts[HVI_MAX_LAYERS + 1];
    real da[MAXW], db[MAXW], x[MAXW], xbar[MAXW];
    int l;
    int64_t i;
    for (l = 0; l <= HVI_MAX_LAYERS; l++) acts[l] = actbuf + (size_t)l * MAXW;
    memset(UMb, 0, sizeof(UMb));
#pragma omp for schedule(static)
    for (i = 0; i < nb; i++) {
      real mu[HVI_MAX_PAR], q[HVI_MAX_PAR], sig[HVI_MAX_PAR], ls[HVI_MAX_PAR];
      real mu_bar0[HVI_MAX_PAR]; /* cotangent of the pass-1 mean (through σ only when pbm covars) */
      double lsbar[HVI_MAX_PAR], mubar_sum[HVI_MAX_PAR];
      const real *S1 = xP + (int64_t)2 * nO * i, *S2 = S1 + nO, *yo = y_o + (int64_t)nO * i, *yu = y_unc + (int64_t)nO * i;
      int c, kk;
      for (c = 0; c < nC; c++) x[c] = xM[(int64_t)nC * i + c];
      for (c = 0; c < npc; c++) x[nC + c] = phi[L.o_muP + d->pbm_covar_idx[c]];
      g_forward(d, &L, phi, x, acts, mu, q);       /* pass 1: μ_ζMs0 (src/elbo.jl:491-492) */
      const int64_t isite = i_sites ? i_sites[i] : i;   /* src/elbo_sepvec.jl:23 */
      for (k = 0; k < nM; k++) {
        ls[k] = L.sepvar ? phi[L.o_coef + isite * nM + k] : phi[L.o_coef + 2 * k] + phi[L.o_coef + 2 * k + 1] * mu[k];
        sig[k] = (real)exp((double)ls[k] / 2);
        t_ls += (double)ls[k] / 2;
        lsbar[k] = 0; mubar_sum[k] = 0;
      }
      for (m = 0; m < M; m++) {
        real z[HVI_MAX_PAR], r[HVI_MAX_PAR], zeta[HVI_MAX_PAR], zbar[HVI_MAX_PAR], thPm[HVI_MAX_PAR], thPbar[HVI_MAX_PAR];
        real mu_m[HVI_MAX_PAR], q_m[HVI_MAX_PAR];
        draw_sums s = {0, 0, 0};
        for (k = 0; k < nM; k++) z[k] = zMs[m + (int64_t)M * (i * nM + k)];
        for (j = 0; j < nP; j++) { thPm[j] = thP[j * M + m]; thPbar[j] = 0; }
        if (npc > 0) { /* pass 2: g([xM; ζP_m[idx]]) per draw (src/elbo.jl:508-515) */
          for (c = 0; c < npc; c++) x[nC + c] = zetaP[d->pbm_covar_idx[c] * M + m];
          g_forward(d, &L, phi, x, acts, mu_m, q_m);
        } else for (k = 0; k < nM; k++) mu_m[k] = mu[k];
        for (k = 0; k < nM; k++) {
          real t = 0;
          for (kk = L.blk_start_M[k]; kk <= k; kk++) t += UM[kk][k] * z[kk];
          r[k] = sig[k] * t;
          zeta[k] = mu_m[k] + r[k];
        }
        draw_eval(d, &L, thPm, zeta, S1, S2, yo, yu, &s, zbar, thPbar);
        t_nly += s.nly; t_nlp += s.nlp; t_lj += s.lj;
        for (k = 0; k < nM; k++) {
          double zb = w * (double)zbar[k];
          lsbar[k] += 0.5 * zb * (double)r[k];
          for (kk = L.blk_start_M[k]; kk <= k; kk++) UMb[kk][k] += zb * (double)sig[k] * (double)z[kk];
          mubar_sum[k] += zb;
        }
        for (j = 0; j < nP; j++) zPb[j * M + m] += w * (double)(thPbar[j] * gP[j * M + m]);
        if (npc > 0) { /* pass-2 network backward with this draw's cotangent */
          real mb[HVI_MAX_PAR];
          for (k = 0; k < nM; k++) mb[k] = (real)(w * (double)zbar[k]);
          g_backward(d, &L, phi, acts, q_m, mb, Gt, xbar, da, db);
          for (c = 0; c < npc; c++) zPb[d->pbm_covar_idx[c] * M + m] += (double)xbar[nC + c];
        }
      }
      for (k = 0; k < nM; k++) {
        lsbar[k] -= 0.5; /* entropy: −Σ log σ = −Σ ls/2 */
        if (L.sepvar) {
          Gt[L.o_coef + isite * nM + k] += lsbar[k];
          mu_bar0[k] = (real)(npc > 0 ? 0.0 : mubar_sum[k]);
        } else {
          Gt[L.o_coef + 2 * k] += lsbar[k];
          Gt[L.o_coef + 2 * k + 1] += lsbar[k] * (double)mu[k];
          mu_bar0[k] = (real)(lsbar[k] * (double)phi[L.o_coef + 2 * k + 1] + (npc > 0 ? 0.0 : mubar_sum[k]));
        }
      }
      if (npc > 0) { /* redo pass 1 to have its activations for the backward */
        for (c = 0; c < npc; c++) x[nC + c] = phi[L.o_muP + d->pbm_covar_idx[c]];
        g_forward(d, &L, phi, x, acts, mu, q);
      }
      g_backward(d, &L, phi, acts, q, mu_bar0, Gt, xbar, da, db);
      for (c = 0; c < npc; c++) Gt[L.o_muP + d->pbm_covar_idx[c]] += (double)xbar[nC + c];
    }
#pragma omp critical
    {
      int e, a, b;
      for (e = 0; e < L.nphi; e++) G[e] += Gt[e];
      for (e = 0; e < nP * M; e++) zetaP_bar[e] += zPb[e];
      for (a = 0; a < nM; a++) for (b = 0; b < nM; b++) UMbar[a][b] += UMb[a][b];
      tot_nly += t_nly; tot_nlp += t_nlp; tot_lj += t_lj; tot_logsig += t_ls;
    }
    free(Gt); free(zPb); free(actbuf);
  }
  /* global block adjoint */
  for (j = 0; j < nP; j++) {
    double lsb = 0;
    for (m = 0; m < M; m++) {
      double zb = zetaP_bar[j * M + m];
      G[L.o_muP + j] += zb;
      lsb += 0.5 * zb * (double)rP[j * M + m];
      for (jj = L.blk_start_P[j]; jj <= j; jj++) UPbar[jj][j] += zb * (double)sigP[j] * (double)zP[m + (int64_t)M * jj];
    }
    if (d->count_globals) { lsb -= 0.5; tot_logsig += (double)phi[L.o_ls2P + j] / 2; }
    G[L.o_ls2P + j] += lsb;
  }
  build_U_adj(nP, L.blk_start_P, L.rho_off_P, UP, nrmP, UPbar, G + L.o_rhoP);
  build_U_adj(nM, L.blk_start_M, L.rho_off_M, UM, nrmM, UMbar, G + L.o_rhoM);
  {
    double Kdim = (double)nM * (double)nb + (d->count_globals ? nP : 0);
    double nLy = tot_nly * w, nlp = tot_nlp * w, nlj = -tot_lj * w;
    double ent = 0.5 * Kdim * 2.8378770664093454836 + tot_logsig; /* (K log(2πe) + 2 Σ log σ)/2 */
    comps[0] = nLy + nlp + nlj; comps[1] = ent; comps[2] = 0; comps[3] = nLy; comps[4] = nlp; comps[5] = nlj;
    *nL_out = comps[0] - comps[1] + comps[2];
  }
  if (grad) { int e; for (e = 0; e < L.nphi; e++) grad[e] = (real)G[e]; }
  free(G); free(rP); free(zetaP_bar);
  (void)k;
  return HVI_OK;
}

int FN(hvi_oracle_phi_len)(const hvi_desc* d) { layout_t L; make_layout(d, &L); return L.nphi; }
