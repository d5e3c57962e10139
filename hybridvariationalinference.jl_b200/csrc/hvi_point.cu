// hvi_point.cu -- the point-solver loss loss_gf and its gradient.
//
// loss_gf(ϕ = [ϕg; ϕP]) (src/gf.jl:227-295) = py(y_o, f(θP, θMs, xP), y_unc) + neg_log_prior(θP, θMs) with
// θMs = transM(g([xM; ϕP[covars]], ϕg)') and θP = transP(ϕP) (gf, src/gf.jl:153-177; gtrans :183-199):
// the negative-ELBO path without draws, entropy, log-Jacobian and without the clamp of
// transform_and_logjac_ζ.  g forward/backward are the kernels of the ELBO path (the caller places ϕP in the
// μP slot of a full-layout ϕ); this file adds the per-site kernel between them and the final reduction.
#include <math.h>

#include "hvi_internal.cuh"

namespace hvi {
namespace {

template <typename T>
__device__ __forceinline__ T warp_sum_(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// θ = T(ζ) without clamp, dθ/dζ; −logpdf(prior, θ) including its constant, and d(−logpdf)/dζ
template <typename T>
__device__ __forceinline__ void trans_prior_point(int tr, int pk, T a, T ib, double c0, T zeta, T& th, T& dth, double& nl,
                                                  T& dnl) {
  if (tr == HVI_TR_EXP) { th = exp(zeta); dth = th; }
  else if (tr == HVI_TR_LOGISTIC) { th = T(1) / (T(1) + exp(-zeta)); dth = th * (T(1) - th); }
  else { th = zeta; dth = T(1); }
  nl = 0.0; dnl = T(0);
  if (pk == HVI_PR_LOGNORMAL) {
    const T lt = tr == HVI_TR_EXP ? zeta : log(th);
    const T u = (lt - a) * ib;
    nl = (double)lt + 0.5 * (double)u * (double)u + c0;
    dnl = (T(1) + u * ib) / th * dth;
  } else if (pk == HVI_PR_NORMAL) {
    const T u = (th - a) * ib;
    nl = 0.5 * (double)u * (double)u + c0;
    dnl = u * ib * dth;
  }
}

constexpr int PT_THREADS = 256;

// one thread per site: θM = transM(μ_ζ), f_doubleMM over the site's observations, likelihood, priors of θM and
// the adjoint w.r.t. μ_ζ (→ mubar) and θP (→ per-block partial sums).  part row: [Σ w·res², Σ prior θM, θ̄P[nP], non-finite]
template <typename T>
__global__ void __launch_bounds__(PT_THREADS) k_point(const __grid_constant__ Cfg<T> cfg, const T* __restrict__ phi,
                                                       const T* __restrict__ rec, const T* __restrict__ mu,
                                                       T* __restrict__ mubar, int64_t nb, double* __restrict__ part) {
  __shared__ double sred[PT_THREADS / 32][3 + MAXP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nP = cfg.nP, nM = cfg.nM, nO = cfg.nO;
  T thP[MAXP], dthP[MAXP];
  for (int j = 0; j < nP; j++) {
    double nl; T dnl;
    trans_prior_point<T>(cfg.transP[j], HVI_PR_NONE, T(0), T(0), 0.0, phi[cfg.o_muP + j], thP[j], dthP[j], nl, dnl);
  }
  double a_nly = 0.0, a_pr = 0.0, a_bad = 0.0, a_thP[MAXP];
  for (int j = 0; j < MAXP; j++) a_thP[j] = 0.0;
  const int64_t ntiles = (nb + 31) / 32;
  const int64_t nwarps = (int64_t)gridDim.x * (PT_THREADS / 32);
  for (int64_t tile = (int64_t)blockIdx.x * (PT_THREADS / 32) + warp; tile < ntiles; tile += nwarps) {
    const int64_t site = tile * 32 + lane;
    if (site >= nb) continue;
    const T* r = rec + (tile * 4 * nO) * 32 + lane;
    T thM[MAXP], dthM[MAXP], dprM[MAXP];
    for (int k = 0; k < nM; k++) {
      double nl;
      const int pk = cfg.omit_priors ? HVI_PR_NONE : cfg.prM_kind[k];
      trans_prior_point<T>(cfg.transM[k], pk, cfg.prM_a[k], cfg.prM_ib[k], cfg.prM_c0[k], mu[site * nM + k], thM[k], dthM[k],
                           nl, dprM[k]);
      a_pr += nl;
    }
    T par[4];
    for (int s = 0; s < 4; s++) {
      const int c = cfg.slot_code[s];
      par[s] = c < 0 ? cfg.slot_fix[s] : (c < nP ? thP[c] : thM[c - nP]);
    }
    T g0 = T(0), g1 = T(0), gA = T(0), gB = T(0), nly = T(0);
    for (int o = 0; o < nO; o++) {   // y = r0 + r1·S1/(K1+S1)·S2/(K2+S2)  (src/DoubleMM/f_doubleMM.jl:137)
      const T S1 = r[(0 * nO + o) * 32], S2 = r[(1 * nO + o) * 32], yo = r[(2 * nO + o) * 32], w = r[(3 * nO + o) * 32];
      const T d1 = par[2] + S1, d2 = par[3] + S2;
      const T ab = S1 / d1 * S2 / d2;
      const T res = par[0] + par[1] * ab - yo, dy = w * res;
      nly += res * dy;
      g0 += dy; g1 += dy * ab; gA += dy * ab / d1; gB += dy * ab / d2;
    }
    const T pbar[4] = {g0, g1, -par[1] * gA, -par[1] * gB};
    a_nly += (double)nly;
    T thbar[2 * MAXP];
    for (int q = 0; q < nP + nM; q++) {
      T v = T(0);
      for (int s = 0; s < 4; s++) v += cfg.slot_code[s] == q ? pbar[s] : T(0);
      thbar[q] = v;
    }
    for (int j = 0; j < nP; j++) a_thP[j] += (double)thbar[j];
    for (int k = 0; k < nM; k++) {
      const T mb = thbar[nP + k] * dthM[k] + dprM[k];
      mubar[site * nM + k] = mb;
      if (!isfinite(mb)) a_bad += 1.0;
    }
  }
  // block partials, fixed order
  a_nly = warp_sum_(a_nly); a_pr = warp_sum_(a_pr); a_bad = warp_sum_(a_bad);
  for (int j = 0; j < nP; j++) a_thP[j] = warp_sum_(a_thP[j]);
  if (lane == 0) {
    sred[warp][0] = a_nly; sred[warp][1] = a_pr; sred[warp][2] = a_bad;
    for (int j = 0; j < nP; j++) sred[warp][3 + j] = a_thP[j];
  }
  __syncthreads();
  if (threadIdx.x < 3 + nP) {
    double s = 0.0;
    for (int w = 0; w < PT_THREADS / 32; w++) s += sred[w][threadIdx.x];
    part[(size_t)blockIdx.x * (3 + MAXP) + threadIdx.x] = s;
  }
}

// one block: out = [∇ϕg (nphig); ∇ϕP (nP); nLjoint_pen, nLy, neg_log_prior, non-finite count]
template <typename T>
__global__ void __launch_bounds__(256) k_point_finalize(const __grid_constant__ Cfg<T> cfg, const T* __restrict__ phi,
                                                         const double* __restrict__ part, int nrows,
                                                         const double* __restrict__ part_mlp, int nblk,
                                                         const double* __restrict__ part_x, const double* __restrict__ batch_consts,
                                                         double* __restrict__ out, T* __restrict__ grad_T) {
  __shared__ double tot[3 + MAXP];
  __shared__ double bad_g;
  const int nP = cfg.nP, ng = cfg.nphig;
  if (threadIdx.x == 0) bad_g = 0.0;
  if (threadIdx.x < 3 + nP) {
    double s = 0.0;
    for (int r = 0; r < nrows; r++) s += part[(size_t)r * (3 + MAXP) + threadIdx.x];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  double bad = 0.0;
  for (int e = threadIdx.x; e < ng; e += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblk; b++) s += part_mlp[(size_t)b * ng + e];
    out[e] = s;
    if (grad_T) grad_T[e] = (T)s;
    if (!isfinite(s)) bad += 1.0;
  }
  if (bad != 0.0) atomicAdd(&bad_g, bad);
  __syncthreads();
  if (threadIdx.x == 0) {
    double nlpP = 0.0, nbad = tot[2] + bad_g;
    for (int j = 0; j < nP; j++) {
      T th, dth, dnl;
      double nl;
      const int pk = cfg.omit_priors ? HVI_PR_NONE : cfg.prP_kind[j];
      trans_prior_point<T>(cfg.transP[j], pk, cfg.prP_a[j], cfg.prP_ib[j], cfg.prP_c0[j], phi[cfg.o_muP + j], th, dth, nl, dnl);
      nlpP += nl;
      double g = tot[3 + j] * (double)dth + (double)dnl;
      // cotangent of the pbm covariate inputs of g (ϕP[idx] appended to xM, src/gf.jl:165)
      for (int c = 0; c < cfg.npc; c++)
        if (cfg.pbm_covar_idx[c] == j && part_x)
          for (int b = 0; b < nblk; b++) g += part_x[(size_t)b * cfg.npc + c];
      out[ng + j] = g;
      if (grad_T) grad_T[ng + j] = (T)g;
      if (!isfinite(g)) nbad += 1.0;
    }
    const double nLy = 0.5 * tot[0] + batch_consts[0], nlp = tot[1] + nlpP;
    out[ng + nP] = nLy + nlp; out[ng + nP + 1] = nLy; out[ng + nP + 2] = nlp;
    if (!isfinite(nLy + nlp)) nbad += 1.0;
    out[ng + nP + 3] = nbad;
  }
}

}  // namespace

int point_max_rows(int sm_count) { return sm_count * 8; }
int point_row_len() { return 3 + MAXP; }

template <typename T>
cudaError_t launch_point(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* rec, const T* mu, T* mubar, int64_t nb,
                         double* part, int* nrows_out) {
  const int64_t ntiles = (nb + 31) / 32;
  int64_t blocks = (ntiles + PT_THREADS / 32 - 1) / (PT_THREADS / 32);
  if (blocks > point_max_rows(L.sm_count)) blocks = point_max_rows(L.sm_count);
  if (blocks < 1) blocks = 1;
  k_point<T><<<(int)blocks, PT_THREADS, 0, L.stream>>>(cfg, phi, rec, mu, mubar, nb, part);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (L.counter) ++*L.counter;
  *nrows_out = (int)blocks;
  return cudaSuccess;
}

template <typename T>
cudaError_t launch_point_finalize(const Launch& L, const Cfg<T>& cfg, const T* phi, const double* part, int nrows,
                                  const double* part_mlp, int nblk, const double* part_x, const double* batch_consts, double* out,
                                  T* grad_T) {
  k_point_finalize<T><<<1, 256, 0, L.stream>>>(cfg, phi, part, nrows, part_mlp, nblk, part_x, batch_consts, out, grad_T);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (L.counter) ++*L.counter;
  return cudaSuccess;
}

template cudaError_t launch_point<float>(const Launch&, const Cfg<float>&, const float*, const float*, const float*, float*,
                                         int64_t, double*, int*);
template cudaError_t launch_point<double>(const Launch&, const Cfg<double>&, const double*, const double*, const double*,
                                          double*, int64_t, double*, int*);
template cudaError_t launch_point_finalize<float>(const Launch&, const Cfg<float>&, const float*, const double*, int,
                                                  const double*, int, const double*, const double*, double*, float*);
template cudaError_t launch_point_finalize<double>(const Launch&, const Cfg<double>&, const double*, const double*, int,
                                                   const double*, int, const double*, const double*, double*, double*);

}  // namespace hvi
