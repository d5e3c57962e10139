// hvi_pbm.cuh -- the generic process-based-model seam of the streaming kernel.
//
// The reference's process model is any callable `f(θP, θMs_tr, xP) -> y` (AbstractPBMApplicator,
// src/PBMApplicator.jl:26-32; per-site form `fθ(θ, xP_site)` of PBMSiteApplicator, :122-195), differentiated by Zygote.
// Here a PBM is a struct with one function template, written once for a generic scalar S:
//
//     struct UserPbm {
//       template <class S, class T>   // S: T or a forward-mode dual number over T;  T: float | double
//       __device__ static void f(const S* thP, const S* thM, const T* fix, const T* x, S* y);
//       // optional: a site-level penalty added to the loss (floss_penalty, src/elbo.jl:153,282-286), declared with
//       //   static constexpr bool HAS_PENALTY = true;
//       //   template <class S, class T> __device__ static S penalty(const S* thP, const S* thM, const T* fix, const T* x, const S* y);
//     };
//
// thP (n_θP) / thM (n_θM): the constrained parameters of one draw in the problem's order, fix: the fixed parameters
// (θFix), x: the site's driver column (n_x values), y: the n_obs predictions.  The library instantiates f with dual
// numbers to obtain ∂y/∂θ (forward mode: n_θP + n_θM directional derivatives per call), so the user writes no adjoint.
// A literal must be written in the working type: T(2) * th, not 2.0 * th.
//
// This header is compiled at run time by NVRTC together with the PBM source (hvi_generic.cu), so it includes nothing
// and uses fundamental types only.  Everything around the PBM call -- ζ = μ + σ(μ)·Uᵀξ, bijectors + clamp + log-Jacobian,
// priors, the likelihood with NaN-observation masking, the adjoint and the per-block partial sums -- mirrors k_stream
// (hvi_kernels.cu) and feeds the same k_reduce / k_finalize.
#pragma once

namespace hvipbm {

typedef unsigned int u32;
typedef unsigned long long u64;
typedef long long i64;

constexpr int MAXP = 8;

// ---- forward-mode dual numbers -----------------------------------------------------------
template <typename T, int N>
struct Dual {
  T v;
  T d[N];
  __device__ Dual() {}
  __device__ Dual(T x) : v(x) {
#pragma unroll
    for (int i = 0; i < N; i++) d[i] = T(0);
  }
};
#define HVI_DUAL_BIN(OP, VAL, DER_AB, DER_A, DER_B)                                                   \
  template <typename T, int N>                                                                        \
  __device__ __forceinline__ Dual<T, N> operator OP(const Dual<T, N>& a, const Dual<T, N>& b) {       \
    Dual<T, N> r;                                                                                     \
    r.v = VAL(a.v, b.v);                                                                              \
    _Pragma("unroll") for (int i = 0; i < N; i++) r.d[i] = DER_AB;                                    \
    return r;                                                                                         \
  }                                                                                                   \
  template <typename T, int N>                                                                        \
  __device__ __forceinline__ Dual<T, N> operator OP(const Dual<T, N>& a, T bv) {                      \
    Dual<T, N> r;                                                                                     \
    r.v = VAL(a.v, bv);                                                                               \
    _Pragma("unroll") for (int i = 0; i < N; i++) r.d[i] = DER_A;                                     \
    return r;                                                                                         \
  }                                                                                                   \
  template <typename T, int N>                                                                        \
  __device__ __forceinline__ Dual<T, N> operator OP(T av, const Dual<T, N>& b) {                      \
    Dual<T, N> r;                                                                                     \
    r.v = VAL(av, b.v);                                                                               \
    _Pragma("unroll") for (int i = 0; i < N; i++) r.d[i] = DER_B;                                     \
    return r;                                                                                         \
  }
#define HVI_ADD(x, y) ((x) + (y))
#define HVI_SUB(x, y) ((x) - (y))
#define HVI_MUL(x, y) ((x) * (y))
#define HVI_DIV(x, y) ((x) / (y))
HVI_DUAL_BIN(+, HVI_ADD, a.d[i] + b.d[i], a.d[i], b.d[i])
HVI_DUAL_BIN(-, HVI_SUB, a.d[i] - b.d[i], a.d[i], -b.d[i])
HVI_DUAL_BIN(*, HVI_MUL, a.d[i] * b.v + a.v * b.d[i], a.d[i] * bv, av * b.d[i])
// quotient q = a/b: dq = (da − q·db)/b
HVI_DUAL_BIN(/, HVI_DIV, (a.d[i] - r.v * b.d[i]) / b.v, a.d[i] / bv, -r.v * b.d[i] / b.v)
#undef HVI_DUAL_BIN
template <typename T, int N>
__device__ __forceinline__ Dual<T, N> operator-(const Dual<T, N>& a) {
  Dual<T, N> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = -a.d[i];
  return r;
}
template <typename T, int N>
__device__ __forceinline__ Dual<T, N> chain(const Dual<T, N>& a, T val, T der) {
  Dual<T, N> r;
  r.v = val;
#pragma unroll
  for (int i = 0; i < N; i++) r.d[i] = der * a.d[i];
  return r;
}
__device__ __forceinline__ float exp_(float x) { return expf(x); }
__device__ __forceinline__ double exp_(double x) { return exp(x); }
__device__ __forceinline__ float log_(float x) { return logf(x); }
__device__ __forceinline__ double log_(double x) { return log(x); }
__device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ float tanh_(float x) { return tanhf(x); }
__device__ __forceinline__ double tanh_(double x) { return tanh(x); }
__device__ __forceinline__ float log1p_(float x) { return log1pf(x); }
__device__ __forceinline__ double log1p_(double x) { return log1p(x); }
// the same spellings for plain scalars and duals, so that a PBM written for S compiles for both
__device__ __forceinline__ float hexp(float x) { return expf(x); }
__device__ __forceinline__ double hexp(double x) { return exp(x); }
__device__ __forceinline__ float hlog(float x) { return logf(x); }
__device__ __forceinline__ double hlog(double x) { return log(x); }
__device__ __forceinline__ float hsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double hsqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float htanh(float x) { return tanhf(x); }
__device__ __forceinline__ double htanh(double x) { return tanh(x); }
template <typename T, int N> __device__ __forceinline__ Dual<T, N> hexp(const Dual<T, N>& a) { const T e = exp_(a.v); return chain(a, e, e); }
template <typename T, int N> __device__ __forceinline__ Dual<T, N> hlog(const Dual<T, N>& a) { return chain(a, log_(a.v), T(1) / a.v); }
template <typename T, int N> __device__ __forceinline__ Dual<T, N> hsqrt(const Dual<T, N>& a) { const T s = sqrt_(a.v); return chain(a, s, T(0.5) / s); }
template <typename T, int N> __device__ __forceinline__ Dual<T, N> htanh(const Dual<T, N>& a) { const T t = tanh_(a.v); return chain(a, t, T(1) - t * t); }
template <typename T> __device__ __forceinline__ T value_of(T x) { return x; }
template <typename T, int N> __device__ __forceinline__ T value_of(const Dual<T, N>& a) { return a.v; }

// ---- built-in process models ---------------------------------------------------------------
// f_doubleMM (src/DoubleMM/f_doubleMM.jl:47-65,101-139) behind the generic seam: x = [S1[0..nO); S2[0..nO)];
// (r0, r1, K1, K2) gathered from θP / θM / θFix by the slot map of the description (src/PBMApplicator.jl:234-239);
// slot code: −1 fixed (value in fix[s]), j < n_θP: θP[j], else θM[code − n_θP].  Used to check the seam against the
// hand-written DoubleMM kernel.
template <int NP, int NO>
struct PbmDoubleMM {
  template <class S, class T>
  __device__ static void f(const S* thP, const S* thM, const T* fix, const int* slot, const T* x, S* y) {
    S par[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int c = slot[s];
      par[s] = c < 0 ? S(fix[s]) : (c < NP ? thP[c < NP ? c : 0] : thM[c - NP]);
    }
#pragma unroll
    for (int o = 0; o < NO; o++) {
      const T S1 = x[o], S2 = x[NO + o];
      y[o] = par[0] + par[1] * S1 / (par[2] + S1) * S2 / (par[3] + S2);   // left to right as :137
    }
  }
};

// ---- arguments (fundamental types only: the same bytes under nvcc and NVRTC) ------------------
template <typename T>
struct EvalConstsG {   // = hvi::EvalConsts<T> (hvi_internal.cuh), checked by a static_assert in hvi_generic.cu
  T UM[MAXP][MAXP], UP[MAXP][MAXP];
  T nrmM[MAXP], nrmP[MAXP];
  T coefA[MAXP], coefB[MAXP];
  T sigP[MAXP], muP[MAXP];
};

template <typename T>
struct GenArgs {
  const T* xP;        // drivers (n_x × nb), column = site
  const T* y_o;       // observations (n_obs × nb), NaN = missing
  const T* y_unc;     // log σ² of the observations (n_obs × nb)
  const T* mu;        // μ_ζ (n_θM × nb)
  const T* zMs;       // draws (M × n_θM·nb) column-major, or NULL with rng
  const T* pd;        // per-draw records of the global block [M][pd_stride]: θP, dθP/dζP, zP
  const EvalConstsG<T>* ec;
  T* mubar;
  const T* MU;        // pbm covariates: per-draw means [M][n_θM][nb]
  T* MUBAR;
  double* partials;   // [gridDim.x][nacc]
  double* pen_partials;   // [gridDim.x] Σ penalty (NULL: none)
  i64 nb;
  int M, pd_stride;
  const T* phi;
  const i64* i_sites;
  double* out;
  T* grad_T;
  int rng;
  u64 seed;
  const u64* seed_dev;
  i64 col_offset;
  // problem constants
  int transM[MAXP], prM_kind[MAXP];
  T prM_a[MAXP], prM_ib[MAXP];
  T clamp_lo, clamp_hi;
  int sepvar, o_coef;
  int slot[4];
  T fix[MAXP];
};

enum { TR_IDENTITY = 0, TR_EXP = 1, TR_LOGISTIC = 2, PR_NONE = 0, PR_NORMAL = 1, PR_LOGNORMAL = 2 };
enum { ACC_NLY = 0, ACC_PRIOR = 1, ACC_LJ = 2, ACC_LOGSIG = 3, ACC_FIRST_VEC = 4 };
__host__ __device__ constexpr int tri(int n) { return n * (n + 1) / 2; }
__host__ __device__ constexpr int nacc(int nP, int nM) { return 4 + 2 * nM + tri(nM) + nP + tri(nP); }
__host__ __device__ constexpr int ut(int j, int k) { return k * (k + 1) / 2 + j; }

// θ = T(ζ) with the reference's clamp (src/elbo.jl:783-808) fused with the prior on θ (src/elbo.jl:203-260): the same
// mathematics as transform_prior of hvi_kernels.cu
template <typename T>
__device__ __forceinline__ void transform_prior(int tr, int pk, T a, T ib, T lo, T hi, T zeta, T& th, T& dth, T& pv, T& dz,
                                                T& lj) {
  T raw, d, dlj, dpz;
  if (tr == TR_EXP) {
    raw = exp_(zeta); d = raw; lj = zeta; dlj = T(1);
  } else if (tr == TR_LOGISTIC) {
    raw = T(1) / (T(1) + exp_(-zeta)); d = raw * (T(1) - raw);
    lj = -(log1p_(exp_(-zeta)) + log1p_(exp_(zeta))); dlj = T(1) - T(2) * raw;
  } else {
    raw = zeta; d = T(1); lj = T(0); dlj = T(0);
  }
  th = raw < hi ? raw : hi;
  th = th > lo ? th : lo;
  const bool uncl = (th == raw);
  dth = uncl ? d : T(0);
  if (pk == PR_LOGNORMAL) {
    const T lt = (tr == TR_EXP && uncl) ? zeta : log_(th);
    const T u = (lt - a) * ib;
    pv = lt + T(0.5) * u * u;
    if (tr == TR_EXP) { dz = uncl ? u * ib : T(-1); return; }
    dpz = (T(1) + u * ib) / th * dth;
  } else if (pk == PR_NORMAL) {
    const T u = (th - a) * ib;
    pv = T(0.5) * u * u;
    dpz = u * ib * dth;
  } else {
    pv = T(0); dpz = T(0);
  }
  dz = dpz - dlj;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <class P, class = void> struct HasPenalty { static constexpr bool value = false; };
template <class P> struct HasPenalty<P, decltype((void)P::HAS_PENALTY)> { static constexpr bool value = P::HAS_PENALTY; };

// PBM call adaptor: built-in models take the slot map, user models do not
template <class PBM, bool WITH_SLOT> struct CallPbm;
template <class PBM> struct CallPbm<PBM, true> {
  template <class S, class T> __device__ static void f(const S* thP, const S* thM, const GenArgs<T>& a, const T* x, S* y) {
    PBM::f(thP, thM, a.fix, a.slot, x, y);
  }
};
template <class PBM> struct CallPbm<PBM, false> {
  template <class S, class T> __device__ static void f(const S* thP, const S* thM, const GenArgs<T>& a, const T* x, S* y) {
    PBM::f(thP, thM, a.fix, x, y);
  }
};

constexpr int GEN_THREADS = 128;

// four standard normals of column gcol, draws 4·m4 … 4·m4+3: the library's generator (hvi_rng.cuh) restated for this
// self-contained header -- Philox4x32-10 keyed by the seed, Box-Muller
__device__ __forceinline__ void normal4(u64 seed, u64 gcol, int m4, float (&n4)[4]) {
  u32 c0 = (u32)gcol, c1 = (u32)(gcol >> 32), c2 = (u32)m4, c3 = 0x5eedu, k0 = (u32)seed, k1 = (u32)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const u32 hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const u32 hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const u32 n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const u32 r[4] = {c0, c1, c2, c3};
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const float u1 = fmaf((float)(r[2 * h] >> 8), 5.9604644775390625e-08f, 5.9604644775390625e-08f);
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-1.3862943611198906f * __log2f(u1)));
    float sn, cs;
    __sincosf(fmaf((float)(r[2 * h + 1] >> 8), 6.283185307179586f * 5.9604644775390625e-08f, -3.141592653589793f), &sn, &cs);
    n4[2 * h] = rad * cs; n4[2 * h + 1] = rad * sn;
  }
}

// One thread per site, all draws; per-block row of double partial sums in the layout k_reduce / k_finalize read.
template <class PBM, bool WITH_SLOT, typename T, int NP, int NM, int NO, int NX, bool HASMU>
__device__ void generic_stream(const GenArgs<T>& a) {
  constexpr int NQ = NP + NM;                 // differentiated parameters
  constexpr int NUM = tri(NM), NUP = NP > 0 ? tri(NP) : 1, NACC = nacc(NP, NM), NWARP = GEN_THREADS / 32;
  typedef Dual<T, NQ> D;
  __shared__ double wacc[NWARP][NACC + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int i = 0; i <= NACC; i++) wacc[warp][i] = 0.0;
  __syncwarp();
  const int M = a.M;
  const T wM = T(1) / T(M);
  const u64 seed = (a.rng && a.seed_dev) ? *a.seed_dev : a.seed;
  T cA[NM], cB[NM], U[NUM];
#pragma unroll
  for (int k = 0; k < NM; k++) {
    cA[k] = a.ec->coefA[k]; cB[k] = a.ec->coefB[k];
#pragma unroll
    for (int j = 0; j <= k; j++) U[ut(j, k)] = a.ec->UM[j][k];
  }
  const i64 ntiles = (a.nb + 31) / 32;
  for (i64 tile = (i64)blockIdx.x * NWARP + warp; tile < ntiles; tile += (i64)gridDim.x * NWARP) {
    const i64 site = tile * 32 + lane;
    const bool active = site < a.nb;
    const i64 isite = (a.sepvar && active && a.i_sites) ? a.i_sites[site] : site;
    T acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = T(0);
    T pen = T(0);
    if (active) {
      T x[NX], yo[NO], w[NO];
#pragma unroll
      for (int i = 0; i < NX; i++) x[i] = a.xP[site * NX + i];
#pragma unroll
      for (int o = 0; o < NO; o++) {
        yo[o] = a.y_o[site * NO + o];
        w[o] = exp_(-a.y_unc[site * NO + o]);
      }
      T mu[NM], ls[NM], sig[NM], sU[NUM], smu[NM], sz[NUM], aP[NP > 0 ? NP : 1], cP[NUP];
#pragma unroll
      for (int k = 0; k < NM; k++) {
        mu[k] = a.mu[site * NM + k];
        ls[k] = a.sepvar ? a.phi[a.o_coef + isite * NM + k] : cB[k] * mu[k] + cA[k];
        sig[k] = exp_(T(0.5) * ls[k]);
        smu[k] = T(0);
#pragma unroll
        for (int j = 0; j <= k; j++) { sU[ut(j, k)] = sig[k] * U[ut(j, k)]; sz[ut(j, k)] = T(0); }
      }
#pragma unroll
      for (int i = 0; i < (NP > 0 ? NP : 1); i++) aP[i] = T(0);
#pragma unroll
      for (int i = 0; i < NUP; i++) cP[i] = T(0);
      T nly = T(0), prior = T(0), ljs = T(0);
      for (int m = 0; m < M; m++) {
        T z[NM];
        if (a.rng) {
          float n4[4];
#pragma unroll
          for (int k = 0; k < NM; k++) {
            normal4(seed, (u64)(a.col_offset + site * NM + k), m / 4, n4);
            z[k] = (T)(m % 4 == 0 ? n4[0] : m % 4 == 1 ? n4[1] : m % 4 == 2 ? n4[2] : n4[3]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < NM; k++) z[k] = a.zMs[((size_t)site * NM + k) * M + m];
        }
        const T* pd = a.pd + (size_t)m * a.pd_stride;
        D thP[NP > 0 ? NP : 1], thM[NM], y[NO];
        T dth[NM], dz[NM];
#pragma unroll
        for (int j = 0; j < NP; j++) { thP[j] = D(pd[j]); thP[j].d[j] = T(1); }
#pragma unroll
        for (int k = 0; k < NM; k++) {
          T zeta = HASMU ? a.MU[((size_t)m * NM + k) * a.nb + site] : mu[k];
#pragma unroll
          for (int j = 0; j <= k; j++) zeta += sU[ut(j, k)] * z[j];
          T th, pv, lj;
          transform_prior<T>(a.transM[k], a.prM_kind[k], a.prM_a[k], a.prM_ib[k], a.clamp_lo, a.clamp_hi, zeta, th, dth[k], pv,
                             dz[k], lj);
          prior += pv; ljs += lj;
          thM[k] = D(th); thM[k].d[NP + k] = T(1);
        }
        CallPbm<PBM, WITH_SLOT>::f(thP, thM, a, x, y);
        T thbar[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) thbar[q] = T(0);
#pragma unroll
        for (int o = 0; o < NO; o++) {
          if (yo[o] == yo[o]) {   // observation present (NaN = missing, src/logden_normal.jl:34); a NaN prediction here
            const T res = y[o].v - yo[o];   // makes the result non-finite, as in the reference
            const T dy = w[o] * res;
            nly += res * dy;
#pragma unroll
            for (int q = 0; q < NQ; q++) thbar[q] += dy * y[o].d[q];
          }
        }
        if constexpr (HasPenalty<PBM>::value) {
          const D pn = PBM::penalty(thP, thM, a.fix, x, y);
          pen += pn.v;
          // the loss adds mean_m(penalty): same 1/n_MC scaling as the likelihood cotangent dy (unscaled here)
#pragma unroll
          for (int q = 0; q < NQ; q++) thbar[q] += pn.d[q];
        }
#pragma unroll
        for (int k = 0; k < NM; k++) {
          const T zb = thbar[NP + k] * dth[k] + dz[k];
          if (HASMU) a.MUBAR[((size_t)m * NM + k) * a.nb + site] = wM * zb;
          smu[k] += zb;
#pragma unroll
          for (int j = 0; j <= k; j++) sz[ut(j, k)] += zb * z[j];
        }
#pragma unroll
        for (int j = 0; j < NP; j++) {
          const T zb = thbar[j] * pd[NP + j];
          aP[j] += zb;
#pragma unroll
          for (int jj = 0; jj <= j; jj++) cP[ut(jj, j)] += zb * pd[2 * NP + jj];
        }
      }
      // end of site: σ-path, entropy, μ̄_ζ (SURVEY Appendix A adjoint), as in k_stream
      acc[ACC_NLY] = nly; acc[ACC_LJ] = ljs; acc[ACC_PRIOR] = prior;
#pragma unroll
      for (int k = 0; k < NM; k++) {
        T sr = T(0);
#pragma unroll
        for (int j = 0; j <= k; j++) sr += sU[ut(j, k)] * sz[ut(j, k)];
        const T lsb = T(0.5) * wM * sr - T(0.5);
        acc[ACC_LOGSIG] += T(0.5) * ls[k];
        if (a.sepvar) {
          const i64 o = (i64)a.o_coef + isite * NM + k;
          a.out[o] = (double)lsb;
          if (a.grad_T) a.grad_T[o] = lsb;
          acc[ACC_FIRST_VEC + k] = (lsb == lsb && lsb - lsb == T(0)) ? T(0) : T(1);
          a.mubar[site * NM + k] = HASMU ? T(0) : wM * smu[k];
        } else {
          acc[ACC_FIRST_VEC + k] = lsb;
          acc[ACC_FIRST_VEC + NM + k] = lsb * mu[k];
          a.mubar[site * NM + k] = HASMU ? lsb * cB[k] : wM * smu[k] + lsb * cB[k];
        }
#pragma unroll
        for (int j = 0; j <= k; j++) acc[ACC_FIRST_VEC + 2 * NM + ut(j, k)] = sig[k] * sz[ut(j, k)];
      }
#pragma unroll
      for (int j = 0; j < NP; j++) acc[ACC_FIRST_VEC + 2 * NM + NUM + j] = aP[j];
      if (NP > 0) {
#pragma unroll
        for (int i = 0; i < tri(NP); i++) acc[ACC_FIRST_VEC + 2 * NM + NUM + NP + i] = cP[i];
      }
    }
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      const T v = warp_sum(acc[i]);
      if (lane == 0) wacc[warp][i] += (double)v;
    }
    {
      const T v = warp_sum(pen);
      if (lane == 0) wacc[warp][NACC] += (double)v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= NACC; i += GEN_THREADS) {
    double s = 0.0;
    for (int w_ = 0; w_ < NWARP; w_++) s += wacc[w_][i];
    if (i < NACC) a.partials[(size_t)blockIdx.x * NACC + i] = s;
    else if (a.pen_partials) a.pen_partials[blockIdx.x] = s;
  }
}

}  // namespace hvipbm
