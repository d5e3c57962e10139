// hvi_stream.cuh -- the fused streaming body (one thread per site, draws through warp-private shared memory);
// included by hvi_stream.cu (k_stream) and hvi_kernels.cu (phase of the single-launch small-batch iteration)
#pragma once
#include "hvi_dev.cuh"
#include "hvi_rng.cuh"

namespace hvi {

// ---------------------------------------------------------------------------------------
// The fused streaming kernel.  One thread owns one site for all draws; a warp owns a tile of
// 32 consecutive sites.  The tile's draws ξ (32·NM rows of M values, contiguous in HBM) are
// staged through warp-private shared memory with coalesced 16-byte cp.async and read back
// transposed; 16-byte chunks are rotated by the owning lane so that the transposed LDS.128 is
// bank-conflict free.  Everything between μ_ζ and (μ̄_ζ, partial sums of the shared-parameter
// cotangents) stays in registers.
//
// Traits: the PBM slot map, the bijector codes and the prior kinds are either read from the
// constant bank at run time (DynTraits: any configuration) or baked in at compile time
// (StaTraits: the named configurations below), which removes every select from the draw loop.
// ---------------------------------------------------------------------------------------
__host__ __device__ constexpr int ut(int j, int k) { return k * (k + 1) / 2 + j; }  // packed upper incl. diag

struct DynTraits {
  static constexpr bool STATIC = false;
};
// SLOTS: 4 nibbles (r0, r1, K1, K2), nibble = slot code + 1 (0 = fixed value);
// TRM: 2 bits per θM column (bijector code); PRM: 2 bits per θM column (prior kind, 0 if omitted)
template <uint32_t SLOTS, uint32_t TRM, uint32_t PRM>
struct StaTraits {
  static constexpr bool STATIC = true;
  static constexpr uint32_t slots = SLOTS, trm = TRM, prm = PRM;
};

template <class TR, typename T>
__device__ __forceinline__ int tr_slot(const Cfg<T>& cfg, int s) {
  if constexpr (TR::STATIC) return (int)((TR::slots >> (4 * s)) & 15u) - 1;
  else return cfg.slot_code[s];
}
template <class TR, typename T>
__device__ __forceinline__ int tr_transM(const Cfg<T>& cfg, int k) {
  if constexpr (TR::STATIC) return (int)((TR::trm >> (2 * k)) & 3u);
  else return cfg.transM[k];
}
template <class TR, typename T>
__device__ __forceinline__ int tr_priorM(const Cfg<T>& cfg, int k) {
  if constexpr (TR::STATIC) return (int)((TR::prm >> (2 * k)) & 3u);
  else return cfg.omit_priors ? HVI_PR_NONE : cfg.prM_kind[k];
}

// all θM columns are Exp-transformed with LogNormal priors (known at compile time)
template <class TR, int NM>
__host__ __device__ constexpr bool all_exp_ln() {
  if constexpr (TR::STATIC) {
    for (int k = 0; k < NM; k++)
      if (((TR::trm >> (2 * k)) & 3u) != HVI_TR_EXP || ((TR::prm >> (2 * k)) & 3u) != HVI_PR_LOGNORMAL) return false;
    return true;
  } else {
    return false;
  }
}

// per-draw block of the global parameters, one record per draw: θP[NP], dθP/dζP[NP], zP[NP]
template <typename T, int NP, int NM, int NO>
struct SiteState {
  static constexpr int NOP = (NO + 1) / 2;   // observation pairs
  typename PairT<T>::type S1[NOP], S2[NOP], S12[NOP], nyo[NOP], w[NOP];   // nyo = −y_o
  // Σ over the draws of res² per observation (the weight is applied once per site at the end: res·res reads two
  // register pairs where res·(w·res) + acc reads three, and operand reads bound this kernel)
  typename PairT<T>::type q[NOP];
  T mu[NM];
  T sU[tri(NM)];                 // σ[k]·U[j,k]
  T smu[NM], sz[tri(NM)];        // Σ_m ζ̄[k],  Σ_m ζ̄[k]·z[j]
  T aP[NP > 0 ? NP : 1], cP[NP > 0 ? tri(NP) : 1];
  T prior, lj;                   // Σ of the generic prior terms / of all log-Jacobian terms
  T u2, ljln;                    // Exp+LogNormal columns: Σ u² and Σ ζ  (prior = ζ + u²/2 when unclamped)
};

// f_doubleMM + logden_normal + adjoint w.r.t. (r0, r1, K1, K2), all observations of one site
// and one draw:  y = r0 + r1·S1/(K1+S1)·S2/(K2+S2)   (src/DoubleMM/f_doubleMM.jl:137)
// accumulates res² per observation into st.q; pbar = (Σ dy, Σ dy·ab, −r1 Σ dy·ab/d1, −r1 Σ dy·ab/d2),
// dy = w·res.  14 FP32 operations + 1 reciprocal per observation, two observations per
// instruction.
template <typename T, int NP, int NM, int NO>
__device__ __forceinline__ void obs_loop(SiteState<T, NP, NM, NO>& st, const T (&par)[4], T (&pbar)[4]) {
  using P = typename PairT<T>::type;
  const P r0 = mk2(par[0], par[0]), r1 = mk2(par[1], par[1]);
  const P k1 = mk2(par[2], par[2]), k2 = mk2(par[3], par[3]);
  P g0 = mk2(T(0), T(0)), g1 = g0, gA = g0, gB = g0;
#pragma unroll
  for (int o = 0; o < SiteState<T, NP, NM, NO>::NOP; o++) {
    const P d1 = padd(k1, st.S1[o]), d2 = padd(k2, st.S2[o]);
    const P p = pmul(d1, d2);
    const P rp = mk2(rcp_fast(p.x), rcp_fast(p.y));
    const P ab = pmul(st.S12[o], rp);
    const P y = pfma(r1, ab, r0);
    const P res = padd(y, st.nyo[o]);
    const P dy = pmul(res, st.w[o]);
    st.q[o] = pfma(res, res, st.q[o]);
    g0 = padd(g0, dy);
    const P t = pmul(dy, ab);
    g1 = padd(g1, t);
    const P u = pmul(t, rp);
    gA = pfma(u, d2, gA);
    gB = pfma(u, d1, gB);
  }
  pbar[0] = g0.x + g0.y;
  pbar[1] = g1.x + g1.y;
  pbar[2] = -par[1] * (gA.x + gA.y);
  pbar[3] = -par[1] * (gB.x + gB.y);
}

// the same for two draws at once: the statements of the two draws alternate in program order, which gives the
// scheduler eight independent observation chains instead of four (two warps per scheduler do not cover the
// dependent FMUL2 → MUFU.RCP → FMUL2 chain of a single draw)
template <typename T, int NP, int NM, int NO>
__device__ __forceinline__ void obs_loop2(SiteState<T, NP, NM, NO>& st, const T (&parA)[4], const T (&parB)[4],
                                          T (&pbarA)[4], T (&pbarB)[4]) {
  using P = typename PairT<T>::type;
  const P r0A = mk2(parA[0], parA[0]), r1A = mk2(parA[1], parA[1]), k1A = mk2(parA[2], parA[2]), k2A = mk2(parA[3], parA[3]);
  const P r0B = mk2(parB[0], parB[0]), r1B = mk2(parB[1], parB[1]), k1B = mk2(parB[2], parB[2]), k2B = mk2(parB[3], parB[3]);
  P g0A = mk2(T(0), T(0)), g1A = g0A, gAA = g0A, gBA = g0A, g0B = g0A, g1B = g0A, gAB = g0A, gBB = g0A;
#pragma unroll
  for (int o = 0; o < SiteState<T, NP, NM, NO>::NOP; o++) {
    const P d1A = padd(k1A, st.S1[o]), d1B = padd(k1B, st.S1[o]);
    const P d2A = padd(k2A, st.S2[o]), d2B = padd(k2B, st.S2[o]);
    const P pA = pmul(d1A, d2A), pB = pmul(d1B, d2B);
    const P rpA = mk2(rcp_fast(pA.x), rcp_fast(pA.y)), rpB = mk2(rcp_fast(pB.x), rcp_fast(pB.y));
    const P abA = pmul(st.S12[o], rpA), abB = pmul(st.S12[o], rpB);
    const P yA = pfma(r1A, abA, r0A), yB = pfma(r1B, abB, r0B);
    const P resA = padd(yA, st.nyo[o]), resB = padd(yB, st.nyo[o]);
    const P dyA = pmul(resA, st.w[o]), dyB = pmul(resB, st.w[o]);
    st.q[o] = pfma(resB, resB, pfma(resA, resA, st.q[o]));
    g0A = padd(g0A, dyA); g0B = padd(g0B, dyB);
    const P tA = pmul(dyA, abA), tB = pmul(dyB, abB);
    g1A = padd(g1A, tA); g1B = padd(g1B, tB);
    const P uA = pmul(tA, rpA), uB = pmul(tB, rpB);
    gAA = pfma(uA, d2A, gAA); gAB = pfma(uB, d2B, gAB);
    gBA = pfma(uA, d1A, gBA); gBB = pfma(uB, d1B, gBB);
  }
  pbarA[0] = g0A.x + g0A.y; pbarB[0] = g0B.x + g0B.y;
  pbarA[1] = g1A.x + g1A.y; pbarB[1] = g1B.x + g1B.y;
  pbarA[2] = -parA[1] * (gAA.x + gAA.y); pbarB[2] = -parB[1] * (gAB.x + gAB.y);
  pbarA[3] = -parA[1] * (gBA.x + gBA.y); pbarB[3] = -parB[1] * (gBB.x + gBB.y);
}

// what one draw carries from the part before the observation loop to the part after it
template <typename T, int NP, int NM>
struct DrawTmp {
  T dth[NM], dz[NM];                       // dθ/dζ, d(prior − logjac)/dζ of the site parameters
  T gp[NP > 0 ? NP : 1], zp[NP > 0 ? NP : 1];   // dθP/dζP and zP of this draw
  T par[4];                                // (r0, r1, K1, K2)
};

// before the observations: ζ = μ + σ(μ)·Uᵀz, θ = T(ζ) with clamp, priors, log-Jacobian, slot gather
// mud: the mean μ_ζ of this draw (st.mu without pbm covariates; g([xM; ζP_m]) with them, src/elbo.jl:508-515)
template <class TR, typename T, int NP, int NM, int NO>
__device__ __forceinline__ void draw_pre(const Cfg<T>& cfg, SiteState<T, NP, NM, NO>& st, const T (&z)[NM],
                                         const T* __restrict__ pd, const T (&mud)[NM], DrawTmp<T, NP, NM>& d) {
  T th[NM];
#pragma unroll
  for (int k = 0; k < NM; k++) {
    T zeta = mud[k];
#pragma unroll
    for (int j = 0; j <= k; j++) zeta = fma(st.sU[ut(j, k)], z[j], zeta);
    const int trk = tr_transM<TR>(cfg, k), pkk = tr_priorM<TR>(cfg, k);
    if (trk == HVI_TR_EXP && pkk == HVI_PR_LOGNORMAL) {
      // θ = exp(ζ) with a LogNormal prior: −logpdf = ζ + ½u² + const, u = (ζ − μ)/σ; minus the
      // log-Jacobian ζ the ζ's cancel in the gradient: d/dζ = u/σ
      // the clamp θ ∈ [√floatmin, √floatmax] (src/elbo.jl:797-798) is applied to ζ before the
      // exponential (same θ, no branch, no logarithm): ζc = clamp(ζ, ln lo, ln hi)
      const T zc = fmax(cfg.ln_clamp_lo, fmin(cfg.ln_clamp_hi, zeta));
      th[k] = exp_t(zc);
      const bool uncl = (zc == zeta);
      if (!uncl) st.prior += zc - zeta;   // log θ = ζc when clamped; predicated, almost never taken
      const T u = fma(zc, cfg.prM_ib[k], cfg.prM_nab[k]);
      st.u2 = fma(u, u, st.u2);
      st.lj += zeta;
      if constexpr (!all_exp_ln<TR, NM>()) st.ljln += zeta;
      d.dth[k] = uncl ? th[k] : T(0);
      d.dz[k] = uncl ? u * cfg.prM_ib[k] : T(-1);
    } else {
      T pv, lj;
      transform_prior<T>(trk, pkk, cfg.prM_a[k], cfg.prM_ib[k], cfg.clamp_lo, cfg.clamp_hi, zeta, th[k], d.dth[k], pv,
                         d.dz[k], lj);
      st.prior += pv;
      st.lj += lj;
    }
  }
  T thp[NP > 0 ? NP : 1];
  if constexpr (NP == 2 && sizeof(T) == 4) {
    const float4 v = *reinterpret_cast<const float4*>(pd);
    const float2 v2 = *reinterpret_cast<const float2*>(pd + 4);
    thp[0] = v.x; thp[1] = v.y; d.gp[0] = v.z; d.gp[1] = v.w; d.zp[0] = v2.x; d.zp[1] = v2.y;
  } else {
#pragma unroll
    for (int j = 0; j < NP; j++) { thp[j] = pd[j]; d.gp[j] = pd[NP + j]; d.zp[j] = pd[2 * NP + j]; }
  }
  // gather (r0, r1, K1, K2) by slot (src/PBMApplicator.jl:234-239,283-287)
#pragma unroll
  for (int s = 0; s < 4; s++) {
    T v = cfg.slot_fix[s];
    const int c = tr_slot<TR>(cfg, s);
#pragma unroll
    for (int j = 0; j < NP; j++) v = (c == j) ? thp[j] : v;
#pragma unroll
    for (int k = 0; k < NM; k++) v = (c == NP + k) ? th[k] : v;
    d.par[s] = v;
  }
}

// after the observations: cotangents of ζM and ζP of this draw into the per-site sums
// zbout: cotangent ζ̄M of this draw (unscaled by 1/n_MC)
template <class TR, typename T, int NP, int NM, int NO>
__device__ __forceinline__ void draw_post(const Cfg<T>& cfg, SiteState<T, NP, NM, NO>& st, const T (&z)[NM],
                                          const DrawTmp<T, NP, NM>& d, const T (&pbar)[4], T (&zbout)[NM]) {
  T thbar[NP + NM];
#pragma unroll
  for (int q = 0; q < NP + NM; q++) {
    T v = T(0);
    bool first = true;
#pragma unroll
    for (int s = 0; s < 4; s++) {
      if constexpr (TR::STATIC) {
        if (tr_slot<TR>(cfg, s) == q) { v = first ? pbar[s] : v + pbar[s]; first = false; }
      } else {
        v += (tr_slot<TR>(cfg, s) == q) ? pbar[s] : T(0);
      }
    }
    thbar[q] = v;
  }
#pragma unroll
  for (int k = 0; k < NM; k++) {
    const T zb = fma(thbar[NP + k], d.dth[k], d.dz[k]);
    zbout[k] = zb;
    st.smu[k] += zb;
#pragma unroll
    for (int j = 0; j <= k; j++) st.sz[ut(j, k)] = fma(zb, z[j], st.sz[ut(j, k)]);
  }
#pragma unroll
  for (int j = 0; j < NP; j++) {
    const T zb = thbar[j] * d.gp[j];
    st.aP[j] += zb;
#pragma unroll
    for (int jj = 0; jj <= j; jj++) st.cP[ut(jj, j)] = fma(zb, d.zp[jj], st.cP[ut(jj, j)]);
  }
}

template <class TR, typename T, int NP, int NM, int NO>
__device__ __forceinline__ void do_draw(const Cfg<T>& cfg, SiteState<T, NP, NM, NO>& st, const T (&z)[NM],
                                        const T* __restrict__ pd, const T (&mud)[NM], T (&zbout)[NM]) {
  DrawTmp<T, NP, NM> d;
  T pbar[4];
  draw_pre<TR, T, NP, NM, NO>(cfg, st, z, pd, mud, d);
  obs_loop<T, NP, NM, NO>(st, d.par, pbar);
  draw_post<TR, T, NP, NM, NO>(cfg, st, z, d, pbar, zbout);
}

constexpr int STREAM_THREADS_MAX = 512;

// Σ over the 32 lanes of N ≤ 32 values per lane at once: at every level half of the values go to the partner lane and
// the other half is kept, so the tree costs NPAD − 1 + (32/NPAD... ) shuffles instead of 5·N.  Returns, in lane l, the total of
// value index warp_multi_index<NPAD>(l) (fixed order of additions: deterministic).
template <int NPAD>
__device__ __forceinline__ int warp_multi_index(int lane) {
  // level with offset 16 keeps the upper half of the values in lanes with bit 4 set, and so on down the bits
  int idx = 0, n = NPAD;
#pragma unroll
  for (int b = 4; b >= 0 && n > 1; b--) { n >>= 1; idx += ((lane >> b) & 1) * n; }
  return idx;
}
template <typename T, int NPAD>
__device__ __forceinline__ T warp_multi_sum(T (&v)[NPAD], int lane) {
  static_assert(NPAD == 16 || NPAD == 32, "16 or 32 values");
  int n = NPAD;
#pragma unroll
  for (int b = 4; b >= 0; b--) {
    const int off = 1 << b;
    if (n > 1) {
      n >>= 1;
      const bool up = (lane >> b) & 1;
#pragma unroll
      for (int i = 0; i < NPAD / 2; i++) {
        if (i < n) {
          const T send = up ? v[i] : v[i + n];
          const T keep = up ? v[i + n] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
  return v[0];
}

// dynamic shared memory of the streaming body: per warp the two half-row stages of ξ, one tile of site records
// and means (prefetched one tile ahead), one row of double accumulators; per block the per-draw records of the
// global parameters (pd_elems = n_MC·pd_stride(n_θP), or 0 when they stay in global memory)
__host__ __device__ constexpr size_t stream_smem_z(size_t szT, int nwarp, int nM) { return (size_t)nwarp * 32 * nM * (8 * (16 / szT)) * szT; }
__host__ __device__ constexpr size_t stream_smem_site(size_t szT, int nwarp, int nM, int nO) { return (size_t)nwarp * (4 * nO + nM) * 32 * szT; }
__host__ __device__ constexpr size_t stream_smem_bytes(size_t szT, int nwarp, int nP, int nM, int nO, size_t pd_elems, bool spf = true) {
  return stream_smem_z(szT, nwarp, nM) + (spf ? stream_smem_site(szT, nwarp, nM, nO) : 0) +
         (size_t)nwarp * nacc(nP, nM) * sizeof(double) + pd_elems * szT + 16;
}

// PDS: the per-draw records are copied to shared memory (compile time, so that they are read with LDS, not with
// generic loads); ILV: two draws share one pass of the observation loop
// SPF: the site records and means of the next tile are prefetched into shared memory (costs (4·n_obs + n_θM)·128
// bytes per warp; switched off by the launcher where that would cost a resident block)
// COMPACT: every draw goes through the rolled one-draw loop (one copy of the per-draw code instead of five): the
// single-launch iteration of small batches is bound by instruction fetch, not by issue
template <class TR, typename T, int NP, int NM, int NO, int STREAM_THREADS, bool PBM, bool RNG, bool PDS = true, bool ILV = true,
          bool SPF = true, bool COMPACT = false>
__device__ __forceinline__ void stream_body(const Cfg<T>& cfg, const StreamArgs<T>& a, unsigned char* smem_raw, int bid,
                                            int nblk) {
  constexpr int CH = 16 / sizeof(T);  // elements per 16-byte chunk
  constexpr int MC = 8 * CH;          // draws per stage (one 128-byte row)
  constexpr int NUM = tri(NM), NUP = tri(NP);
  constexpr int NACC = nacc(NP, NM);
  constexpr int NWARP = STREAM_THREADS / 32;
  constexpr int PD = pd_stride(NP);
  constexpr int REC = 4 * NO * 32;     // elements of one tile of site records
  constexpr size_t OFF_SITE = stream_smem_z(sizeof(T), NWARP, NM);
  constexpr size_t OFF_ACC = OFF_SITE + (SPF ? stream_smem_site(sizeof(T), NWARP, NM, NO) : 0);
  constexpr size_t OFF_PD = OFF_ACC + (size_t)NWARP * NACC * sizeof(double);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* ztile = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (32 * NM * MC);
  T* srec = reinterpret_cast<T*>(smem_raw + OFF_SITE) + (size_t)warp * (REC + NM * 32);
  T* smu = srec + REC;
  double* wacc = reinterpret_cast<double*>(smem_raw + OFF_ACC) + warp * NACC;
  T* spd = reinterpret_cast<T*>(smem_raw + OFF_PD);
  const int M = a.M;
  // per-draw global-parameter records: shared memory if they fit, else read from global
  if constexpr (PDS)
    for (int e = threadIdx.x; e < M * PD; e += STREAM_THREADS) spd[e] = a.pd[e];
  if (lane == 0)
    for (int i = 0; i < NACC; i++) wacc[i] = 0.0;
  __syncthreads();

  SiteState<T, NP, NM, NO> st;
  T cA[NM], cB[NM], U[NUM];
#pragma unroll
  for (int k = 0; k < NM; k++) {
    cA[k] = a.ec->coefA[k];
    cB[k] = a.ec->coefB[k];
#pragma unroll
    for (int j = 0; j <= k; j++) U[ut(j, k)] = a.ec->UM[j][k];
  }
  const T wM = T(1) / T(M);
  const bool staged = a.staged != 0;
  const uint64_t seed = (RNG && a.seed_dev) ? (uint64_t)*a.seed_dev : a.seed;
  const int64_t ntiles = (a.nb + 31) / 32;
  const bool mu_al = (reinterpret_cast<uintptr_t>(a.mu) & 15) == 0;
  constexpr int MCS = MC / 2, QS = 4;
  const int mode = RNG ? 2 : (staged ? 1 : 0);
  const int nst = mode == 1 ? (M + MCS - 1) / MCS : 1;

  // cp.async of one tile of site records and means into this warp's buffer (no commit)
  auto issue_site = [&](int64_t tl) {
    if constexpr (!SPF) return;
    const T* r = a.rec + tl * REC;
#pragma unroll
    for (int c0 = 0; c0 < REC / CH; c0 += 32) cp_async16(srec + (c0 + lane) * CH, r + (size_t)(c0 + lane) * CH);
    if (mu_al && tl * 32 + 32 <= a.nb) {
      const T* m = a.mu + tl * 32 * NM;
#pragma unroll
      for (int c0 = 0; c0 < NM * 32 / CH; c0 += 32)
        if (c0 + lane < NM * 32 / CH) cp_async16(smu + (c0 + lane) * CH, m + (size_t)(c0 + lane) * CH);
    }
  };
  // cp.async of stage sg (MCS draws) of the draws ξ of tile tl into half-row buffer b (no commit); 16-byte chunks
  // are rotated by the owning lane so that the transposed LDS.128 read-back is conflict free
  auto issue_xi = [&](int64_t tl, int sg, int b) {
    const T* gz = a.zMs + (size_t)tl * 32 * NM * M;  // row r of the tile at gz + r*M
    const int rows_valid = (int)(a.nb - tl * 32 < 32 ? a.nb - tl * 32 : 32) * NM;
    const int m0 = sg * MCS;
    const int nch = ((M - m0) < MCS ? (M - m0) : MCS) / CH;
    T* buf = ztile + b * (32 * NM * MCS);
    const int q = lane & 3, r0 = lane >> 2;   // this lane's chunk of the row, and its first row (rows r0, r0 + 8, ...)
    if (rows_valid == 32 * NM && nch == QS) {
      // whole tile, whole stage: no guards, one pointer increment per copy
      const T* src = gz + (size_t)r0 * M + m0 + q * CH;
      const size_t step = (size_t)8 * M;
#pragma unroll
      for (int j = 0; j < 4 * NM; j++) {
        const int row = r0 + 8 * j;
        const int rot = NM == 2 ? ((q + (r0 >> 1)) & 3) : ((q + row / NM) & 3);   // (r0 + 8j)/2 = r0/2 + 4j
        cp_async16(buf + row * MCS + rot * CH, src);
        src += step;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4 * NM; j++) {
        const int row = r0 + 8 * j, owner = row / NM;
        if (q < nch && row < rows_valid)
          cp_async16(buf + row * MCS + ((q + owner) & 3) * CH, gz + (size_t)row * M + m0 + q * CH);
      }
    }
  };

  const int64_t tstride = (int64_t)nblk * NWARP;
  int64_t tile = (int64_t)bid * NWARP + warp;
  // software pipeline over the warp's tiles: the records and means of tile t+1 and the first stage of its draws
  // are in flight while tile t is computed, so that a tile starts without a round trip to HBM
  if (tile < ntiles) {
    issue_site(tile);
    if (mode == 1) issue_xi(tile, 0, 0);
    cp_async_commit();
  }
  int g = 0;            // running stage count of this warp: stage g lives in half-row buffer g & 1
  bool first = true;
  for (; tile < ntiles; tile += tstride) {
    const int64_t tnext = tile + tstride;
    const bool has_next = tnext < ntiles;
    const int64_t site = tile * 32 + lane;
    const bool active = site < a.nb;
    const bool full = tile * 32 + 32 <= a.nb;
    const int64_t isite = (cfg.sepvar && active && a.i_sites) ? a.i_sites[site] : site;   // global site index
    // the group that carried this tile's records is complete, except for the first tile, single-stage tiles and
    // the modes without staged draws, where it is the most recent one
    if (first || nst == 1 || mode != 1) cp_async_wait_group<0>();
    __syncwarp();
    first = false;
    {
      const T* r = SPF ? srec + lane : a.rec + tile * REC + lane;
      constexpr int NOP = SiteState<T, NP, NM, NO>::NOP;
#pragma unroll
      for (int o = 0; o < NOP; o++) {
        const bool two = (2 * o + 1 < NO);   // odd n_obs: the last pair is padded with weight 0
        const int o0 = 2 * o, o1 = two ? 2 * o + 1 : 2 * o;
        st.S1[o] = mk2(r[(0 * NO + o0) * 32], r[(0 * NO + o1) * 32]);
        st.S2[o] = mk2(r[(1 * NO + o0) * 32], r[(1 * NO + o1) * 32]);
        st.nyo[o] = mk2(-r[(2 * NO + o0) * 32], -r[(2 * NO + o1) * 32]);
        st.w[o] = mk2(r[(3 * NO + o0) * 32], two ? r[(3 * NO + o1) * 32] : T(0));
        st.S12[o] = pmul(st.S1[o], st.S2[o]);
      }
    }
    T ls[NM], sig[NM];
#pragma unroll
    for (int k = 0; k < NM; k++) {
      st.mu[k] = (SPF && mu_al && full) ? smu[lane * NM + k] : (active ? a.mu[site * NM + k] : T(0));
      // log σ²: intercept + slope·μ_ζ (src/elbo.jl:652-653) or the site's own parameter (src/elbo_sepvec.jl:23)
      ls[k] = (cfg.sepvar && active) ? a.phi[cfg.o_coef + isite * NM + k] : fma(cB[k], st.mu[k], cA[k]);
      sig[k] = exp_t(T(0.5) * ls[k]);
      st.smu[k] = T(0);
#pragma unroll
      for (int j = 0; j <= k; j++) st.sU[ut(j, k)] = sig[k] * U[ut(j, k)];
    }
#pragma unroll
    for (int i = 0; i < NUM; i++) st.sz[i] = T(0);
#pragma unroll
    for (int i = 0; i < (NP > 0 ? NP : 1); i++) st.aP[i] = T(0);
#pragma unroll
    for (int i = 0; i < (NP > 0 ? NUP : 1); i++) st.cP[i] = T(0);
    st.prior = st.lj = st.u2 = st.ljln = T(0);
#pragma unroll
    for (int o = 0; o < SiteState<T, NP, NM, NO>::NOP; o++) st.q[o] = mk2(T(0), T(0));
    __syncwarp();   // every lane has taken its records: the buffer may be refilled
    if (mode != 1) {
      if (has_next) issue_site(tnext);
      cp_async_commit();
    }

    // three sources of the draws ξ of this tile, one copy of the per-draw code:
    //   mode 1: cp.async staging through shared memory (two half-row stages of MCS draws in flight)
    //   mode 2: generated in registers (Philox + Box-Muller), no memory traffic at all
    //   mode 0: direct global loads (n_MC not a multiple of the 16-byte chunk)
    {
      for (int st_ = 0; st_ < nst; st_++, g++) {
        if (mode == 1) {
          if (st_ + 1 < nst) issue_xi(tile, st_ + 1, (g + 1) & 1);
          else if (has_next) issue_xi(tnext, 0, (g + 1) & 1);
          if (st_ == 0 && has_next) issue_site(tnext);
          cp_async_commit();
          cp_async_wait_group<1>();
          __syncwarp();
        }
        const int m0 = mode == 1 ? st_ * MCS : 0;
        const int nd = mode == 1 ? ((M - m0) < MCS ? (M - m0) : MCS) : M;   // draws of this stage
        const T* buf = ztile + (g & 1) * (32 * NM * MCS);
        if (active) {
          const int nfull = COMPACT ? 0 : nd / CH;
          // pbm covariates: the per-draw means of the NEXT chunk travel while this chunk is computed
          T mun[CH][NM];
          if constexpr (PBM) {
            if (nfull > 0) {
#pragma unroll
              for (int d = 0; d < CH; d++)
#pragma unroll
                for (int k = 0; k < NM; k++) mun[d][k] = a.MU[((size_t)(m0 + d) * NM + k) * a.nb + site];
            }
          }
          for (int q = 0; q < nfull; q++) {   // full chunks of CH draws, no guards
            const int mq = m0 + q * CH;
            T zc[NM][CH];
            if (mode == 1) {
#pragma unroll
              for (int k = 0; k < NM; k++) {
                const T* src = buf + (lane * NM + k) * MCS + ((q + lane) & 3) * CH;
                if constexpr (sizeof(T) == 4) {
                  const float4 v = *reinterpret_cast<const float4*>(src);
                  zc[k][0] = v.x; zc[k][1] = v.y; zc[k][2] = v.z; zc[k][3] = v.w;
                } else {
                  const double2 v = *reinterpret_cast<const double2*>(src);
                  zc[k][0] = v.x; zc[k][1] = v.y;
                }
              }
            } else if (mode == 2) {
#pragma unroll
              for (int k = 0; k < NM; k++) {
                float n4[4];
                normal4(seed, (uint64_t)(a.col_offset + site * NM + k), mq / 4, n4);
#pragma unroll
                for (int d = 0; d < CH; d++) zc[k][d] = (T)n4[(mq % 4) + d];
              }
            } else {
#pragma unroll
              for (int k = 0; k < NM; k++)
#pragma unroll
                for (int d = 0; d < CH; d++) zc[k][d] = a.zMs[((size_t)site * NM + k) * M + mq + d];
            }
            T mud[CH][NM], z[CH][NM], zb[CH][NM];
#pragma unroll
            for (int d = 0; d < CH; d++)
#pragma unroll
              for (int k = 0; k < NM; k++) {
                if constexpr (PBM) mud[d][k] = mun[d][k]; else mud[d][k] = st.mu[k];
                z[d][k] = zc[k][d];
              }
            if constexpr (PBM) {
              if (q + 1 < nfull) {
#pragma unroll
                for (int d = 0; d < CH; d++)
#pragma unroll
                  for (int k = 0; k < NM; k++) mun[d][k] = a.MU[((size_t)(mq + CH + d) * NM + k) * a.nb + site];
              }
            }
            if constexpr (ILV) {
              DrawTmp<T, NP, NM> dt[CH];
              T pbar[CH][4];
#pragma unroll
              for (int d = 0; d < CH; d++) {
                const T* pd = PDS ? spd + (size_t)(mq + d) * PD : a.pd + (size_t)(mq + d) * PD;
                draw_pre<TR, T, NP, NM, NO>(cfg, st, z[d], pd, mud[d], dt[d]);
              }
#pragma unroll
              for (int d = 0; d < CH; d += 2) obs_loop2<T, NP, NM, NO>(st, dt[d].par, dt[d + 1].par, pbar[d], pbar[d + 1]);
#pragma unroll
              for (int d = 0; d < CH; d++) draw_post<TR, T, NP, NM, NO>(cfg, st, z[d], dt[d], pbar[d], zb[d]);
            } else {
#pragma unroll
              for (int d = 0; d < CH; d++) {
                const T* pd = PDS ? spd + (size_t)(mq + d) * PD : a.pd + (size_t)(mq + d) * PD;
                do_draw<TR, T, NP, NM, NO>(cfg, st, z[d], pd, mud[d], zb[d]);
              }
            }
            if constexpr (PBM) {
#pragma unroll
              for (int d = 0; d < CH; d++)
#pragma unroll
                for (int k = 0; k < NM; k++) a.MUBAR[((size_t)(mq + d) * NM + k) * a.nb + site] = wM * zb[d][k];
            }
          }
          // remaining draws (n_MC not a multiple of the chunk; never in mode 1), one at a time
#pragma unroll 1
          for (int m = m0 + nfull * CH; m < m0 + nd; m++) {
            T z[NM], zb[NM], mud[NM];
#pragma unroll
            for (int k = 0; k < NM; k++) {
              if (mode == 2) {
                float n4[4];
                normal4(seed, (uint64_t)(a.col_offset + site * NM + k), m / 4, n4);
                z[k] = (T)(m % 4 == 0 ? n4[0] : m % 4 == 1 ? n4[1] : m % 4 == 2 ? n4[2] : n4[3]);
              } else {
                z[k] = a.zMs[((size_t)site * NM + k) * M + m];
              }
              mud[k] = PBM ? a.MU[((size_t)m * NM + k) * a.nb + site] : st.mu[k];
            }
            const T* pd = PDS ? spd + (size_t)m * PD : a.pd + (size_t)m * PD;
            do_draw<TR, T, NP, NM, NO>(cfg, st, z, pd, mud, zb);
            if constexpr (PBM) {
#pragma unroll
              for (int k = 0; k < NM; k++) a.MUBAR[((size_t)m * NM + k) * a.nb + site] = wM * zb[k];
            }
          }
        }
        if (mode == 1) __syncwarp();
      }
    }

    // end of site: σ-path, entropy, μ̄_ζ  (SURVEY Appendix A adjoint)
    constexpr int NPAD = NACC <= 16 ? 16 : 32;
    static_assert(NACC <= 32, "at most 32 accumulators");
    T acc[NPAD];
#pragma unroll
    for (int i = 0; i < NPAD; i++) acc[i] = T(0);
    if (active) {
      {
        typename PairT<T>::type s2 = mk2(T(0), T(0));
#pragma unroll
        for (int o = 0; o < SiteState<T, NP, NM, NO>::NOP; o++) s2 = pfma(st.w[o], st.q[o], s2);
        acc[ACC_NLY] = s2.x + s2.y;
      }
      acc[ACC_LJ] = st.lj;
      acc[ACC_PRIOR] = st.prior + T(0.5) * st.u2 + (all_exp_ln<TR, NM>() ? st.lj : st.ljln);
#pragma unroll
      for (int k = 0; k < NM; k++) {
        T sr = T(0);                                      // Σ_m ζ̄[k,m]·r[k,m]
#pragma unroll
        for (int j = 0; j <= k; j++) sr = fma(st.sU[ut(j, k)], st.sz[ut(j, k)], sr);
        const T lsb = T(0.5) * wM * sr - T(0.5);          // −½ from the entropy
        acc[ACC_LOGSIG] += T(0.5) * ls[k];
        if (cfg.sepvar) {
          // the gradient w.r.t. this site's own log-variance goes straight to its slot of ∇ϕ; the
          // slot of ā counts non-finite entries instead
          const int64_t o = (int64_t)cfg.o_coef + isite * NM + k;
          a.out[o] = (double)lsb;
          if (a.grad_T) a.grad_T[o] = lsb;
          acc[ACC_FIRST_VEC + k] = isfinite(lsb) ? T(0) : T(1);
          a.mubar[site * NM + k] = PBM ? T(0) : wM * st.smu[k];
        } else {
          acc[ACC_FIRST_VEC + k] = lsb;
          acc[ACC_FIRST_VEC + NM + k] = lsb * st.mu[k];
          // with pbm covariates the pass-1 mean feeds σ only (src/elbo.jl:491-495)
          a.mubar[site * NM + k] = PBM ? lsb * cB[k] : fma(wM, st.smu[k], lsb * cB[k]);
        }
#pragma unroll
        for (int j = 0; j <= k; j++) acc[ACC_FIRST_VEC + 2 * NM + ut(j, k)] = sig[k] * st.sz[ut(j, k)];
      }
#pragma unroll
      for (int j = 0; j < NP; j++) acc[ACC_FIRST_VEC + 2 * NM + NUM + j] = st.aP[j];
#pragma unroll
      for (int i = 0; i < NUP; i++) acc[ACC_FIRST_VEC + 2 * NM + NUM + NP + i] = st.cP[i];
    }
    {
      // all accumulators at once: lane l ends up with the warp's total of accumulator warp_multi_index(l)
      const T v = warp_multi_sum<T, NPAD>(acc, lane);
      const int ia = warp_multi_index<NPAD>(lane);
      if (ia < NACC && (NPAD == 32 || !(lane & 1))) wacc[ia] += (double)v;
    }
  }
  cp_async_wait_group<0>();
  // one row of partial sums per block: the warps' double accumulators added in warp order
  __syncthreads();
  {
    const double* w0 = reinterpret_cast<const double*>(smem_raw + OFF_ACC);
    for (int i = threadIdx.x; i < NACC; i += STREAM_THREADS) {
      double s = 0.0;
      for (int w = 0; w < NWARP; w++) s += w0[w * NACC + i];
      a.partials[(size_t)bid * NACC + i] = s;
    }
  }
}


// ---------------------------------------------------------------------------------------
// The streaming phase of a SMALL batch inside one block (k_small_iter): one thread per (site, draw) unit instead of one
// per site -- 20 sites x 12 draws keep 240 threads busy instead of 20 lanes of one warp.  A unit runs the per-draw code
// once (do_draw, device draws) and hands its contributions to shared memory; one thread per site then adds its draws
// in draw order and applies the end-of-site formulas of stream_body; the sites are added in site order (double).
// No MeanVarSep, no pbm covariates (k_small_iter supports neither).
// ---------------------------------------------------------------------------------------
template <int NP, int NM>
__host__ __device__ constexpr int stream_units_nv() { return NM + tri(NM) + (NP > 0 ? NP + tri(NP) : 0) + 5; }
__host__ __device__ constexpr size_t stream_units_smem(size_t szT, int64_t nb, int M, int nv, int nacc_) {
  return ((size_t)nb * M * nv + (size_t)nb * nacc_) * szT;
}

template <class TR, typename T, int NP, int NM, int NO, int THREADS>
__device__ __forceinline__ void stream_units_body(const Cfg<T>& cfg, const StreamArgs<T>& a, unsigned char* smem_raw, uint64_t seed) {
  constexpr int NUM = tri(NM), NUP = tri(NP), NACC = nacc(NP, NM);
  constexpr int NV = stream_units_nv<NP, NM>();
  constexpr int PD = pd_stride(NP);
  constexpr int REC = 4 * NO * 32;
  using ST = SiteState<T, NP, NM, NO>;
  const int t = threadIdx.x;
  const int M = a.M;
  const int nb = (int)a.nb, units = nb * M;
  T* vals = reinterpret_cast<T*>(smem_raw);          // [units][NV]
  T* accs = vals + (size_t)units * NV;               // [nb][NACC]
  T cA[NM], cB[NM], U[NUM];
#pragma unroll
  for (int k = 0; k < NM; k++) {
    cA[k] = a.ec->coefA[k];
    cB[k] = a.ec->coefB[k];
#pragma unroll
    for (int j = 0; j <= k; j++) U[ut(j, k)] = a.ec->UM[j][k];
  }
  const T wM = T(1) / T(M);
  // ---- one (site, draw) unit per thread
  for (int u = t; u < units; u += THREADS) {
    const int site = u / M, m = u - site * M;
    ST st;
    {
      const T* r = a.rec + (size_t)(site >> 5) * REC + (site & 31);
#pragma unroll
      for (int o = 0; o < ST::NOP; o++) {
        const bool two = (2 * o + 1 < NO);
        const int o0 = 2 * o, o1 = two ? 2 * o + 1 : 2 * o;
        st.S1[o] = mk2(r[(0 * NO + o0) * 32], r[(0 * NO + o1) * 32]);
        st.S2[o] = mk2(r[(1 * NO + o0) * 32], r[(1 * NO + o1) * 32]);
        st.nyo[o] = mk2(-r[(2 * NO + o0) * 32], -r[(2 * NO + o1) * 32]);
        st.w[o] = mk2(r[(3 * NO + o0) * 32], two ? r[(3 * NO + o1) * 32] : T(0));
        st.S12[o] = pmul(st.S1[o], st.S2[o]);
        st.q[o] = mk2(T(0), T(0));
      }
    }
    T z[NM], zb[NM];
#pragma unroll
    for (int k = 0; k < NM; k++) {
      st.mu[k] = a.mu[(size_t)site * NM + k];
      const T sig = exp_t(T(0.5) * fma(cB[k], st.mu[k], cA[k]));
      st.smu[k] = T(0);
#pragma unroll
      for (int j = 0; j <= k; j++) st.sU[ut(j, k)] = sig * U[ut(j, k)];
      float n4[4];
      normal4(seed, (uint64_t)(a.col_offset + (int64_t)site * NM + k), m / 4, n4);
      z[k] = (T)((m & 3) == 0 ? n4[0] : (m & 3) == 1 ? n4[1] : (m & 3) == 2 ? n4[2] : n4[3]);
    }
#pragma unroll
    for (int i = 0; i < NUM; i++) st.sz[i] = T(0);
#pragma unroll
    for (int i = 0; i < (NP > 0 ? NP : 1); i++) st.aP[i] = T(0);
#pragma unroll
    for (int i = 0; i < (NP > 0 ? NUP : 1); i++) st.cP[i] = T(0);
    st.prior = st.lj = st.u2 = st.ljln = T(0);
    do_draw<TR, T, NP, NM, NO>(cfg, st, z, a.pd + (size_t)m * PD, st.mu, zb);
    T* v = vals + (size_t)u * NV;
    int c = 0;
#pragma unroll
    for (int k = 0; k < NM; k++) v[c++] = st.smu[k];
#pragma unroll
    for (int i = 0; i < NUM; i++) v[c++] = st.sz[i];
    if constexpr (NP > 0) {
#pragma unroll
      for (int j = 0; j < NP; j++) v[c++] = st.aP[j];
#pragma unroll
      for (int i = 0; i < NUP; i++) v[c++] = st.cP[i];
    }
    v[c++] = st.prior; v[c++] = st.lj; v[c++] = st.u2; v[c++] = st.ljln;
    {
      typename PairT<T>::type s2 = mk2(T(0), T(0));
#pragma unroll
      for (int o = 0; o < ST::NOP; o++) s2 = pfma(st.w[o], st.q[o], s2);
      v[c++] = s2.x + s2.y;
    }
  }
  __syncthreads();
  // ---- one thread per site: its draws in draw order, then the end of site (σ-path, entropy, μ̄_ζ) as in stream_body
  for (int site = t; site < nb; site += THREADS) {
    T sum[NV];
#pragma unroll
    for (int c = 0; c < NV; c++) sum[c] = T(0);
    for (int m = 0; m < M; m++) {
      const T* v = vals + ((size_t)site * M + m) * NV;
#pragma unroll
      for (int c = 0; c < NV; c++) sum[c] += v[c];
    }
    const T* smu = sum;
    const T* sz = sum + NM;
    const T* aP = sz + NUM;
    const T* cP = aP + (NP > 0 ? NP : 0);
    const T* tail = sum + NV - 5;   // prior, lj, u2, ljln, nly
    T acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = T(0);
    acc[ACC_NLY] = tail[4];
    acc[ACC_LJ] = tail[1];
    acc[ACC_PRIOR] = tail[0] + T(0.5) * tail[2] + (all_exp_ln<TR, NM>() ? tail[1] : tail[3]);
#pragma unroll
    for (int k = 0; k < NM; k++) {
      const T mu = a.mu[(size_t)site * NM + k];
      const T ls = fma(cB[k], mu, cA[k]);
      const T sig = exp_t(T(0.5) * ls);
      T sr = T(0);
#pragma unroll
      for (int j = 0; j <= k; j++) sr = fma(sig * U[ut(j, k)], sz[ut(j, k)], sr);
      const T lsb = T(0.5) * wM * sr - T(0.5);
      acc[ACC_LOGSIG] += T(0.5) * ls;
      acc[ACC_FIRST_VEC + k] = lsb;
      acc[ACC_FIRST_VEC + NM + k] = lsb * mu;
      a.mubar[(size_t)site * NM + k] = fma(wM, smu[k], lsb * cB[k]);
#pragma unroll
      for (int j = 0; j <= k; j++) acc[ACC_FIRST_VEC + 2 * NM + ut(j, k)] = sig * sz[ut(j, k)];
    }
    if constexpr (NP > 0) {
#pragma unroll
      for (int j = 0; j < NP; j++) acc[ACC_FIRST_VEC + 2 * NM + NUM + j] = aP[j];
#pragma unroll
      for (int i = 0; i < NUP; i++) acc[ACC_FIRST_VEC + 2 * NM + NUM + NP + i] = cP[i];
    }
#pragma unroll
    for (int i = 0; i < NACC; i++) accs[(size_t)site * NACC + i] = acc[i];
  }
  __syncthreads();
  for (int i = t; i < NACC; i += THREADS) {
    double tot = 0.0;
    for (int site = 0; site < nb; site++) tot += (double)accs[(size_t)site * NACC + i];
    a.partials[i] = tot;
  }
}

// (n_θP, n_θM, n_obs) shapes with a compiled streaming kernel
#define HVI_STREAM_CASES(X) \
  X(0, 2, 8) X(1, 2, 8) X(2, 2, 8) X(0, 3, 8) X(1, 3, 8) X(2, 3, 8) X(3, 2, 8) X(2, 1, 8) X(2, 2, 4)

}  // namespace hvi
