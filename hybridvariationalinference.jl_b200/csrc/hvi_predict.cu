// hvi_predict.cu -- predict_hvi / sample_posterior forward path.
// Reference citations (file:line) are relative to /root/reference.
#include <math.h>
#include <stdlib.h>

#include "hvi_internal.cuh"
#include "hvi_rng.cuh"

namespace hvi {

template <typename T>
__device__ __forceinline__ T warp_sum_p(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// predict_hvi / sample_posterior forward path (src/elbo.jl:320-458; transform_ζs :815-835; 3-D PBM
// applicator src/PBMApplicator.jl:51-57): for every posterior sample m and site i
//   θsMs_tr[i,k,m] = T_k(ζMs[i,k,m]),  y[o,i,m] = f_doubleMM(θP_m, θMs_{i,m}, xP[:,i]),
// no clamp (transform_ζs applies none), NaN where a driver is not finite.  One thread per site,
// writes coalesced over sites; output-bandwidth bound (4·(n_θM + n_obs) bytes per site·sample).
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T transform_only(int tr, T zeta) {
  if (tr == HVI_TR_EXP) return exp(zeta);
  if (tr == HVI_TR_LOGISTIC) return T(1) / (T(1) + exp(-zeta));
  return zeta;
}

template <typename T>
__global__ void __launch_bounds__(256) k_predict(const __grid_constant__ Cfg<T> cfg,
                                                 const __grid_constant__ StreamArgs<T> a, const T* __restrict__ pdraw,
                                                 const T* __restrict__ xP, T* __restrict__ thetaP,
                                                 T* __restrict__ thetaMs_tr, T* __restrict__ y,
                                                 double* __restrict__ part_logsig) {
  const int nP = cfg.nP, nM = cfg.nM, nO = cfg.nO, M = a.M;
  const EvalConsts<T>& ec = *a.ec;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const T* zetaP = pdraw + (size_t)nP * M;
  if (gid < (int64_t)nP * M) {
    const int j = (int)(gid / M), m = (int)(gid % M);
    thetaP[j + (size_t)nP * m] = transform_only<T>(cfg.transP[j], zetaP[gid]);
  }
  double logsig = 0.0;
  for (int64_t i = gid; i < a.nb; i += (int64_t)gridDim.x * blockDim.x) {
    T mu[MAXP], sg[MAXP], S1[32], S2[32];
    for (int k = 0; k < nM; k++) {
      mu[k] = a.mu[i * nM + k];
      const T ls = cfg.sepvar ? a.phi[cfg.o_coef + (a.i_sites ? a.i_sites[i] : i) * nM + k] : ec.coefA[k] + ec.coefB[k] * mu[k];
      sg[k] = exp(T(0.5) * ls);
      logsig += 0.5 * (double)ls;
    }
    for (int o = 0; o < nO; o++) { S1[o] = xP[i * 2 * nO + o]; S2[o] = xP[i * 2 * nO + nO + o]; }
    for (int m = 0; m < M; m++) {
      T th[MAXP], par[4];
      for (int k = 0; k < nM; k++) {
        T t = T(0);
        for (int j = cfg.blkM_start[k]; j <= k; j++) {
          T z;
          if (a.rng) {
            float n4[4];
            normal4(a.seed, (uint64_t)(a.col_offset + i * nM + j), m / 4, n4);
            z = (T)(m % 4 == 0 ? n4[0] : m % 4 == 1 ? n4[1] : m % 4 == 2 ? n4[2] : n4[3]);
          } else {
            z = a.zMs[((size_t)i * nM + j) * M + m];
          }
          t += ec.UM[j][k] * z;
        }
        const T mum = a.MU ? a.MU[((size_t)m * nM + k) * a.nb + i] : mu[k];
        th[k] = transform_only<T>(cfg.transM[k], mum + sg[k] * t);
        thetaMs_tr[i + a.nb * (k + (size_t)nM * m)] = th[k];
      }
      for (int s = 0; s < 4; s++) {
        const int c = cfg.slot_code[s];
        par[s] = c < 0 ? cfg.slot_fix[s] : c < nP ? transform_only<T>(cfg.transP[c], zetaP[(size_t)c * M + m]) : th[c - nP];
      }
      T* yo = y + (size_t)nO * (i + a.nb * (size_t)m);
      for (int o = 0; o < nO; o++) {
        // is_valid .* p: invalid drivers give NaN predictions (src/DoubleMM/f_doubleMM.jl:111-137)
        const bool valid = isfinite(S1[o]) && isfinite(S2[o]);
        yo[o] = valid ? par[0] + par[1] * S1[o] / (par[2] + S1[o]) * S2[o] / (par[3] + S2[o]) : T(NAN);
      }
    }
  }
  logsig = warp_sum_p(logsig);
  if ((threadIdx.x & 31) == 0) part_logsig[gid >> 5] = logsig;
}


// ---------------------------------------------------------------------------------------
// The same path sized for its roofline: it writes 4·(n_obs + n_θM) bytes per site·sample and reads next to nothing,
// so the kernel is built around the stores.  One thread per site, 32 consecutive sites per warp: for a fixed sample
// the warp's predictions are one contiguous block of 32·n_obs values, written with 16-byte streaming stores; the
// site parameters are written one coalesced row per (k, sample).  Samples are taken four at a time (one Philox call
// yields the four draws of a column), the samples are split over blockIdx.y so that small batches with many samples
// still fill the machine, θP of the block's samples is staged in shared memory, drivers and S1·S2 live in registers,
// and Float32 replaces the two divisions by one approximate reciprocal of (K1+S1)(K2+S2) (2^-22 relative).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float rcp_p(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ double rcp_p(double x) { return 1.0 / x; }
// Float32: exp through ex2.approx (≈2e-7 relative, as in the streaming kernel); Float64: libm
__device__ __forceinline__ float exp_p(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}
__device__ __forceinline__ double exp_p(double x) { return exp(x); }
template <typename T>
__device__ __forceinline__ T transform_fast(int tr, T zeta) {
  if (tr == HVI_TR_EXP) return exp_p(zeta);
  if (tr == HVI_TR_LOGISTIC) return rcp_p(T(1) + exp_p(-zeta));
  return zeta;
}

// 256-bit streaming stores (sm_100): one full 32-byte sector per lane and instruction, so a warp's store is one
// contiguous kilobyte of full sectors (two 16-byte stores per lane reach L2 as twice as many half-filled requests)
__device__ __forceinline__ void st256(float* p, const float* v) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void st256(double* p, const double* v) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

constexpr int PR_THREADS = 256;
constexpr int PR_MAX_SMEM_DRAWS = 4096;

// SLOTS: where (r0, r1, K1, K2) come from, one nibble each = slot code + 1 (0 = fixed value), and TRM: bijector code of
// every θM column, 2 bits each -- compile-time for the named DoubleMM problems (as StaTraits of the streaming kernel),
// PR_DYN = read from the problem description at run time
constexpr uint32_t PR_DYN = 0xffffffffu;

template <typename T, int NM, int NO, bool RNG, bool PBM, uint32_t SLOTS, uint32_t TRM>
__global__ void __launch_bounds__(PR_THREADS) k_predict_fast(const __grid_constant__ Cfg<T> cfg,
                                                             const __grid_constant__ StreamArgs<T> a,
                                                             const T* __restrict__ pdraw, const T* __restrict__ xP,
                                                             T* __restrict__ thetaP, T* __restrict__ thetaMs_tr,
                                                             T* __restrict__ y, double* __restrict__ part_logsig,
                                                             int draws_per_split) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sthP = reinterpret_cast<T*>(smem_raw);            // [draws of this split][nP]
  const int nP = cfg.nP, M = a.M;
  const EvalConsts<T>& ec = *a.ec;
  const int m0 = blockIdx.y * draws_per_split, m1 = min(M, m0 + draws_per_split);
  const T* zetaP = pdraw + (size_t)nP * M;
  for (int e = threadIdx.x; e < nP * (m1 - m0); e += PR_THREADS) {
    const int j = e % nP, m = m0 + e / nP;
    const T v = transform_only<T>(cfg.transP[j], zetaP[(size_t)j * M + m]);
    sthP[e] = v;
    if (blockIdx.x == 0) thetaP[j + (size_t)nP * m] = v;
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * PR_THREADS + threadIdx.x;
  const bool active = i < a.nb;
  double logsig = 0.0;
  if (active) {
    T mu[NM], sg[NM], U[NM][NM];
#pragma unroll
    for (int k = 0; k < NM; k++) {
      mu[k] = a.mu[i * NM + k];
      const T ls = cfg.sepvar ? a.phi[cfg.o_coef + (a.i_sites ? a.i_sites[i] : i) * NM + k] : ec.coefA[k] + ec.coefB[k] * mu[k];
      sg[k] = exp(T(0.5) * ls);
      logsig += 0.5 * (double)ls;
#pragma unroll
      for (int j = 0; j < NM; j++) U[j][k] = (j >= cfg.blkM_start[k] && j <= k) ? ec.UM[j][k] * sg[k] : T(0);
    }
    // drivers: rows S1[0..NO), S2[0..NO) of this site's column of xP.  is_valid .* p (src/DoubleMM/f_doubleMM.jl:111-137):
    // a non-finite driver gives a NaN prediction -- folded into the numerator S1·S2, so the sample loop has no select
    T S1[NO], S2[NO], S12[NO];
    {
      const T* col = xP + i * 2 * NO;
#pragma unroll
      for (int o = 0; o < NO; o++) {
        const T s1 = col[o], s2 = col[NO + o];
        const bool valid = isfinite(s1) && isfinite(s2);
        S1[o] = valid ? s1 : T(1); S2[o] = valid ? s2 : T(1);
        S12[o] = valid ? s1 * s2 : T(NAN);
      }
    }
    // where each of (r0, r1, K1, K2) comes from: a fixed value, θP[c] of the sample (shared memory) or θM[k]
    int pidx[4], kidx[4];
    T fix[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int c = SLOTS == PR_DYN ? cfg.slot_code[s] : (int)((SLOTS >> (4 * s)) & 15u) - 1;
      fix[s] = cfg.slot_fix[s];
      pidx[s] = (c >= 0 && c < nP) ? c : -1;
      kidx[s] = c >= nP ? c - nP : -1;
    }
    int trM[NM];
#pragma unroll
    for (int k = 0; k < NM; k++) trM[k] = TRM == PR_DYN ? cfg.transM[k] : (int)((TRM >> (2 * k)) & 3u);
    T* yo = y + (size_t)NO * (i + a.nb * (size_t)m0);
    T* tho = thetaMs_tr + i + a.nb * ((size_t)NM * m0);
    const size_t ystep = (size_t)NO * a.nb, tstep = (size_t)a.nb;
    // one sample: ζ = μ + σ·Uᵀz, θ = T(ζ), the site's predictions; all stores are streaming (never re-read)
    auto one_sample = [&](int m, const T (&zd)[NM]) {
      T th[NM];
#pragma unroll
      for (int k = 0; k < NM; k++) {
        T r = PBM ? a.MU[((size_t)m * NM + k) * a.nb + i] : mu[k];
#pragma unroll
        for (int j = 0; j <= k; j++) r = fma(U[j][k], zd[j], r);
        th[k] = transform_fast<T>(trM[k], r);
        __stcs(tho, th[k]);
        tho += tstep;
      }
      T par[4];
#pragma unroll
      for (int s = 0; s < 4; s++) {
        T v = fix[s];
        if (pidx[s] >= 0) v = sthP[(m - m0) * nP + pidx[s]];
#pragma unroll
        for (int k = 0; k < NM; k++) v = (kidx[s] == k) ? th[k] : v;
        par[s] = v;
      }
      T yv[NO];
#pragma unroll
      for (int o = 0; o < NO; o++) yv[o] = fma(par[1], S12[o] * rcp_p((par[2] + S1[o]) * (par[3] + S2[o])), par[0]);
      constexpr int V32 = 32 / (int)sizeof(T);   // elements per 32-byte sector
      if constexpr (NO % V32 == 0) {
#pragma unroll
        for (int o = 0; o < NO; o += V32) st256(yo + o, yv + o);
      } else {
#pragma unroll
        for (int o = 0; o < NO; o++) __stcs(yo + o, yv[o]);
      }
      yo += ystep;
    };
    int mq = m0;
    for (; mq + 4 <= m1; mq += 4) {   // full groups of four samples (one Philox call per column)
      T z[NM][4];
#pragma unroll
      for (int j = 0; j < NM; j++) {
        if constexpr (RNG) {
          float n4[4];
          normal4(a.seed, (uint64_t)(a.col_offset + i * NM + j), mq / 4, n4);
#pragma unroll
          for (int d = 0; d < 4; d++) z[j][d] = (T)n4[d];
        } else {
#pragma unroll
          for (int d = 0; d < 4; d++) z[j][d] = a.zMs[((size_t)i * NM + j) * M + mq + d];
        }
      }
#pragma unroll
      for (int d = 0; d < 4; d++) {
        T zd[NM];
#pragma unroll
        for (int j = 0; j < NM; j++) zd[j] = z[j][d];
        one_sample(mq + d, zd);
      }
    }
    for (; mq < m1; mq++) {           // the last one to three samples
      T zd[NM];
#pragma unroll
      for (int j = 0; j < NM; j++) {
        if constexpr (RNG) {
          float n4[4];
          normal4(a.seed, (uint64_t)(a.col_offset + i * NM + j), mq / 4, n4);
          zd[j] = (T)(mq % 4 == 0 ? n4[0] : mq % 4 == 1 ? n4[1] : mq % 4 == 2 ? n4[2] : n4[3]);
        } else {
          zd[j] = a.zMs[((size_t)i * NM + j) * M + mq];
        }
      }
      one_sample(mq, zd);
    }
  }
  logsig = warp_sum_p(logsig);
  if (blockIdx.y == 0 && (threadIdx.x & 31) == 0) part_logsig[(size_t)blockIdx.x * (PR_THREADS / 32) + (threadIdx.x >> 5)] = logsig;
}

template <typename T>
static uint32_t slots_key(const Cfg<T>& cfg) {
  uint32_t k = 0;
  for (int s = 0; s < 4; s++) k |= (uint32_t)(cfg.slot_code[s] + 1) << (4 * s);
  return k;
}
template <typename T>
static uint32_t trm_key(const Cfg<T>& cfg) {
  uint32_t k = 0;
  for (int i = 0; i < cfg.nM; i++) k |= (uint32_t)cfg.transM[i] << (2 * i);
  return k;
}

static int predict_dps_override() {   // experiments: samples per block (multiple of 4)
  static int v = -1;
  if (v < 0) { const char* e = getenv("HVI_PREDICT_DPS"); v = e ? atoi(e) / 4 * 4 : 0; }
  return v;
}

#define HVI_PREDICT_CASES(X) X(2, 8) X(1, 8) X(3, 8) X(2, 4)

template <typename T>
cudaError_t launch_predict(const Launch& L, const Cfg<T>& cfg, const StreamArgs<T>& a, const T* pdraw, const T* xP,
                           T* thetaP, T* thetaMs_tr, T* y, double* part_logsig, int* nparts) {
  if (cfg.nO > 32) return cudaErrorInvalidConfiguration;
  const int64_t sblocks = (a.nb + PR_THREADS - 1) / PR_THREADS;
  const bool aligned16 = ((uintptr_t)y & 31) == 0 && ((uintptr_t)xP & 15) == 0;   // 256-bit stores of y
#define X(NM_, NO_)                                                                                                  \
  if (cfg.nM == NM_ && cfg.nO == NO_ && aligned16 && sblocks <= 0x7fffffff) {                                         \
    /* split the samples until the grid has about 8 blocks per SM; splits are multiples of 4 samples */              \
    int64_t splits = ((int64_t)L.sm_count * 8 + sblocks - 1) / sblocks;                                               \
    const int64_t max_splits = (a.M + 3) / 4;                                                                         \
    if (splits > max_splits) splits = max_splits;                                                                     \
    if (splits > 65535) splits = 65535;                                                                               \
    if (splits < 1) splits = 1;                                                                                       \
    int dps = (int)((((a.M + splits - 1) / splits) + 3) / 4 * 4);                                                     \
    if (predict_dps_override() > 0) dps = predict_dps_override();                                                     \
    if (dps > PR_MAX_SMEM_DRAWS) dps = PR_MAX_SMEM_DRAWS;                                                             \
    splits = (a.M + dps - 1) / dps;                                                                                   \
    if (splits <= 65535) {                                                                                            \
      const size_t smem = sizeof(T) * (size_t)(cfg.nP > 0 ? cfg.nP : 1) * dps + 16;                                   \
      dim3 grid((unsigned)sblocks, (unsigned)splits);                                                                 \
      const bool rng_ = a.rng != 0, pbm_ = a.MU != nullptr;                                                           \
      /* DoubleMMCase default: θP = (r0, K2), θM = (r1, K1), both Exp (src/DoubleMM/f_doubleMM.jl:161-243) */         \
      const bool named = NM_ == 2 && cfg.nP == 2 && slots_key(cfg) == 0x2431u && trm_key(cfg) == 0x5u;                \
      auto kern = k_predict_fast<T, NM_, NO_, false, false, PR_DYN, PR_DYN>;                                          \
      if (named) kern = rng_ ? (pbm_ ? k_predict_fast<T, NM_, NO_, true, true, 0x2431u, 0x5u>                         \
                                     : k_predict_fast<T, NM_, NO_, true, false, 0x2431u, 0x5u>)                       \
                             : (pbm_ ? k_predict_fast<T, NM_, NO_, false, true, 0x2431u, 0x5u>                        \
                                     : k_predict_fast<T, NM_, NO_, false, false, 0x2431u, 0x5u>);                     \
      else kern = rng_ ? (pbm_ ? k_predict_fast<T, NM_, NO_, true, true, PR_DYN, PR_DYN>                              \
                               : k_predict_fast<T, NM_, NO_, true, false, PR_DYN, PR_DYN>)                            \
                       : (pbm_ ? k_predict_fast<T, NM_, NO_, false, true, PR_DYN, PR_DYN>                             \
                               : k_predict_fast<T, NM_, NO_, false, false, PR_DYN, PR_DYN>);                          \
      kern<<<grid, PR_THREADS, smem, L.stream>>>(cfg, a, pdraw, xP, thetaP, thetaMs_tr, y, part_logsig, dps);         \
      cudaError_t e_ = cudaGetLastError();                                                                            \
      if (e_ != cudaSuccess) return e_;                                                                               \
      if (L.counter) ++*L.counter;                                                                                    \
      *nparts = (int)sblocks * (PR_THREADS / 32);                                                                     \
      return cudaSuccess;                                                                                             \
    }                                                                                                                 \
  }
  HVI_PREDICT_CASES(X)
#undef X
  int64_t n = a.nb > (int64_t)cfg.nP * a.M ? a.nb : (int64_t)cfg.nP * a.M;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)L.sm_count * 16 && (int64_t)L.sm_count * 16 * 256 >= (int64_t)cfg.nP * a.M) blocks = (int64_t)L.sm_count * 16;
  if (blocks < 1) blocks = 1;
  k_predict<T><<<(int)blocks, 256, 0, L.stream>>>(cfg, a, pdraw, xP, thetaP, thetaMs_tr, y, part_logsig);
  cudaError_t e_ = cudaGetLastError();
  if (e_ != cudaSuccess) return e_;
  if (L.counter) ++*L.counter;
  *nparts = (int)blocks * 8;
  return cudaSuccess;
}
template cudaError_t launch_predict<float>(const Launch&, const Cfg<float>&, const StreamArgs<float>&, const float*,
                                           const float*, float*, float*, float*, double*, int*);
template cudaError_t launch_predict<double>(const Launch&, const Cfg<double>&, const StreamArgs<double>&, const double*,
                                            const double*, double*, double*, double*, double*, int*);

}  // namespace hvi
