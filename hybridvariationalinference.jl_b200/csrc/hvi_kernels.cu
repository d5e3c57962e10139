// hvi_kernels.cu -- sm_100a kernels of the negative-ELBO + gradient path.
//
// Pipeline per evaluation (one CUDA stream, no host round trips):
//   k_prologue   ϕq → U_P, U_M, σ_P, per-draw θP / dθP      (src/elbo.jl:633-660, cholesky.jl:156-174)
//   k_mlp_fwd    μ_ζ = g(xM; ϕg)                              (src/ModelApplicator.jl:163-176)
//   k_stream     fused sample_ζ → T(ζ)+logjac → f_doubleMM → logden_normal → priors → adjoint,
//                one thread per site, draws streamed through warp-private shared memory
//                (src/elbo.jl:132-201,472-521,633-678,783-808; f_doubleMM.jl:101-139; logden_normal.jl:16-43)
//   k_mlp_bwd    μ̄_ζ → ∇ϕg
//   k_finalize   deterministic reduction of the per-warp partial sums, Cholesky adjoint,
//                global-block terms, scatter into the ϕ layout (src/init_hybrid_params.jl:119-124)
// All file:line citations are relative to /root/reference.
#include <math.h>
#include <stdlib.h>

#include "hvi_internal.cuh"
#include "hvi_rng.cuh"
#include "hvi_dev.cuh"
#include "hvi_stream.cuh"

namespace hvi {


// ---------------------------------------------------------------------------------------
// set_data: re-tile one batch into site records and fold the ϕ-independent part of the
// likelihood, ½ Σ_{finite obs} y_unc (src/logden_normal.jl:34-39), into one constant.
// record layout: [tile of 32 sites][field f = 0..4nO)[lane]; fields S1[o], S2[o], y_o[o], w[o]
// w = exp(−y_unc) where the observation is finite and both drivers are finite, else 0
// (is_valid masking, src/DoubleMM/f_doubleMM.jl:111-137).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_preprocess(int nO, int64_t nb, const T* __restrict__ xP,
                                                    const T* __restrict__ y_o, const T* __restrict__ y_unc,
                                                    T* __restrict__ rec, double* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t ntiles = (nb + 31) / 32;
  double cst = 0.0, anomalies = 0.0;
  for (int64_t tile = gw; tile < ntiles; tile += nw) {
    const int64_t site = tile * 32 + lane;
    T* r = rec + (tile * 4 * nO) * 32 + lane;
    for (int o = 0; o < nO; o++) {
      T s1 = T(1), s2 = T(1), yo = T(0), w = T(0);
      if (site < nb) {
        const T a = xP[site * 2 * nO + o], b = xP[site * 2 * nO + nO + o];
        // prediction-only batches carry no observations (y_o == NULL): every entry counts as missing
        const T y = y_o ? y_o[site * nO + o] : T(NAN), u = y_unc ? y_unc[site * nO + o] : T(0);
        const bool valid = isfinite(a) && isfinite(b);
        const bool obs = !isnan(y);
        if (valid) { s1 = a; s2 = b; }
        if (obs && valid) { yo = y; w = exp_t(-u); cst += 0.5 * (double)u; }
        // reference: invalid driver with a finite observation gives a NaN loss (SURVEY C-2)
        if (obs && !valid) anomalies += 1.0;
      }
      r[(0 * nO + o) * 32] = s1;
      r[(1 * nO + o) * 32] = s2;
      r[(2 * nO + o) * 32] = yo;
      r[(3 * nO + o) * 32] = w;
    }
  }
  cst = warp_sum(cst);
  anomalies = warp_sum(anomalies);
  if (lane == 0) { part[2 * gw] = cst; part[2 * gw + 1] = anomalies; }
}

// Minibatch of a device-resident dataset (gdev_hybridproblem_dataloader, src/AbstractHybridProblem.jl:220-250, walked
// by the DataLoader of solve, src/HybridSolver.jl:259-260): batch site i is dataset site idx[i] (or first + i).  One
// pass gathers the columns into the batch arrays (xM for g, raw xP for the prediction path) and builds the tiled
// records exactly as k_preprocess does.  A site's columns are 4·nC … 8·nO contiguous bytes: every lane reads whole
// sectors of its own site; the record stores are coalesced.
template <typename T>
__device__ __forceinline__ void gather_preprocess_body(int nC, int nO, int64_t nb, const int64_t* __restrict__ idx,
                                                       int64_t first, const T* __restrict__ ds_xM,
                                                       const T* __restrict__ ds_xP, const T* __restrict__ ds_yo,
                                                       const T* __restrict__ ds_yu, T* __restrict__ xM_b,
                                                       T* __restrict__ xP_b, T* __restrict__ rec,
                                                       double* __restrict__ part,
                                                       const unsigned long long* __restrict__ ctr,
                                                       int64_t n_batches, int64_t* __restrict__ isites_out, int bid, int nblk) {
  // optimiser loops replayed as CUDA graphs: the minibatch is chosen by the device-side iteration counter ctr[1]
  if (ctr) first = (first + (int64_t)(ctr[1] % (unsigned long long)n_batches)) % n_batches * nb;
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)bid * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)nblk * blockDim.x) >> 5;
  const int64_t ntiles = (nb + 31) / 32;
  double cst = 0.0, anomalies = 0.0;
  for (int64_t tile = gw; tile < ntiles; tile += nw) {
    const int64_t site = tile * 32 + lane;
    const int64_t src = site < nb ? (idx ? idx[site] : first + site) : 0;
    T* r = rec + (tile * 4 * nO) * 32 + lane;
    if (site < nb) {
      for (int c = 0; c < nC; c++) xM_b[site * nC + c] = ds_xM[src * nC + c];
      if (isites_out) isites_out[site] = src;   // MeanVarSep reads its per-site parameters through this map
    }
    // Float32 rows of 8 observations (DoubleMM): a site's drivers, observations and uncertainties are 64 + 32 + 32 contiguous,
    // 16-byte aligned bytes -- fetched with 128-bit loads (whole sectors per lane), the raw drivers stored the same way
    T va[8], vb[8], vy[8], vu[8];
    bool vec = false;
    if constexpr (sizeof(T) == 4) {
      vec = nO == 8 && ((reinterpret_cast<uintptr_t>(ds_xP) | reinterpret_cast<uintptr_t>(xP_b) | reinterpret_cast<uintptr_t>(ds_yo) |
                         reinterpret_cast<uintptr_t>(ds_yu)) & 15) == 0;
      if (vec && site < nb) {
        const float4* xp = reinterpret_cast<const float4*>(ds_xP + src * 16);
        const float4 a0 = xp[0], a1 = xp[1], b0 = xp[2], b1 = xp[3];
        float4* xo = reinterpret_cast<float4*>(xP_b + site * 16);
        xo[0] = a0; xo[1] = a1; xo[2] = b0; xo[3] = b1;
        va[0] = a0.x; va[1] = a0.y; va[2] = a0.z; va[3] = a0.w; va[4] = a1.x; va[5] = a1.y; va[6] = a1.z; va[7] = a1.w;
        vb[0] = b0.x; vb[1] = b0.y; vb[2] = b0.z; vb[3] = b0.w; vb[4] = b1.x; vb[5] = b1.y; vb[6] = b1.z; vb[7] = b1.w;
        if (ds_yo) {
          const float4* yp = reinterpret_cast<const float4*>(ds_yo + src * 8);
          const float4 y0 = yp[0], y1 = yp[1];
          vy[0] = y0.x; vy[1] = y0.y; vy[2] = y0.z; vy[3] = y0.w; vy[4] = y1.x; vy[5] = y1.y; vy[6] = y1.z; vy[7] = y1.w;
        }
        if (ds_yu) {
          const float4* up = reinterpret_cast<const float4*>(ds_yu + src * 8);
          const float4 u0 = up[0], u1 = up[1];
          vu[0] = u0.x; vu[1] = u0.y; vu[2] = u0.z; vu[3] = u0.w; vu[4] = u1.x; vu[5] = u1.y; vu[6] = u1.z; vu[7] = u1.w;
        }
      }
    }
    // one observation of this lane's site into the tile's records (mask folded into the weight, as k_preprocess)
    auto emit = [&](int o, T a, T b, T y, T u) {
      T s1 = T(1), s2 = T(1), yo = T(0), w = T(0);
      if (site < nb) {
        const bool valid = isfinite(a) && isfinite(b);
        const bool obs = !isnan(y);
        if (valid) { s1 = a; s2 = b; }
        if (obs && valid) { yo = y; w = exp_t(-u); cst += 0.5 * (double)u; }
        if (obs && !valid) anomalies += 1.0;
      }
      r[(0 * nO + o) * 32] = s1;
      r[(1 * nO + o) * 32] = s2;
      r[(2 * nO + o) * 32] = yo;
      r[(3 * nO + o) * 32] = w;
    };
    if (vec) {
#pragma unroll
      for (int o = 0; o < 8; o++) emit(o, va[o], vb[o], ds_yo ? vy[o] : T(NAN), ds_yu ? vu[o] : T(0));
    } else {
      for (int o = 0; o < nO; o++) {
        T a = T(0), b = T(0), y = T(NAN), u = T(0);
        if (site < nb) {
          a = ds_xP[src * 2 * nO + o]; b = ds_xP[src * 2 * nO + nO + o];
          xP_b[site * 2 * nO + o] = a;
          xP_b[site * 2 * nO + nO + o] = b;
          if (ds_yo) y = ds_yo[src * nO + o];
          if (ds_yu) u = ds_yu[src * nO + o];
        }
        emit(o, a, b, y, u);
      }
    }
  }
  cst = warp_sum(cst);
  anomalies = warp_sum(anomalies);
  if (lane == 0) { part[2 * gw] = cst; part[2 * gw + 1] = anomalies; }
}

template <typename T>
__global__ void __launch_bounds__(256) k_gather_preprocess(int nC, int nO, int64_t nb, const int64_t* __restrict__ idx,
                                                           int64_t first, const T* __restrict__ ds_xM,
                                                           const T* __restrict__ ds_xP, const T* __restrict__ ds_yo,
                                                           const T* __restrict__ ds_yu, T* __restrict__ xM_b,
                                                           T* __restrict__ xP_b, T* __restrict__ rec,
                                                           double* __restrict__ part,
                                                           const unsigned long long* __restrict__ ctr,
                                                           int64_t n_batches, int64_t* __restrict__ isites_out) {
  gather_preprocess_body<T>(nC, nO, nb, idx, first, ds_xM, ds_xP, ds_yo, ds_yu, xM_b, xP_b, rec, part, ctr, n_batches,
                            isites_out, blockIdx.x, gridDim.x);
}

// the two constants of a registered batch, summed in a fixed order (run-to-run bit-identical), left on the device so
// that registering a batch needs no host round trip
__global__ void __launch_bounds__(256) k_batch_consts(const double* __restrict__ part, int nrows, double* __restrict__ consts) {
  __shared__ double sm[2][256];
  const int t = threadIdx.x;
  double c = 0.0, a = 0.0;
  for (int r = t; r < nrows; r += 256) { c += part[2 * r]; a += part[2 * r + 1]; }
  sm[0][t] = c; sm[1][t] = a;
  __syncthreads();
  for (int s_ = 128; s_ > 0; s_ >>= 1) {
    if (t < s_) { sm[0][t] += sm[0][t + s_]; sm[1][t] += sm[1][t + s_]; }
    __syncthreads();
  }
  if (t == 0) { consts[0] = sm[0][0]; consts[1] = sm[1][0]; }
}

// ---------------------------------------------------------------------------------------
// prologue: one block.  thread 0 builds the Cholesky factors; all threads do the draws of
// the global block (src/elbo.jl:647-657,664-670,496; src/cholesky.jl:156-174,262-313).
// pdraw = 4 arrays [nP][M]: rP, ζP, θP, gP = dθP/dζP.
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __noinline__ void build_U(int n, const int* blk_start, const int* rho_off, const T* rho, T (*U)[MAXP], T* nrm) {
  for (int k = 0; k < n; k++) {
    T ss = T(1);
    for (int i = 0; i < n; i++) U[i][k] = T(0);
    for (int i = blk_start[k]; i < k; i++) {
      const T v = rho[rho_off[k] + (i - blk_start[k])];
      U[i][k] = v;
      ss += v * v;
    }
    U[k][k] = T(1);
    nrm[k] = sqrt(ss);
    for (int i = blk_start[k]; i <= k; i++) U[i][k] /= nrm[k];
  }
}

template <typename T>
__device__ __forceinline__ void prologue_body(const Cfg<T>& cfg, const T* __restrict__ phi, const T* __restrict__ zP,
                                              int M, EvalConsts<T>* __restrict__ ec, T* __restrict__ pdraw) {
  __shared__ EvalConsts<T> s;
  const int nP = cfg.nP, nM = cfg.nM;
  // two serial sections on two warps: the M block's factor and the log σ² coefficients (thread 0), the P block's factor, σP
  // and μP (thread 32); the constants then go to global memory with one word per thread
  if (threadIdx.x == 0) {
    build_U<T>(nM, cfg.blkM_start, cfg.rhoM_off, phi + cfg.o_rhoM, s.UM, s.nrmM);
    _Pragma("unroll 1") for (int k = 0; k < nM; k++) {   // intercept and slope of log σ² (src/elbo.jl:652-653); MeanVarSep has neither
      s.coefA[k] = cfg.sepvar ? T(0) : phi[cfg.o_coef + 2 * k];
      s.coefB[k] = cfg.sepvar ? T(0) : phi[cfg.o_coef + 2 * k + 1];
    }
  } else if (threadIdx.x == 32) {
    build_U<T>(nP, cfg.blkP_start, cfg.rhoP_off, phi + cfg.o_rhoP, s.UP, s.nrmP);
    _Pragma("unroll 1") for (int j = 0; j < nP; j++) { s.sigP[j] = exp_t(phi[cfg.o_ls2P + j] / T(2)); s.muP[j] = phi[cfg.o_muP + j]; }
  }
  __syncthreads();
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&s);
    uint32_t* dst = reinterpret_cast<uint32_t*>(ec);
    _Pragma("unroll 1") for (int e = threadIdx.x; e < (int)(sizeof(EvalConsts<T>) / 4); e += blockDim.x) dst[e] = src[e];
  }
  T* rP = pdraw;
  T* zetaP = pdraw + (size_t)nP * M;
  T* thP = pdraw + (size_t)2 * nP * M;
  T* gP = pdraw + (size_t)3 * nP * M;
  _Pragma("unroll 1") for (int e = threadIdx.x; e < nP * M; e += blockDim.x) {
    const int j = e / M, m = e % M;
    T t = T(0);
    _Pragma("unroll 1") for (int jj = cfg.blkP_start[j]; jj <= j; jj++) t += s.UP[jj][j] * zP[m + (size_t)M * jj];
    const T r = s.sigP[j] * t;
    const T zeta = s.muP[j] + r;
    T th, dth, pv, dz, lj;
    transform_prior<T>(cfg.transP[j], HVI_PR_NONE, T(0), T(0), cfg.clamp_lo, cfg.clamp_hi, zeta, th, dth, pv, dz, lj);
    rP[e] = r; zetaP[e] = zeta; thP[e] = th; gP[e] = dth;
    // packed per-draw record read by the streaming kernel
    T* pd = pdraw + (((size_t)4 * (nP > 0 ? nP : 1) * M + 3) & ~(size_t)3) + (size_t)m * pd_stride(nP);
    pd[j] = th; pd[nP + j] = dth; pd[2 * nP + j] = zP[m + (size_t)M * j];
  }
}

template <typename T>
__global__ void k_prologue(const __grid_constant__ Cfg<T> cfg, const T* __restrict__ phi, const T* __restrict__ zP,
                           int M, EvalConsts<T>* __restrict__ ec, T* __restrict__ pdraw) {
  prologue_body<T>(cfg, phi, zP, M, ec, pdraw);
}

// ---------------------------------------------------------------------------------------
// g forward, CUDA-core version for narrow networks: one thread per site, weights and the
// activation column of every thread in shared memory ([neuron][thread], conflict-free).
// (src/ModelApplicator.jl:163-176; ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52)
// ---------------------------------------------------------------------------------------
constexpr int MLP_BS = 128;

template <typename T>
__device__ __forceinline__ void mlp_forward_block(const Cfg<T>& cfg, const T* __restrict__ sW, T* __restrict__ acts,
                                                  int t) {
  // acts: stack of activation rows, layer l input at row offset Σ_{l'<l} n_in(l'), [row][MLP_BS]
  int row_in = 0;
  for (int l = 0; l < cfg.nL; l++) {
    const int ni = cfg.l_in[l], no = cfg.l_out[l];
    const int row_out = row_in + ni;
    const T* W = sW + cfg.woff[l];
    for (int j = 0; j < no; j++) {
      T z = cfg.l_bias[l] ? sW[cfg.boff[l] + j] : T(0);
      for (int i = 0; i < ni; i++) z = fma(W[i * no + j], acts[(row_in + i) * MLP_BS + t], z);
      acts[(row_out + j) * MLP_BS + t] = act_f<T>(cfg.l_act[l], z);
    }
    row_in = row_out;
  }
}

// body: any block of >= MLP_BS threads (the first MLP_BS own the sites of a tile); (bid, nblk) = the block's place in the grid
template <typename T>
__device__ __forceinline__ void mlp_fwd_body(const Cfg<T>& cfg, const T* __restrict__ phi, const T* __restrict__ xM,
                                             int64_t nb, T* __restrict__ mu, unsigned char* smem_raw, int bid, int nblk) {
  T* sW = reinterpret_cast<T*>(smem_raw);
  T* acts = sW + ((cfg.nphig + 3) & ~3);
  const int t = threadIdx.x;
  for (int e = t; e < cfg.nphig; e += blockDim.x) sW[e] = phi[e];
  __syncthreads();
  const int64_t ntiles = (nb + MLP_BS - 1) / MLP_BS;
  const int out_row = cfg.sum_width - cfg.l_out[cfg.nL - 1];
  for (int64_t tile = bid; tile < ntiles; tile += nblk) {
    const int64_t site = tile * MLP_BS + t;
    if (t < MLP_BS && site < nb) {
      for (int i = 0; i < cfg.nC; i++) acts[i * MLP_BS + t] = xM[site * cfg.nC + i];
      for (int c = 0; c < cfg.npc; c++) acts[(cfg.nC + c) * MLP_BS + t] = phi[cfg.o_muP + cfg.pbm_covar_idx[c]];
      mlp_forward_block<T>(cfg, sW, acts, t);
      for (int k = 0; k < cfg.nM; k++) {
        const T p = acts[(out_row + k) * MLP_BS + t];
        mu[site * cfg.nM + k] = (cfg.scale_kind == HVI_SCALE_RANGE) ? p * cfg.scale_b[k] + cfg.scale_a[k]
                                                                  : cfg.scale_a[k] + cfg.scale_b[k] * ndtri_t(p);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(MLP_BS) k_mlp_fwd(const __grid_constant__ Cfg<T> cfg, const T* __restrict__ phi,
                                                    const T* __restrict__ xM, int64_t nb, T* __restrict__ mu) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  mlp_fwd_body<T>(cfg, phi, xM, nb, mu, smem_raw, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------------------------------
// g backward: recompute the forward per tile of 128 sites, back-propagate δ, then form the
// weight gradient as (δ stack)·(activation stack)ᵀ over the tile with each thread owning
// fixed gradient entries (deterministic; no atomics).  Per-block double partial sums.
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void mlp_bwd_body(const Cfg<T>& cfg, const T* __restrict__ phi, const T* __restrict__ xM,
                                             const T* __restrict__ mubar, int64_t nb, double* __restrict__ part,
                                             unsigned char* smem_raw, int bid, int nblk) {
  const int NT = blockDim.x;
  double* gacc = reinterpret_cast<double*>(smem_raw);                 // [nphig]
  int2* tab = reinterpret_cast<int2*>(gacc + cfg.nphig);              // [nphig] (delta row, act row | -1)
  T* sW = reinterpret_cast<T*>(tab + cfg.nphig);                      // [nphig]
  T* acts = sW + ((cfg.nphig + 3) & ~3);                              // [sum_width][BS]
  T* dl = acts + (size_t)cfg.sum_width * MLP_BS;                      // [sum_out][BS]
  const int t = threadIdx.x;
  for (int e = t; e < cfg.nphig; e += NT) { sW[e] = phi[e]; gacc[e] = 0.0; }
  {  // gradient entry e ↔ (δ row, activation row)
    int row_in = 0, drow = 0;
    for (int l = 0; l < cfg.nL; l++) {
      const int ni = cfg.l_in[l], no = cfg.l_out[l];
      for (int e = t; e < ni * no; e += NT) tab[cfg.woff[l] + e] = make_int2(drow + e % no, row_in + e / no);
      if (cfg.l_bias[l])
        for (int e = t; e < no; e += NT) tab[cfg.boff[l] + e] = make_int2(drow + e, -1);
      row_in += ni; drow += no;
    }
  }
  __syncthreads();
  const int64_t ntiles = (nb + MLP_BS - 1) / MLP_BS;
  const int out_row = cfg.sum_width - cfg.l_out[cfg.nL - 1];
  for (int64_t tile = bid; tile < ntiles; tile += nblk) {
    const int64_t site = tile * MLP_BS + t;
    const bool active = site < nb;
    if (t < MLP_BS) {
    // forward
    for (int i = 0; i < cfg.nC; i++) acts[i * MLP_BS + t] = active ? xM[site * cfg.nC + i] : T(0);
    for (int c = 0; c < cfg.npc; c++) acts[(cfg.nC + c) * MLP_BS + t] = phi[cfg.o_muP + cfg.pbm_covar_idx[c]];
    mlp_forward_block<T>(cfg, sW, acts, t);
    // output delta: μ̄_ζ · σ_k · √(2π) e^{q²/2} · p(1−p)
    int drow = cfg.sum_out - cfg.l_out[cfg.nL - 1];
    for (int k = 0; k < cfg.l_out[cfg.nL - 1]; k++) {
      T d = T(0);
      if (active && k < cfg.nM) {
        const T p = acts[(out_row + k) * MLP_BS + t];
        const T mb = mubar[site * cfg.nM + k];
        if (cfg.scale_kind == HVI_SCALE_RANGE) d = mb * cfg.scale_b[k];
        else { const T q = ndtri_t(p); d = mb * cfg.scale_b[k] * T(2.50662827463100050242) * exp_t(T(0.5) * q * q); }
        d *= act_d<T>(cfg.l_act[cfg.nL - 1], p);
      }
      dl[(drow + k) * MLP_BS + t] = d;
    }
    // hidden deltas
    int row_out = out_row;
    for (int l = cfg.nL - 1; l >= 1; l--) {
      const int ni = cfg.l_in[l], no = cfg.l_out[l];
      const int row_in = row_out - ni;         // activations feeding layer l
      const int drow_prev = drow - ni;         // delta rows of layer l-1 (n_out(l-1) == ni)
      const T* W = sW + cfg.woff[l];
      for (int i = 0; i < ni; i++) {
        T s = T(0);
        for (int j = 0; j < no; j++) s = fma(W[i * no + j], dl[(drow + j) * MLP_BS + t], s);
        dl[(drow_prev + i) * MLP_BS + t] = s * act_d<T>(cfg.l_act[l - 1], acts[(row_in + i) * MLP_BS + t]);
      }
      row_out = row_in; drow = drow_prev;
    }
    }
    __syncthreads();
    // weight gradient: entry e owned by thread e % BS; site index rotated per thread so that the
    // 32 lanes of a warp hit 32 different banks
    for (int e = t; e < cfg.nphig; e += NT) {
      const int2 rc = tab[e];
      const T* dr = dl + (size_t)rc.x * MLP_BS;
      T acc = T(0);
      if (rc.y >= 0) {
        const T* ar = acts + (size_t)rc.y * MLP_BS;
        for (int s0 = 0; s0 < MLP_BS; s0++) { const int s = (s0 + t) & (MLP_BS - 1); acc = fma(dr[s], ar[s], acc); }
      } else {
        for (int s0 = 0; s0 < MLP_BS; s0++) { const int s = (s0 + t) & (MLP_BS - 1); acc += dr[s]; }
      }
      gacc[e] += (double)acc;
    }
    __syncthreads();
  }
  for (int e = t; e < cfg.nphig; e += NT) part[(size_t)bid * cfg.nphig + e] = gacc[e];
}

template <typename T>
__global__ void __launch_bounds__(MLP_BS) k_mlp_bwd(const __grid_constant__ Cfg<T> cfg, const T* __restrict__ phi,
                                                    const T* __restrict__ xM, const T* __restrict__ mubar, int64_t nb,
                                                    double* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  mlp_bwd_body<T>(cfg, phi, xM, mubar, nb, part, smem_raw, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------------------------------
// Register-tiled applicator for the reference's canonical network
//   n_in → H1 tanh → H2 tanh → n_out logistic (no bias)   (ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52)
// with every dimension a compile-time constant.  Forward: one thread per site, activations in
// registers, weights read from shared memory as broadcast LDS.128 (4 outputs per load).
// Backward: phase A (thread per site) recomputes the forward, back-propagates δ and stores the
// activation rows [x;1], [h1;1], [h2;1] and the δ rows of the tile's 128 sites in shared
// memory; phase B forms the weight gradient Σ_sites δ·[a;1]ᵀ as 4×4 register tiles, one or two
// tiles per lane, 2 LDS.128 per 16 FMA; the four warps split the tile's sites.  Row strides
// are an odd number of 16-byte units so that the per-site STS.128/LDS.128 are conflict-free.
// ---------------------------------------------------------------------------------------
__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }
__host__ __device__ constexpr int oddpad(int n) { return (((n + 3) / 4) | 1) * 4; }

template <typename T> struct V4 { T v[4]; };
__device__ __forceinline__ V4<float> ld4(const float* p) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  return {{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ V4<double> ld4(const double* p) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  return {{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
  *reinterpret_cast<double2*>(p + 2) = make_double2(c, d);
}

// tanh accurate to ≈1e-6 relative from 2 MUFU + a short odd polynomial near 0
__device__ __forceinline__ float tanh_fast(float x) {
  const float ax = fabsf(x), x2 = x * x;
  const float p = x * fmaf(x2, fmaf(x2, fmaf(x2, -17.0f / 315.0f, 2.0f / 15.0f), -1.0f / 3.0f), 1.0f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.885390081777927f));
  const float r = copysignf(1.0f - 2.0f * rcp_fast(e + 1.0f), x);
  return ax < 0.25f ? p : r;
}
__device__ __forceinline__ double tanh_fast(double x) { return tanh(x); }
__device__ __forceinline__ float logistic_fast(float z) { return rcp_fast(1.0f + exp_t(-z)); }
__device__ __forceinline__ double logistic_fast(double z) { return 1.0 / (1.0 + exp(-z)); }

template <typename T, int NI, int H1, int H2, int NOUT>
struct Mlp3 {
  static constexpr int H1P = pad4(H1), H2P = pad4(H2), OP = pad4(NOUT);
  // padded weights in shared memory: W_l as [i][out padded], biases padded with zeros
  // (rows padded to a multiple of 4 inputs, zero-filled)
  static constexpr int oW1 = 0, ob1 = oW1 + pad4(NI) * H1P, oW2 = ob1 + H1P, ob2 = oW2 + pad4(H1) * H2P,
                       oW3 = ob2 + H2P, nW = oW3 + pad4(H2) * OP;
  // per-site rows of the backward tile: [x;1] [h1;1] [h2;1] δ1 δ2 δ3
  static constexpr int sA0 = oddpad(NI + 1), sA1 = oddpad(H1 + 1), sA2 = oddpad(H2 + 1), sD1 = oddpad(H1),
                       sD2 = oddpad(H2), sD3 = oddpad(NOUT);
  static constexpr int oA0 = 0, oA1 = oA0 + sA0, oA2 = oA1 + sA1, oD1 = oA2 + sA2, oD2 = oD1 + sD1, oD3 = oD2 + sD2,
                       ROW = oddpad(oD3 + sD3);   // odd number of 16-byte units between consecutive sites
  // 4×4 gradient tiles: layer l has ceil(out/4) × ceil((in+1)/4)
  static constexpr int T1j = H1P / 4, T1i = pad4(NI + 1) / 4, T2j = H2P / 4, T2i = pad4(H1 + 1) / 4, T3j = OP / 4,
                       T3i = pad4(H2 + 1) / 4;
  static constexpr int NT1 = T1j * T1i, NT2 = T2j * T2i, NT3 = T3j * T3i, NT = NT1 + NT2 + NT3;
  static constexpr int TPL = (NT + 31) / 32;   // tiles per lane

  __device__ static void load_weights(const Cfg<T>& cfg, const T* __restrict__ phi, T* __restrict__ sW, int t, int nthr) {
    for (int e = t; e < nW; e += nthr) sW[e] = T(0);
    __syncthreads();
    for (int e = t; e < NI * H1; e += nthr) sW[oW1 + (e / H1) * H1P + e % H1] = phi[cfg.woff[0] + e];  // [i][j]
    for (int e = t; e < H1; e += nthr) sW[ob1 + e] = phi[cfg.boff[0] + e];
    for (int e = t; e < H1 * H2; e += nthr) sW[oW2 + (e / H2) * H2P + e % H2] = phi[cfg.woff[1] + e];
    for (int e = t; e < H2; e += nthr) sW[ob2 + e] = phi[cfg.boff[1] + e];
    for (int e = t; e < H2 * NOUT; e += nthr) sW[oW3 + (e / NOUT) * OP + e % NOUT] = phi[cfg.woff[2] + e];
    __syncthreads();
  }

  template <int NIN, int NOP_, bool BIAS>
  __device__ __forceinline__ static void dense(const T* __restrict__ W, const T* __restrict__ b, const T* x, T* y) {
#pragma unroll
    for (int j4 = 0; j4 < NOP_ / 4; j4++) {
      if (BIAS) { const V4<T> bb = ld4(b + 4 * j4); y[4 * j4] = bb.v[0]; y[4 * j4 + 1] = bb.v[1]; y[4 * j4 + 2] = bb.v[2]; y[4 * j4 + 3] = bb.v[3]; }
      else { y[4 * j4] = y[4 * j4 + 1] = y[4 * j4 + 2] = y[4 * j4 + 3] = T(0); }
    }
#pragma unroll
    for (int i = 0; i < NIN; i++) {
#pragma unroll
      for (int j4 = 0; j4 < NOP_ / 4; j4++) {
        const V4<T> w = ld4(W + i * NOP_ + 4 * j4);
#pragma unroll
        for (int c = 0; c < 4; c++) y[4 * j4 + c] = fma(w.v[c], x[i], y[4 * j4 + c]);
      }
    }
  }

  // dprev[i] = Σ_j W[i][j]·d[j]
  template <int NIN, int NOP_>
  __device__ __forceinline__ static void dense_T(const T* __restrict__ W, const T* d, T* dprev) {
#pragma unroll
    for (int i = 0; i < NIN; i++) {
      T s = T(0);
#pragma unroll
      for (int j4 = 0; j4 < NOP_ / 4; j4++) {
        const V4<T> w = ld4(W + i * NOP_ + 4 * j4);
#pragma unroll
        for (int c = 0; c < 4; c++) s = fma(w.v[c], d[4 * j4 + c], s);
      }
      dprev[i] = s;
    }
  }

  // shared-memory-row variants used by the backward kernel: inputs are read four at a time from
  // this thread's site row, outputs stay in registers; loops over the inputs are not unrolled, so
  // the register footprint is the output vector only
  template <int NIN, int NOP_, bool BIAS>
  __device__ __forceinline__ static void dense_row(const T* __restrict__ W, const T* __restrict__ b, const T* in_row, T* y) {
#pragma unroll
    for (int j4 = 0; j4 < NOP_ / 4; j4++) {
      if (BIAS) { const V4<T> bb = ld4(b + 4 * j4); y[4 * j4] = bb.v[0]; y[4 * j4 + 1] = bb.v[1]; y[4 * j4 + 2] = bb.v[2]; y[4 * j4 + 3] = bb.v[3]; }
      else { y[4 * j4] = y[4 * j4 + 1] = y[4 * j4 + 2] = y[4 * j4 + 3] = T(0); }
    }
#pragma unroll 1
    for (int i4 = 0; i4 < pad4(NIN) / 4; i4++) {
      const V4<T> x4 = ld4(in_row + 4 * i4);
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int j4 = 0; j4 < NOP_ / 4; j4++) {
          const V4<T> w = ld4(W + (4 * i4 + c) * NOP_ + 4 * j4);
#pragma unroll
          for (int cc = 0; cc < 4; cc++) y[4 * j4 + cc] = fma(w.v[cc], x4.v[c], y[4 * j4 + cc]);
        }
    }
  }
  // δ_prev[i] = (Σ_j W[i][j]·d[j])·(1 − a[i]²), written to the site row (a = tanh activations)
  template <int NIN, int NOP_>
  __device__ __forceinline__ static void dense_T_row(const T* __restrict__ W, const T* d, const T* act_row, T* dprev_row) {
#pragma unroll 1
    for (int i4 = 0; i4 < pad4(NIN) / 4; i4++) {
      const V4<T> a4 = ld4(act_row + 4 * i4);
      T s[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        s[c] = T(0);
#pragma unroll
        for (int j4 = 0; j4 < NOP_ / 4; j4++) {
          const V4<T> w = ld4(W + (4 * i4 + c) * NOP_ + 4 * j4);
#pragma unroll
          for (int cc = 0; cc < 4; cc++) s[c] = fma(w.v[cc], d[4 * j4 + cc], s[c]);
        }
        s[c] = (4 * i4 + c < NIN) ? s[c] * (T(1) - a4.v[c] * a4.v[c]) : T(0);
      }
      st4(dprev_row + 4 * i4, s[0], s[1], s[2], s[3]);
    }
  }

  __device__ __forceinline__ static void forward(const T* __restrict__ sW, const T* x, T* h1, T* h2, T* p) {
    dense<NI, H1P, true>(sW + oW1, sW + ob1, x, h1);
#pragma unroll
    for (int j = 0; j < H1P; j++) h1[j] = j < H1 ? tanh_fast(h1[j]) : T(0);
    asm volatile("" ::: "memory");
    dense<H1, H2P, true>(sW + oW2, sW + ob2, h1, h2);
#pragma unroll
    for (int j = 0; j < H2P; j++) h2[j] = j < H2 ? tanh_fast(h2[j]) : T(0);
    dense<H2, OP, false>(sW + oW3, nullptr, h2, p);
#pragma unroll
    for (int j = 0; j < OP; j++) p[j] = j < NOUT ? logistic_fast(p[j]) : T(0);
  }
};

constexpr int MLP3_BS = 128;

// Columns of the applicator.  Site mode (M_units == 0): one column per site, the pbm covariates
// (if any) are μP[idx] -- pass 1 of generate_ζ (src/elbo.jl:491-492).  Unit mode (M_units = n_MC):
// one column per (draw m, site), the pbm covariates are ζP[idx, m] -- pass 2 (src/elbo.jl:508-515);
// results go to / cotangents come from [m][k][site] arrays (site fastest: coalesced for the
// thread-per-site kernels on both sides).
struct MlpCols {
  int M_units;       // 0: site mode
  int64_t nb;
};

template <typename T, int NI, int H1, int H2, int NOUT>
__global__ void __launch_bounds__(MLP3_BS) k_mlp3_fwd(const __grid_constant__ Cfg<T> cfg, const T* __restrict__ phi,
                                                      const T* __restrict__ xM, const T* __restrict__ zetaP,
                                                      MlpCols cols, T* __restrict__ mu) {
  using N = Mlp3<T, NI, H1, H2, NOUT>;
  __shared__ __align__(16) T sW[N::nW];
  const int t = threadIdx.x;
  const int64_t nb = cols.nb;
  N::load_weights(cfg, phi, sW, t, MLP3_BS);
  const int64_t ntiles = (nb + MLP3_BS - 1) / MLP3_BS;
  const int64_t ntot = ntiles * (cols.M_units > 0 ? cols.M_units : 1);
  for (int64_t g = blockIdx.x; g < ntot; g += gridDim.x) {
    const int m = (int)(g / ntiles);
    const int64_t site = (g % ntiles) * MLP3_BS + t;
    if (site >= nb) continue;
    T x[NI], h1[N::H1P], h2[N::H2P], p[N::OP];
#pragma unroll
    for (int i = 0; i < NI; i++) {
      if (i < cfg.nC) x[i] = xM[site * cfg.nC + i];
      else {
        const int j = cfg.pbm_covar_idx[i - cfg.nC];
        x[i] = cols.M_units > 0 ? zetaP[(size_t)j * cols.M_units + m] : phi[cfg.o_muP + j];
      }
    }
    N::forward(sW, x, h1, h2, p);
#pragma unroll
    for (int k = 0; k < NOUT; k++)
      if (k < cfg.nM) {
        const T v = (cfg.scale_kind == HVI_SCALE_RANGE) ? p[k] * cfg.scale_b[k] + cfg.scale_a[k]
                                                        : cfg.scale_a[k] + cfg.scale_b[k] * ndtri_t(p[k]);
        if (cols.M_units > 0) mu[((size_t)m * cfg.nM + k) * nb + site] = v;
        else mu[site * cfg.nM + k] = v;
      }
  }
}

template <typename T, int NI, int H1, int H2, int NOUT>
__global__ void __launch_bounds__(MLP3_BS, 3) k_mlp3_bwd(const __grid_constant__ Cfg<T> cfg, const T* __restrict__ phi,
                                                      const T* __restrict__ xM, const T* __restrict__ zetaP,
                                                      const T* __restrict__ mubar, MlpCols cols,
                                                      double* __restrict__ part, double* __restrict__ part_x) {
  using N = Mlp3<T, NI, H1, H2, NOUT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sW = reinterpret_cast<T*>(smem_raw);
  T* rows = sW + pad4(N::nW);                       // [MLP3_BS][ROW]
  T* sacc = rows + (size_t)MLP3_BS * N::ROW;        // [4 warps][NT tiles][16]: gradient tiles between site tiles
  // cotangent of the pbm covariate inputs, summed over the sites: [max(M_units,1)][npc] doubles
  double* xacc = reinterpret_cast<double*>(sacc + (size_t)4 * N::NT * 16);
  T* xred = reinterpret_cast<T*>(xacc + (size_t)(cols.M_units > 0 ? cols.M_units : 1) * cfg.npc);  // [4 warps][npc]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t nb = cols.nb;
  const int nxa = (cols.M_units > 0 ? cols.M_units : 1) * cfg.npc;
  N::load_weights(cfg, phi, sW, t, MLP3_BS);
  for (int e = t; e < 4 * N::NT * 16; e += MLP3_BS) sacc[e] = T(0);
  for (int e = t; e < nxa; e += MLP3_BS) xacc[e] = 0.0;
  // this lane's gradient tiles: offsets of the δ quad and of the activation quad inside a site row
  int doff[N::TPL], aoff[N::TPL];
#pragma unroll
  for (int q = 0; q < N::TPL; q++) {
    int tl = lane + 32 * q;
    if (tl >= N::NT) tl = 0;   // idle slot: recompute tile 0, result discarded
    if (tl < N::NT1) { doff[q] = N::oD1 + 4 * (tl % N::T1j); aoff[q] = N::oA0 + 4 * (tl / N::T1j); }
    else if (tl < N::NT1 + N::NT2) { const int u = tl - N::NT1; doff[q] = N::oD2 + 4 * (u % N::T2j); aoff[q] = N::oA1 + 4 * (u / N::T2j); }
    else { const int u = tl - N::NT1 - N::NT2; doff[q] = N::oD3 + 4 * (u % N::T3j); aoff[q] = N::oA2 + 4 * (u / N::T3j); }
  }
  const int64_t ntiles = (nb + MLP3_BS - 1) / MLP3_BS;
  const int64_t ntot = ntiles * (cols.M_units > 0 ? cols.M_units : 1);
  for (int64_t g = blockIdx.x; g < ntot; g += gridDim.x) {
    const int m = (int)(g / ntiles);
    const int64_t site = (g % ntiles) * MLP3_BS + t;
    const bool active = site < nb;
    {  // phase A: this thread's site; all vectors live in the site row r
      T* r = rows + (size_t)t * N::ROW;
      {
        T x[pad4(NI + 1)];
#pragma unroll
        for (int i = 0; i < pad4(NI + 1); i++) x[i] = T(0);
#pragma unroll
        for (int i = 0; i < NI; i++) {
          if (!active) x[i] = T(0);
          else if (i < cfg.nC) x[i] = xM[site * cfg.nC + i];
          else {
            const int j = cfg.pbm_covar_idx[i - cfg.nC];
            x[i] = cols.M_units > 0 ? zetaP[(size_t)j * cols.M_units + m] : phi[cfg.o_muP + j];
          }
        }
#pragma unroll
        for (int i = 0; i < pad4(NI + 1); i += 4) st4(r + N::oA0 + i, x[i], x[i + 1], x[i + 2], x[i + 3]);
      }
      {
        T h[pad4(H1 + 1)];
#pragma unroll
        for (int i = N::H1P; i < pad4(H1 + 1); i++) h[i] = T(0);
        N::template dense_row<NI, N::H1P, true>(sW + N::oW1, sW + N::ob1, r + N::oA0, h);
#pragma unroll
        for (int j = 0; j < N::H1P; j++) h[j] = j < H1 ? tanh_fast(h[j]) : T(0);
#pragma unroll
        for (int i = 0; i < pad4(H1 + 1); i += 4) st4(r + N::oA1 + i, h[i], h[i + 1], h[i + 2], h[i + 3]);
      }
      {
        T h[pad4(H2 + 1)];
#pragma unroll
        for (int i = N::H2P; i < pad4(H2 + 1); i++) h[i] = T(0);
        N::template dense_row<H1, N::H2P, true>(sW + N::oW2, sW + N::ob2, r + N::oA1, h);
#pragma unroll
        for (int j = 0; j < N::H2P; j++) h[j] = j < H2 ? tanh_fast(h[j]) : T(0);
#pragma unroll
        for (int i = 0; i < pad4(H2 + 1); i += 4) st4(r + N::oA2 + i, h[i], h[i + 1], h[i + 2], h[i + 3]);
      }
      T d3[N::OP];
      {
        T p[N::OP];
        N::template dense_row<H2, N::OP, false>(sW + N::oW3, nullptr, r + N::oA2, p);
#pragma unroll
        for (int k = 0; k < N::OP; k++) {
          T d = T(0);
          if (k < NOUT && k < cfg.nM && active) {
            const T pk = logistic_fast(p[k]);
            const T mb = cols.M_units > 0 ? mubar[((size_t)m * cfg.nM + k) * nb + site] : mubar[site * cfg.nM + k];
            if (cfg.scale_kind == HVI_SCALE_RANGE) d = mb * cfg.scale_b[k];
            else { const T q = ndtri_t(pk); d = mb * cfg.scale_b[k] * T(2.50662827463100050242) * exp_t(T(0.5) * q * q); }
            d *= pk * (T(1) - pk);
          }
          d3[k] = d;
        }
#pragma unroll
        for (int i = 0; i < N::OP; i += 4) st4(r + N::oD3 + i, d3[i], d3[i + 1], d3[i + 2], d3[i + 3]);
      }
      // δ2 = (W3ᵀ δ3)·(1 − h2²) → row; δ1 = (W2ᵀ δ2)·(1 − h1²) → row
      N::template dense_T_row<H2, N::OP>(sW + N::oW3, d3, r + N::oA2, r + N::oD2);
      {
        T d2[N::H2P];
#pragma unroll
        for (int i4 = 0; i4 < N::H2P / 4; i4++) {
          const V4<T> v = ld4(r + N::oD2 + 4 * i4);
#pragma unroll
          for (int c = 0; c < 4; c++) d2[4 * i4 + c] = v.v[c];
        }
        N::template dense_T_row<H1, N::H2P>(sW + N::oW2, d2, r + N::oA1, r + N::oD1);
      }
      // cotangent of the pbm covariate inputs: x̄[nC + c] = Σ_j W1[nC + c][j]·δ1[j], summed over
      // the tile's sites in a fixed order (warp shuffle tree, then the 4 warps in order)
      for (int c = 0; c < cfg.npc; c++) {
        T xb = T(0);
        const T* w1 = sW + N::oW1 + (cfg.nC + c) * N::H1P;
#pragma unroll
        for (int j4 = 0; j4 < N::H1P / 4; j4++) {
          const V4<T> w = ld4(w1 + 4 * j4), d = ld4(r + N::oD1 + 4 * j4);
#pragma unroll
          for (int cc = 0; cc < 4; cc++) xb = fma(w.v[cc], d.v[cc], xb);
        }
        xb = warp_sum(xb);
        if (lane == 0) xred[warp * cfg.npc + c] = xb;
      }
      // the constant 1 that multiplies the bias
      r[N::oA0 + NI] = T(1); r[N::oA1 + H1] = T(1); r[N::oA2 + H2] = T(1);
    }
    __syncthreads();
    if (t < cfg.npc) {
      double sx = 0.0;
      for (int w = 0; w < MLP3_BS / 32; w++) sx += (double)xred[w * cfg.npc + t];
      xacc[(cols.M_units > 0 ? m : 0) * cfg.npc + t] += sx;
    }
    // phase B: this warp's 32 sites of the tile
    {
      T acc[N::TPL][16];
#pragma unroll
      for (int q = 0; q < N::TPL; q++) {
        const int tl = lane + 32 * q;
        const T* sa = sacc + ((size_t)warp * N::NT + (tl < N::NT ? tl : 0)) * 16;
#pragma unroll
        for (int e4 = 0; e4 < 4; e4++) {
          const V4<T> v = ld4(sa + 4 * e4);
#pragma unroll
          for (int c = 0; c < 4; c++) acc[q][4 * e4 + c] = v.v[c];
        }
      }
#pragma unroll 4
      for (int s = 0; s < 32; s++) {
        const T* r = rows + (size_t)(warp * 32 + s) * N::ROW;
#pragma unroll
        for (int q = 0; q < N::TPL; q++) {
          const V4<T> d = ld4(r + doff[q]), a = ld4(r + aoff[q]);
#pragma unroll
          for (int jj = 0; jj < 4; jj++)
#pragma unroll
            for (int ii = 0; ii < 4; ii++) acc[q][4 * jj + ii] = fma(d.v[jj], a.v[ii], acc[q][4 * jj + ii]);
        }
      }
      __syncwarp();   // idle slots read tile 0 above: all reads before any write
#pragma unroll
      for (int q = 0; q < N::TPL; q++) {
        const int tl = lane + 32 * q;
        if (tl < N::NT) {
          T* sa = sacc + ((size_t)warp * N::NT + tl) * 16;
#pragma unroll
          for (int e4 = 0; e4 < 4; e4++) st4(sa + 4 * e4, acc[q][4 * e4], acc[q][4 * e4 + 1], acc[q][4 * e4 + 2], acc[q][4 * e4 + 3]);
        }
      }
    }
    __syncthreads();
  }
  // block reduction over the 4 warps (fixed order), then scatter to the layout of ϕg
  double* out = part + (size_t)blockIdx.x * cfg.nphig;
  for (int e = t; e < N::NT * 16; e += MLP3_BS) {
    const int tl = e / 16, jj = (e % 16) / 4, ii = e % 4;
    double s = 0;
    for (int w = 0; w < MLP3_BS / 32; w++)
      s += (double)sacc[((size_t)w * N::NT + tl) * 16 + (e % 16)];
    int l, j, i, nin, nout;
    if (tl < N::NT1) { l = 0; j = 4 * (tl % N::T1j) + jj; i = 4 * (tl / N::T1j) + ii; nin = NI; nout = H1; }
    else if (tl < N::NT1 + N::NT2) { const int u = tl - N::NT1; l = 1; j = 4 * (u % N::T2j) + jj; i = 4 * (u / N::T2j) + ii; nin = H1; nout = H2; }
    else { const int u = tl - N::NT1 - N::NT2; l = 2; j = 4 * (u % N::T3j) + jj; i = 4 * (u / N::T3j) + ii; nin = H2; nout = NOUT; }
    if (j < nout) {
      if (i < nin) out[cfg.woff[l] + i * nout + j] = s;
      else if (i == nin && cfg.l_bias[l]) out[cfg.boff[l] + j] = s;
    }
  }
  if (part_x)
    for (int e = t; e < nxa; e += MLP3_BS) part_x[(size_t)blockIdx.x * nxa + e] = xacc[e];
}

#define HVI_MLP3_CASES(X) X(5, 20, 20, 2) X(6, 24, 24, 2)

template <typename T>
static bool mlp3_match(const Cfg<T>& c, int ni, int h1, int h2, int no) {
  return c.nL == 3 && c.l_in[0] == ni && c.l_out[0] == h1 && c.l_out[1] == h2 && c.l_out[2] == no &&
         c.l_act[0] == HVI_ACT_TANH && c.l_act[1] == HVI_ACT_TANH && c.l_act[2] == HVI_ACT_LOGISTIC && c.l_bias[0] &&
         c.l_bias[1] && !c.l_bias[2];
}

// ---------------------------------------------------------------------------------------
// finalize: one block.  Fixed-order reduction of all partial sums, global-block terms,
// Cholesky adjoint, scatter into the layout of ϕ.
// out[0..nphi) gradient, out[nphi..nphi+6) components, out[nphi+6] nL, out[nphi+7] non-finite count
// ---------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;

__device__ __noinline__ double block_sum_fixed(double v, double* sred) {
  // fixed shape (butterfly inside each warp, then the warps' totals in warp order): deterministic run to run
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  constexpr int NW = FIN_THREADS / 32;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();   // sred may still be read by an earlier call's last step
  if (lane == 0) sred[warp] = v;
  __syncthreads();
  double sum = 0.0;
  for (int w = 0; w < NW; w++) sum += sred[w];
  __syncthreads();
  return sum;
}

// the same for n values per thread at once: sums[i] = Σ_threads v[i], bit-identical to n separate calls; sred holds at
// least 8·n doubles, sums is shared memory
__device__ __noinline__ void block_sum_fixed_n(const double* v, int n, double* sred, double* sums) {
  // butterfly inside each warp, then the warps' totals in warp order: two barriers instead of one per tree level
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  constexpr int NW = FIN_THREADS / 32;
  for (int i = 0; i < n; i++) {
    double x = v[i];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sred[i * NW + warp] = x;
  }
  __syncthreads();
  if (t < n) {
    double sum = 0.0;
    for (int w = 0; w < NW; w++) sum += sred[t * NW + w];
    sums[t] = sum;
  }
  __syncthreads();
}

template <typename T>
__device__ __noinline__ void build_U_adj(int n, const int* blk_start, const int* rho_off, const T (*U)[MAXP], const T* nrm,
                            const double (*Ubar)[MAXP], double* rho_bar) {
  for (int k = 0; k < n; k++) {
    double dot = 0;
    for (int i = blk_start[k]; i <= k; i++) dot += (double)U[i][k] * Ubar[i][k];
    for (int i = blk_start[k]; i < k; i++)
      rho_bar[rho_off[k] + (i - blk_start[k])] += (Ubar[i][k] - (double)U[i][k] * dot) / (double)nrm[k];
  }
}

// column sums of a [nrows][ncols] array of per-block partials, 32 columns per block, fixed order
__global__ void __launch_bounds__(256) k_colsum(const double* __restrict__ part, int nrows, int ncols,
                                                double* __restrict__ out) {
  __shared__ double sm[8][33];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + c;
  double s0 = 0;
  if (e < ncols)
    for (int r = g; r < nrows; r += 8) s0 += part[(size_t)r * ncols + e];
  sm[g][c] = s0;
  __syncthreads();
  if (g == 0 && e < ncols) {
    double s = 0;
    for (int gg = 0; gg < 8; gg++) s += sm[gg][c];
    out[e] = s;
  }
}

// stage 1 of the reduction: many blocks, coalesced, fixed summation order.
//   blocks [0, gridDim.x−1): 32 entries of ∇ϕg each, 8 row groups per entry
//   last block            : the streaming kernel's per-warp partial sums, 32 columns at a time
// red[0..NACC) = column sums; red[128 + b] = number of non-finite entries seen by block b
constexpr int RED_THREADS = 256;

// Σ_{r = g, g+8, g+16, …} col[r·ld] in a fixed order with eight loads in flight (the reduction is latency bound:
// a thread that adds one row after the other waits a DRAM/L2 round trip per row)
__device__ __forceinline__ double strided_sum8(const double* __restrict__ col, int ld, int g, int nrows) {
  double s[8];
#pragma unroll
  for (int u = 0; u < 8; u++) s[u] = 0.0;
  int r = g;
  for (; r + 56 < nrows; r += 64) {
#pragma unroll
    for (int u = 0; u < 8; u++) s[u] += col[(size_t)(r + 8 * u) * ld];
  }
#pragma unroll
  for (int u = 0; u < 8; u++)
    if (r + 8 * u < nrows) s[u] += col[(size_t)(r + 8 * u) * ld];
  return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

// body for one (virtual) block `bid` of `ngrid` blocks of RED_THREADS threads; ends with a barrier so that a caller can
// loop over the virtual blocks
template <typename T>
__device__ __forceinline__ void reduce_body(int nphig, const double* __restrict__ part_mlp, int nblk,
                                            const double* __restrict__ part_stream, int nrows, int NACC,
                                            double* __restrict__ out, T* __restrict__ grad_T, double* __restrict__ red,
                                            int bid, int ngrid) {
  __shared__ double sm[8][33];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  if (bid < ngrid - 1) {
    const int e = bid * 32 + c;
    double s0 = 0, s1 = 0;
    if (e < nphig) {
      s0 = strided_sum8(part_mlp + e, nphig, g, nblk);
    }
    sm[g][c] = s0 + s1;
    __syncthreads();
    if (g == 0) {
      double s = 0, bad = 0;
      for (int gg = 0; gg < 8; gg++) s += sm[gg][c];
      if (e < nphig) {
        out[e] = s;
        if (grad_T) grad_T[e] = (T)s;
        bad = isfinite(s) ? 0.0 : 1.0;
      }
      bad = warp_sum(bad);
      if (c == 0) red[128 + bid] = bad;
    }
  } else {
    for (int a0 = 0; a0 < NACC; a0 += 32) {
      const int a = a0 + c;
      double s0 = 0, s1 = 0;
      if (a < NACC) {
        s0 = strided_sum8(part_stream + a, NACC, g, nrows);
      }
      sm[g][c] = s0 + s1;
      __syncthreads();
      if (g == 0 && a < NACC) {
        double s = 0;
        for (int gg = 0; gg < 8; gg++) s += sm[gg][c];
        red[a] = s;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) red[128 + bid] = 0.0;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS) k_reduce(int nphig, const double* __restrict__ part_mlp, int nblk,
                                                        const double* __restrict__ part_stream, int nrows, int NACC,
                                                        double* __restrict__ out, T* __restrict__ grad_T,
                                                        double* __restrict__ red) {
  reduce_body<T>(nphig, part_mlp, nblk, part_stream, nrows, NACC, out, grad_T, red, blockIdx.x, gridDim.x);
}


// k_finalize executes a few thousand instructions once; with a cold instruction cache (after a large streaming kernel)
// its run time is instruction fetch, so loops are not unrolled and helpers are not inlined (4888 -> 2128 instructions)
template <typename T>
__device__ __noinline__ void finalize_body(const Cfg<T>& cfg, const T* __restrict__ phi, const T* __restrict__ zP,
                                           const T* __restrict__ pdraw, const EvalConsts<T>* __restrict__ ecp,
                                           const double* __restrict__ red, int nred_blocks,
                                           const double* __restrict__ xs, const double* __restrict__ xu,
                                           const double* __restrict__ batch_consts, int64_t nb, int M,
                                           double* __restrict__ out, T* __restrict__ grad_T) {
  __shared__ double sred[(3 + MAXP) * FIN_THREADS];
  __shared__ double sacc[128];
  __shared__ double sq[8 * MAXP + 2 * MAXP * MAXP + 16];   // ∇ϕq then the 8 trailing scalars
  const int t = threadIdx.x;
  const int nP = cfg.nP, nM = cfg.nM;
  const int NACC = nacc(nP, nM);
  // MeanVarSep: the block logσ2_ζMs (nM × n_site_total) is written by the streaming kernel; everything
  // here works on ∇ϕq with that block cut out
  const int big = cfg.sepvar ? nM * cfg.n_site_total : 0;
  const int nq = cfg.nphi - cfg.nphig - big;
  // the serial section below reads the evaluation constants and a few entries of ϕ: fetch them with one parallel
  // round trip instead of one dependent global load after the other
  __shared__ EvalConsts<T> sec;
  __shared__ T sls2P[MAXP];
  __shared__ double sbc[2];   // the two constants of the registered batch
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(ecp);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sec);
    _Pragma("unroll 1") for (int e = t; e < (int)(sizeof(EvalConsts<T>) / 4); e += FIN_THREADS) dst[e] = src[e];
    if (t < nP) sls2P[t] = phi[cfg.o_ls2P + t];
    if (t < 2) sbc[t] = batch_consts[t];
  }
  double bad = 0.0;
  _Pragma("unroll 1") for (int e = t; e < nred_blocks; e += FIN_THREADS) bad += red[128 + e];
  _Pragma("unroll 1") for (int e = t; e < NACC; e += FIN_THREADS) sacc[e] = red[e];
  _Pragma("unroll 1") for (int e = t; e < nq + 8; e += FIN_THREADS) sq[e] = 0.0;
  // prior and log-Jacobian of θP and their cotangents, once per evaluation: the
  // draws are spread over the threads, each sum is a fixed-shape tree
  __shared__ double sPg[MAXP][3 + MAXP];   // per j: Σ pv, Σ lj, Σ zb, Σ zb·zP[m, jj]
  __shared__ double sPx[MAXP][1 + MAXP];   // per j: Σ x̄, Σ x̄·zP[m, jj]  (pbm covariates)
  {
    const T* zetaP_ = pdraw + (size_t)nP * M;
    _Pragma("unroll 1") for (int j = 0; j < nP; j++) {
      double v[3 + MAXP];
      _Pragma("unroll 1") for (int i = 0; i < 3 + MAXP; i++) v[i] = 0.0;
      if (cfg.count_globals) {
        _Pragma("unroll 1") for (int m = t; m < M; m += FIN_THREADS) {
          T th, dth, pv, dz, ljm;
          transform_prior<T>(cfg.transP[j], cfg.omit_priors ? HVI_PR_NONE : cfg.prP_kind[j], cfg.prP_a[j],
                             cfg.prP_ib[j], cfg.clamp_lo, cfg.clamp_hi, zetaP_[(size_t)j * M + m], th, dth, pv, dz, ljm);
          v[0] += (double)pv; v[1] += (double)ljm; v[2] += (double)dz;
          _Pragma("unroll 1") for (int jj = 0; jj <= j; jj++) v[3 + jj] += (double)dz * (double)zP[m + (size_t)M * jj];
        }
      }
      block_sum_fixed_n(v, 3 + j + 1, sred, sPg[j]);
      // cotangent that reaches ζP[j, m] through the pbm covariate inputs of g (already scaled by 1/n_MC)
      if (xu) {
        _Pragma("unroll 1") for (int i = 0; i < 1 + MAXP; i++) v[i] = 0.0;
        _Pragma("unroll 1") for (int c = 0; c < cfg.npc; c++) {
          if (cfg.pbm_covar_idx[c] != j) continue;
          _Pragma("unroll 1") for (int m = t; m < M; m += FIN_THREADS) {
            const double x = xu[(size_t)m * cfg.npc + c];
            v[0] += x;
            _Pragma("unroll 1") for (int jj = 0; jj <= j; jj++) v[1 + jj] += x * (double)zP[m + (size_t)M * jj];
          }
        }
        block_sum_fixed_n(v, 1 + j + 1, sred, sPx[j]);
      }
    }
  }
  __syncthreads();
  // the serial part runs on two warps: thread 0 takes the likelihood/prior scalars, the log σ² coefficients and the M block,
  // thread 32 the global (P) block; they write disjoint entries of ∇ϕq, the P block's scalar contributions travel through
  // shared memory and thread 0 assembles the components afterwards
  __shared__ double sPpart[3];   // P block's parts of nlp, lj, logsig
  const double w = 1.0 / (double)M;
  const int NUM = tri(nM);
  const double* abar = sacc + ACC_FIRST_VEC;
  const double* bbar = abar + nM;
  const double* cU = bbar + nM;
  const double* aPs = cU + NUM;
  const double* cPs = aPs + nP;
  // G(o): slot of ϕ position o (≥ nphig, outside the logσ2_ζMs block) in the compressed vector
  auto Gp = [&](int o) -> double* { return sq + (o - cfg.nphig - ((cfg.sepvar && o >= cfg.o_coef + big) ? big : 0)); };
  double nLy = 0.0, nlp = 0.0, lj = 0.0, logsig = 0.0;
  if (t == 0) {
    const EvalConsts<T>& ec = sec;
    nLy = 0.5 * w * sacc[ACC_NLY] + sbc[0];
    nlp = w * sacc[ACC_PRIOR];
    lj = w * sacc[ACC_LJ];
    logsig = sacc[ACC_LOGSIG];
    if (!cfg.omit_priors)
      _Pragma("unroll 1") for (int k = 0; k < nM; k++) nlp += (double)nb * cfg.prM_c0[k];
    // coefficients of log σ²
    if (!cfg.sepvar)
      _Pragma("unroll 1") for (int k = 0; k < nM; k++) { *Gp(cfg.o_coef + 2 * k) = abar[k]; *Gp(cfg.o_coef + 2 * k + 1) = bbar[k]; }
    // M block: Ū = w·cU (columns restricted to their block) → ρ̄sM
    double Ubar[MAXP][MAXP];
    _Pragma("unroll 1") for (int k = 0; k < nM; k++)
      _Pragma("unroll 1") for (int j = 0; j <= k; j++) Ubar[j][k] = w * cU[k * (k + 1) / 2 + j];
    build_U_adj<T>(nM, cfg.blkM_start, cfg.rhoM_off, ec.UM, ec.nrmM, Ubar, Gp(cfg.o_rhoM));
  } else if (t == 32) {
    const EvalConsts<T>& ec = sec;
    // global block
    double nlpP = 0.0, ljP = 0.0, logsigP = 0.0;
    double Ubar[MAXP][MAXP], CP[MAXP][MAXP];
    _Pragma("unroll 1") for (int j = 0; j < nP; j++) {
      double aP = w * aPs[j];
      _Pragma("unroll 1") for (int jj = 0; jj <= j; jj++) CP[jj][j] = w * cPs[j * (j + 1) / 2 + jj];
      if (cfg.count_globals) {
        aP += w * sPg[j][2];
        _Pragma("unroll 1") for (int jj = 0; jj <= j; jj++) CP[jj][j] += w * sPg[j][3 + jj];
        nlpP += w * sPg[j][0] + (cfg.omit_priors ? 0.0 : cfg.prP_c0[j]);
        ljP += w * sPg[j][1];
        logsigP += 0.5 * (double)sls2P[j];
      }
      if (xu) {
        aP += sPx[j][0];
        _Pragma("unroll 1") for (int jj = 0; jj <= j; jj++) CP[jj][j] += sPx[j][1 + jj];
      }
      *Gp(cfg.o_muP + j) += aP;
      // Σ_m ζ̄P[j,m]·rP[j,m] = σP[j] Σ_{jj} UP[jj][j]·CP[jj][j]
      double sr = 0;
      _Pragma("unroll 1") for (int jj = cfg.blkP_start[j]; jj <= j; jj++) sr += (double)ec.UP[jj][j] * CP[jj][j];
      sr *= (double)ec.sigP[j];
      *Gp(cfg.o_ls2P + j) = 0.5 * sr - (cfg.count_globals ? 0.5 : 0.0);
      _Pragma("unroll 1") for (int jj = 0; jj <= j; jj++) Ubar[jj][j] = (double)ec.sigP[j] * CP[jj][j];
    }
    build_U_adj<T>(nP, cfg.blkP_start, cfg.rhoP_off, ec.UP, ec.nrmP, Ubar, Gp(cfg.o_rhoP));
    // pass 1 of g took μP[idx] as input (src/elbo.jl:491)
    if (xs)
      _Pragma("unroll 1") for (int c = 0; c < cfg.npc; c++) *Gp(cfg.o_muP + cfg.pbm_covar_idx[c]) += xs[c];
    sPpart[0] = nlpP; sPpart[1] = ljP; sPpart[2] = logsigP;
  }
  __syncthreads();
  if (t == 0) {
    // the sums in the order of the one-thread form: the M block's terms first, then the P block's
    nlp += sPpart[0]; lj += sPpart[1]; logsig += sPpart[2];
    const double Kdim = (double)nM * (double)nb + (cfg.count_globals ? nP : 0);
    const double ent = 0.5 * Kdim * 2.8378770664093454836 + logsig;   // (K·log(2πe) + 2Σlog σ)/2, logden_normal.jl:74
    const double nlj = -lj;
    double* C = sq + nq;
    C[0] = nLy + nlp + nlj; C[1] = ent; C[2] = 0.0; C[3] = nLy; C[4] = nlp; C[5] = nlj;
    C[6] = C[0] - C[1] + C[2];
  }
  __syncthreads();
  _Pragma("unroll 1") for (int e = t; e < nq + 7; e += FIN_THREADS) {
    const double v = sq[e];
    const int o = cfg.nphig + e + ((cfg.sepvar && cfg.nphig + e >= cfg.o_coef) ? big : 0);   // back to ϕ positions
    out[o] = v;
    if (grad_T && e < nq) grad_T[o] = (T)v;
    bad += isfinite(v) ? 0.0 : 1.0;
  }
  if (cfg.sepvar && t == 0)
    _Pragma("unroll 1") for (int k = 0; k < nM; k++) bad += sacc[ACC_FIRST_VEC + k];
  // an observation that is finite where a driver is not makes the reference's loss NaN: the masked
  // value is evaluated, and the anomaly count joins the non-finite count so that no caller of the device-resident result
  // (optimiser loop, all-reduce) takes the iteration for a clean one
  if (t == 0) bad += sbc[1];
  bad = block_sum_fixed(bad, sred);
  if (t == 0) out[cfg.nphi + 7] = bad;
}

template <typename T>
__global__ void __launch_bounds__(FIN_THREADS) k_finalize(const __grid_constant__ Cfg<T> cfg,
                                                          const T* __restrict__ phi, const T* __restrict__ zP,
                                                          const T* __restrict__ pdraw,
                                                          const EvalConsts<T>* __restrict__ ecp,
                                                          const double* __restrict__ red, int nred_blocks,
                                                          const double* __restrict__ xs, const double* __restrict__ xu,
                                                          const double* __restrict__ batch_consts, int64_t nb, int M,
                                                          double* __restrict__ out, T* __restrict__ grad_T) {
  finalize_body<T>(cfg, phi, zP, pdraw, ecp, red, nred_blocks, xs, xu, batch_consts, nb, M, out, grad_T);
}

// gradient entries of ϕg are written by many threads above; convert them after the block is done
template <typename T>
__global__ void k_cast_grad(int n, const double* __restrict__ out, T* __restrict__ grad_T) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) grad_T[e] = (T)out[e];
}

// ---------------------------------------------------------------------------------------
// counter-based standard normals (Philox4x32-10 + Box-Muller), keyed by
// (seed; global column, draw/4): independent of the sharding.  Stands in for CUDA.randn
// (ext/HybridVariationalInferenceCUDAExt.jl:82-86).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_gen_normal(T* __restrict__ z, int64_t n_cols, int M, uint64_t seed,
                                                    int64_t col_offset, const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev) seed = *seed_dev;
  const int M4 = (M + 3) / 4;
  const int64_t n = n_cols * M4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t col = e / M4;
    const int m4 = (int)(e % M4);
    const uint64_t gcol = (uint64_t)(col + col_offset);
    float n4[4];
    normal4(seed, gcol, m4, n4);
    for (int d = 0; d < 4; d++) {
      const int m = m4 * 4 + d;
      if (m < M) z[(size_t)col * M + m] = (T)n4[d];
    }
  }
}

// ---------------------------------------------------------------------------------------
// generate_ζ forward (src/elbo.jl:472-521): ζsP (nP×M), ζsMs_tr (nb×nM×M), σ.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_generate_zeta(const __grid_constant__ Cfg<T> cfg,
                                                       const __grid_constant__ StreamArgs<T> a,
                                                       const T* __restrict__ pdraw, T* __restrict__ zetaP,
                                                       T* __restrict__ zetaMs_tr, T* __restrict__ sigma) {
  const int nP = cfg.nP, nM = cfg.nM, M = a.M;
  const EvalConsts<T>& ec = *a.ec;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid < (int64_t)nP * M) {
    const int j = (int)(gid / M), m = (int)(gid % M);
    zetaP[j + (size_t)nP * m] = pdraw[(size_t)nP * M + gid];  // ζP array of pdraw
  }
  if (gid < nP) sigma[gid] = ec.sigP[gid];
  for (int64_t e = gid; e < a.nb * nM; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / nM;
    const int k = (int)(e % nM);
    const T mu = a.mu[i * nM + k];
    const T sg = exp_t(T(0.5) * (cfg.sepvar ? a.phi[cfg.o_coef + (a.i_sites ? a.i_sites[i] : i) * nM + k]
                                            : ec.coefA[k] + ec.coefB[k] * mu));
    sigma[nP + i * nM + k] = sg;
    for (int m = 0; m < M; m++) {
      T t = T(0);
      for (int j = cfg.blkM_start[k]; j <= k; j++) t += ec.UM[j][k] * a.zMs[((size_t)i * nM + j) * M + m];
      const T mum = a.MU ? a.MU[((size_t)m * nM + k) * a.nb + i] : mu;   // pass-2 mean with pbm covariates
      zetaMs_tr[i + a.nb * (k + (size_t)nM * m)] = mum + sg * t;
    }
  }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
int mlp_bwd_max_blocks(int sm_count) { return sm_count * 4; }

template <typename T>
cudaError_t launch_preprocess(const Launch& L, int nO, int64_t nb, const T* xP, const T* y_o, const T* y_unc, T* rec,
                              double* part, int* nrows_out) {
  const int64_t ntiles = (nb + 31) / 32;
  int blocks = (int)((ntiles + 7) / 8);
  if (blocks > L.sm_count * 8) blocks = L.sm_count * 8;
  if (blocks < 1) blocks = 1;
  k_preprocess<T><<<blocks, 256, 0, L.stream>>>(nO, nb, xP, y_o, y_unc, rec, part);
  HVI_LAUNCH_CHECK();
  *nrows_out = blocks * 8;
  return cudaSuccess;
}

template <typename T>
cudaError_t launch_gather_preprocess(const Launch& L, int nC, int nO, int64_t nb, const int64_t* idx, int64_t first,
                                     const T* ds_xM, const T* ds_xP, const T* ds_yo, const T* ds_yu, T* xM_b, T* xP_b,
                                     T* rec, double* part, int* nrows_out, const unsigned long long* ctr,
                                     int64_t n_batches, int64_t* isites_out) {
  const int64_t ntiles = (nb + 31) / 32;
  int blocks = (int)((ntiles + 7) / 8);
  if (blocks > L.sm_count * 8) blocks = L.sm_count * 8;
  if (blocks < 1) blocks = 1;
  k_gather_preprocess<T><<<blocks, 256, 0, L.stream>>>(nC, nO, nb, idx, first, ds_xM, ds_xP, ds_yo, ds_yu, xM_b, xP_b, rec, part, ctr,
                                                       n_batches, isites_out);
  HVI_LAUNCH_CHECK();
  *nrows_out = blocks * 8;
  return cudaSuccess;
}

cudaError_t launch_batch_consts(const Launch& L, const double* part, int nrows, double* consts) {
  k_batch_consts<<<1, 256, 0, L.stream>>>(part, nrows, consts);
  HVI_LAUNCH_CHECK();
  return cudaSuccess;
}

template <typename T>
cudaError_t launch_prologue(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* zP, int M, EvalConsts<T>* ec,
                            T* pdraw) {
  k_prologue<T><<<1, 128, 0, L.stream>>>(cfg, phi, zP, M, ec, pdraw);
  HVI_LAUNCH_CHECK();
  return cudaSuccess;
}

template <typename T>
static size_t mlp_fwd_smem(const Cfg<T>& cfg) {
  return sizeof(T) * (((cfg.nphig + 3) & ~3) + (size_t)cfg.sum_width * MLP_BS);
}
template <typename T>
__host__ __device__ static size_t mlp_bwd_smem(const Cfg<T>& cfg) {
  return (sizeof(double) + sizeof(int2)) * cfg.nphig +
         sizeof(T) * (((cfg.nphig + 3) & ~3) + (size_t)(cfg.sum_width + cfg.sum_out) * MLP_BS);
}

template <typename T>
cudaError_t launch_mlp_fwd(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* xM, int64_t nb, T* mu, int M_units,
                           const T* zetaP) {
  if constexpr (sizeof(T) == 4) {   // Float32: the named shapes run on the tensor cores
    const cudaError_t et = launch_mlp3_tc_fwd(L, cfg, phi, xM, nb, mu, M_units, zetaP);
    if (et != cudaErrorNotSupported) return et;
    const cudaError_t em = launch_mlp3_mma_fwd(L, cfg, phi, xM, nb, mu, M_units, zetaP);
    if (em != cudaErrorNotSupported) return em;
  }
#define X(NI, H1, H2, NO_)                                                                       \
  if (mlp3_match(cfg, NI, H1, H2, NO_)) {                                                        \
    int64_t blocks = ((nb + MLP3_BS - 1) / MLP3_BS) * (M_units > 0 ? M_units : 1);               \
    if (blocks > (int64_t)L.sm_count * 8) blocks = (int64_t)L.sm_count * 8;                      \
    MlpCols cols{M_units, nb};                                                                   \
    k_mlp3_fwd<T, NI, H1, H2, NO_><<<(int)blocks, MLP3_BS, 0, L.stream>>>(cfg, phi, xM, zetaP, cols, mu); \
    HVI_LAUNCH_CHECK();                                                                          \
    return cudaSuccess;                                                                          \
  }
  HVI_MLP3_CASES(X)
#undef X
  if (M_units > 0 || cfg.npc > 0) return cudaErrorInvalidConfiguration;  // pbm covariates need the tiled applicator
  const size_t smem = mlp_fwd_smem(cfg);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(k_mlp_fwd<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t ntiles = (nb + MLP_BS - 1) / MLP_BS;
  int blocks = (int)(ntiles < (int64_t)L.sm_count * 4 ? ntiles : (int64_t)L.sm_count * 4);
  if (blocks < 1) blocks = 1;
  k_mlp_fwd<T><<<blocks, MLP_BS, smem, L.stream>>>(cfg, phi, xM, nb, mu);
  HVI_LAUNCH_CHECK();
  return cudaSuccess;
}

bool mlp_supports_pbm(const hvi_desc* d) {
#define X(NI, H1, H2, NO_)                                                                                              \
  if (d->n_layers == 3 && d->layers[0].n_in == NI && d->layers[0].n_out == H1 && d->layers[1].n_out == H2 &&            \
      d->layers[2].n_out == NO_ && d->layers[0].act == HVI_ACT_TANH && d->layers[1].act == HVI_ACT_TANH &&              \
      d->layers[2].act == HVI_ACT_LOGISTIC && d->layers[0].has_bias && d->layers[1].has_bias && !d->layers[2].has_bias) \
    return true;
  HVI_MLP3_CASES(X)
#undef X
  return false;
}

template <typename T>
cudaError_t launch_mlp_bwd(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* xM, const T* mubar, int64_t nb,
                           double* part, int* nblk_out, int nblk_max, int M_units, const T* zetaP, double* part_x,
                           const T* mu_fwd) {
  if constexpr (sizeof(T) == 4) {
    const cudaError_t et = launch_mlp3_tc_bwd(L, cfg, phi, xM, mubar, nb, part, nblk_out, nblk_max, M_units, zetaP, part_x, mu_fwd);
    if (et != cudaErrorNotSupported) return et;
    const cudaError_t em = launch_mlp3_mma_bwd(L, cfg, phi, xM, mubar, nb, part, nblk_out, nblk_max, M_units, zetaP, part_x);
    if (em != cudaErrorNotSupported) return em;
  }
#define X(NI, H1, H2, NO_)                                                                                   \
  if (mlp3_match(cfg, NI, H1, H2, NO_)) {                                                                    \
    using N = Mlp3<T, NI, H1, H2, NO_>;                                                                      \
    const size_t smem3 = sizeof(T) * (pad4(N::nW) + (size_t)MLP3_BS * N::ROW + (size_t)4 * N::NT * 16) + \
                         sizeof(double) * (size_t)(M_units > 0 ? M_units : 1) * cfg.npc + sizeof(T) * 4 * cfg.npc + 16; \
    if (smem3 > 220 * 1024) return cudaErrorInvalidConfiguration;                                            \
    auto kern = k_mlp3_bwd<T, NI, H1, H2, NO_>;                                                              \
    cudaError_t e3 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);    \
    if (e3 != cudaSuccess) return e3;                                                                        \
    const int64_t ntiles3 = ((nb + MLP3_BS - 1) / MLP3_BS) * (M_units > 0 ? M_units : 1);                    \
    int blocks3 = (int)(ntiles3 < (int64_t)nblk_max ? ntiles3 : (int64_t)nblk_max);                          \
    if (blocks3 < 1) blocks3 = 1;                                                                            \
    MlpCols cols{M_units, nb};                                                                               \
    kern<<<blocks3, MLP3_BS, smem3, L.stream>>>(cfg, phi, xM, zetaP, mubar, cols, part, part_x);             \
    HVI_LAUNCH_CHECK();                                                                                      \
    *nblk_out = blocks3;                                                                                     \
    return cudaSuccess;                                                                                      \
  }
  HVI_MLP3_CASES(X)
#undef X
  if (M_units > 0 || cfg.npc > 0) return cudaErrorInvalidConfiguration;
  const size_t smem = mlp_bwd_smem(cfg);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(k_mlp_bwd<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t ntiles = (nb + MLP_BS - 1) / MLP_BS;
  int blocks = (int)(ntiles < (int64_t)nblk_max ? ntiles : (int64_t)nblk_max);
  if (blocks < 1) blocks = 1;
  k_mlp_bwd<T><<<blocks, MLP_BS, smem, L.stream>>>(cfg, phi, xM, mubar, nb, part);
  HVI_LAUNCH_CHECK();
  *nblk_out = blocks;
  return cudaSuccess;
}

template <typename T>
cudaError_t launch_finalize(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* zP, const T* pdraw,
                            const EvalConsts<T>* ec, const double* part_stream, int nrows, const double* part_mlp,
                            int nblk, const double* batch_consts, int64_t nb, int M, double* out, T* grad_T, double* red,
                            const double* part_xs, int nblk_xs, const double* part_xu, int nblk_xu, double* xred) {
  const int nbg = (cfg.nphig + 31) / 32;
  k_reduce<T><<<nbg + 1, RED_THREADS, 0, L.stream>>>(cfg.nphig, part_mlp, nblk, part_stream, nrows, nacc(cfg.nP, cfg.nM),
                                                     out, grad_T, red);
  HVI_LAUNCH_CHECK();
  double *xs = nullptr, *xu = nullptr;
  if (cfg.npc > 0 && part_xs && part_xu) {
    xs = xred;
    xu = xred + cfg.npc;
    k_colsum<<<(cfg.npc + 31) / 32, 256, 0, L.stream>>>(part_xs, nblk_xs, cfg.npc, xs);
    HVI_LAUNCH_CHECK();
    k_colsum<<<(M * cfg.npc + 31) / 32, 256, 0, L.stream>>>(part_xu, nblk_xu, M * cfg.npc, xu);
    HVI_LAUNCH_CHECK();
  }
  k_finalize<T><<<1, FIN_THREADS, 0, L.stream>>>(cfg, phi, zP, pdraw, ec, red, nbg + 1, xs, xu, batch_consts, nb, M, out,
                                                grad_T);
  HVI_LAUNCH_CHECK();
  return cudaSuccess;
}

template <typename T>
cudaError_t launch_gen_normal(const Launch& L, T* z, int64_t n_cols, int M, uint64_t seed, int64_t col_offset,
                              const unsigned long long* seed_dev) {
  if (n_cols <= 0) return cudaSuccess;
  const int64_t n = n_cols * ((M + 3) / 4);
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)L.sm_count * 16) blocks = (int64_t)L.sm_count * 16;
  k_gen_normal<T><<<(int)blocks, 256, 0, L.stream>>>(z, n_cols, M, seed, col_offset, seed_dev);
  HVI_LAUNCH_CHECK();
  return cudaSuccess;
}

// generate_ζ sized for its traffic (it reads ξ once and writes ζ once, n_θM values per site·draw each way): one thread
// per site handles all its parameters, so that for a fixed (k, draw) the warp writes 32 consecutive sites of one row
// of ζMs_tr (the kernel above alternates k between neighbouring threads and half-fills every sector); four draws per
// step, the draws of a column are read as one 16-byte load when the layout allows it.
template <typename T, int NM>
__global__ void __launch_bounds__(256) k_generate_zeta_fast(const __grid_constant__ Cfg<T> cfg,
                                                            const __grid_constant__ StreamArgs<T> a,
                                                            const T* __restrict__ pdraw, T* __restrict__ zetaP,
                                                            T* __restrict__ zetaMs_tr, T* __restrict__ sigma) {
  const int nP = cfg.nP, M = a.M;
  const EvalConsts<T>& ec = *a.ec;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t e = gid; e < (int64_t)nP * M; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e / M), m = (int)(e % M);
    zetaP[j + (size_t)nP * m] = pdraw[(size_t)nP * M + e];  // ζP array of pdraw
  }
  if (gid < nP) sigma[gid] = ec.sigP[gid];
  constexpr int CH = 16 / (int)sizeof(T);
  const bool vec = (M % CH == 0) && (((uintptr_t)a.zMs & 15) == 0);
  for (int64_t i = gid; i < a.nb; i += (int64_t)gridDim.x * blockDim.x) {
    T mu[NM], U[NM][NM];
#pragma unroll
    for (int k = 0; k < NM; k++) {
      mu[k] = a.mu[i * NM + k];
      const T sg = exp_t(T(0.5) * (cfg.sepvar ? a.phi[cfg.o_coef + (a.i_sites ? a.i_sites[i] : i) * NM + k]
                                              : ec.coefA[k] + ec.coefB[k] * mu[k]));
      sigma[nP + i * NM + k] = sg;
#pragma unroll
      for (int j = 0; j < NM; j++) U[j][k] = (j >= cfg.blkM_start[k] && j <= k) ? ec.UM[j][k] * sg : T(0);
    }
    const T* zrow = a.zMs + (size_t)i * NM * M;
    T* out = zetaMs_tr + i;
    for (int m0 = 0; m0 < M; m0 += CH) {
      T z[NM][CH];
#pragma unroll
      for (int j = 0; j < NM; j++) {
        if (vec) {
          if constexpr (sizeof(T) == 4) {
            const float4 v = *reinterpret_cast<const float4*>(zrow + (size_t)j * M + m0);
            z[j][0] = v.x; z[j][1] = v.y; z[j][2] = v.z; z[j][3] = v.w;
          } else {
            const double2 v = *reinterpret_cast<const double2*>(zrow + (size_t)j * M + m0);
            z[j][0] = v.x; z[j][1] = v.y;
          }
        } else {
#pragma unroll
          for (int d = 0; d < CH; d++) z[j][d] = m0 + d < M ? zrow[(size_t)j * M + m0 + d] : T(0);
        }
      }
#pragma unroll
      for (int d = 0; d < CH; d++) {
        const int m = m0 + d;
        if (m >= M) break;
#pragma unroll
        for (int k = 0; k < NM; k++) {
          T r = a.MU ? a.MU[((size_t)m * NM + k) * a.nb + i] : mu[k];   // pass-2 mean with pbm covariates
#pragma unroll
          for (int j = 0; j <= k; j++) r = fma(U[j][k], z[j][d], r);
          __stcs(out + a.nb * (k + (size_t)NM * m), r);
        }
      }
    }
  }
}

template <typename T>
cudaError_t launch_generate_zeta(const Launch& L, const Cfg<T>& cfg, const StreamArgs<T>& a, const T* pdraw, T* zetaP,
                                 T* zetaMs_tr, T* sigma) {
  if (cfg.nM >= 1 && cfg.nM <= 3) {
    int64_t blocks = (a.nb + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)L.sm_count * 16) blocks = (int64_t)L.sm_count * 16;
    if (cfg.nM == 1) k_generate_zeta_fast<T, 1><<<(int)blocks, 256, 0, L.stream>>>(cfg, a, pdraw, zetaP, zetaMs_tr, sigma);
    else if (cfg.nM == 2) k_generate_zeta_fast<T, 2><<<(int)blocks, 256, 0, L.stream>>>(cfg, a, pdraw, zetaP, zetaMs_tr, sigma);
    else k_generate_zeta_fast<T, 3><<<(int)blocks, 256, 0, L.stream>>>(cfg, a, pdraw, zetaP, zetaMs_tr, sigma);
    HVI_LAUNCH_CHECK();
    return cudaSuccess;
  }
  int64_t n = a.nb * cfg.nM;
  if (n < (int64_t)cfg.nP * a.M) n = (int64_t)cfg.nP * a.M;
  int64_t blocks = (n + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > (int64_t)L.sm_count * 16) blocks = (int64_t)L.sm_count * 16;
  // the first nP*M threads also copy ζP; make sure they exist
  if (blocks * 256 < (int64_t)cfg.nP * a.M) return cudaErrorInvalidConfiguration;
  k_generate_zeta<T><<<(int)blocks, 256, 0, L.stream>>>(cfg, a, pdraw, zetaP, zetaMs_tr, sigma);
  HVI_LAUNCH_CHECK();
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------
// The reference's own operating point -- n_batch = 20 sites of 800, n_MC = 3 … 12 (src/DoubleMM/f_doubleMM.jl:315-325,
// src/HybridSolver.jl:147) -- is pure launch latency when it runs as eight kernels.  k_small_iter runs one whole
// iteration of the optimiser loop (src/HybridSolver.jl:256-260) in ONE block of ONE launch: [minibatch gather from
// the device-resident dataset] → draws of the global block → prologue → g forward → fused streaming phase (device
// draws) → g backward → fixed-order reductions → finalize → [Adam update].  Every phase is the body of the kernel the
// large-batch path launches (same arithmetic, same fixed reduction orders within a phase), separated by block
// barriers instead of kernel boundaries.  For batches of a few hundred sites; g must be one of the CUDA-core shapes
// without pbm covariates.
// ---------------------------------------------------------------------------------------
// g for a small batch inside one block: the work is spread over (site, neuron) pairs instead of one thread per site,
// so that a 20-site minibatch keeps all threads busy and every dependent chain is one dot product long.  Same
// shared-memory layout as k_mlp_fwd / k_mlp_bwd (weights, activation stack [row][MLP_BS], delta stack), same
// accumulation order within a dot product as mlp_forward_block.  All threads of the block call these.
template <typename T>
__device__ __forceinline__ void small_mlp_forward_tile(const Cfg<T>& cfg, const T* __restrict__ sW, T* __restrict__ acts, int ns) {
  const int t = threadIdx.x, NT = blockDim.x;
  int row_in = 0;
  for (int l = 0; l < cfg.nL; l++) {
    const int ni = cfg.l_in[l], no = cfg.l_out[l], row_out = row_in + ni;
    const T* W = sW + cfg.woff[l];
    for (int task = t; task < ns * no; task += NT) {
      const int site = task % ns, j = task / ns;
      T z = cfg.l_bias[l] ? sW[cfg.boff[l] + j] : T(0);
      for (int i = 0; i < ni; i++) z = fma(W[i * no + j], acts[(row_in + i) * MLP_BS + site], z);
      acts[(row_out + j) * MLP_BS + site] = act_f<T>(cfg.l_act[l], z);
    }
    __syncthreads();
    row_in = row_out;
  }
}

template <typename T>
__device__ __forceinline__ void small_mlp_fwd(const Cfg<T>& cfg, const T* __restrict__ phi, const T* __restrict__ xM,
                                              int64_t nb, T* __restrict__ mu, unsigned char* smem_raw) {
  // the layout of small_mlp_bwd (gradient accumulator | table | weights | activations | deltas): the backward phase
  // finds weights and activations of a one-tile batch where this phase left them
  T* sW = reinterpret_cast<T*>(reinterpret_cast<int2*>(reinterpret_cast<double*>(smem_raw) + cfg.nphig) + cfg.nphig);
  T* acts = sW + ((cfg.nphig + 3) & ~3);
  const int t = threadIdx.x, NT = blockDim.x;
  for (int e = t; e < cfg.nphig; e += NT) sW[e] = phi[e];
  const int out_row = cfg.sum_width - cfg.l_out[cfg.nL - 1];
  for (int64_t s0 = 0; s0 < nb; s0 += MLP_BS) {
    const int ns = (int)(nb - s0 < MLP_BS ? nb - s0 : MLP_BS);
    for (int e = t; e < ns * cfg.nC; e += NT) acts[(e % cfg.nC) * MLP_BS + e / cfg.nC] = xM[s0 * cfg.nC + e];
    __syncthreads();
    small_mlp_forward_tile<T>(cfg, sW, acts, ns);
    for (int e = t; e < ns * cfg.nM; e += NT) {
      const int site = e / cfg.nM, k = e % cfg.nM;
      const T p = acts[(out_row + k) * MLP_BS + site];
      mu[(s0 + site) * cfg.nM + k] = (cfg.scale_kind == HVI_SCALE_RANGE) ? p * cfg.scale_b[k] + cfg.scale_a[k]
                                                                        : cfg.scale_a[k] + cfg.scale_b[k] * ndtri_t(p);
    }
    __syncthreads();
  }
}

// one row of ∇ϕg partial sums (part[0 .. nphig)), deterministic
template <typename T>
__device__ __forceinline__ void small_mlp_bwd(const Cfg<T>& cfg, const T* __restrict__ phi, const T* __restrict__ xM,
                                              const T* __restrict__ mubar, int64_t nb, double* __restrict__ part,
                                              unsigned char* smem_raw, bool reuse = false, bool tab_ready = false) {
  double* gacc = reinterpret_cast<double*>(smem_raw);                 // [nphig]
  int2* tab = reinterpret_cast<int2*>(gacc + cfg.nphig);              // [nphig] (delta row, act row | -1)
  T* sW = reinterpret_cast<T*>(tab + cfg.nphig);                      // [nphig]
  T* acts = sW + ((cfg.nphig + 3) & ~3);                              // [sum_width][BS]
  T* dl = acts + (size_t)cfg.sum_width * MLP_BS;                      // [sum_out][BS]
  const int t = threadIdx.x, NT = blockDim.x;
  // reuse: weights and the activations of the batch's only tile are still in place from small_mlp_fwd
  reuse = reuse && nb <= MLP_BS;
  for (int e = t; e < cfg.nphig; e += NT) { if (!reuse) sW[e] = phi[e]; gacc[e] = 0.0; }
  if (!(reuse && tab_ready)) {   // the table depends on the network only: inside a launch it is built once
    int row_in = 0, drow = 0;
    for (int l = 0; l < cfg.nL; l++) {
      const int ni = cfg.l_in[l], no = cfg.l_out[l];
      for (int e = t; e < ni * no; e += NT) tab[cfg.woff[l] + e] = make_int2(drow + e % no, row_in + e / no);
      if (cfg.l_bias[l])
        for (int e = t; e < no; e += NT) tab[cfg.boff[l] + e] = make_int2(drow + e, -1);
      row_in += ni; drow += no;
    }
  }
  const int out_row = cfg.sum_width - cfg.l_out[cfg.nL - 1];
  const int nlast = cfg.l_out[cfg.nL - 1];
  for (int64_t s0 = 0; s0 < nb; s0 += MLP_BS) {
    const int ns = (int)(nb - s0 < MLP_BS ? nb - s0 : MLP_BS);
    if (!reuse) {
      for (int e = t; e < ns * cfg.nC; e += NT) acts[(e % cfg.nC) * MLP_BS + e / cfg.nC] = xM[s0 * cfg.nC + e];
      __syncthreads();
      small_mlp_forward_tile<T>(cfg, sW, acts, ns);
    } else {
      __syncthreads();   // the table and the zeroed accumulator
    }
    // output delta: μ̄_ζ · σ_k · √(2π) e^{q²/2} · p(1−p)
    int drow = cfg.sum_out - nlast;
    for (int e = t; e < ns * nlast; e += NT) {
      const int site = e / nlast, k = e % nlast;
      T d = T(0);
      if (k < cfg.nM) {
        const T p = acts[(out_row + k) * MLP_BS + site];
        const T mb = mubar[(s0 + site) * cfg.nM + k];
        if (cfg.scale_kind == HVI_SCALE_RANGE) d = mb * cfg.scale_b[k];
        else { const T q = ndtri_t(p); d = mb * cfg.scale_b[k] * T(2.50662827463100050242) * exp_t(T(0.5) * q * q); }
        d *= act_d<T>(cfg.l_act[cfg.nL - 1], p);
      }
      dl[(drow + k) * MLP_BS + site] = d;
    }
    __syncthreads();
    // hidden deltas, one (site, neuron) pair per task
    int row_out = out_row;
    for (int l = cfg.nL - 1; l >= 1; l--) {
      const int ni = cfg.l_in[l], no = cfg.l_out[l];
      const int row_in = row_out - ni, drow_prev = drow - ni;
      const T* W = sW + cfg.woff[l];
      for (int task = t; task < ns * ni; task += NT) {
        const int site = task % ns, i = task / ns;
        T sacc = T(0);
        for (int j = 0; j < no; j++) sacc = fma(W[i * no + j], dl[(drow + j) * MLP_BS + site], sacc);
        dl[(drow_prev + i) * MLP_BS + site] = sacc * act_d<T>(cfg.l_act[l - 1], acts[(row_in + i) * MLP_BS + site]);
      }
      __syncthreads();
      row_out = row_in; drow = drow_prev;
    }
    // weight gradient: entry e = Σ_sites δ·a over the sites of this tile
    for (int e = t; e < cfg.nphig; e += NT) {
      const int2 rc = tab[e];
      const T* dr = dl + (size_t)rc.x * MLP_BS;
      T acc = T(0);
      if (rc.y >= 0) {
        const T* ar = acts + (size_t)rc.y * MLP_BS;
        for (int si = 0; si < ns; si++) acc = fma(dr[si], ar[si], acc);
      } else {
        for (int si = 0; si < ns; si++) acc += dr[si];
      }
      gacc[e] += (double)acc;
    }
    __syncthreads();
  }
  for (int e = t; e < cfg.nphig; e += NT) part[e] = gacc[e];
}

constexpr int SMALL_THREADS = 256;
static_assert(SMALL_THREADS == FIN_THREADS && SMALL_THREADS == RED_THREADS && SMALL_THREADS >= MLP_BS, "block shapes of the fused phases");

template <class TR, typename T, int NP, int NM, int NO>
__global__ void __launch_bounds__(SMALL_THREADS, 1) k_small_iter(const __grid_constant__ Cfg<T> cfg,
                                                                 const __grid_constant__ SmallArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int t = threadIdx.x;
  const int M = a.st.M;
  const int64_t nb = a.st.nb;
  // the iterations of the optimiser loop run INSIDE the launch (n_inner > 1): from the second one on the code comes out
  // of the instruction cache instead of L2 -- the single iteration is bound by instruction fetch.  Iteration counter,
  // seed and minibatch index live in device memory (a.ctr) and are advanced by the Adam phase, as for graph replays.
  const int n_inner = (a.n_inner > 1 && a.adam_m && a.ctr) ? a.n_inner : 1;
#define HVI_STAMP() do { if (a.clocks && t == 0) a.clocks[stamp] = clock64(); ++stamp; } while (0)
  _Pragma("unroll 1") for (int inner = 0; inner < n_inner; inner++) {
  const uint64_t seed = a.ctr ? (uint64_t)a.ctr[2] : a.st.seed;
  int stamp = 0;
  HVI_STAMP();
  // 0. minibatch of the device-resident dataset (hvi_adam_run_minibatches): gather + re-tile + batch constants
  if (a.ds_xM) {
    gather_preprocess_body<T>(cfg.nC, NO, nb, nullptr, a.first, a.ds_xM, a.ds_xP, a.ds_yo, a.ds_yu, a.xM_b, a.xP_b, a.rec_w,
                              a.st.partials, a.ctr, a.n_batches, a.isites_out, 0, 1);
    __syncthreads();
    if (t == 0) {
      double c = 0.0, an = 0.0;
      for (int r = 0; r < SMALL_THREADS / 32; r++) { c += a.st.partials[2 * r]; an += a.st.partials[2 * r + 1]; }
      a.bconst[0] = c; a.bconst[1] = an;
    }
    __syncthreads();
  }
  HVI_STAMP();
  // 1. draws of the global block (k_gen_normal: columns keyed far above any site column)
  {
    const int M4 = (M + 3) / 4;
    for (int e = t; e < NP * M4; e += SMALL_THREADS) {
      const int col = e / M4, m4 = e % M4;
      float n4[4];
      normal4(seed, (uint64_t)((int64_t)col + ((int64_t)1 << 62)), m4, n4);
      for (int d = 0; d < 4; d++)
        if (m4 * 4 + d < M) a.zP[(size_t)col * M + m4 * 4 + d] = (T)n4[d];
    }
  }
  __syncthreads();
  HVI_STAMP();
  // 2. ϕq → Cholesky factors, σP, per-draw θP
  prologue_body<T>(cfg, a.phi, a.zP, M, const_cast<EvalConsts<T>*>(a.st.ec), a.pdraw);
  __syncthreads();
  HVI_STAMP();
  // few sites: the streaming phase runs one thread per (site, draw) unit in shared memory BEHIND the applicator's, so that
  // the backward phase finds the forward phase's weights and activations in place (no recomputation)
  const size_t mlp_bytes = (mlp_bwd_smem<T>(cfg) + 15) & ~(size_t)15;
  const bool units_mode = a.units_mode && nb < 128 &&
                          mlp_bytes + stream_units_smem(sizeof(T), nb, M, stream_units_nv<NP, NM>(), nacc(NP, NM)) <= (size_t)a.units_mode;
  // 3. μ_ζ = g(xM; ϕg)
  small_mlp_fwd<T>(cfg, a.phi, a.xM, nb, const_cast<T*>(a.st.mu), smem_raw);
  __syncthreads();
  HVI_STAMP();
  // 4. the fused streaming phase (draws generated in registers)
  {
    StreamArgs<T> sa = a.st;
    sa.seed = seed; sa.seed_dev = nullptr; sa.rng = 1; sa.phi = a.phi;
    // few sites: one thread per (site, draw) unit (stream_units_body); otherwise one thread per site (stream_body)
    if (units_mode) stream_units_body<TR, T, NP, NM, NO, SMALL_THREADS>(cfg, sa, smem_raw + mlp_bytes, seed);
    else stream_body<TR, T, NP, NM, NO, SMALL_THREADS, false, true, (NP > 0), false, true, true>(cfg, sa, smem_raw, 0, 1);
  }
  __syncthreads();
  HVI_STAMP();
  // 5. μ̄_ζ → ∇ϕg (one row of partial sums)
  small_mlp_bwd<T>(cfg, a.phi, a.xM, a.st.mubar, nb, a.part_mlp, smem_raw, units_mode, inner > 0);
  __syncthreads();
  HVI_STAMP();
  // 6. what k_reduce does for a single row of partial sums: ∇ϕg and its non-finite count per group of 32 entries, the
  // streaming phase's accumulators
  const int nbg = (cfg.nphig + 31) / 32;
  for (int e0 = 0; e0 < nbg * 32; e0 += SMALL_THREADS) {
    const int e = e0 + t;
    double bad = 0.0;
    if (e < cfg.nphig) {
      const double v = a.part_mlp[e];
      a.out[e] = v;
      if (a.grad_T) a.grad_T[e] = (T)v;
      bad = isfinite(v) ? 0.0 : 1.0;
    }
    bad = warp_sum(bad);
    if ((t & 31) == 0 && e < nbg * 32) a.red[128 + e / 32] = bad;
  }
  for (int i = t; i < nacc(NP, NM); i += SMALL_THREADS) a.red[i] = a.st.partials[i];
  if (t == 0) a.red[128 + nbg] = 0.0;
  __syncthreads();
  HVI_STAMP();
  // 7. global-block terms, Cholesky adjoint, scatter into the layout of ϕ
  finalize_body<T>(cfg, a.phi, a.zP, a.pdraw, a.st.ec, a.red, nbg + 1, nullptr, nullptr, a.bconst, nb, M, a.out, a.grad_T);
  __syncthreads();
  HVI_STAMP();
  // 8. Adam (Optimisers.jl rule), as k_adam
  if (a.adam_m) {
    const int64_t n = cfg.nphi;
    const unsigned long long t_applied = a.ctr[0], it = a.ctr[1];
    const bool bad = a.out[n + 7] != 0.0 || !isfinite(a.out[n + 6]);
    if (!bad) {
      // β^t by repeated squaring (t is an integer): ≈ 2 log2(t) multiplications instead of two double-precision pow()
      double p1 = 1.0, p2 = 1.0, q1 = a.b1, q2 = a.b2;
      _Pragma("unroll 1") for (unsigned long long e = t_applied + 1; e > 0; e >>= 1) {
        if (e & 1) { p1 *= q1; p2 *= q2; }
        q1 *= q1; q2 *= q2;
      }
      const double c1 = 1.0 / (1.0 - p1), c2 = 1.0 / (1.0 - p2);
      for (int64_t i = t; i < n; i += SMALL_THREADS) {
        const double g = a.out[i];
        const double mi = a.b1 * a.adam_m[i] + (1.0 - a.b1) * g, vi = a.b2 * a.adam_v[i] + (1.0 - a.b2) * g * g;
        a.adam_m[i] = mi; a.adam_v[i] = vi;
        a.phi[i] = (T)((double)a.phi[i] - a.eta * (mi * c1) / (sqrt(vi * c2) + a.eps));
      }
    }
    __syncthreads();
    if (t == 0) {
      if (a.trace) a.trace[it] = bad ? nan("") : a.out[n + 6];
      if (!bad) a.ctr[0] = t_applied + 1;
      a.ctr[1] = it + 1;
      a.ctr[2] += 1;
    }
  }
  HVI_STAMP();
  __syncthreads();   // counters and ϕ of this iteration are visible to every thread of the block before the next one
  }
#undef HVI_STAMP
}

template <typename T>
static size_t small_iter_smem(const Cfg<T>& cfg, int M) {
  const size_t stream = stream_smem_bytes(sizeof(T), SMALL_THREADS / 32, cfg.nP, cfg.nM, cfg.nO,
                                          cfg.nP > 0 ? (size_t)M * pd_stride(cfg.nP) : 0, true);
  const size_t mlp = mlp_bwd_smem(cfg);
  return stream > mlp ? stream : mlp;
}

template <typename T>
bool small_iter_supported(const Cfg<T>& cfg, int64_t nb, int M) {
  if (cfg.npc > 0 || cfg.sepvar || nb < 1 || nb > 4096 || M < 1) return false;
  if (!stream_supported(cfg.nP, cfg.nM, cfg.nO)) return false;
  if (cfg.max_width > HVI_MAX_WIDTH) return false;
  if ((size_t)M * pd_stride(cfg.nP) * sizeof(T) > 16 * 1024) return false;
  return small_iter_smem(cfg, M) <= 200 * 1024;
}

template <typename T>
cudaError_t launch_small_iter(const Launch& L, const Cfg<T>& cfg, SmallArgs<T> a) {
  a.st.pd_in_smem = cfg.nP > 0 ? 1 : 0;
  a.st.staged = 0; a.st.rng = 1;
  size_t smem = small_iter_smem(cfg, a.st.M);
  if (a.st.nb < 128) {   // room for the (site, draw) form of the streaming phase behind the applicator's shared memory
    const size_t need = ((mlp_bwd_smem(cfg) + 15) & ~(size_t)15) +
                        stream_units_smem(sizeof(T), a.st.nb, a.st.M, cfg.nM + tri(cfg.nM) + (cfg.nP > 0 ? cfg.nP + tri(cfg.nP) : 0) + 5,
                                          nacc(cfg.nP, cfg.nM));
    if (need > smem && need <= 200 * 1024) smem = need;
  }
  // bytes of shared memory the (site, draw) form of the streaming phase may use (0: thread-per-site form, HVI_SMALL_UNITS=0)
  { const char* e = getenv("HVI_SMALL_UNITS"); a.units_mode = (e && atoi(e) == 0) ? 0 : (int)smem; }
#define X(P, M_, O)                                                                              \
  if (cfg.nP == P && cfg.nM == M_ && cfg.nO == O) {                                              \
    auto kern = k_small_iter<DynTraits, T, P, M_, O>;                                            \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                              \
    kern<<<1, SMALL_THREADS, smem, L.stream>>>(cfg, a);                                          \
    HVI_LAUNCH_CHECK();                                                                          \
    return cudaSuccess;                                                                          \
  }
  HVI_STREAM_CASES(X)
#undef X
  return cudaErrorInvalidConfiguration;
}

#define HVI_INSTANTIATE(T)                                                                                              \
  template cudaError_t launch_preprocess<T>(const Launch&, int, int64_t, const T*, const T*, const T*, T*, double*, int*); \
  template cudaError_t launch_prologue<T>(const Launch&, const Cfg<T>&, const T*, const T*, int, EvalConsts<T>*, T*);    \
  template cudaError_t launch_gather_preprocess<T>(const Launch&, int, int, int64_t, const int64_t*, int64_t, const T*,  \
                                                   const T*, const T*, const T*, T*, T*, T*, double*, int*,               \
                                                   const unsigned long long*, int64_t, int64_t*);                         \
  template cudaError_t launch_mlp_fwd<T>(const Launch&, const Cfg<T>&, const T*, const T*, int64_t, T*, int, const T*);  \
  template cudaError_t launch_mlp_bwd<T>(const Launch&, const Cfg<T>&, const T*, const T*, const T*, int64_t, double*,   \
                                         int*, int, int, const T*, double*, const T*);                                   \
  template cudaError_t launch_finalize<T>(const Launch&, const Cfg<T>&, const T*, const T*, const T*,                    \
                                          const EvalConsts<T>*, const double*, int, const double*, int, const double*, int64_t, \
                                          int, double*, T*, double*, const double*, int, const double*, int, double*);                                                             \
  template cudaError_t launch_gen_normal<T>(const Launch&, T*, int64_t, int, uint64_t, int64_t, const unsigned long long*); \
  template cudaError_t launch_generate_zeta<T>(const Launch&, const Cfg<T>&, const StreamArgs<T>&, const T*, T*, T*, T*); \
  template bool small_iter_supported<T>(const Cfg<T>&, int64_t, int);                                                    \
  template cudaError_t launch_small_iter<T>(const Launch&, const Cfg<T>&, SmallArgs<T>);
HVI_INSTANTIATE(float)
HVI_INSTANTIATE(double)

}  // namespace hvi
