// hvi_stream.cu -- k_stream, the dominant kernel of the negative-ELBO + gradient path, and its launch shapes
#include <stdlib.h>

#include "hvi_stream.cuh"

namespace hvi {

// site-record prefetch only where it does not cost a resident block (64 draws of per-draw records assumed)
template <typename T, int NP, int NM, int NO, int STREAM_THREADS, int MINB>
__host__ __device__ constexpr bool stream_spf() {
  return (size_t)MINB * stream_smem_bytes(sizeof(T), STREAM_THREADS / 32, NP, NM, NO, (size_t)64 * pd_stride(NP), true) <= (size_t)227 * 1024;
}

template <class TR, typename T, int NP, int NM, int NO, int STREAM_THREADS, int MINB, bool PBM, bool RNG, bool ILVP = true>
__global__ void __launch_bounds__(STREAM_THREADS, MINB) k_stream(const __grid_constant__ Cfg<T> cfg,
                                                                 const __grid_constant__ StreamArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // two draws per pass of the observation loop only where the register budget carries it (Float32, one block per SM)
  constexpr bool ILV = ILVP && sizeof(T) == 4 && MINB == 1;
  stream_body<TR, T, NP, NM, NO, STREAM_THREADS, PBM, RNG, (NP > 0), ILV, stream_spf<T, NP, NM, NO, STREAM_THREADS, MINB>()>(
      cfg, a, smem_raw, blockIdx.x, gridDim.x);
}

int stream_max_rows(int sm_count) { return sm_count * 8 * (STREAM_THREADS_MAX / 32); }

// largest n_MC whose per-draw records of the global parameters fit into the shared memory of one block per SM
int stream_max_draws(size_t szT, int nP, int nM, int nO) {
  if (nP == 0) return 1 << 30;
  const size_t fixed = stream_smem_bytes(szT, 8, nP, nM, nO, 0, true);
  return (int)(((size_t)227 * 1024 - fixed) / (pd_stride(nP) * szT));
}

template <class TR, typename T, int NP, int NM, int NO, int STREAM_THREADS, int MINB, bool PBM, bool RNG, bool ILVP = true>
static cudaError_t launch_stream_cfg2(const Launch& L, const Cfg<T>& cfg, StreamArgs<T> a, int* nrows_out, int nrows_max) {
  constexpr int NWARP = STREAM_THREADS / 32;
  constexpr bool SPF = stream_spf<T, NP, NM, NO, STREAM_THREADS, MINB>();
  a.pd_in_smem = NP > 0 ? 1 : 0;
  const size_t smem = stream_smem_bytes(sizeof(T), NWARP, NP, NM, NO, NP > 0 ? (size_t)a.M * pd_stride(NP) : 0, SPF);
  if (smem > (size_t)227 * 1024) return cudaErrorInvalidConfiguration;   // n_MC beyond stream_max_draws
  auto kern = k_stream<TR, T, NP, NM, NO, STREAM_THREADS, MINB, PBM, RNG, ILVP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, STREAM_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const int64_t ntiles = (a.nb + 31) / 32;
  int64_t blocks = (ntiles + NWARP - 1) / NWARP;
  if (blocks > (int64_t)L.sm_count * occ) blocks = (int64_t)L.sm_count * occ;
  if (blocks > nrows_max) blocks = nrows_max;
  if (blocks < 1) blocks = 1;
  kern<<<(int)blocks, STREAM_THREADS, smem, L.stream>>>(cfg, a);
  HVI_LAUNCH_CHECK();
  *nrows_out = (int)blocks;
  return cudaSuccess;
}

template <class TR, typename T, int NP, int NM, int NO, int STREAM_THREADS, int MINB, bool PBM, bool ILVP = true>
static cudaError_t launch_stream_cfg(const Launch& L, const Cfg<T>& cfg, StreamArgs<T> a, int* nrows_out, int nrows_max) {
  if (a.rng) return launch_stream_cfg2<TR, T, NP, NM, NO, STREAM_THREADS, MINB, PBM, true, ILVP>(L, cfg, a, nrows_out, nrows_max);
  return launch_stream_cfg2<TR, T, NP, NM, NO, STREAM_THREADS, MINB, PBM, false, ILVP>(L, cfg, a, nrows_out, nrows_max);
}

// launch shape (Float32, compile-time traits, 1e7 sites x 64 draws): one block of 384 threads per SM (168 registers,
// three warps per scheduler) 3.65 ms; one block of 512 (128 registers) 3.70 ms; two blocks of 256 (128 registers)
// 3.87 ms; one block of 256 with up to 255 registers 3.99 ms, 3.91 ms with two draws per pass of the observation loop.
// Until the per-tile and per-stage overheads were cut (cross-tile prefetch, one-pass accumulator reduction, unguarded
// cp.async addresses) the order was the reverse: registers for interleaving draws were worth more than resident
// warps.  HVI_STREAM_VARIANT = 2, 4, 5, 6, 7 selects the other shapes for experiments.
static int stream_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HVI_STREAM_VARIANT"); v = e ? atoi(e) : 0; }
  return v;
}

template <class TR, typename T, int NP, int NM, int NO>
static cudaError_t launch_stream_inst(const Launch& L, const Cfg<T>& cfg, const StreamArgs<T>& a, int* nrows_out,
                                      int nrows_max) {
  const bool pbm = a.MU != nullptr;
  if constexpr (sizeof(T) == 4) {
    if constexpr (TR::STATIC) {
      if (pbm) return launch_stream_cfg<TR, T, NP, NM, NO, 256, 2, true>(L, cfg, a, nrows_out, nrows_max);
      if (stream_variant() == 2) return launch_stream_cfg<TR, T, NP, NM, NO, 256, 2, false>(L, cfg, a, nrows_out, nrows_max);
      if (stream_variant() == 4) return launch_stream_cfg<TR, T, NP, NM, NO, 256, 1, false, false>(L, cfg, a, nrows_out, nrows_max);
      if (stream_variant() == 5) return launch_stream_cfg<TR, T, NP, NM, NO, 384, 1, false, true>(L, cfg, a, nrows_out, nrows_max);
      if (stream_variant() == 6) return launch_stream_cfg<TR, T, NP, NM, NO, 512, 1, false, false>(L, cfg, a, nrows_out, nrows_max);
      if (stream_variant() == 7) return launch_stream_cfg<TR, T, NP, NM, NO, 256, 1, false, true>(L, cfg, a, nrows_out, nrows_max);
      return launch_stream_cfg<TR, T, NP, NM, NO, 384, 1, false, false>(L, cfg, a, nrows_out, nrows_max);
    } else {
      if (pbm) return launch_stream_cfg<TR, T, NP, NM, NO, 256, 2, true>(L, cfg, a, nrows_out, nrows_max);
      return launch_stream_cfg<TR, T, NP, NM, NO, 256, 2, false>(L, cfg, a, nrows_out, nrows_max);   // run-time traits: 1.41 ms vs 1.55 ms with 256 x 1
    }
  } else {
    if (pbm) return launch_stream_cfg<TR, T, NP, NM, NO, 256, 1, true>(L, cfg, a, nrows_out, nrows_max);
    return launch_stream_cfg<TR, T, NP, NM, NO, 256, 1, false>(L, cfg, a, nrows_out, nrows_max);
  }
}


bool stream_supported(int nP, int nM, int nO) {
#define X(P, M_, O) if (nP == P && nM == M_ && nO == O) return true;
  HVI_STREAM_CASES(X)
#undef X
  return false;
}

// compile-time specialisations: (slots, bijectors of θM, priors of θM) of the named problems
//   DoubleMMCase default (src/DoubleMM/f_doubleMM.jl:161-243): θP=(r0,K2), θM=(r1,K1), Exp, LogNormal
//   test/test_HybridProblem.jl:32-111: same slots, transM=(identity, Exp), priors (Normal, LogNormal)
using TraitsDoubleMM = StaTraits<0x2431u, 0x5u, 0xAu>;
using TraitsDoubleMMNoPrior = StaTraits<0x2431u, 0x5u, 0x0u>;
using TraitsTestHybridProblem = StaTraits<0x2431u, 0x4u, 0x9u>;

template <typename T>
static void traits_key(const Cfg<T>& cfg, uint32_t& slots, uint32_t& trm, uint32_t& prm) {
  slots = trm = prm = 0;
  for (int s = 0; s < 4; s++) slots |= (uint32_t)(cfg.slot_code[s] + 1) << (4 * s);
  for (int k = 0; k < cfg.nM; k++) {
    trm |= (uint32_t)cfg.transM[k] << (2 * k);
    prm |= (uint32_t)(cfg.omit_priors ? 0 : cfg.prM_kind[k]) << (2 * k);
  }
}

template <typename T>
cudaError_t launch_stream(const Launch& L, const Cfg<T>& cfg, const StreamArgs<T>& a, int* nrows_out, int nrows_max) {
  if (cfg.nP == 2 && cfg.nM == 2 && cfg.nO == 8) {
    uint32_t slots, trm, prm;
    traits_key(cfg, slots, trm, prm);
#define Y(TRAITS) \
    if (slots == TRAITS::slots && trm == TRAITS::trm && prm == TRAITS::prm) \
      return launch_stream_inst<TRAITS, T, 2, 2, 8>(L, cfg, a, nrows_out, nrows_max);
    Y(TraitsDoubleMM) Y(TraitsDoubleMMNoPrior) Y(TraitsTestHybridProblem)
#undef Y
  }
#define X(P, M_, O) \
  if (cfg.nP == P && cfg.nM == M_ && cfg.nO == O) \
    return launch_stream_inst<DynTraits, T, P, M_, O>(L, cfg, a, nrows_out, nrows_max);
  HVI_STREAM_CASES(X)
#undef X
  return cudaErrorInvalidConfiguration;
}

template cudaError_t launch_stream<float>(const Launch&, const Cfg<float>&, const StreamArgs<float>&, int*, int);
template cudaError_t launch_stream<double>(const Launch&, const Cfg<double>&, const StreamArgs<double>&, int*, int);

}  // namespace hvi
