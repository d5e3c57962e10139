// hvi_dense_tma.cu -- the wide applicator (BASELINE config 5: 4x512 hidden) as a chain of TMA-fed,
// warp-specialised, persistent tcgen05 GEMMs with error-compensated bf16 operands ("3xBF16").
//
// Data layout in HBM.  Every activation H_l and every cotangent δ_l of a chunk of columns is kept as
// two bf16 planes [row][feature], hi = bf16(x) and lo = bf16(x − hi) (16 mantissa bits together, the
// footprint of one fp32 array).  The epilogue that produces a tensor writes the planes, so no GEMM
// converts or splits anything on its way into shared memory: TMA (cp.async.bulk.tensor, SWIZZLE_128B)
// drops the tiles where tcgen05.mma reads them.  The layer and data-gradient GEMMs contract over the
// features (operands K-major); the weight-gradient GEMM contracts over the rows and reads the SAME
// planes as MN-major operands (64-feature x 8-row swizzle atoms, a_major = b_major = 1 in the
// instruction descriptor), so nothing is ever transposed in memory.
//
// Kernel k_gemm_bf3<BN, MODE>:  D[M x N] = Σ_k A[m][k]·B[n][k]  with A ≈ A_hi + A_lo, B ≈ B_hi + B_lo and
// three MMAs per K step (hi·hi + hi·lo + lo·hi) into one fp32 TMEM accumulator.
//   warp 0 (one lane)  TMA producer: A hi/lo (128 x 64) and B hi/lo (BN x 64) per stage → full barrier
//   warp 1 (one lane)  MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M = 128, N = BN, K = 16;
//                      tcgen05.commit frees the stage (empty barrier) and publishes the accumulator
//   warps 2-9          epilogue: tcgen05.ld of one of the TWO accumulators (2·BN TMEM columns) while the
//                      MMA warp fills the other one for the CTA's next tile (persistent, static schedule);
//                      two warps per TMEM lane quarter, each half of the columns
//   MODE 0  H = act(D + bias)                          layer forward
//   MODE 1  δ_{l-1} = D ⊙ act'(H_{l-1})                data gradient (D = δ_l W_l)
//   MODE 2  P[split] = D over this split's K range     weight gradient, split over the rows
//
// Reference: the dense layers of g and their Zygote pullbacks
// (ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52, ext/HybridVariationalInferenceFluxExt.jl:26-31).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hvi_internal.cuh"
#include "hvi_tc_common.cuh"

namespace hvi {

typedef __nv_bfloat16 bf16;

constexpr int G_BM = 128;            // rows per tile = TMEM lanes
constexpr int G_BK = 64;             // K per stage: 64 bf16 = one 128-byte swizzle row
constexpr int G_EPI_WARPS = 8;
constexpr int G_THREADS = 64 + 32 * G_EPI_WARPS;   // TMA warp, MMA warp, epilogue warps

template <int BN>
struct GemmTile {
  static constexpr int A_BYTES = G_BM * 128;                       // one plane of A per stage (16 KB)
  static constexpr int B_BYTES = BN * 128;                         // one plane of B per stage
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);      // hi + lo of both
  static constexpr int STAGES = BN == 256 ? 2 : 3;                 // 192 KB of shared memory either way
  static constexpr int TMEM_COLS = 2 * BN;                         // two accumulators
  // D = F32 [4,6) = 1, A = B = BF16 [7,10) = [10,13) = 1, a_major [15], b_major [16] (1 = MN-major), N >> 3 [17,23), M >> 4 [24,29)
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(G_BM >> 4) << 24);
  static constexpr uint32_t IDESC_MN = IDESC | (1u << 15) | (1u << 16);
};
constexpr int S_LD = 64;             // width of the skinny operand planes (inputs + ones, δ of the last layer)
constexpr int G_MN_BOX = 64 * 128;   // bytes of one MN-major TMA box: 64 rows (K) x 64 features

// MN-major SWIZZLE_128B descriptor: atoms of 64 features (128 B) x 8 rows; LBO = distance between 64-feature
// chunks (one TMA box), SBO = distance between 8-row groups (1024 B)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(G_MN_BOX >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// the two planes of one tensor, [rows][ld]
struct Planes {
  bf16 *hi, *lo;
  int64_t ld;
};

struct GemmShape {
  int64_t M;             // rows of D (MODE 0/1: rows of A; MODE 2: features of A)
  int N;                 // columns of D (MODE 0/1: rows of B; MODE 2: features of B)
  int64_t K;             // contraction length (MODE 2: the rows of both operands)
  int nsplit;            // MODE 2: K is cut into nsplit ranges of k_per_split
  int64_t k_per_split;
  int dbg;               // experiments (HVI_GEMM_DBG): 1 no stores, 2 no bias loads, 4 no TMEM loads
};

struct GemmEpi {
  const float* bias;     // MODE 0 (nullable)
  int act;               // MODE 0: activation; MODE 1: activation of the layer below
  Planes out;            // MODE 0/1
  const bf16 *h_hi, *h_lo;   // MODE 1: activation of the layer below, normal planes
  int64_t ldh;
  float* P;              // MODE 2: [split][M][N]
  float* colsum;         // MODE 1 (nullable): [row group of 32][N] column sums of the output (bias gradient of the layer below)
};

// ---- PTX ---------------------------------------------------------------------------------
// bounded wait: a protocol error traps after ~2 s instead of hanging the device
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c_inner, int c_outer, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- plane helpers -------------------------------------------------------------------------
// 256-bit global accesses (sm_100): one full 32-byte sector per lane and instruction
__device__ __forceinline__ void stg256(void* p, const uint32_t* a) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]),
               "r"(a[5]), "r"(a[6]), "r"(a[7])
               : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint32_t* a) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7])
               : "l"(p));
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// 32 consecutive features of one row (this lane's), fp32 → hi/lo planes: 64 contiguous bytes per lane and plane
__device__ __forceinline__ void store_planes32(const Planes& o, int64_t row, int col0, const float (&x)[32], bool ok) {
  uint32_t h[16], l[16];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * j] - __low2float(hh), x[2 * j + 1] - __high2float(hh));
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
    l[j] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  if (!ok) return;
  bf16* ph = o.hi + row * o.ld + col0;
  bf16* pl = o.lo + row * o.ld + col0;
  stg256(ph, h); stg256(ph + 16, h + 8);
  stg256(pl, l); stg256(pl + 16, l + 8);
}
// tanh to ~2e-7 absolute with two MUFU ops: 1 − 2/(1 + e^{2x})
__device__ __forceinline__ float tanh_fast(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}
// raw hi/lo words of 32 consecutive features of one row (issue early, combine later)
__device__ __forceinline__ void load_planes32_raw(const bf16* hi, const bf16* lo, int64_t ld, int64_t row, int col0, uint32_t (&a)[16],
                                                  uint32_t (&b)[16]) {
  ldg256(hi + row * ld + col0, a); ldg256(hi + row * ld + col0 + 16, a + 8);
  ldg256(lo + row * ld + col0, b); ldg256(lo + row * ld + col0 + 16, b + 8);
}
// 32 consecutive features of one row from the normal planes, recombined to fp32
__device__ __forceinline__ void load_planes32(const bf16* hi, const bf16* lo, int64_t ld, int64_t row, int col0, float (&x)[32]) {
  uint32_t a[16], b[16];
  ldg256(hi + row * ld + col0, a); ldg256(hi + row * ld + col0 + 16, a + 8);
  ldg256(lo + row * ld + col0, b); ldg256(lo + row * ld + col0 + 16, b + 8);
#pragma unroll
  for (int j = 0; j < 16; j++) {
    x[2 * j] = bf_lo(a[j]) + bf_lo(b[j]);
    x[2 * j + 1] = bf_hi(a[j]) + bf_hi(b[j]);
  }
}

// Σ over the 32 lanes of 32 values per lane: at every level half of the values go to the partner lane, so the tree costs
// 31 shuffles; lane j ends with the total of value j (fixed order: deterministic).  Destroys v.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
  int n = 32;
#pragma unroll
  for (int b = 4; b >= 0; b--) {
    n >>= 1;
    const bool up = (lane >> b) & 1;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (i < n) {
        const float send = up ? v[i] : v[i + n];
        const float keep = up ? v[i + n] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << b);
      }
    }
  }
  return v[0];
}

// ---- the GEMM ------------------------------------------------------------------------------
template <int BN, int MODE>
__global__ void __launch_bounds__(G_THREADS, 1)
k_gemm_bf3(const __grid_constant__ CUtensorMap mA_hi, const __grid_constant__ CUtensorMap mA_lo,
           const __grid_constant__ CUtensorMap mB_hi, const __grid_constant__ CUtensorMap mB_lo, const GemmShape sh,
           const GemmEpi ep) {
  typedef GemmTile<BN> TL;
  constexpr int S = TL::STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[S], bar_empty[S], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

  const int m_tiles = (int)((sh.M + G_BM - 1) / G_BM), n_tiles = sh.N / BN;
  const int mn_tiles = m_tiles * n_tiles;
  const int total = mn_tiles * sh.nsplit;

  if (t == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], G_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TL::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int split = tile / mn_tiles, mn = tile % mn_tiles;
        const int m0 = (mn / n_tiles) * G_BM, n0 = (mn % n_tiles) * BN;
        const int64_t k0 = (int64_t)split * sh.k_per_split;
        const int64_t k1 = k0 + sh.k_per_split < sh.K ? k0 + sh.k_per_split : sh.K;
        const int nk = (int)((k1 - k0 + G_BK - 1) / G_BK);
        for (int kb = 0; kb < nk; kb++, it++) {
          const int s = it % S;
          mbar_wait_b(&bar_empty[s], ((it / S) & 1) ^ 1);
          mbar_expect_tx(&bar_full[s], TL::STAGE_BYTES);
          const uint32_t base = smem_u32(tiles + (size_t)s * TL::STAGE_BYTES);
          const int kc = (int)(k0 + (int64_t)kb * G_BK);
          if (MODE != 2) {      // K-major: one box per plane, [tile rows][64 K values]
            tma_load_2d(base, &mA_hi, kc, m0, &bar_full[s]);
            tma_load_2d(base + TL::A_BYTES, &mA_lo, kc, m0, &bar_full[s]);
            tma_load_2d(base + 2 * TL::A_BYTES, &mB_hi, kc, n0, &bar_full[s]);
            tma_load_2d(base + 2 * TL::A_BYTES + TL::B_BYTES, &mB_lo, kc, n0, &bar_full[s]);
          } else {              // MN-major: boxes of [64 rows (K)][64 features], one per 64 features of the tile
#pragma unroll
            for (int b = 0; b < G_BM / 64; b++) {
              tma_load_2d(base + b * G_MN_BOX, &mA_hi, m0 + 64 * b, kc, &bar_full[s]);
              tma_load_2d(base + TL::A_BYTES + b * G_MN_BOX, &mA_lo, m0 + 64 * b, kc, &bar_full[s]);
            }
#pragma unroll
            for (int b = 0; b < BN / 64; b++) {
              tma_load_2d(base + 2 * TL::A_BYTES + b * G_MN_BOX, &mB_hi, n0 + 64 * b, kc, &bar_full[s]);
              tma_load_2d(base + 2 * TL::A_BYTES + TL::B_BYTES + b * G_MN_BOX, &mB_lo, n0 + 64 * b, kc, &bar_full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t it = 0, tc = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, tc++) {
        const int split = tile / mn_tiles;
        const int64_t k0 = (int64_t)split * sh.k_per_split;
        const int64_t k1 = k0 + sh.k_per_split < sh.K ? k0 + sh.k_per_split : sh.K;
        const int nk = (int)((k1 - k0 + G_BK - 1) / G_BK);
        const uint32_t acc = tc & 1;
        mbar_wait_b(&bar_tempty[acc], ((tc >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + acc * BN;
        for (int kb = 0; kb < nk; kb++, it++) {
          const int s = it % S;
          mbar_wait_b(&bar_full[s], (it / S) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_u32(tiles + (size_t)s * TL::STAGE_BYTES), a_lo = a_hi + TL::A_BYTES,
                         b_hi = a_hi + 2 * TL::A_BYTES, b_lo = b_hi + TL::B_BYTES;
#pragma unroll
          for (int ks = 0; ks < G_BK / 16; ks++) {
            const uint32_t first = (kb > 0 || ks > 0) ? 1u : 0u;
            if (MODE != 2) {     // UMMA_K = 16 bf16 = 32 bytes along the swizzled 128-byte row
              const uint32_t o = ks * 32;
              umma_bf16(d, make_desc(a_hi + o), make_desc(b_hi + o), TL::IDESC, first);
              umma_bf16(d, make_desc(a_hi + o), make_desc(b_lo + o), TL::IDESC, 1u);
              umma_bf16(d, make_desc(a_lo + o), make_desc(b_hi + o), TL::IDESC, 1u);
            } else {             // 16 rows = two 8-row groups of 1024 bytes
              const uint32_t o = ks * 2048;
              umma_bf16(d, make_desc_mn(a_hi + o), make_desc_mn(b_hi + o), TL::IDESC_MN, first);
              umma_bf16(d, make_desc_mn(a_hi + o), make_desc_mn(b_lo + o), TL::IDESC_MN, 1u);
              umma_bf16(d, make_desc_mn(a_lo + o), make_desc_mn(b_hi + o), TL::IDESC_MN, 1u);
            }
          }
          umma_commit(&bar_empty[s]);               // stage free once these MMAs have read it
        }
        umma_commit(&bar_tfull[acc]);               // accumulator complete
      }
    }
  } else {
    // ===== epilogue: warp w may only touch TMEM lanes 32·(w mod 4) …; two warps per quarter split the columns =====
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    uint32_t tc = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, tc++) {
      const int split = tile / mn_tiles, mn = tile % mn_tiles;
      const int m0 = (mn / n_tiles) * G_BM, n0 = (mn % n_tiles) * BN;
      const int64_t k0 = (int64_t)split * sh.k_per_split;
      const bool empty_k = k0 >= sh.K;
      const uint32_t acc = tc & 1;
      const int64_t row = (int64_t)m0 + q * 32 + lane;
      const bool ok = row < sh.M;
      // MODE 1: the activation planes of the first block are requested before the accumulator is awaited,
      // those of block b+1 while block b is converted and stored
      uint32_t ha[16], hb[16];
      if (MODE == 1 && ok) load_planes32_raw(ep.h_hi, ep.h_lo, ep.ldh, row, n0 + chalf * (BN / 2), ha, hb);
      mbar_wait_b(&bar_tfull[acc], (tc >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cb = chalf * (BN / 2); cb < (chalf + 1) * (BN / 2); cb += 32) {
        uint32_t v[32];
        if (!(sh.dbg & 4)) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cb, v);
        else {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = j + lane;
        }
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; j++) x[j] = __uint_as_float(v[j]);
        if (empty_k) {
#pragma unroll
          for (int j = 0; j < 32; j++) x[j] = 0.f;
        }
        const int col0 = n0 + cb;
        if (MODE == 0) {
          if (ep.bias && !(sh.dbg & 2)) {   // the 32 biases of this block: four 256-bit loads, same address in every lane
            uint32_t bb[32];
#pragma unroll
            for (int g = 0; g < 4; g++) ldg256(ep.bias + col0 + 8 * g, bb + 8 * g);
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] += __uint_as_float(bb[j]);
          }
          // the activation is chosen once per block, not per element
          if (ep.act == HVI_ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = tanh_fast(x[j]);
          } else if (ep.act == HVI_ACT_LOGISTIC) {
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = __fdividef(1.f, 1.f + __expf(-x[j]));
          }
          store_planes32(ep.out, row, col0, x, ok && (!(sh.dbg & 1) || x[3] == 1234.5f));
        } else if (MODE == 1) {
          if (ep.act == HVI_ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const float h0 = bf_lo(ha[j]) + bf_lo(hb[j]), h1 = bf_hi(ha[j]) + bf_hi(hb[j]);
              x[2 * j] *= fmaf(-h0, h0, 1.f);
              x[2 * j + 1] *= fmaf(-h1, h1, 1.f);
            }
          } else if (ep.act == HVI_ACT_LOGISTIC) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const float h0 = bf_lo(ha[j]) + bf_lo(hb[j]), h1 = bf_hi(ha[j]) + bf_hi(hb[j]);
              x[2 * j] *= h0 * (1.f - h0);
              x[2 * j + 1] *= h1 * (1.f - h1);
            }
          }
          if (ok && cb + 32 < (chalf + 1) * (BN / 2)) load_planes32_raw(ep.h_hi, ep.h_lo, ep.ldh, row, col0 + 32, ha, hb);
          store_planes32(ep.out, row, col0, x, ok);
          if (ep.colsum) {   // column sums over this warp's 32 rows (rows beyond M count as zero)
            if (!ok) {
#pragma unroll
              for (int j = 0; j < 32; j++) x[j] = 0.f;
            }
            const float cs = warp_colsum32(x, lane);
            ep.colsum[(size_t)(m0 / 32 + q) * sh.N + col0 + lane] = cs;
          }
        } else if (ok) {
          float4* out = reinterpret_cast<float4*>(ep.P + ((size_t)split * sh.M + row) * sh.N + col0);
#pragma unroll
          for (int j = 0; j < 8; j++) out[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[acc]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TL::TMEM_COLS));
}

// ---- host side: tensor maps and the launcher -------------------------------------------------
static PFN_cuTensorMapEncodeTiled tma_encoder() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }
  return fn;
}
// bf16 matrix [outer][ld] with `inner` valid elements per row, boxes of 64 x box_outer, 128-byte swizzle;
// out-of-range elements read as zero
static bool make_map(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  PFN_cuTensorMapEncodeTiled enc = tma_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {64, box_outer};
  const cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// row splits of a weight-gradient GEMM: one CTA per (tile, split); at most 24 full tiles of partials
// (the work space holds 24·maxw² floats), any number when the left operand is skinny (Md ≤ S_LD)
int bf3_wgrad_splits(int Md, int Nd, int sm_count) {
  const int bn = Nd % 256 == 0 ? 256 : 128;
  const int tiles = ((Md + G_BM - 1) / G_BM) * (Nd / bn);
  int ns = sm_count / (tiles > 0 ? tiles : 1);
  const int cap = Md <= 64 ? 148 : 24;
  return ns < 1 ? 1 : (ns > cap ? cap : ns);
}
// the splits launch_gemm_bf3<2> really uses for a contraction of K rows (every split a multiple of 64, none empty)
static int effective_splits(int64_t K, int nsplit, int64_t* per_out) {
  int64_t per = (K + nsplit - 1) / nsplit;
  per = (per + G_BK - 1) / G_BK * G_BK;
  if (per_out) *per_out = per;
  return (int)((K + per - 1) / per);
}

// MODE 0/1: D[M x N] = A·Bᵀ with A [M][lda], B [N][ldb], K features per row (K-major operands).
// MODE 2:   D[M x N] = AᵀB  with A [K][lda] (M features per row), B [K][ldb] (N features per row): MN-major operands.
// N % 128 == 0, lda/ldb % 8 == 0, 16-byte aligned bases.
template <int MODE>
static cudaError_t launch_gemm_bf3(const Launch& L, const bf16* A_hi, const bf16* A_lo, int64_t lda, int64_t M, const bf16* B_hi,
                                   const bf16* B_lo, int64_t ldb, int N, int64_t K, int nsplit, const GemmEpi& ep) {
  if (N % 128 != 0 || lda % 8 != 0 || ldb % 8 != 0 || M < 1 || K < 1) return cudaErrorInvalidValue;
  if (MODE == 0 && ep.bias && (reinterpret_cast<uintptr_t>(ep.bias) & 31)) return cudaErrorInvalidValue;   // 256-bit bias loads
  const int bn = N % 256 == 0 ? 256 : 128;
  GemmShape sh;
  sh.M = M; sh.N = N; sh.K = K;
  sh.nsplit = effective_splits(K, nsplit < 1 ? 1 : nsplit, &sh.k_per_split);
  { const char* d = getenv("HVI_GEMM_DBG"); sh.dbg = d ? atoi(d) : 0; }
  CUtensorMap ma_h, ma_l, mb_h, mb_l;
  bool ok;
  if (MODE != 2)
    ok = make_map(&ma_h, A_hi, (uint64_t)K, (uint64_t)M, (uint64_t)lda, G_BM) && make_map(&ma_l, A_lo, (uint64_t)K, (uint64_t)M, (uint64_t)lda, G_BM) &&
         make_map(&mb_h, B_hi, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, (uint32_t)bn) &&
         make_map(&mb_l, B_lo, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, (uint32_t)bn);
  else
    ok = make_map(&ma_h, A_hi, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64) && make_map(&ma_l, A_lo, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64) &&
         make_map(&mb_h, B_hi, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64) && make_map(&mb_l, B_lo, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64);
  if (!ok) return cudaErrorInvalidValue;
  const int m_tiles = (int)((M + G_BM - 1) / G_BM);
  const int64_t total = (int64_t)m_tiles * (N / bn) * sh.nsplit;
  const int sms = L.sm_count > 0 ? L.sm_count : 148;
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  cudaError_t e;
  if (bn == 256) {
    const size_t smem = (size_t)GemmTile<256>::STAGES * GemmTile<256>::STAGE_BYTES + 1024;
    auto kern = k_gemm_bf3<256, MODE>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, G_THREADS, smem, L.stream>>>(ma_h, ma_l, mb_h, mb_l, sh, ep);
  } else {
    const size_t smem = (size_t)GemmTile<128>::STAGES * GemmTile<128>::STAGE_BYTES + 1024;
    auto kern = k_gemm_bf3<128, MODE>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, G_THREADS, smem, L.stream>>>(ma_h, ma_l, mb_h, mb_l, sh, ep);
  }
  if (L.counter) ++*L.counter;
  return cudaGetLastError();
}

// ---- CUDA-core kernels around the GEMM chain -------------------------------------------------
// weights of one layer, reference layout [in][out] fp32 → planes in the same layout (B operand of the
// data-gradient GEMM: rows = in, K = out) and transposed [out][in] (B operand of the forward GEMM)
__global__ void k_split_w(const float* __restrict__ W, int n_in, int n_out, bf16* __restrict__ n_hi, bf16* __restrict__ n_lo,
                          bf16* __restrict__ t_hi, bf16* __restrict__ t_lo) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_in * n_out) return;
  const int i = e / n_out, j = e % n_out;
  unsigned short h, l;
  bf16_split(W[e], h, l);
  reinterpret_cast<unsigned short*>(n_hi)[e] = h;
  reinterpret_cast<unsigned short*>(n_lo)[e] = l;
  reinterpret_cast<unsigned short*>(t_hi)[(size_t)j * n_in + i] = h;
  reinterpret_cast<unsigned short*>(t_lo)[(size_t)j * n_in + i] = l;
}

// first layer as a GEMM over the S_LD-wide input planes: B[j][i] = W_0[i][j] (i < K0), b_0[j] (i = K0, the ones
// column of Sx), 0 otherwise
__global__ void k_split_w0(const float* __restrict__ W0, const float* __restrict__ b0, int K0, int N, bf16* __restrict__ hi,
                           bf16* __restrict__ lo) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * S_LD) return;
  const int j = e / S_LD, i = e % S_LD;
  const float v = i < K0 ? W0[(size_t)i * N + j] : (i == K0 && b0) ? b0[j] : 0.f;
  unsigned short h, l;
  bf16_split(v, h, l);
  reinterpret_cast<unsigned short*>(hi)[e] = h;
  reinterpret_cast<unsigned short*>(lo)[e] = l;
}

// fp32 [rows][W] → planes (test harness); one warp per 32 rows x 32 features
__global__ void __launch_bounds__(256) k_split_planes(const float* __restrict__ X, int64_t rows, int W, Planes o) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * 8 + warp) * 32 + lane;
  const int col0 = blockIdx.y * 32;
  const bool ok = row < rows;
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; j++) x[j] = ok ? X[row * W + col0 + j] : 0.f;
  store_planes32(o, row, col0, x, ok);
}
// planes → fp32 (test harness)
__global__ void k_join_planes(Planes p, int64_t rows, int W, float* __restrict__ X) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * W) return;
  const int64_t r = e / W;
  const int c = (int)(e % W);
  X[e] = __bfloat162float(p.hi[r * p.ld + c]) + __bfloat162float(p.lo[r * p.ld + c]);
}

// [x ; pbm covariates ; 1 ; 0 …] of every row of the chunk as hi/lo planes [rows][S_LD]: the skinny left operand
// of the first layer's weight gradient (rows 0 … K-1 → ∇W_0, row K → ∇b_0) and, through its ones column,
// of the hidden layers' bias gradients
__global__ void k_build_sx(const float* __restrict__ xM, int nC, const float* __restrict__ extra, int npc, int64_t site0, int64_t rows,
                           bf16* __restrict__ s_hi, bf16* __restrict__ s_lo) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  uint32_t h[S_LD / 2], l[S_LD / 2];
#pragma unroll
  for (int j = 0; j < S_LD / 2; j++) {
    float v[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int i = 2 * j + u;
      v[u] = i < nC ? xM[(site0 + r) * nC + i] : i < nC + npc ? extra[i - nC] : i == nC + npc ? 1.f : 0.f;
    }
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[0], v[1]);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[0] - __low2float(hh), v[1] - __high2float(hh));
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
    l[j] = *reinterpret_cast<const uint32_t*>(&ll);
  }
#pragma unroll
  for (int g = 0; g < S_LD / 16; g++) { stg256(s_hi + r * S_LD + 16 * g, h + 8 * g); stg256(s_lo + r * S_LD + 16 * g, l + 8 * g); }
}

// last layer (N = n_θM ≤ 8): a warp takes 32 rows; per row the lanes share the features (8 consecutive ones per
// lane and step) and a butterfly leaves the sums in lane (row mod 32), so that the output scaling (Φ⁻¹ in double)
// runs with all lanes busy.
// BWD == 0: μ_ζ = scale(act(z));  BWD == 1: δ_L[k] = μ̄·scale'·act' → dlast[r][8] and the planes [r][S_LD]
template <int BWD, int NOUT>
__global__ void __launch_bounds__(256) k_last_bf3(const Cfg<float> cfg, const bf16* __restrict__ hi, const bf16* __restrict__ lo, int64_t ld,
                                                  const float* __restrict__ Wref, int act, int64_t rows, int K,
                                                  float* __restrict__ mu, const float* __restrict__ mubar, int64_t site0, int m,
                                                  int M_units, int64_t nb, float* __restrict__ dlast, bf16* __restrict__ d_hi,
                                                  bf16* __restrict__ d_lo) {
  const int lane = threadIdx.x & 31;
  const int64_t r0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
  if (r0 >= rows) return;
  float mine[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; k++) mine[k] = 0.f;
  const int nr = rows - r0 < 32 ? (int)(rows - r0) : 32;
  // this lane's slice of W_L (features lane·8 … +7 of every 256): 8·NOUT consecutive floats per step, kept in
  // registers across the rows when the layer is at most 512 wide
  constexpr int WSTEPS = 2;
  uint32_t wv[WSTEPS][8 * NOUT];
  const bool wreg = K <= 256 * WSTEPS;
  if (wreg) {
#pragma unroll
    for (int st = 0; st < WSTEPS; st++)
      if (lane * 8 + 256 * st < K) {
#pragma unroll
        for (int g = 0; g < NOUT; g++) ldg256(Wref + (size_t)(lane * 8 + 256 * st) * NOUT + 8 * g, wv[st] + 8 * g);
      }
  }
  for (int i = 0; i < nr; i++) {
    const int64_t r = r0 + i;
    float acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; k++) acc[k] = 0.f;
    if (wreg) {
#pragma unroll
      for (int st = 0; st < WSTEPS; st++) {
        const int i0 = lane * 8 + 256 * st;
        if (i0 < K) {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + r * ld + i0)), b = __ldg(reinterpret_cast<const uint4*>(lo + r * ld + i0));
          const float x[8] = {bf_lo(a.x) + bf_lo(b.x), bf_hi(a.x) + bf_hi(b.x), bf_lo(a.y) + bf_lo(b.y), bf_hi(a.y) + bf_hi(b.y),
                              bf_lo(a.z) + bf_lo(b.z), bf_hi(a.z) + bf_hi(b.z), bf_lo(a.w) + bf_lo(b.w), bf_hi(a.w) + bf_hi(b.w)};
#pragma unroll
          for (int c = 0; c < 8; c++)
#pragma unroll
            for (int k = 0; k < NOUT; k++) acc[k] = fmaf(x[c], __uint_as_float(wv[st][c * NOUT + k]), acc[k]);
        }
      }
    } else {
      for (int i0 = lane * 8; i0 < K; i0 += 256) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + r * ld + i0)), b = __ldg(reinterpret_cast<const uint4*>(lo + r * ld + i0));
        const float x[8] = {bf_lo(a.x) + bf_lo(b.x), bf_hi(a.x) + bf_hi(b.x), bf_lo(a.y) + bf_lo(b.y), bf_hi(a.y) + bf_hi(b.y),
                            bf_lo(a.z) + bf_lo(b.z), bf_hi(a.z) + bf_hi(b.z), bf_lo(a.w) + bf_lo(b.w), bf_hi(a.w) + bf_hi(b.w)};
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
          for (int k = 0; k < NOUT; k++) acc[k] = fmaf(x[c], __ldg(Wref + (size_t)(i0 + c) * NOUT + k), acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
      float s = acc[k];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == i) mine[k] = s;
    }
  }
  if (lane >= nr) return;
  const int64_t r = r0 + lane, site = site0 + r;
  float d[8];
#pragma unroll
  for (int k = 0; k < 8; k++) d[k] = 0.f;
#pragma unroll
  for (int k = 0; k < NOUT; k++) {
    if (k >= cfg.nM) continue;
    const float p = act_apply<float>(act, mine[k]);
    if (BWD == 0) {
      const float v = (cfg.scale_kind == HVI_SCALE_RANGE) ? p * cfg.scale_b[k] + cfg.scale_a[k]
                                                         : cfg.scale_a[k] + cfg.scale_b[k] * (float)normcdfinv((double)p);
      if (M_units > 0) mu[((size_t)m * cfg.nM + k) * nb + site] = v;
      else mu[site * cfg.nM + k] = v;
    } else {
      const float mb = M_units > 0 ? mubar[((size_t)m * cfg.nM + k) * nb + site] : mubar[site * cfg.nM + k];
      float dv;
      if (cfg.scale_kind == HVI_SCALE_RANGE) dv = mb * cfg.scale_b[k];
      else { const float qn = (float)normcdfinv((double)p); dv = mb * cfg.scale_b[k] * 2.50662827463100050242f * expf(0.5f * qn * qn); }
      dv *= (act == HVI_ACT_LOGISTIC) ? p * (1.f - p) : (act == HVI_ACT_TANH) ? 1.f - p * p : 1.f;
      if (k < 8) d[k] = dv;
    }
  }
  if (BWD == 1) {
    float4* dl = reinterpret_cast<float4*>(dlast + r * 8);
    dl[0] = make_float4(d[0], d[1], d[2], d[3]);
    dl[1] = make_float4(d[4], d[5], d[6], d[7]);
    uint32_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(d[2 * j], d[2 * j + 1]);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(d[2 * j] - __low2float(hh), d[2 * j + 1] - __high2float(hh));
      h[j] = *reinterpret_cast<const uint32_t*>(&hh);
      l[j] = *reinterpret_cast<const uint32_t*>(&ll);
      h[4 + j] = l[4 + j] = 0u;
    }
    const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    stg256(d_hi + r * S_LD, h); stg256(d_lo + r * S_LD, l);
#pragma unroll
    for (int g = 1; g < S_LD / 16; g++) { stg256(d_hi + r * S_LD + 16 * g, zero); stg256(d_lo + r * S_LD + 16 * g, zero); }
  }
}

// δ_{L-1}[r][i] = (Σ_k W_L[i][k] δ_L[r][k]) · act'(H_{L-1}[r][i]); lane ↔ row, 64 features per warp
template <int NOUT>
__global__ void __launch_bounds__(256) k_thin_bwd_bf3(const float* __restrict__ dlast, const float* __restrict__ Wref,
                                                      const bf16* __restrict__ h_hi, const bf16* __restrict__ h_lo, int64_t ldh,
                                                      int act_prev, int64_t rows, Planes o) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * 8 + warp) * 32 + lane;
  const bool ok = row < rows;
  float dl[8];
#pragma unroll
  for (int k = 0; k < 8; k++) dl[k] = ok ? dlast[row * 8 + k] : 0.f;
#pragma unroll 1
  for (int cb = blockIdx.y * 64; cb < blockIdx.y * 64 + 64; cb += 32) {
    float h[32], z[32];
    if (ok) load_planes32(h_hi, h_lo, ldh, row, cb, h);
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += 8) {      // W_L[(cb+j)·n_out + k], 8 features at a time: n_out 256-bit loads, one address per warp
      uint32_t wv[8 * NOUT];
#pragma unroll
      for (int g = 0; g < NOUT; g++) ldg256(Wref + (size_t)(cb + j0) * NOUT + 8 * g, wv + 8 * g);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NOUT; k++) s = fmaf(__uint_as_float(wv[j * NOUT + k]), dl[k], s);
        z[j0 + j] = ok ? s * act_deriv_f(act_prev, h[j0 + j]) : 0.f;
      }
    }
    store_planes32(o, row, cb, z, ok);
  }
}

// column sums, stage 1: block b adds its share of the R partial rows (one per 32 rows of the GEMM output) for every column
__global__ void __launch_bounds__(256) k_colsum_rows(const float* __restrict__ cs, int R, int N, float* __restrict__ P2) {
  const int per = (R + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = r0 + per < R ? r0 + per : R;
  for (int n = threadIdx.x; n < N; n += 256) {
    float s = 0.f;
    for (int r = r0; r < r1; r++) s += cs[(size_t)r * N + n];
    P2[(size_t)blockIdx.x * N + n] = s;
  }
}

constexpr int LAST_STAGES = 4, LAST_STAGE_BYTES = 2048 + 32;   // ring of k_last_bwd_fused_bf3: rows in flight per warp
__device__ __forceinline__ void cp_async16_g(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4_g(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
// Last layer of the backward pass in ONE pass over H_{L-1} (K ≤ 512): a warp takes one row at a time, the lanes share its
// features (8 consecutive ones per lane and 256-feature step, as in k_last_bf3); z_L by a butterfly that leaves the sums
// in every lane, δ_L = μ̄·scale'·act' in every lane (Φ⁻¹ in single precision: it only enters through exp(q²/2)), then
// δ_{L-1} = (δ_L W_Lᵀ) ⊙ act'(H_{L-1}) written as hi/lo planes, and the sums over the rows
// ∇W_L[i][k] = Σ_r H_{L-1}[r][i]·δ_L[r][k] and ∇b_{L-1}[i] = Σ_r δ_{L-1}[r][i] in registers of the lane that owns
// feature i.  Replaces k_last_bf3<1> + k_thin_bwd_bf3 + two skinny tensor-core products (three more reads of H_{L-1}
// and one of δ_{L-1}).  The planes of the warp's next row travel while the current one is computed.  Persistent; one
// slice of partial sums per block: P[block][K·NOUT (∇W_L, index i·NOUT + k) | K (∇b_{L-1})], summed in fixed order by k_accum.
template <int NOUT>
__global__ void __launch_bounds__(256, 2) k_last_bwd_fused_bf3(const Cfg<float> cfg, const bf16* __restrict__ hi, const bf16* __restrict__ lo,
                                                            int64_t ld, const float* __restrict__ Wref, int act, int act_prev,
                                                            int64_t rows, int K, const float* __restrict__ mubar, int64_t site0,
                                                            int m, int M_units, int64_t nb, Planes o, float* __restrict__ P) {
  __shared__ float sacc[512 * (NOUT + 1)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int WSTEPS = 2;
  uint32_t wv[WSTEPS][8 * NOUT];
  float gW[WSTEPS][8][NOUT], gb[WSTEPS][8];
#pragma unroll
  for (int st = 0; st < WSTEPS; st++) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
      gb[st][c] = 0.f;
#pragma unroll
      for (int k = 0; k < NOUT; k++) gW[st][c][k] = 0.f;
    }
    if (lane * 8 + 256 * st < K) {
#pragma unroll
      for (int g = 0; g < NOUT; g++) ldg256(Wref + (size_t)(lane * 8 + 256 * st) * NOUT + 8 * g, wv[st] + 8 * g);
    } else {
#pragma unroll
      for (int g = 0; g < 8 * NOUT; g++) wv[st][g] = 0u;
    }
  }
  // per warp a ring of LAST_STAGES rows in shared memory (hi plane 1 KB | lo plane 1 KB | μ̄ of the row), filled with
  // cp.async: every lane copies the 16-byte chunks it will read itself, lanes k < n_θM the row's cotangents
  extern __shared__ __align__(16) unsigned char ring_raw[];
  unsigned char* ring = ring_raw + (size_t)warp * LAST_STAGES * LAST_STAGE_BYTES;
  const int64_t nw = (int64_t)gridDim.x * 8;
  auto issue = [&](int64_t r, int stage) {
    if (r < rows) {
      unsigned char* dst = ring + (size_t)stage * LAST_STAGE_BYTES;
#pragma unroll
      for (int st = 0; st < WSTEPS; st++) {
        const int i0 = lane * 8 + 256 * st;
        if (i0 < K) {
          cp_async16_g(dst + 2 * i0, hi + r * ld + i0);
          cp_async16_g(dst + 1024 + 2 * i0, lo + r * ld + i0);
        }
      }
      if (lane < NOUT && lane < cfg.nM) {
        const int64_t site = site0 + r;
        cp_async4_g(dst + 2048 + 4 * lane, M_units > 0 ? mubar + ((size_t)m * cfg.nM + lane) * nb + site : mubar + site * cfg.nM + lane);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int64_t r = (int64_t)blockIdx.x * 8 + warp;
#pragma unroll
  for (int sg = 0; sg < LAST_STAGES - 1; sg++) issue(r + sg * nw, sg);
  for (int it = 0; r < rows; r += nw, it++) {
    issue(r + (LAST_STAGES - 1) * nw, (it + LAST_STAGES - 1) % LAST_STAGES);
    asm volatile("cp.async.wait_group %0;" ::"n"(LAST_STAGES - 1) : "memory");
    __syncwarp();
    const unsigned char* src = ring + (size_t)(it % LAST_STAGES) * LAST_STAGE_BYTES;
    float mb[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; k++) mb[k] = k < cfg.nM ? *reinterpret_cast<const float*>(src + 2048 + 4 * k) : 0.f;
    float x[WSTEPS][8];
    float acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; k++) acc[k] = 0.f;
#pragma unroll
    for (int st = 0; st < WSTEPS; st++) {
      const int i0 = lane * 8 + 256 * st;
      uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
      if (i0 < K) {
        a = *reinterpret_cast<const uint4*>(src + 2 * i0);
        b = *reinterpret_cast<const uint4*>(src + 1024 + 2 * i0);
      }
      x[st][0] = bf_lo(a.x) + bf_lo(b.x); x[st][1] = bf_hi(a.x) + bf_hi(b.x);
      x[st][2] = bf_lo(a.y) + bf_lo(b.y); x[st][3] = bf_hi(a.y) + bf_hi(b.y);
      x[st][4] = bf_lo(a.z) + bf_lo(b.z); x[st][5] = bf_hi(a.z) + bf_hi(b.z);
      x[st][6] = bf_lo(a.w) + bf_lo(b.w); x[st][7] = bf_hi(a.w) + bf_hi(b.w);
#pragma unroll
      for (int c = 0; c < 8; c++)
#pragma unroll
        for (int k = 0; k < NOUT; k++) acc[k] = fmaf(x[st][c], __uint_as_float(wv[st][c * NOUT + k]), acc[k]);
    }
    // z_L in every lane, then δ_L[k] = μ̄·scale'·act'
    float dl[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
      float sum = acc[k];
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, of);
      float dv = 0.f;
      if (k < cfg.nM) {
        const float p = act_apply<float>(act, sum);
        if (cfg.scale_kind == HVI_SCALE_RANGE) dv = mb[k] * cfg.scale_b[k];
        else { const float qn = normcdfinvf(p); dv = mb[k] * cfg.scale_b[k] * 2.50662827463100050242f * expf(0.5f * qn * qn); }
        dv *= (act == HVI_ACT_LOGISTIC) ? p * (1.f - p) : (act == HVI_ACT_TANH) ? 1.f - p * p : 1.f;
      }
      dl[k] = dv;
    }
#pragma unroll
    for (int st = 0; st < WSTEPS; st++) {
      const int i0 = lane * 8 + 256 * st;
      if (i0 < K) {
        float z[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          float sum = 0.f;
#pragma unroll
          for (int k = 0; k < NOUT; k++) {
            sum = fmaf(__uint_as_float(wv[st][c * NOUT + k]), dl[k], sum);
            gW[st][c][k] = fmaf(x[st][c], dl[k], gW[st][c][k]);
          }
          z[c] = sum * act_deriv_f(act_prev, x[st][c]);
          gb[st][c] += z[c];
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const __nv_bfloat162 hh = __floats2bfloat162_rn(z[2 * j], z[2 * j + 1]);
          const __nv_bfloat162 ll = __floats2bfloat162_rn(z[2 * j] - __low2float(hh), z[2 * j + 1] - __high2float(hh));
          h[j] = *reinterpret_cast<const uint32_t*>(&hh);
          l[j] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        *reinterpret_cast<uint4*>(o.hi + r * o.ld + i0) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(o.lo + r * o.ld + i0) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
    __syncwarp();   // every lane has read this stage's μ̄ before the stage is refilled
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // ---- one slice of sums per block: the warps add their registers in warp order (deterministic)
  const int slice = K * (NOUT + 1);
  for (int e = threadIdx.x; e < slice; e += 256) sacc[e] = 0.f;
  __syncthreads();
  for (int w = 0; w < 8; w++) {
    if (warp == w) {
#pragma unroll
      for (int st = 0; st < WSTEPS; st++) {
        const int i0 = lane * 8 + 256 * st;
        if (i0 < K) {
#pragma unroll
          for (int c = 0; c < 8; c++) {
#pragma unroll
            for (int k = 0; k < NOUT; k++) sacc[(i0 + c) * NOUT + k] += gW[st][c][k];
            sacc[K * NOUT + i0 + c] += gb[st][c];
          }
        }
      }
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < slice; e += 256) P[(size_t)blockIdx.x * slice + e] = sacc[e];
}

#define HVI_NOUT_SWITCH(n, CALL)                                                               \
  switch (n) {                                                                                 \
    case 1: { constexpr int NOUT = 1; CALL; } break;                                           \
    case 2: { constexpr int NOUT = 2; CALL; } break;                                           \
    case 3: { constexpr int NOUT = 3; CALL; } break;                                           \
    case 4: { constexpr int NOUT = 4; CALL; } break;                                           \
    case 5: { constexpr int NOUT = 5; CALL; } break;                                           \
    case 6: { constexpr int NOUT = 6; CALL; } break;                                           \
    case 7: { constexpr int NOUT = 7; CALL; } break;                                           \
    default: { constexpr int NOUT = 8; CALL; } break;                                          \
  }

// ---- eligibility and work space ---------------------------------------------------------------
bool bf3_eligible(const Cfg<float>& cfg) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("HVI_WIDE_IMPL"); off = (e && strcmp(e, "tf32") == 0) ? 1 : 0; }
  if (off || !tma_encoder()) return false;
  if (cfg.nL < 3 || cfg.l_bias[cfg.nL - 1] || cfg.l_out[cfg.nL - 1] > 8 || cfg.nC + cfg.npc + 1 > 8) return false;
  for (int l = 0; l < cfg.nL - 1; l++)
    if (cfg.l_out[l] % 128 != 0) return false;
  return true;
}

struct Bf3Work {
  Planes H[8];          // outputs of layers 0 … nL-2
  Planes D[2];          // cotangent ping-pong
  bf16 *wn_hi, *wn_lo, *wt_hi, *wt_lo;   // weight planes, indexed like ϕg
  bf16 *w0_hi, *w0_lo;                   // first layer [N][S_LD] incl. the bias column
  float* extra_h;       // pbm covariate values of the current draw
  Planes Sx;            // [chunk][S_LD]: inputs, pbm covariates, 1 (k_build_sx)
  Planes Sd;            // [chunk][S_LD]: δ of the last layer
  float* dlast;         // [chunk][8], the same in fp32
  float* P;             // weight-gradient partial tiles / skinny partials
  double* db0;
};
// partial tiles of the split-K GEMMs: 24 splits of a full weight gradient, or 148 splits of a skinny one
static size_t bf3_partial_floats(int maxw) {
  const size_t pf = (size_t)24 * maxw * maxw, sk = (size_t)148 * 64 * maxw;
  return pf > sk ? pf : sk;
}
// carve the work space of launch_mlp_bwd_wide (wide_bwd_work_bytes) / launch_mlp_fwd_wide (wide_work_elems·4)
static void carve(Bf3Work& w, void* work, int nL, int maxw, int nphig, bool bwd) {
  const size_t cw = (size_t)WIDE_CHUNK * maxw;
  bf16* p = reinterpret_cast<bf16*>(work);
  auto planes = [&](Planes& pl) { pl.hi = p; p += cw; pl.lo = p; p += cw; pl.ld = 0; };
  if (bwd) {
    for (int l = 0; l < nL - 1; l++) planes(w.H[l]);
    planes(w.D[0]); planes(w.D[1]);
  } else {
    planes(w.H[0]); planes(w.H[1]);
  }
  w.wn_hi = p; p += nphig; w.wn_lo = p; p += nphig; w.wt_hi = p; p += nphig; w.wt_lo = p; p += nphig;
  p += (16 - (4 * (size_t)nphig) % 16) % 16;   // keep 32-byte alignment
  w.w0_hi = p; p += (size_t)maxw * S_LD; w.w0_lo = p; p += (size_t)maxw * S_LD;
  const size_t sn = (size_t)WIDE_CHUNK * S_LD;
  w.Sx = Planes{p, p + sn, S_LD}; p += 2 * sn;
  float* f = reinterpret_cast<float*>(p);
  w.extra_h = f; f += 16;
  if (bwd) {
    bf16* sp = reinterpret_cast<bf16*>(f);
    w.Sd = Planes{sp, sp + sn, S_LD};
    f += sn;              // 2 planes of sn bf16
    w.dlast = f; f += (size_t)WIDE_CHUNK * 8;
    w.P = f;
    f += bf3_partial_floats(maxw);
    f += ((uintptr_t)f % 8) ? 1 : 0;
    w.db0 = reinterpret_cast<double*>(f);
  }
}
size_t bf3_cache_slot_bytes(int nL, int maxw, int64_t rows) { return (size_t)(nL - 1) * (size_t)(rows > 0 ? rows : WIDE_CHUNK) * maxw * 4; }
int64_t bf3_chunks(int64_t nb) { return (nb + WIDE_CHUNK - 1) / WIDE_CHUNK; }
int64_t bf3_chunk_rows() { return WIDE_CHUNK; }
// the planes of slot `slot` of the activation cache, or false when the slot does not exist
static bool cache_planes(const WideCache& c, int64_t slot, int nL, int maxw, Planes* H) {
  if (!c.base || slot < 0 || slot >= c.n_slots) return false;
  const size_t cw = (size_t)(c.rows > 0 ? c.rows : WIDE_CHUNK) * maxw;
  bf16* p = reinterpret_cast<bf16*>(reinterpret_cast<unsigned char*>(c.base) + (size_t)slot * bf3_cache_slot_bytes(nL, maxw, c.rows));
  for (int l = 0; l < nL - 1; l++) { H[l].hi = p; p += cw; H[l].lo = p; p += cw; H[l].ld = 0; }
  return true;
}
size_t bf3_fwd_work_bytes(int maxw, int nphig) {
  return (size_t)WIDE_CHUNK * maxw * 8 + (size_t)nphig * 8 + 4 * (size_t)S_LD * (maxw + WIDE_CHUNK) + 512;
}
size_t bf3_bwd_work_bytes(int nL, int maxw, int nphig) {
  const size_t cw = (size_t)WIDE_CHUNK * maxw;
  return (size_t)(nL + 1) * cw * 4 + (size_t)nphig * 8 + 4 * (size_t)S_LD * maxw +
         4 * ((size_t)WIDE_CHUNK * (2 * S_LD + 8) + bf3_partial_floats(maxw)) + 8 * (size_t)maxw + 1024;
}

static cudaError_t split_weights(const Launch& L, const Cfg<float>& cfg, const float* phi, Bf3Work& w) {
  for (int l = 1; l < cfg.nL - 1; l++) {
    const int n = cfg.l_in[l] * cfg.l_out[l];
    const int o = cfg.woff[l];
    k_split_w<<<(n + 255) / 256, 256, 0, L.stream>>>(phi + o, cfg.l_in[l], cfg.l_out[l], w.wn_hi + o, w.wn_lo + o, w.wt_hi + o, w.wt_lo + o);
    if (L.counter) ++*L.counter;
  }
  {
    const int N = cfg.l_out[0];
    k_split_w0<<<(N * S_LD + 255) / 256, 256, 0, L.stream>>>(phi + cfg.woff[0], cfg.l_bias[0] ? phi + cfg.boff[0] : nullptr, cfg.l_in[0], N,
                                                             w.w0_hi, w.w0_lo);
    if (L.counter) ++*L.counter;
  }
  return cudaGetLastError();
}
static cudaError_t set_extra(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* zetaP, int M_units, int m, float* extra_h) {
  for (int c = 0; c < cfg.npc; c++) {
    const float* src = M_units > 0 ? zetaP + (size_t)cfg.pbm_covar_idx[c] * M_units + m : phi + cfg.o_muP + cfg.pbm_covar_idx[c];
    cudaError_t e = cudaMemcpyAsync(extra_h + c, src, sizeof(float), cudaMemcpyDeviceToDevice, L.stream);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
// forward chain of one chunk: first layer on CUDA cores, hidden layers on the tensor cores; outs[l] = output of layer l
static cudaError_t chain_fwd(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM, int64_t s0, int64_t rows,
                             Bf3Work& w, Planes* outs /*nL-1 entries*/) {
  {  // first layer: [x ; covariates ; 1] (Sx planes, K = S_LD) against [W_0 ; b_0]ᵀ on the tensor cores
    const int N = cfg.l_out[0];
    outs[0].ld = N;
    k_build_sx<<<(unsigned)((rows + 127) / 128), 128, 0, L.stream>>>(xM, cfg.nC, w.extra_h, cfg.npc, s0, rows, w.Sx.hi, w.Sx.lo);
    if (L.counter) ++*L.counter;
    GemmEpi ep;
    memset(&ep, 0, sizeof(ep));
    ep.act = cfg.l_act[0];
    ep.out = outs[0];
    cudaError_t e = launch_gemm_bf3<0>(L, w.Sx.hi, w.Sx.lo, S_LD, rows, w.w0_hi, w.w0_lo, S_LD, N, S_LD, 1, ep);
    if (e != cudaSuccess) return e;
  }
  for (int l = 1; l < cfg.nL - 1; l++) {
    const int N = cfg.l_out[l], K = cfg.l_in[l];
    GemmEpi ep;
    memset(&ep, 0, sizeof(ep));
    ep.bias = cfg.l_bias[l] ? phi + cfg.boff[l] : nullptr;
    ep.act = cfg.l_act[l];
    outs[l].ld = N;
    ep.out = outs[l];
    cudaError_t e = launch_gemm_bf3<0>(L, outs[l - 1].hi, outs[l - 1].lo, K, rows, w.wt_hi + cfg.woff[l], w.wt_lo + cfg.woff[l], K, N, K, 1, ep);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

cudaError_t launch_mlp_fwd_bf3(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM, int64_t nb, float* mu,
                               int M_units, const float* zetaP, void* work, WideCache cache) {
  Bf3Work w;
  carve(w, work, cfg.nL, cfg.max_width, cfg.nphig, false);
  cudaError_t e = split_weights(L, cfg, phi, w);
  if (e != cudaSuccess) return e;
  const int nm = M_units > 0 ? M_units : 1;
  const int nL = cfg.nL;
  for (int m = 0; m < nm; m++) {
    e = set_extra(L, cfg, phi, zetaP, M_units, m, w.extra_h);
    if (e != cudaSuccess) return e;
    for (int64_t s0 = 0; s0 < nb; s0 += WIDE_CHUNK) {
      const int64_t rows = nb - s0 < WIDE_CHUNK ? nb - s0 : WIDE_CHUNK;
      Planes outs[8];
      if (!cache_planes(cache, cache.slot0 + (int64_t)m * bf3_chunks(nb) + s0 / WIDE_CHUNK, nL, cfg.max_width, outs))
        for (int l = 0; l < nL - 1; l++) outs[l] = w.H[l & 1];    // not kept: ping-pong
      e = chain_fwd(L, cfg, phi, xM, s0, rows, w, outs);
      if (e != cudaSuccess) return e;
      const int l = nL - 1;
      HVI_NOUT_SWITCH(cfg.l_out[l], (k_last_bf3<0, NOUT><<<(unsigned)((rows + 255) / 256), 256, 0, L.stream>>>(
                                        cfg, outs[l - 1].hi, outs[l - 1].lo, cfg.l_in[l], phi + cfg.woff[l], cfg.l_act[l], rows, cfg.l_in[l], mu,
                                        nullptr, s0, m, M_units, nb, nullptr, nullptr, nullptr)));
      if (L.counter) ++*L.counter;
    }
  }
  return cudaGetLastError();
}

cudaError_t launch_mlp_bwd_bf3(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM, int64_t nb, const float* mubar,
                               int M_units, const float* zetaP, void* work, double* grad_acc, double* xacc, WideCache cache) {
  Bf3Work w;
  const int nL = cfg.nL;
  carve(w, work, nL, cfg.max_width, cfg.nphig, true);
  auto count = [&]() { if (L.counter) ++*L.counter; };
  cudaError_t e = split_weights(L, cfg, phi, w);
  if (e != cudaSuccess) return e;
  const int sms = L.sm_count > 0 ? L.sm_count : 148;
  const int nm = M_units > 0 ? M_units : 1;
  // skinny product G[S_w x N] = Sᵀ·B over the rows of the chunk on the tensor cores (S: planes [rows][S_LD], only the
  // first S_w features used; B: planes [rows][N]); returns the number of partial slices in w.P ([slice][S_w][N])
  auto skinny = [&](const Planes& S, int S_w, const Planes& B, int N, int64_t rows, int* nslice) -> cudaError_t {
    GemmEpi ep;
    memset(&ep, 0, sizeof(ep));
    ep.P = w.P;
    const int ns = bf3_wgrad_splits(S_w, N, sms);
    *nslice = effective_splits(rows, ns, nullptr);
    return launch_gemm_bf3<2>(L, S.hi, S.lo, S_LD, S_w, B.hi, B.lo, N, N, rows, ns, ep);
  };
  for (int m = 0; m < nm; m++) {
    e = set_extra(L, cfg, phi, zetaP, M_units, m, w.extra_h);
    if (e != cudaSuccess) return e;
    for (int64_t s0 = 0; s0 < nb; s0 += WIDE_CHUNK) {
      const int64_t rows = nb - s0 < WIDE_CHUNK ? nb - s0 : WIDE_CHUNK;
      const int K0 = cfg.l_in[0];          // inputs of the first layer; Sx column K0 is the constant 1
      int nslice = 0;
      // ---- activations of the chunk: kept by the forward pass, or recomputed (which also builds Sx)
      Planes H[8];
      if (cache_planes(cache, cache.slot0 + (int64_t)m * bf3_chunks(nb) + s0 / WIDE_CHUNK, nL, cfg.max_width, H)) {
        for (int l = 0; l < nL - 1; l++) H[l].ld = cfg.l_out[l];    // kept by the forward pass of this evaluation
        k_build_sx<<<(unsigned)((rows + 127) / 128), 128, 0, L.stream>>>(xM, cfg.nC, w.extra_h, cfg.npc, s0, rows, w.Sx.hi, w.Sx.lo);
        count();
      } else {
        for (int l = 0; l < nL - 1; l++) H[l] = w.H[l];
        e = chain_fwd(L, cfg, phi, xM, s0, rows, w, H);
        if (e != cudaSuccess) return e;
      }
      // ---- last layer: δ_L, ∇W_L, δ_{L-1} (and ∇b_{L-1} where the fused kernel applies)
      Planes Dcur = w.D[0], Dnxt = w.D[1];
      bool bias_below_done = false;
      {
        const int l = nL - 1, K = cfg.l_in[l], no = cfg.l_out[l];
        if (K <= 512 && !getenv("HVI_WIDE_LAST_UNFUSED")) {
          int nblk = (int)((rows + 255) / 256);
          if (nblk > 2 * sms) nblk = 2 * sms;
          Dcur.ld = K;
          const size_t ring_bytes = (size_t)8 * LAST_STAGES * LAST_STAGE_BYTES;
          HVI_NOUT_SWITCH(no, (cudaFuncSetAttribute(k_last_bwd_fused_bf3<NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes)));
          HVI_NOUT_SWITCH(no, (k_last_bwd_fused_bf3<NOUT><<<nblk, 256, ring_bytes, L.stream>>>(
                                  cfg, H[l - 1].hi, H[l - 1].lo, K, phi + cfg.woff[l], cfg.l_act[l], cfg.l_act[l - 1], rows, K, mubar, s0, m,
                                  M_units, nb, Dcur, w.P)));
          count();
          const size_t sst = (size_t)K * (no + 1);
          k_accum<float><<<(no * K + 255) / 256, 256, 0, L.stream>>>(w.P, nblk, sst, 1, no * K, grad_acc, cfg.woff[l], 0, 1, nullptr);
          count();
          if (cfg.l_bias[l - 1]) {
            k_accum<float><<<(K + 255) / 256, 256, 0, L.stream>>>(w.P + (size_t)no * K, nblk, sst, 1, K, grad_acc, cfg.boff[l - 1], 0, 1,
                                                                  l - 1 == 0 ? w.db0 : nullptr);
            count();
          }
          bias_below_done = true;
        } else {
          HVI_NOUT_SWITCH(no, (k_last_bf3<1, NOUT><<<(unsigned)((rows + 255) / 256), 256, 0, L.stream>>>(
                                  cfg, H[l - 1].hi, H[l - 1].lo, K, phi + cfg.woff[l], cfg.l_act[l], rows, K, nullptr, mubar, s0, m, M_units, nb,
                                  w.dlast, w.Sd.hi, w.Sd.lo)));
          count();
          Dcur.ld = K;
          HVI_NOUT_SWITCH(no, (k_thin_bwd_bf3<NOUT><<<dim3((unsigned)((rows + 255) / 256), K / 64), 256, 0, L.stream>>>(
                                  w.dlast, phi + cfg.woff[l], H[l - 1].hi, H[l - 1].lo, K, cfg.l_act[l - 1], rows, Dcur)));
          count();
          e = skinny(w.Sd, no, H[l - 1], K, rows, &nslice);      // ∇W_L[i][k] = Σ_r δ_L[r][k] H_{L-1}[r][i]
          if (e != cudaSuccess) return e;
          k_accum<float><<<(no * K + 255) / 256, 256, 0, L.stream>>>(w.P, nslice, (size_t)no * K, no, K, grad_acc, cfg.woff[l], 1, no, nullptr);
          count();
        }
      }
      // ---- hidden layers l = nL-2 … 1: bias, weight gradient, δ of the layer below
      bool bias_done = bias_below_done;   // ∇b_l already summed by the kernel that produced δ_l
      const bool colsum_on = !getenv("HVI_WIDE_BIAS_SKINNY");
      for (int l = nL - 2; l >= 1; l--) {
        const int N = cfg.l_out[l], K = cfg.l_in[l];   // Dcur: [rows][N]
        if (cfg.l_bias[l] && !bias_done) {             // ∇b_l = 1ᵀ δ_l: row K0 of Sxᵀ δ_l
          e = skinny(w.Sx, K0 + 1, Dcur, N, rows, &nslice);
          if (e != cudaSuccess) return e;
          k_accum<float><<<(N + 255) / 256, 256, 0, L.stream>>>(w.P + (size_t)K0 * N, nslice, (size_t)(K0 + 1) * N, 1, N, grad_acc, cfg.boff[l], 0, 1, nullptr);
          count();
        }
        {  // ∇W_l[in][out] = H_{l-1}ᵀ δ_l: the same planes as MN-major operands, contraction over the rows
          GemmEpi ep;
          memset(&ep, 0, sizeof(ep));
          ep.P = w.P;
          const int ns = bf3_wgrad_splits(K, N, sms);
          e = launch_gemm_bf3<2>(L, H[l - 1].hi, H[l - 1].lo, K, K, Dcur.hi, Dcur.lo, N, N, rows, ns, ep);
          if (e != cudaSuccess) return e;
          const int ns_eff = effective_splits(rows, ns, nullptr);
          k_accum<float><<<(K * N + 255) / 256, 256, 0, L.stream>>>(w.P, ns_eff, (size_t)K * N, 1, K * N, grad_acc, cfg.woff[l], 0, 1, nullptr);
          count();
        }
        {  // δ_{l-1} = (δ_l W_l) ⊙ act'(H_{l-1}): A = δ_l [rows x N], B = W_l in its own [in][out] layout; where the layer
           // below is a hidden one with a bias, the epilogue also leaves the column sums of δ_{l-1} per 32 rows (∇b_{l-1})
          GemmEpi ep;
          memset(&ep, 0, sizeof(ep));
          ep.act = cfg.l_act[l - 1];
          ep.h_hi = H[l - 1].hi; ep.h_lo = H[l - 1].lo; ep.ldh = K;
          Dnxt.ld = K;
          ep.out = Dnxt;
          const bool cs = colsum_on && l - 1 >= 1 && cfg.l_bias[l - 1];
          const int R = (int)((rows + G_BM - 1) / G_BM) * (G_BM / 32);
          if (cs) ep.colsum = w.P;            // the partial tiles of the weight gradient have been summed: P is free
          e = launch_gemm_bf3<1>(L, Dcur.hi, Dcur.lo, N, rows, w.wn_hi + cfg.woff[l], w.wn_lo + cfg.woff[l], N, K, N, 1, ep);
          if (e != cudaSuccess) return e;
          if (cs) {
            const int nb1 = R < 128 ? R : 128;
            float* P2 = w.P + (size_t)R * K;
            k_colsum_rows<<<nb1, 256, 0, L.stream>>>(w.P, R, K, P2);
            k_accum<float><<<(K + 255) / 256, 256, 0, L.stream>>>(P2, nb1, (size_t)K, 1, K, grad_acc, cfg.boff[l - 1], 0, 1, nullptr);
            count(); count();
          }
          bias_done = cs;
        }
        Planes tmp = Dcur; Dcur = Dnxt; Dnxt = tmp;
      }
      // ---- first layer: ∇W_0 = [x; covariates]ᵀ δ_0 (rows 0 … K0-1), ∇b_0 (row K0), and the cotangent of the covariate inputs
      {
        const int N = cfg.l_out[0], ns = K0 + 1;
        e = skinny(w.Sx, ns, Dcur, N, rows, &nslice);
        if (e != cudaSuccess) return e;
        const size_t sst = (size_t)ns * N;
        double* acc_b = cfg.l_bias[0] ? grad_acc : nullptr;
        k_accum<float><<<(K0 * N + 255) / 256, 256, 0, L.stream>>>(w.P, nslice, sst, K0, N, grad_acc, cfg.woff[0], N, 1, nullptr);
        k_accum<float><<<(N + 255) / 256, 256, 0, L.stream>>>(w.P + (size_t)K0 * N, nslice, sst, 1, N, acc_b, cfg.boff[0], 0, 1, w.db0);
        count(); count();
        if (cfg.npc > 0 && xacc) {
          k_xbar<float><<<1, 32 * cfg.npc, 0, L.stream>>>(phi + cfg.woff[0], cfg.nC, cfg.npc, N, w.db0, xacc + (size_t)(M_units > 0 ? m : 0) * cfg.npc);
          count();
        }
      }
      e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

// ---- test harness: one GEMM on device fp32 arrays (split → GEMM → join) ----------------------
// mode 0: C = act(A·Bᵀ + bias), A [M][K], B [N][K];  mode 1: C = (A·Bᵀ) ⊙ act'(Hd);
// mode 2: C = AᵀB through the split-K path, A [K][M], B [K][N] (the weight-gradient orientation).  C [M][N].
cudaError_t gemm_bf3_device(cudaStream_t stream, int sm_count, int mode, int64_t M, int N, int64_t K, const float* A, const float* B,
                            const float* bias, int act, const float* Hd, float* C) {
  if (N % 128 != 0 || !tma_encoder()) return cudaErrorInvalidValue;
  if (mode == 2 ? (M % 32 != 0) : (K % 32 != 0)) return cudaErrorInvalidValue;
  Launch L{stream, sm_count, nullptr};
  bf16* buf = nullptr;
  const size_t nA = (size_t)M * K, nB = (size_t)N * K, nC = (size_t)M * N;
  const int ns = mode == 2 ? bf3_wgrad_splits((int)M, N, sm_count) : 1;
  cudaError_t e = cudaMalloc((void**)&buf, 2 * (2 * nA + 2 * nB + 4 * nC) + sizeof(float) * (size_t)ns * M * N + 64);
  if (e != cudaSuccess) return e;
  const int64_t a_rows = mode == 2 ? K : M, a_w = mode == 2 ? M : K, b_rows = mode == 2 ? K : N, b_w = mode == 2 ? N : K;
  Planes pa{buf, buf + nA, a_w};
  Planes pb{buf + 2 * nA, buf + 2 * nA + nB, b_w};
  bf16* q = buf + 2 * nA + 2 * nB;
  Planes pc{q, q + nC, N};
  Planes ph{q + 2 * nC, q + 3 * nC, N};
  float* P = reinterpret_cast<float*>(q + 4 * nC);
  k_split_planes<<<dim3((unsigned)((a_rows + 255) / 256), (unsigned)(a_w / 32)), 256, 0, stream>>>(A, a_rows, (int)a_w, pa);
  k_split_planes<<<dim3((unsigned)((b_rows + 255) / 256), (unsigned)(b_w / 32)), 256, 0, stream>>>(B, b_rows, (int)b_w, pb);
  GemmEpi ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = bias; ep.act = act; ep.out = pc; ep.P = P;
  if (mode == 1) {
    k_split_planes<<<dim3((unsigned)((M + 255) / 256), (unsigned)(N / 32)), 256, 0, stream>>>(Hd, M, N, ph);
    ep.h_hi = ph.hi; ep.h_lo = ph.lo; ep.ldh = N;
  }
  auto run = [&]() {
    if (mode == 0) return launch_gemm_bf3<0>(L, pa.hi, pa.lo, K, M, pb.hi, pb.lo, K, N, K, 1, ep);
    if (mode == 1) return launch_gemm_bf3<1>(L, pa.hi, pa.lo, K, M, pb.hi, pb.lo, K, N, K, 1, ep);
    return launch_gemm_bf3<2>(L, pa.hi, pa.lo, M, M, pb.hi, pb.lo, N, N, K, ns, ep);
  };
  e = run();
  if (e == cudaSuccess && getenv("HVI_GEMM_TIME")) {   // device time of the GEMM alone (profiles/bench_gemm_bf3.py)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 20;
    cudaEventRecord(e0, stream);
    for (int i = 0; i < reps && e == cudaSuccess; i++) e = run();
    cudaEventRecord(e1, stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double us = 1e3 * ms / reps, flop = 2.0 * (double)M * N * (double)K;
    fprintf(stderr, "hvi_gemm_bf3 mode %d M=%lld N=%d K=%lld: %.1f us/launch, %.1f TFLOP/s fp32-equivalent (x3 bf16 MMA work)\n", mode,
            (long long)M, N, (long long)K, us, flop / us * 1e-6);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  if (e == cudaSuccess) {
    const int64_t n = M * N;
    if (mode == 2) {
      const int ns_eff = effective_splits(K, ns, nullptr);
      double* acc = nullptr;
      e = cudaMalloc((void**)&acc, sizeof(double) * n);
      if (e == cudaSuccess) {
        cudaMemsetAsync(acc, 0, sizeof(double) * n, stream);
        k_accum<float><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(P, ns_eff, (size_t)n, 1, (int)n, acc, 0, 0, 1, nullptr);
        double* hacc = (double*)malloc(sizeof(double) * n);
        e = cudaMemcpyAsync(hacc, acc, sizeof(double) * n, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        float* hc = (float*)malloc(sizeof(float) * n);
        for (int64_t i = 0; i < n; i++) hc[i] = (float)hacc[i];
        if (e == cudaSuccess) e = cudaMemcpy(C, hc, sizeof(float) * n, cudaMemcpyHostToDevice);
        free(hacc); free(hc);
        cudaFree(acc);
      }
    } else {
      k_join_planes<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pc, M, N, C);
      e = cudaGetLastError();
    }
  }
  cudaError_t e2 = cudaStreamSynchronize(stream);
  cudaFree(buf);
  return e != cudaSuccess ? e : e2;
}

}  // namespace hvi
