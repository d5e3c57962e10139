// hvi_mlp_tc.cu -- the small site-batched ML applicator g(xM; ϕg) on the 5th-generation tensor cores (Float32).
//
// g = NormalScaling ∘ MLP(tanh, tanh, logistic) (src/ModelApplicator.jl:163-176,
// ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52) for the 3-layer shapes of DoubleMM (5-20-20-2 and, with one
// pbm covariate, 6-24-24-2), evaluated on sites or -- pass 2 of generate_ζ, src/elbo.jl:508-515 -- on (draw, site) units.
//
// One CTA of 256 threads owns a tile of 128 columns.  Column r of the tile = TMEM lane r = row r of the operand tiles;
// it is served by TWO threads (warps q and q + 4 share TMEM lane quarter q), each owning one half (12) of the padded
// 24 features of a layer: half the registers and half the dependency chain per thread, twice the warps per SM to hide
// the MUFU / TMEM / barrier latencies (the first, thread-per-column version issued 47 % of its cycles).
//   * layer 1 (K = n_covar + pbm covariates) runs on the CUDA cores in exact fp32.  A CTA walks a contiguous range of
//     units u = site_tile·n_MC + draw, so with pbm covariates the site part of the pre-activation (W1[:nC]ᵀ xM + b1) is
//     computed once per site tile and each draw adds the rank-1 term of its covariates (only ζP[idx] changes between
//     the draws of a site, src/elbo.jl:508-515).
//   * layer 2 and every product of the backward pass are tcgen05.mma instructions issued by ONE thread, operands in
//     shared memory as 128-byte-swizzled tiles split hi + lo, accumulators in TMEM, completion by tcgen05.commit →
//     mbarrier; the epilogues read their TMEM lane with tcgen05.ld.  Tensor work is asynchronous to the FP32/MUFU
//     epilogues (warp-level mma.sync blocks FP32 issue of its scheduler on B200: profiles/microbench/mma_pipes_b200.txt).
//   * the 2-output last layer, logistic and Φ⁻¹ run on the CUDA cores; the two threads of a column exchange their
//     partial sums through shared memory (named barrier of the 64 threads of a lane quarter).
// Forward: 3xTF32 (hi = x truncated to tf32, lo = x − hi exactly, of which the tensor core reads the top 19 bits):
// μ_ζ feeds the σ- and ρ-gradients of the streaming kernel, which amplify its rounding.
// Backward: 3xBF16 like the mma.sync kernel it replaces; see the block comment above k_mlp3_tc_bwd.
// The mma.sync kernels (hvi_mlp_mma.cu) remain the fallback (HVI_MLP_TC=0) and the cross-check in the tests.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "hvi_internal.cuh"
#include "hvi_tc_common.cuh"

namespace hvi {
namespace {

constexpr int TC_COLS = 128;           // columns per tile = TMEM lanes
constexpr int TC_THREADS = 256;        // two threads per column
constexpr int TC_HP = 24;              // padded hidden width
constexpr int TC_HF = TC_HP / 2;       // features per thread
constexpr int TC_PLANE = TC_COLS * 128;   // one plane of an operand tile: 128 rows of one swizzle row (16 KB)

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// MN-major SWIZZLE_128B descriptor: atoms of 64 bf16 features (128 B) x 8 rows; LBO = distance between 64-feature atoms,
// SBO = distance between 8-row groups (1024 B)
__device__ __forceinline__ uint64_t make_desc_mn_(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 4; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// this thread's 12 accumulator columns
__device__ __forceinline__ void tmem_ld12(uint32_t taddr, float (&v)[TC_HF]) {
  tmem_ld8(taddr, v);
  tmem_ld4(taddr + 8, v + 8);
  tmem_wait_ld();
}

__device__ __forceinline__ float rcpa(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2a(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
constexpr float K2LOG2E = 2.885390081777927f;   // 2·log2(e): tanh(z) = 1 − 2/(1 + 2^(K2LOG2E·z))
__device__ __forceinline__ float tanh_scaled(float a) { return fmaf(-2.0f, rcpa(ex2a(a) + 1.0f), 1.0f); }   // a = K2LOG2E·z

// shared-memory accesses by 32-bit shared-window address (the tiles are carved out of an aligned-up dynamic buffer)
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts64u(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
// barrier of the 64 threads that serve lane quarter q (warps q and q + 4); barrier 0 is the whole CTA.  Issued as PTX:
// the two halves of the CTA run different instantiations of the loop body and meet at different program counters.
__device__ __forceinline__ void pair_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); }
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }

// ---- tf32 tiles [row][32 floats]: four consecutive features (one 16-byte chunk); hi = trunc_tf32(x), lo = x − hi
__device__ __forceinline__ float trunc_tf(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void store4_tf(uint32_t hi_tile, uint32_t lo_tile, int row, int chunk, float a, float b, float c, float d) {
  const uint32_t off = row * 128 + ((chunk ^ (row & 7)) << 4);
  const float4 h = make_float4(trunc_tf(a), trunc_tf(b), trunc_tf(c), trunc_tf(d));
  sts128(hi_tile + off, h);
  sts128(lo_tile + off, make_float4(a - h.x, b - h.y, c - h.z, d - h.w));
}
__device__ __forceinline__ void store1_tf(uint32_t hi_tile, uint32_t lo_tile, int row, int col, float x) {
  const uint32_t off = row * 128 + ((((col >> 2) ^ (row & 7))) << 4) + ((col & 3) << 2);
  const float h = trunc_tf(x);
  sts32(hi_tile + off, h);
  sts32(lo_tile + off, x - h);
}
// ---- bf16 tiles [row][64 bf16]: hi = bf16(x), lo = bf16(x − hi); x0 in the low half of the packed word
__device__ __forceinline__ void split_bf2(float x0, float x1, uint32_t& h, uint32_t& l) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - h1), "f"(x0 - h0));
}
// four consecutive features starting at feature f (multiple of 4) of row `row`: half a 16-byte chunk
__device__ __forceinline__ void store4_bf(uint32_t hi_tile, uint32_t lo_tile, int row, int f, float a, float b, float c, float d) {
  uint32_t h0, l0, h1, l1;
  split_bf2(a, b, h0, l0);
  split_bf2(c, d, h1, l1);
  const uint32_t off = row * 128 + ((((f >> 3) ^ (row & 7))) << 4) + ((f & 4) << 1);
  sts64u(hi_tile + off, h0, h1);
  sts64u(lo_tile + off, l0, l1);
}
__device__ __forceinline__ void store1_bf(uint32_t hi_tile, uint32_t lo_tile, int row, int col, float x) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  const uint32_t off = row * 128 + ((((col >> 3) ^ (row & 7))) << 4) + ((col & 7) << 1);
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(hi_tile + off), "h"(__bfloat16_as_ushort(h)) : "memory");
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(lo_tile + off), "h"(__bfloat16_as_ushort(l)) : "memory");
}

// tables shared by both kernels (floats): W1 rows [8][HP] (row NI = b1) | b2·K2LOG2E [HP] | W3 [HP][2]
constexpr int T_W1 = 0, T_B2 = 8 * TC_HP, T_W3 = T_B2 + TC_HP, TAB_FLOATS = T_W3 + 2 * TC_HP;
constexpr int TAB_BYTES = (TAB_FLOATS * 4 + 15) & ~15;

template <int NI, int H1, int H2, int NOUT>
struct MlpTc {
  static_assert(NI + 1 <= 8 && H1 <= TC_HP && H2 <= TC_HP && H1 % 4 == 0 && H2 % 4 == 0 && NOUT <= 2, "shape");

  __device__ static float wfetch(const Cfg<float>& cfg, const float* __restrict__ phi, int l, int nin, int nout, int i, int j) {
    return (i < nin && j < nout) ? phi[cfg.woff[l] + i * nout + j] : 0.0f;
  }
  __device__ static void build_tables(const Cfg<float>& cfg, const float* __restrict__ phi, uint32_t tab, int tid, int nthr) {
    for (int e = tid; e < 8 * TC_HP; e += nthr) {
      const int i = e / TC_HP, j = e % TC_HP;
      float w = wfetch(cfg, phi, 0, NI, H1, i, j);
      if (i == NI && j < H1) w = phi[cfg.boff[0] + j];
      sts32(tab + 4 * (T_W1 + e), w);
    }
    for (int e = tid; e < TC_HP; e += nthr) sts32(tab + 4 * (T_B2 + e), e < H2 ? phi[cfg.boff[1] + e] * K2LOG2E : 0.0f);
    for (int e = tid; e < 2 * TC_HP; e += nthr) {
      const int j = e >> 1, k = e & 1;
      sts32(tab + 4 * (T_W3 + e), (j < H2 && k < NOUT) ? phi[cfg.woff[2] + j * NOUT + k] : 0.0f);
    }
  }

  // site part of the layer-1 pre-activation of this thread's features: b1 + Σ_{i < nC} x_i W1[i][f]; also returns the inputs
  template <int HALF>
  __device__ __forceinline__ static void site_base(const Cfg<float>& cfg, uint32_t tab, const float* __restrict__ xM, int64_t site,
                                                   int64_t nb, float (&xs)[8], float (&base)[TC_HF]) {
    const bool in = site < nb;
    const float* row = xM + (in ? site : nb - 1) * cfg.nC;
#pragma unroll
    for (int i = 0; i < 8; i++) xs[i] = (i < NI && i < cfg.nC && in) ? row[i] : (i == NI ? 1.0f : 0.0f);
#pragma unroll
    for (int j4 = 0; j4 < TC_HF / 4; j4++) {
      const float4 b = lds128(tab + 4 * (T_W1 + NI * TC_HP + HALF * TC_HF + 4 * j4));
      base[4 * j4] = b.x; base[4 * j4 + 1] = b.y; base[4 * j4 + 2] = b.z; base[4 * j4 + 3] = b.w;
    }
#pragma unroll
    for (int i = 0; i < NI; i++) {
      if (i < cfg.nC) {
#pragma unroll
        for (int j4 = 0; j4 < TC_HF / 4; j4++) {
          const float4 w = lds128(tab + 4 * (T_W1 + i * TC_HP + HALF * TC_HF + 4 * j4));
          base[4 * j4] = fmaf(xs[i], w.x, base[4 * j4]); base[4 * j4 + 1] = fmaf(xs[i], w.y, base[4 * j4 + 1]);
          base[4 * j4 + 2] = fmaf(xs[i], w.z, base[4 * j4 + 2]); base[4 * j4 + 3] = fmaf(xs[i], w.w, base[4 * j4 + 3]);
        }
      }
    }
  }
  // h1 of this thread's features for one unit: tanh(base + Σ_covariates ζP[idx]·W1[nC + c][f]); records the covariates in xs
  template <int HALF>
  __device__ __forceinline__ static void layer1(const Cfg<float>& cfg, uint32_t tab, const float* __restrict__ phi,
                                                const float* __restrict__ zetaP, int M_units, int m, bool in, const float (&base)[TC_HF],
                                                float (&xs)[8], float (&h1)[TC_HF]) {
#pragma unroll
    for (int j = 0; j < TC_HF; j++) h1[j] = base[j];
#pragma unroll
    for (int i = 0; i < NI; i++) {
      if (i >= cfg.nC) {   // pbm covariates: ζP[idx] of this draw (pass 1: μP[idx])
        const int idx = cfg.pbm_covar_idx[i - cfg.nC];
        const float sv = in ? (M_units > 0 ? zetaP[(size_t)idx * M_units + m] : phi[cfg.o_muP + idx]) : 0.0f;
        xs[i] = sv;
#pragma unroll
        for (int j4 = 0; j4 < TC_HF / 4; j4++) {
          const float4 w = lds128(tab + 4 * (T_W1 + i * TC_HP + HALF * TC_HF + 4 * j4));
          h1[4 * j4] = fmaf(sv, w.x, h1[4 * j4]); h1[4 * j4 + 1] = fmaf(sv, w.y, h1[4 * j4 + 1]);
          h1[4 * j4 + 2] = fmaf(sv, w.z, h1[4 * j4 + 2]); h1[4 * j4 + 3] = fmaf(sv, w.w, h1[4 * j4 + 3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < TC_HF; j++) h1[j] = (HALF * TC_HF + j < H1) ? tanh_scaled(h1[j] * K2LOG2E) : 0.0f;
  }
  // h2 of this thread's features from its accumulator columns, and its partial sums of the last layer
  template <int HALF>
  __device__ __forceinline__ static void layer2_epi(uint32_t tab, const float (&z)[TC_HF], float (&h2)[TC_HF], float (&z3)[2]) {
    float b[TC_HF];
#pragma unroll
    for (int j4 = 0; j4 < TC_HF / 4; j4++) {
      const float4 v = lds128(tab + 4 * (T_B2 + HALF * TC_HF + 4 * j4));
      b[4 * j4] = v.x; b[4 * j4 + 1] = v.y; b[4 * j4 + 2] = v.z; b[4 * j4 + 3] = v.w;
    }
    z3[0] = z3[1] = 0.0f;
#pragma unroll
    for (int j = 0; j < TC_HF; j++) {
      if (HALF * TC_HF + j < H2) {
        h2[j] = tanh_scaled(fmaf(z[j], K2LOG2E, b[j]));
        const float2 w = lds64(tab + 4 * (T_W3 + 2 * (HALF * TC_HF + j)));
        z3[0] = fmaf(h2[j], w.x, z3[0]);
        z3[1] = fmaf(h2[j], w.y, z3[1]);
      } else h2[j] = 0.0f;
    }
  }
};

__device__ __forceinline__ void tc_prologue(uint64_t* bar0, uint64_t* bar1, uint32_t* tmem_slot, int ncols_pow2) {
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    if (bar1) mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols_pow2));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
}
// generic-proxy writes of the tiles → visible to the tensor core; TMEM reads ordered before the next MMA
__device__ __forceinline__ void tc_publish() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cta_sync();
}
__device__ __forceinline__ void tc_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// =====================================================================================================
// Forward
// =====================================================================================================
// shared memory: A hi | A lo ([128][32 tf32]: h1) | W2 hi | W2 lo ([32 n][32 k]) | tables | exchange float2 [2][128]
constexpr int TF_O_A_HI = 0, TF_O_A_LO = TC_PLANE, TF_O_W_HI = 2 * TC_PLANE, TF_O_W_LO = TF_O_W_HI + 32 * 128,
              TF_O_TAB = TF_O_W_LO + 32 * 128, TF_O_XCH = TF_O_TAB + TAB_BYTES, TF_SMEM = TF_O_XCH + 2 * TC_COLS * 8 + 1024;
// D = F32 [4,6) = 1, A = B = TF32 [7,10) = [10,13) = 2, K-major both, N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t TF_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr int TF_CTAS_PER_SM = 4;

template <int NI, int H1, int H2, int NOUT, int HALF>
__device__ __forceinline__ void tc_fwd_body(const Cfg<float>& cfg, const float* __restrict__ phi, const float* __restrict__ xM,
                                            const float* __restrict__ zetaP, int M_units, int64_t nb, float* __restrict__ mu,
                                            uint32_t sm, uint32_t tmem, uint64_t* bar) {
  using N = MlpTc<NI, H1, H2, NOUT>;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, q = warp & 3;
  const int r = q * 32 + lane;
  const uint32_t a_hi = sm + TF_O_A_HI, a_lo = sm + TF_O_A_LO, w_hi = sm + TF_O_W_HI, w_lo = sm + TF_O_W_LO, tab = sm + TF_O_TAB;
  const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + HALF * TC_HF;
  const int Mu = M_units > 0 ? M_units : 1;
  const int64_t ntiles = (nb + TC_COLS - 1) / TC_COLS;
  const int64_t ntot = ntiles * Mu;
  const int64_t u0 = ntot * blockIdx.x / gridDim.x, u1 = ntot * (blockIdx.x + 1) / gridDim.x;
  constexpr int KS = (H1 + 7) / 8;
  uint32_t phase = 0;
  float base[TC_HF], xs[8], h1[TC_HF];
  int64_t st_prev = -1;
  auto prepare = [&](int64_t u) {   // layer 1 of unit u into h1
    const int64_t st = u / Mu;
    const int m = (int)(u - st * Mu);
    const int64_t site = st * TC_COLS + r;
    if (st != st_prev) { st_prev = st; N::template site_base<HALF>(cfg, tab, xM, site, nb, xs, base); }
    N::template layer1<HALF>(cfg, tab, phi, zetaP, M_units, m, site < nb, base, xs, h1);
  };
  if (u0 < u1) prepare(u0);
  for (int64_t u = u0; u < u1; u++) {
    const int64_t st = u / Mu;
    const int m = (int)(u - st * Mu);
    const int64_t site = st * TC_COLS + r;
#pragma unroll
    for (int c = 0; c < TC_HF / 4; c++)
      store4_tf(a_hi, a_lo, r, (HALF * TC_HF) / 4 + c, h1[4 * c], h1[4 * c + 1], h1[4 * c + 2], h1[4 * c + 3]);
    tc_publish();
    if (HALF == 0 && t == 0) {   // z2 = h1·W2: three products per K step of 8
      tc_after_sync();
#pragma unroll
      for (int s = 0; s < KS; s++) {
        const uint32_t o = s * 32;
        umma_tf32(tmem, make_desc(a_lo + o), make_desc(w_hi + o), TF_IDESC, s > 0 ? 1u : 0u);
        umma_tf32(tmem, make_desc(a_hi + o), make_desc(w_lo + o), TF_IDESC, 1u);
        umma_tf32(tmem, make_desc(a_hi + o), make_desc(w_hi + o), TF_IDESC, 1u);
      }
      umma_commit_(bar);
    }
    if (u + 1 < u1) prepare(u + 1);   // layer 1 of the next unit while the tensor core works
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_after_sync();
    float z[TC_HF], h2[TC_HF], z3[2];
    tmem_ld12(trow, z);
    N::template layer2_epi<HALF>(tab, z, h2, z3);
    // the column's two threads exchange their partial sums; thread HALF finishes output k = HALF
    sts64(sm + TF_O_XCH + 8 * (HALF * TC_COLS + r), z3[0], z3[1]);
    pair_sync(q);
    const float2 o = lds64(sm + TF_O_XCH + 8 * ((1 - HALF) * TC_COLS + r));
    constexpr int k = HALF;
    if (k < NOUT && k < cfg.nM && site < nb) {
      const float zk = (k == 0 ? z3[0] + o.x : z3[1] + o.y);
      const float p = rcpa(1.0f + ex2a(-1.4426950408889634f * zk));
      const float v = (cfg.scale_kind == HVI_SCALE_RANGE) ? p * cfg.scale_b[k] + cfg.scale_a[k]
                                                           : cfg.scale_a[k] + cfg.scale_b[k] * normcdfinvf(p);
      if (M_units > 0) mu[((size_t)m * cfg.nM + k) * nb + site] = v;
      else mu[site * cfg.nM + k] = v;
    }
  }
}

template <int NI, int H1, int H2, int NOUT>
__global__ void __launch_bounds__(TC_THREADS, TF_CTAS_PER_SM)
k_mlp3_tc_fwd(const __grid_constant__ Cfg<float> cfg, const float* __restrict__ phi, const float* __restrict__ xM,
              const float* __restrict__ zetaP, int M_units, int64_t nb, float* __restrict__ mu) {
  using N = MlpTc<NI, H1, H2, NOUT>;
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sm = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int t = threadIdx.x;
  tc_prologue(&bar, nullptr, &tmem_base_s, 32);
  // the A tile must hold finite values everywhere the MMAs read
  for (int e = t; e < 2 * TC_PLANE / 16; e += TC_THREADS) sts128u(sm + 16 * e, make_uint4(0u, 0u, 0u, 0u));
  for (int e = t; e < 32 * 32; e += TC_THREADS) {   // B operand [n = output feature j][k = input feature i]
    const int j = e >> 5, i = e & 31;
    store1_tf(sm + TF_O_W_HI, sm + TF_O_W_LO, j, i, N::wfetch(cfg, phi, 1, H1, H2, i, j));
  }
  N::build_tables(cfg, phi, sm + TF_O_TAB, t, TC_THREADS);
  tc_publish();
  tc_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (t < TC_THREADS / 2) tc_fwd_body<NI, H1, H2, NOUT, 0>(cfg, phi, xM, zetaP, M_units, nb, mu, sm, tmem, &bar);
  else tc_fwd_body<NI, H1, H2, NOUT, 1>(cfg, phi, xM, zetaP, M_units, nb, mu, sm, tmem, &bar);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cta_sync();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

// =====================================================================================================
// Backward (hand adjoint of g): recompute forward, δ3 from μ̄_ζ through Φ⁻¹ and the logistic, δ2, δ1, and EVERY weight
// gradient of the tile as ONE tensor-core product.
//
// Shared-memory tiles (bf16 hi/lo planes, [column of the tile][64 features] = one 128-byte swizzle row):
//     U = [ h1 (24) | h2 (24) | x, 1 (8) | - ]          V = [ δ2 (24) | δ3 (8) | δ1 (24) | scratch ]
//   P2  z2    = U[:, h1]·W2          K-major A (first 32 features of U; the weight rows beyond H1 are zero), 3xBF16
//   P3  δ1pre = V[:, δ2]·W2ᵀ         K-major A (first 32 features of V),                                  3xBF16
//   P4  G    += [U_hi | U_lo]ᵀ·V     MN-major operands: the SAME tiles, contraction over the tile's 128 columns.
//       M = 128 = the 64 features of the hi plane and the 64 of the lo plane (descriptor LBO = distance of the
//       planes), N = 56, so the two MMAs per K step (B = V_hi, V_lo) give all four hi/lo products; the halves are
//       added when the accumulator is read, once per CTA.  G holds ∇W2 = h1ᵀδ2, ∇b2 = 1ᵀδ2, ∇W3 = h2ᵀδ3, ∇W1 = xᵀδ1,
//       ∇b1 = 1ᵀδ1 (and cross blocks nobody reads) and stays in TMEM over all tiles of the CTA.
//   The last 16 bytes of a V row (features 56..63) are read by no product: they are the exchange scratch of the column's
//   two threads.
// =====================================================================================================
constexpr int TB_NV = 2 * TC_HP + 8;           // features of V = N of the gradient product (56)
constexpr int TB_OX = 2 * TC_HP;               // first feature of [x, 1] in U
constexpr int TB_OD3 = TC_HP, TB_OD1 = TC_HP + 8;
// shared memory: U hi | U lo | V hi | V lo | W hi | W lo (rows n: [W2[k][n], k < 32 | W2[n][k − 32]]) | tables | slots | xw
constexpr int TB_O_U_HI = 0, TB_O_U_LO = TC_PLANE, TB_O_V_HI = 2 * TC_PLANE, TB_O_V_LO = 3 * TC_PLANE, TB_O_W_HI = 4 * TC_PLANE,
              TB_O_W_LO = TB_O_W_HI + 32 * 128, TB_O_TAB = TB_O_W_LO + 32 * 128, TB_O_SLOT = TB_O_TAB + TAB_BYTES,
              TB_O_XW = TB_O_SLOT + 8 * 8 * 4;
__host__ __device__ constexpr int tb_smem_bytes(int nxa) { return TB_O_XW + nxa * 8 + 1024; }
// D = F32, A = B = BF16 (format 1), N >> 3 [17,23), M >> 4 [24,29); bits 15/16: A/B MN-major
constexpr uint32_t TB_IDESC_K = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t TB_IDESC_G = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TB_NV >> 3) << 17) |
                                ((uint32_t)(128 >> 4) << 24);
constexpr int TB_TMEM_COLS = 128;              // G (64) | z2 (32) | δ1pre (32)
constexpr int TB_CTAS_PER_SM = 3;

template <int NI, int H1, int H2, int NOUT, int HALF>
__device__ __forceinline__ void tc_bwd_body(const Cfg<float>& cfg, const float* __restrict__ phi, const float* __restrict__ xM,
                                            const float* __restrict__ zetaP, const float* __restrict__ mubar, int M_units, int64_t nb,
                                            bool want_x, uint32_t sm, uint32_t tmem, uint64_t* bar, uint64_t* barG, int64_t u0, int64_t u1) {
  using N = MlpTc<NI, H1, H2, NOUT>;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, q = warp & 3;
  const int r = q * 32 + lane;
  const uint32_t u_hi = sm + TB_O_U_HI, u_lo = sm + TB_O_U_LO, v_hi = sm + TB_O_V_HI, v_lo = sm + TB_O_V_LO;
  const uint32_t w_hi = sm + TB_O_W_HI, w_lo = sm + TB_O_W_LO, tab = sm + TB_O_TAB;
  const uint32_t tG = tmem, tZ2 = tmem + 64, tD1 = tmem + 96;
  const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
  // exchange scratch of this column: chunk 7 of its V rows (hi plane: partial sums of the last layer, lo plane: δ3)
  const uint32_t xch_z = v_hi + r * 128 + ((7 ^ (r & 7)) << 4), xch_d = v_lo + r * 128 + ((7 ^ (r & 7)) << 4);
  const int Mu = M_units > 0 ? M_units : 1;
  constexpr int F0 = HALF * TC_HF;   // this thread's first feature
  uint32_t phase = 0, phaseG = 0;
  bool g_pending = false, first = true;
  float base[TC_HF], xs[8];
  int64_t st_prev = -1;
  for (int64_t u = u0; u < u1; u++) {
    const int64_t st = u / Mu;
    const int m = (int)(u - st * Mu);
    const int64_t site = st * TC_COLS + r;
    const bool in = site < nb;
    // cotangent of this thread's output k = HALF (in flight during layer 1)
    float mb = 0.0f;
    if (in && HALF < NOUT && HALF < cfg.nM)
      mb = M_units > 0 ? mubar[((size_t)m * cfg.nM + HALF) * nb + site] : mubar[site * cfg.nM + HALF];
    // ---- E1: layer 1 on the CUDA cores
    if (st != st_prev) { st_prev = st; N::template site_base<HALF>(cfg, tab, xM, site, nb, xs, base); }
    float h1[TC_HF];
    N::template layer1<HALF>(cfg, tab, phi, zetaP, M_units, m, in, base, xs, h1);
    // U is free once the gradient product of the previous tile has read it
    if (g_pending) { mbar_wait(barG, phaseG); phaseG ^= 1; g_pending = false; }
#pragma unroll
    for (int c = 0; c < TC_HF / 4; c++) store4_bf(u_hi, u_lo, r, F0 + 4 * c, h1[4 * c], h1[4 * c + 1], h1[4 * c + 2], h1[4 * c + 3]);
    if (HALF == 0) {
      store4_bf(u_hi, u_lo, r, TB_OX, xs[0], xs[1], xs[2], xs[3]);
      store4_bf(u_hi, u_lo, r, TB_OX + 4, xs[4], xs[5], xs[6], xs[7]);
    }
    tc_publish();
    if (HALF == 0 && t == 0) {   // P2: z2 = h1·W2
      tc_after_sync();
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const uint32_t o = s * 32;
        umma_bf16_(tZ2, make_desc(u_lo + o), make_desc(w_hi + o), TB_IDESC_K, s > 0 ? 1u : 0u);
        umma_bf16_(tZ2, make_desc(u_hi + o), make_desc(w_lo + o), TB_IDESC_K, 1u);
        umma_bf16_(tZ2, make_desc(u_hi + o), make_desc(w_hi + o), TB_IDESC_K, 1u);
      }
      umma_commit_(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_after_sync();
    // ---- E2: h2, outputs, δ3, δ2
    {
      float z[TC_HF], h2[TC_HF], z3[2];
      tmem_ld12(tZ2 + lane_sel + F0, z);
      N::template layer2_epi<HALF>(tab, z, h2, z3);
#pragma unroll
      for (int c = 0; c < TC_HF / 4; c++)
        store4_bf(u_hi, u_lo, r, TC_HP + F0 + 4 * c, h2[4 * c], h2[4 * c + 1], h2[4 * c + 2], h2[4 * c + 3]);
      sts64(xch_z + 8 * HALF, z3[0], z3[1]);
      pair_sync(q);
      const float2 o = lds64(xch_z + 8 * (1 - HALF));
      // δ3[k] = μ̄ · dμ/dp · p(1−p)   (NormalScaling: dμ/dp = σ·√(2π)·exp(q²/2), q = Φ⁻¹(p); src/ModelApplicator.jl:168), k = HALF
      constexpr int k = HALF;
      float d3k = 0.0f;
      if (k < NOUT && k < cfg.nM) {
        const float zk = (k == 0 ? z3[0] + o.x : z3[1] + o.y);
        const float pk = rcpa(1.0f + ex2a(-1.4426950408889634f * zk));
        float d;
        if (cfg.scale_kind == HVI_SCALE_RANGE) d = mb * cfg.scale_b[k];
        else {
          const float qn = normcdfinvf(pk);
          d = mb * cfg.scale_b[k] * 2.50662827463100050242f * ex2a(0.72134752044448170368f * qn * qn);
        }
        d3k = in ? d * pk * (1.0f - pk) : 0.0f;
      }
      sts32(xch_d + 4 * HALF, d3k);
      pair_sync(q);
      const float2 d3 = lds64(xch_d);
      if (HALF == 0) {
        store4_bf(v_hi, v_lo, r, TB_OD3, d3.x, d3.y, 0.0f, 0.0f);
        store4_bf(v_hi, v_lo, r, TB_OD3 + 4, 0.0f, 0.0f, 0.0f, 0.0f);
      }
      float d2[TC_HF];
#pragma unroll
      for (int j = 0; j < TC_HF; j++) {
        if (F0 + j < H2) {
          const float2 w = lds64(tab + 4 * (T_W3 + 2 * (F0 + j)));
          d2[j] = fmaf(d3.x, w.x, d3.y * w.y) * fmaf(-h2[j], h2[j], 1.0f);
        } else d2[j] = 0.0f;
      }
#pragma unroll
      for (int c = 0; c < TC_HF / 4; c++) store4_bf(v_hi, v_lo, r, F0 + 4 * c, d2[4 * c], d2[4 * c + 1], d2[4 * c + 2], d2[4 * c + 3]);
    }
    tc_publish();
    if (HALF == 0 && t == 0) {   // P3: δ1pre = δ2·W2ᵀ  (the transposed weights sit in features 32..63 of the W rows)
      tc_after_sync();
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const uint32_t o = s * 32;
        umma_bf16_(tD1, make_desc(v_lo + o), make_desc(w_hi + 64 + o), TB_IDESC_K, s > 0 ? 1u : 0u);
        umma_bf16_(tD1, make_desc(v_hi + o), make_desc(w_lo + 64 + o), TB_IDESC_K, 1u);
        umma_bf16_(tD1, make_desc(v_hi + o), make_desc(w_hi + 64 + o), TB_IDESC_K, 1u);
      }
      umma_commit_(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_after_sync();
    // ---- E3: δ1 = δ1pre ⊙ (1 − h1²), cotangent of the pbm covariates
    {
      float d1[TC_HF];
      tmem_ld12(tD1 + lane_sel + F0, d1);
#pragma unroll
      for (int j = 0; j < TC_HF; j++) d1[j] = (F0 + j < H1) ? d1[j] * fmaf(-h1[j], h1[j], 1.0f) : 0.0f;
#pragma unroll
      for (int c = 0; c < TC_HF / 4; c++)
        store4_bf(v_hi, v_lo, r, TB_OD1 + F0 + 4 * c, d1[4 * c], d1[4 * c + 1], d1[4 * c + 2], d1[4 * c + 3]);
      if (want_x) {
#pragma unroll
        for (int i = 0; i < NI; i++) {
          if (i >= cfg.nC) {   // x̄[i] = Σ_j W1[i][j] δ1[j]: this thread's features, summed over the warp's columns (fixed-order butterfly)
            float sx = 0.0f;
#pragma unroll
            for (int j4 = 0; j4 < TC_HF / 4; j4++) {
              const float4 w = lds128(tab + 4 * (T_W1 + i * TC_HP + F0 + 4 * j4));
              sx = fmaf(w.x, d1[4 * j4], sx); sx = fmaf(w.y, d1[4 * j4 + 1], sx);
              sx = fmaf(w.z, d1[4 * j4 + 2], sx); sx = fmaf(w.w, d1[4 * j4 + 3], sx);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o);
            if (lane == 0) sts32(sm + TB_O_SLOT + 4 * (warp * 8 + (i - cfg.nC)), sx);
          }
        }
      }
    }
    tc_publish();
    if (HALF == 0 && t == 0) {   // P4: G += [U_hi | U_lo]ᵀ·(V_hi + V_lo), contraction over the 128 columns of the tile
      tc_after_sync();
#pragma unroll
      for (int s = 0; s < TC_COLS / 16; s++) {
        const uint32_t o = s * 2048;
        umma_bf16_(tG, make_desc_mn_(u_hi + o, TC_PLANE), make_desc_mn_(v_lo + o, TC_PLANE), TB_IDESC_G, (first && s == 0) ? 0u : 1u);
        umma_bf16_(tG, make_desc_mn_(u_hi + o, TC_PLANE), make_desc_mn_(v_hi + o, TC_PLANE), TB_IDESC_G, 1u);
      }
      umma_commit_(barG);
    }
    if (want_x && HALF == 1 && t == TC_THREADS / 2) {   // the eight warps' covariate cotangents of this tile, fixed order
      for (int c = 0; c < cfg.npc; c++) {
        float sx = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; w++) sx += lds32(sm + TB_O_SLOT + 4 * (w * 8 + c));
        const uint32_t a = sm + TB_O_XW + 8 * ((M_units > 0 ? m : 0) * cfg.npc + c);
        double acc;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(acc) : "r"(a));
        acc += (double)sx;
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(acc) : "memory");
      }
    }
    first = false;
    g_pending = true;
  }
  if (g_pending) { mbar_wait(barG, phaseG); phaseG ^= 1; }
}

template <int NI, int H1, int H2, int NOUT>
__global__ void __launch_bounds__(TC_THREADS, TB_CTAS_PER_SM)
k_mlp3_tc_bwd(const __grid_constant__ Cfg<float> cfg, const float* __restrict__ phi, const float* __restrict__ xM,
              const float* __restrict__ zetaP, const float* __restrict__ mubar, int M_units, int64_t nb,
              double* __restrict__ part, double* __restrict__ part_x) {
  using N = MlpTc<NI, H1, H2, NOUT>;
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t bar, barG;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sm = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int Mu = M_units > 0 ? M_units : 1;
  const int nxa = Mu * cfg.npc;
  tc_prologue(&bar, &barG, &tmem_base_s, TB_TMEM_COLS);
  // zero the operand tiles: padding features must be finite (they meet zero weight rows in P2 / P3)
  for (int e = t; e < (4 * TC_PLANE) / 16; e += TC_THREADS) sts128u(sm + 16 * e, make_uint4(0u, 0u, 0u, 0u));
  for (int e = t; e < 32 * 64; e += TC_THREADS) {
    const int n = e >> 6, k = e & 63;
    const float w = k < 32 ? N::wfetch(cfg, phi, 1, H1, H2, k, n) : N::wfetch(cfg, phi, 1, H1, H2, n, k - 32);
    store1_bf(sm + TB_O_W_HI, sm + TB_O_W_LO, n, k, w);
  }
  N::build_tables(cfg, phi, sm + TB_O_TAB, t, TC_THREADS);
  for (int e = t; e < nxa * 2; e += TC_THREADS) sts32(sm + TB_O_XW + 4 * e, 0.0f);
  tc_publish();
  tc_after_sync();
  const uint32_t tmem = tmem_base_s;
  // this CTA's contiguous range of units u = site_tile·Mu + m (draws of one site tile are consecutive)
  const int64_t ntot = ((nb + TC_COLS - 1) / TC_COLS) * Mu;
  const int64_t u0 = ntot * blockIdx.x / gridDim.x, u1 = ntot * (blockIdx.x + 1) / gridDim.x;
  if (t < TC_THREADS / 2) tc_bwd_body<NI, H1, H2, NOUT, 0>(cfg, phi, xM, zetaP, mubar, M_units, nb, part_x != nullptr, sm, tmem, &bar, &barG, u0, u1);
  else tc_bwd_body<NI, H1, H2, NOUT, 1>(cfg, phi, xM, zetaP, mubar, M_units, nb, part_x != nullptr, sm, tmem, &bar, &barG, u0, u1);
  tc_after_sync();
  // ---- read G: lane r < 64 holds the hi-plane row of feature r, lane 64 + r the lo-plane row; add, scatter into ϕg's layout.
  // warps 0-3 read columns 0..27 of their lane quarter, warps 4-7 columns 28..55
  float* red = reinterpret_cast<float*>(smem_raw) + ((sm - smem_u32(smem_raw)) >> 2);   // reuse of the U planes: [128][TB_NV]
  cta_sync();   // every thread has seen the last gradient product complete before U is overwritten
  {
    const int q = warp & 3, hf = warp >> 2, r = q * 32 + lane;
    float g[28];
    if (u1 > u0) {
      const uint32_t a = tmem + ((uint32_t)(q * 32) << 16) + 28 * hf;
      tmem_ld8(a, g); tmem_ld8(a + 8, g + 8); tmem_ld8(a + 16, g + 16); tmem_ld4(a + 24, g + 24);
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int c = 0; c < 28; c++) g[c] = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < 28; c++) red[r * TB_NV + 28 * hf + c] = g[c];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  double* out = part + (size_t)blockIdx.x * cfg.nphig;
  for (int e = t; e < 64 * TB_NV; e += TC_THREADS) {
    const int f = e / TB_NV, c = e % TB_NV;   // feature f of U (row of G), feature c of V (column of G)
    const double v = (double)red[f * TB_NV + c] + (double)red[(64 + f) * TB_NV + c];
    if (f < TC_HP) {                                   // row h1[f]: ∇W2[f][j]
      if (f < H1 && c < H2) out[cfg.woff[1] + f * H2 + c] = v;
    } else if (f < 2 * TC_HP) {                        // row h2[f − HP]: ∇W3[·][k]
      const int j = f - TC_HP, k = c - TB_OD3;
      if (j < H2 && k >= 0 && k < NOUT) out[cfg.woff[2] + j * NOUT + k] = v;
    } else {                                           // rows x[i] and the constant 1
      const int i = f - TB_OX;
      if (i < NI) {
        const int j = c - TB_OD1;
        if (j >= 0 && j < H1) out[cfg.woff[0] + i * H1 + j] = v;
      } else if (i == NI) {                            // ∇b1 = 1ᵀδ1, ∇b2 = 1ᵀδ2
        if (c < H2) out[cfg.boff[1] + c] = v;
        const int j = c - TB_OD1;
        if (j >= 0 && j < H1) out[cfg.boff[0] + j] = v;
      }
    }
  }
  if (part_x) {
    for (int e = t; e < nxa; e += TC_THREADS) {
      double v;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sm + TB_O_XW + 8 * (uint32_t)e));
      part_x[(size_t)blockIdx.x * nxa + e] = v;
    }
  }
  __syncthreads();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TB_TMEM_COLS));
}

bool tc_match(const Cfg<float>& c, int ni, int h1, int h2, int no) {
  return c.nL == 3 && c.l_in[0] == ni && c.l_out[0] == h1 && c.l_out[1] == h2 && c.l_out[2] == no &&
         c.l_act[0] == HVI_ACT_TANH && c.l_act[1] == HVI_ACT_TANH && c.l_act[2] == HVI_ACT_LOGISTIC && c.l_bias[0] &&
         c.l_bias[1] && !c.l_bias[2] && c.nM <= no && c.nC <= ni && c.npc <= 8;
}
int tc_mode() {   // HVI_MLP_TC: 0 off, 1 forward only, 2 (default) forward and backward
  const char* e = getenv("HVI_MLP_TC");
  return e ? atoi(e) : 2;
}

#define HVI_TC_CASES(X) X(5, 20, 20, 2) X(6, 24, 24, 2)

}  // namespace

#define HVI_TC_CHECK()                             \
  do {                                             \
    cudaError_t e_ = cudaGetLastError();           \
    if (e_ != cudaSuccess) return e_;              \
    if (L.counter) ++*L.counter;                   \
  } while (0)

// cudaErrorNotSupported: not one of the shapes (or switched off): the caller falls back to the mma.sync kernels
cudaError_t launch_mlp3_tc_fwd(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM, int64_t nb,
                               float* mu, int M_units, const float* zetaP) {
  if (tc_mode() < 1) return cudaErrorNotSupported;
#define X(NI, H1, H2, NO_)                                                                                  \
  if (tc_match(cfg, NI, H1, H2, NO_)) {                                                                     \
    auto kern = k_mlp3_tc_fwd<NI, H1, H2, NO_>;                                                             \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM);       \
    if (e != cudaSuccess) return e;                                                                         \
    const int64_t ntot = ((nb + TC_COLS - 1) / TC_COLS) * (M_units > 0 ? M_units : 1);                      \
    int64_t blocks = ntot;                                                                                  \
    if (blocks > (int64_t)L.sm_count * TF_CTAS_PER_SM) blocks = (int64_t)L.sm_count * TF_CTAS_PER_SM;       \
    if (blocks < 1) blocks = 1;                                                                             \
    kern<<<(int)blocks, TC_THREADS, TF_SMEM, L.stream>>>(cfg, phi, xM, zetaP, M_units, nb, mu);             \
    HVI_TC_CHECK();                                                                                         \
    return cudaSuccess;                                                                                     \
  }
  HVI_TC_CASES(X)
#undef X
  return cudaErrorNotSupported;
}

cudaError_t launch_mlp3_tc_bwd(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM,
                               const float* mubar, int64_t nb, double* part, int* nblk_out, int nblk_max, int M_units,
                               const float* zetaP, double* part_x) {
  if (tc_mode() < 2) return cudaErrorNotSupported;
#define X(NI, H1, H2, NO_)                                                                                  \
  if (tc_match(cfg, NI, H1, H2, NO_)) {                                                                     \
    const int nxa = (M_units > 0 ? M_units : 1) * cfg.npc;                                                  \
    const int smem = tb_smem_bytes(nxa);                                                                    \
    if (smem > 200 * 1024) return cudaErrorNotSupported;                                                    \
    auto kern = k_mlp3_tc_bwd<NI, H1, H2, NO_>;                                                             \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
    if (e != cudaSuccess) return e;                                                                         \
    const int64_t ntot = ((nb + TC_COLS - 1) / TC_COLS) * (M_units > 0 ? M_units : 1);                      \
    int64_t blocks = ntot;                                                                                  \
    const int64_t cap = nblk_max < L.sm_count * TB_CTAS_PER_SM ? nblk_max : L.sm_count * TB_CTAS_PER_SM;    \
    if (blocks > cap) blocks = cap;                                                                         \
    if (blocks < 1) blocks = 1;                                                                             \
    kern<<<(int)blocks, TC_THREADS, smem, L.stream>>>(cfg, phi, xM, zetaP, mubar, M_units, nb, part, part_x); \
    HVI_TC_CHECK();                                                                                         \
    *nblk_out = (int)blocks;                                                                                \
    return cudaSuccess;                                                                                     \
  }
  HVI_TC_CASES(X)
#undef X
  return cudaErrorNotSupported;
}

}  // namespace hvi
