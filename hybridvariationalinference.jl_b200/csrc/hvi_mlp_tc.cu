// hvi_mlp_tc.cu -- the small site-batched ML applicator g(xM; ϕg) on the 5th-generation tensor cores (Float32).
//
// g = NormalScaling ∘ MLP(tanh, tanh, logistic) (src/ModelApplicator.jl:163-176,
// ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52) for the 3-layer shapes of DoubleMM (5-20-20-2 and, with one
// pbm covariate, 6-24-24-2), evaluated on sites or -- pass 2 of generate_ζ, src/elbo.jl:508-515 -- on (draw, site) units.
//
// One CTA of 128 threads owns a tile of 128 columns: thread r = TMEM lane r = column r of the tile.
//   * layer 1 (K = n_covar + pbm covariates) runs on the CUDA cores in exact fp32.  A CTA walks a contiguous range of
//     units u = site_tile·n_MC + draw, so with pbm covariates the site part of the pre-activation (W1[:nC]ᵀ xM + b1) is
//     computed once per site tile and each draw adds the rank-1 term of its covariates (only ζP[idx] changes between
//     the draws of a site, src/elbo.jl:508-515).
//   * layer 2 and every product of the backward pass are tcgen05.mma instructions issued by ONE thread, accumulators in
//     TMEM, completion by tcgen05.commit → mbarrier; the epilogues read their TMEM lane with tcgen05.ld.  Tensor work is
//     asynchronous to the FP32/MUFU epilogues (warp-level mma.sync blocks FP32 issue of its scheduler on B200:
//     profiles/microbench/mma_pipes_b200.txt).
//   * the A operand of every product that contracts over FEATURES (layer 2, δ2·W2ᵀ) is written by its thread straight
//     into TMEM (tcgen05.st, row = lane) and read from there by the MMA: with N = 32 an MMA that fetches its 128-row A
//     tile from shared memory moves 5 KB for 16 cycles of math, and the first two versions of this kernel were bound by
//     shared-memory bandwidth (profiles/r02d_k_mlp3_tc_v1.txt).  Only the small weight tiles (B operands) and, in the
//     backward pass, the tiles of the weight-gradient product (which contracts over COLUMNS, so both operands are
//     MN-major in shared memory) are read from shared memory.
//   * operands are split hi + lo and every product is issued three times (hi·hi + lo·hi + hi·lo).
//     Forward: 3xTF32 (hi = x truncated to tf32, lo = x − hi exactly, of which the tensor core reads the top 19 bits):
//     μ_ζ feeds the σ- and ρ-gradients of the streaming kernel, which amplify its rounding.
//     Backward: 3xBF16 like the mma.sync kernel it replaces; see the block comment above k_mlp3_tc_bwd.
//   * the 2-output last layer, logistic and Φ⁻¹ run on the CUDA cores.
// The mma.sync kernels (hvi_mlp_mma.cu) remain the path of batches below 5·10⁵ columns (the set-up of a CTA costs more than it
// saves there; HVI_MLP_TC=3 takes the tcgen05 kernels at every size, HVI_MLP_TC=0 never) and the cross-check in the tests.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "hvi_internal.cuh"
#include "hvi_tc_common.cuh"

namespace hvi {
namespace {

constexpr int TC_COLS = 128;           // columns per tile = TMEM lanes
constexpr int TC_THREADS = 128;        // epilogue threads (warps 0-3: one per TMEM lane quarter)
constexpr int TC_BLOCK = TC_THREADS + 32;   // + warp 4, whose lane 0 issues every tcgen05.mma of the CTA
constexpr int TC_HP = 24;              // padded hidden width
constexpr int TC_PLANE = TC_COLS * 128;   // one plane of an operand tile: 128 rows of one swizzle row (16 KB)

// D[tmem] (+)= A[tmem]·B[smem]: A is read from tensor memory (lane = row, K along the columns, bf16 packed in pairs)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld24(uint32_t taddr, float (&v)[TC_HP]) {
  tmem_ld8(taddr, v);
  tmem_ld8(taddr + 8, v + 8);
  tmem_ld8(taddr + 16, v + 16);
  tmem_wait_ld();
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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
// two at a time: the FP32 parts as packed instructions (FADD2 / FFMA2: half the issue slots), the MUFU parts per element
__device__ __forceinline__ float2 tanh2_scaled(float2 a) {
  const float2 d = __fadd2_rn(make_float2(ex2a(a.x), ex2a(a.y)), make_float2(1.0f, 1.0f));
  return __ffma2_rn(make_float2(-2.0f, -2.0f), make_float2(rcpa(d.x), rcpa(d.y)), make_float2(1.0f, 1.0f));
}
// the same with ONE reciprocal for the pair: 1/d.x = d.y/(d.x·d.y), 1/d.y = d.x/(d.x·d.y) -- a MUFU.RCP traded for three
// FMUL where the MUFU pipe is the limit (forward kernel).  The arguments are capped at 60 so that the product of the two
// denominators (≤ 2^120) stays finite; tanh is 1 to the last bit far below that.
__device__ __forceinline__ float2 tanh2_scaled_shared(float2 a) {
  const float2 d = __fadd2_rn(make_float2(ex2a(fminf(a.x, 60.0f)), ex2a(fminf(a.y, 60.0f))), make_float2(1.0f, 1.0f));
  const float r = rcpa(d.x * d.y);
  return __ffma2_rn(make_float2(-2.0f, -2.0f), make_float2(r * d.y, r * d.x), make_float2(1.0f, 1.0f));
}
template <int SHR>
__device__ __forceinline__ float2 tanh2_sel(float2 a) {
  if constexpr (SHR > 0) return tanh2_scaled_shared(a);
  else return tanh2_scaled(a);
}
// four tanh with ONE reciprocal: r = 1/(d0 d1 d2 d3), 1/(d0 d1) = r·d2 d3, 1/d0 = d1/(d0 d1), ... (nine FMUL for three MUFU.RCP);
// arguments capped at 30 (product of the four denominators ≤ 2^120; tanh is 1 to the last bit from 2x·log2(e) = 25 on)
template <int SHR>
__device__ __forceinline__ void tanh4_sel(float2 a, float2 b, float2& ta, float2& tb) {
  if constexpr (SHR < 2) { ta = tanh2_sel<SHR>(a); tb = tanh2_sel<SHR>(b); }
  else {
    const float2 one = make_float2(1.0f, 1.0f), m2 = make_float2(-2.0f, -2.0f);
    const float2 da = __fadd2_rn(make_float2(ex2a(fminf(a.x, 30.0f)), ex2a(fminf(a.y, 30.0f))), one);
    const float2 db = __fadd2_rn(make_float2(ex2a(fminf(b.x, 30.0f)), ex2a(fminf(b.y, 30.0f))), one);
    const float pa = da.x * da.y, pb = db.x * db.y;
    const float r = rcpa(pa * pb);
    const float ra = r * pb, rb = r * pa;
    ta = __ffma2_rn(m2, make_float2(ra * da.y, ra * da.x), one);
    tb = __ffma2_rn(m2, make_float2(rb * db.y, rb * db.x), one);
  }
}
// x·(1 − h²) for a pair
__device__ __forceinline__ float2 times_dtanh2(float2 x, float2 h) {
  return __fmul2_rn(x, __ffma2_rn(make_float2(-h.x, -h.y), h, make_float2(1.0f, 1.0f)));
}

// shared-memory accesses by 32-bit shared-window address (the tiles are carved out of an aligned-up dynamic buffer)
__device__ __forceinline__ void sts128u(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }
// hand-off epilogue warps → issuer warp: named barrier `id` counts the 128 non-blocking arrivals and the issuer warp's sync
__device__ __forceinline__ void ready_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(TC_BLOCK) : "memory"); }
__device__ __forceinline__ void ready_wait(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(TC_BLOCK) : "memory"); }

// ---- tf32: hi = trunc_tf32(x), lo = x − hi
__device__ __forceinline__ float trunc_tf(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void store1_tf(uint32_t hi_tile, uint32_t lo_tile, int row, int col, float x) {
  const uint32_t off = row * 128 + ((((col >> 2) ^ (row & 7))) << 4) + ((col & 3) << 2);
  const float h = trunc_tf(x);
  sts32(hi_tile + off, h);
  sts32(lo_tile + off, x - h);
}
// ---- bf16: hi = bf16(x), lo = bf16(x − hi); x0 in the low half of the packed word (even feature = lower K index)
__device__ __forceinline__ void split_bf2(float x0, float x1, uint32_t& h, uint32_t& l) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const float2 d = __ffma2_rn(make_float2(__uint_as_float(h << 16), __uint_as_float(h & 0xffff0000u)), make_float2(-1.0f, -1.0f),
                              make_float2(x0, x1));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(d.y), "f"(d.x));
}
// NW packed words (2·NW features) of row `row` starting at 16-byte chunk `chunk0` of a [128][64 bf16] SWIZZLE_128B tile
template <int NW>
__device__ __forceinline__ void store_words(uint32_t tile, int row, int chunk0, const uint32_t (&w)[NW]) {
#pragma unroll
  for (int c = 0; c < NW / 4; c++)
    sts128u(tile + row * 128 + (((chunk0 + c) ^ (row & 7)) << 4), make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
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

  // site part of the layer-1 pre-activation: b1 + Σ_{i < nC} x_i W1[i][·]; also returns the inputs (and the constant 1)
  __device__ __forceinline__ static void site_base(const Cfg<float>& cfg, uint32_t tab, const float* __restrict__ xM, int64_t site,
                                                   int64_t nb, float (&xs)[8], float (&base)[TC_HP]) {
    const bool in = site < nb;
    const float* row = xM + (in ? site : nb - 1) * cfg.nC;
#pragma unroll
    for (int i = 0; i < 8; i++) xs[i] = (i < NI && i < cfg.nC && in) ? row[i] : (i == NI ? 1.0f : 0.0f);
#pragma unroll
    for (int j4 = 0; j4 < H1 / 4; j4++) {
      const float4 b = lds128(tab + 4 * (T_W1 + NI * TC_HP + 4 * j4));
      base[4 * j4] = b.x; base[4 * j4 + 1] = b.y; base[4 * j4 + 2] = b.z; base[4 * j4 + 3] = b.w;
    }
#pragma unroll
    for (int i = 0; i < NI; i++) {
      if (i < cfg.nC) {
#pragma unroll
        for (int j4 = 0; j4 < H1 / 4; j4++) {
          const float4 w = lds128(tab + 4 * (T_W1 + i * TC_HP + 4 * j4));
          base[4 * j4] = fmaf(xs[i], w.x, base[4 * j4]); base[4 * j4 + 1] = fmaf(xs[i], w.y, base[4 * j4 + 1]);
          base[4 * j4 + 2] = fmaf(xs[i], w.z, base[4 * j4 + 2]); base[4 * j4 + 3] = fmaf(xs[i], w.w, base[4 * j4 + 3]);
        }
      }
    }
  }
  // h1 for one unit: tanh(base + Σ_covariates ζP[idx]·W1[nC + c][·]); records the covariates in xs; features >= H1 are zero
  template <int SHR = 0>
  __device__ __forceinline__ static void layer1(const Cfg<float>& cfg, uint32_t tab, const float* __restrict__ phi,
                                                const float* __restrict__ zetaP, int M_units, int m, bool in, const float (&base)[TC_HP],
                                                float (&xs)[8], float (&h1)[TC_HP]) {
#pragma unroll
    for (int j = 0; j < H1; j++) h1[j] = base[j];
#pragma unroll
    for (int i = 0; i < NI; i++) {
      if (i >= cfg.nC) {   // pbm covariates: ζP[idx] of this draw (pass 1: μP[idx])
        const int idx = cfg.pbm_covar_idx[i - cfg.nC];
        const float sv = in ? (M_units > 0 ? zetaP[(size_t)idx * M_units + m] : phi[cfg.o_muP + idx]) : 0.0f;
        xs[i] = sv;
#pragma unroll
        for (int j4 = 0; j4 < H1 / 4; j4++) {
          const float4 w = lds128(tab + 4 * (T_W1 + i * TC_HP + 4 * j4));
          h1[4 * j4] = fmaf(sv, w.x, h1[4 * j4]); h1[4 * j4 + 1] = fmaf(sv, w.y, h1[4 * j4 + 1]);
          h1[4 * j4 + 2] = fmaf(sv, w.z, h1[4 * j4 + 2]); h1[4 * j4 + 3] = fmaf(sv, w.w, h1[4 * j4 + 3]);
        }
      }
    }
    static_assert(H1 % 4 == 0 && TC_HP % 4 == 0, "hidden widths in fours");
#pragma unroll
    for (int j = 0; j < TC_HP; j += 4) {
      if (j < H1) {
        const float2 k2 = make_float2(K2LOG2E, K2LOG2E);
        float2 va, vb;
        tanh4_sel<SHR>(__fmul2_rn(make_float2(h1[j], h1[j + 1]), k2), __fmul2_rn(make_float2(h1[j + 2], h1[j + 3]), k2), va, vb);
        h1[j] = va.x; h1[j + 1] = va.y; h1[j + 2] = vb.x; h1[j + 3] = vb.y;
      } else h1[j] = h1[j + 1] = h1[j + 2] = h1[j + 3] = 0.0f;
    }
  }
  // h2 from the accumulator row and the pre-activations of the last layer
  template <int SHR = 0>
  __device__ __forceinline__ static void layer2_epi(uint32_t tab, const float (&z)[TC_HP], float (&h2)[TC_HP], float (&z3)[2]) {
    z3[0] = z3[1] = 0.0f;
#pragma unroll
    for (int j4 = 0; j4 < TC_HP / 4; j4++) {
      if (4 * j4 < H2) {
        const float4 b = lds128(tab + 4 * (T_B2 + 4 * j4));
        const float4 wa = lds128(tab + 4 * (T_W3 + 8 * j4)), wb = lds128(tab + 4 * (T_W3 + 8 * j4 + 4));
        const float2 k2 = make_float2(K2LOG2E, K2LOG2E);
        float2 ha, hb;
        tanh4_sel<SHR>(__ffma2_rn(make_float2(z[4 * j4], z[4 * j4 + 1]), k2, make_float2(b.x, b.y)),
                       __ffma2_rn(make_float2(z[4 * j4 + 2], z[4 * j4 + 3]), k2, make_float2(b.z, b.w)), ha, hb);
        h2[4 * j4] = ha.x; h2[4 * j4 + 1] = ha.y; h2[4 * j4 + 2] = hb.x; h2[4 * j4 + 3] = hb.y;
        z3[0] = fmaf(h2[4 * j4], wa.x, z3[0]); z3[1] = fmaf(h2[4 * j4], wa.y, z3[1]);
        z3[0] = fmaf(h2[4 * j4 + 1], wa.z, z3[0]); z3[1] = fmaf(h2[4 * j4 + 1], wa.w, z3[1]);
        z3[0] = fmaf(h2[4 * j4 + 2], wb.x, z3[0]); z3[1] = fmaf(h2[4 * j4 + 2], wb.y, z3[1]);
        z3[0] = fmaf(h2[4 * j4 + 3], wb.z, z3[0]); z3[1] = fmaf(h2[4 * j4 + 3], wb.w, z3[1]);
      } else h2[4 * j4] = h2[4 * j4 + 1] = h2[4 * j4 + 2] = h2[4 * j4 + 3] = 0.0f;
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
// tcgen05.st / generic-proxy writes of this thread → visible to the tensor core; TMEM reads ordered before the next MMA;
// then the non-blocking arrival on the issuer's barrier `id` (id 0: block-wide barrier of the set-up phase instead)
__device__ __forceinline__ void tc_publish(bool smem_tiles, int id) {
  tmem_wait_st();
  if (smem_tiles) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (id == 0) cta_sync();
  else ready_arrive(id);
}
__device__ __forceinline__ void tc_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// =====================================================================================================
// Forward.  TMEM (128 columns): z2 accumulator 0..31 | h1 hi 32..55 | h1 lo 64..87.
// Shared memory: W2 hi | W2 lo ([32 n][32 k] tf32, K-major SWIZZLE_128B) | tables.
// =====================================================================================================
constexpr int TF_O_W_HI = 0, TF_O_W_LO = 32 * 128, TF_O_TAB = 2 * 32 * 128, TF_SMEM = TF_O_TAB + TAB_BYTES + 1024;
constexpr int TF_TMEM_COLS = 128, TF_T_AHI = 32, TF_T_ALO = 64;
// D = F32 [4,6) = 1, A = B = TF32 [7,10) = [10,13) = 2, K-major both, N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t TF_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr int TF_CTAS_PER_SM = 4;

template <int NI, int H1, int H2, int NOUT, int SHR>
__global__ void __launch_bounds__(TC_BLOCK, TF_CTAS_PER_SM)
k_mlp3_tc_fwd(const __grid_constant__ Cfg<float> cfg, const float* __restrict__ phi, const float* __restrict__ xM,
              const float* __restrict__ zetaP, int M_units, int64_t nb, float* __restrict__ mu) {
  using N = MlpTc<NI, H1, H2, NOUT>;
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sm = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int t = threadIdx.x, warp = t >> 5;
  tc_prologue(&bar, nullptr, &tmem_base_s, TF_TMEM_COLS);
  for (int e = t; e < 32 * 32; e += TC_BLOCK) {   // B operand [n = output feature j][k = input feature i]
    const int j = e >> 5, i = e & 31;
    store1_tf(sm + TF_O_W_HI, sm + TF_O_W_LO, j, i, N::wfetch(cfg, phi, 1, H1, H2, i, j));
  }
  N::build_tables(cfg, phi, sm + TF_O_TAB, t, TC_BLOCK);
  tc_publish(true, 0);
  tc_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t w_hi = sm + TF_O_W_HI, w_lo = sm + TF_O_W_LO, tab = sm + TF_O_TAB;
  const int Mu = M_units > 0 ? M_units : 1;
  const int64_t ntot = ((nb + TC_COLS - 1) / TC_COLS) * Mu;
  const int64_t u0 = ntot * blockIdx.x / gridDim.x, u1 = ntot * (blockIdx.x + 1) / gridDim.x;
  constexpr int KS = (H1 + 7) / 8;
  if (warp == TC_THREADS / 32) {
    // ===== issuer warp: z2 = h1·W2, three products per K step of 8; descriptors are loop-invariant
    uint64_t dh[KS], dl[KS];
#pragma unroll
    for (int s_ = 0; s_ < KS; s_++) { dh[s_] = make_desc(w_hi + 32 * s_); dl[s_] = make_desc(w_lo + 32 * s_); }
    for (int64_t u = u0; u < u1; u++) {
      ready_wait(1);
      if (t == TC_THREADS) {
        tc_after_sync();
#pragma unroll
        for (int s_ = 0; s_ < KS; s_++) {
          umma_tf32_ts(tmem, tmem + TF_T_ALO + 8 * s_, dh[s_], TF_IDESC, s_ > 0 ? 1u : 0u);
          umma_tf32_ts(tmem, tmem + TF_T_AHI + 8 * s_, dl[s_], TF_IDESC, 1u);
          umma_tf32_ts(tmem, tmem + TF_T_AHI + 8 * s_, dh[s_], TF_IDESC, 1u);
        }
        umma_commit_(&bar);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue warps: thread t = column t of the tile = TMEM lane t
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;
    float base[TC_HP], xs[8], h1[TC_HP];
    int64_t st_prev = -1;
    auto prepare = [&](int64_t u) {   // layer 1 of unit u into h1
      const int64_t st = u / Mu;
      const int m = (int)(u - st * Mu);
      const int64_t site = st * TC_COLS + t;
      if (st != st_prev) { st_prev = st; N::site_base(cfg, tab, xM, site, nb, xs, base); }
      N::template layer1<SHR>(cfg, tab, phi, zetaP, M_units, m, site < nb, base, xs, h1);
    };
    if (u0 < u1) prepare(u0);
    for (int64_t u = u0; u < u1; u++) {
      const int64_t st = u / Mu;
      const int m = (int)(u - st * Mu);
      const int64_t site = st * TC_COLS + t;
      {   // h1 → TMEM as the A operand: hi = trunc_tf32, lo = the exact residual
        uint32_t hw[TC_HP], lw[TC_HP];
#pragma unroll
        for (int j = 0; j < TC_HP; j++) {
          const float h = trunc_tf(h1[j]);
          hw[j] = __float_as_uint(h);
          lw[j] = __float_as_uint(h1[j] - h);
        }
#pragma unroll
        for (int c = 0; c < KS; c++) { tmem_st8(trow + TF_T_AHI + 8 * c, hw + 8 * c); tmem_st8(trow + TF_T_ALO + 8 * c, lw + 8 * c); }
      }
      tc_publish(false, 1);
      if (u + 1 < u1) prepare(u + 1);   // layer 1 of the next unit while the tensor core works
      mbar_wait(&bar, phase);
      phase ^= 1;
      tc_after_sync();
      float z[TC_HP], h2[TC_HP], z3[2];
      tmem_ld24(trow, z);
      N::template layer2_epi<SHR>(tab, z, h2, z3);
      if (site < nb) {
#pragma unroll
        for (int k = 0; k < NOUT; k++) {
          if (k < cfg.nM) {
            const float p = rcpa(1.0f + ex2a(-1.4426950408889634f * z3[k]));
            const float v = (cfg.scale_kind == HVI_SCALE_RANGE) ? p * cfg.scale_b[k] + cfg.scale_a[k]
                                                                 : cfg.scale_a[k] + cfg.scale_b[k] * normcdfinvf(p);
            if (M_units > 0) mu[((size_t)m * cfg.nM + k) * nb + site] = v;
            else mu[site * cfg.nM + k] = v;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cta_sync();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TF_TMEM_COLS));
}

// =====================================================================================================
// Backward (hand adjoint of g): recompute forward, δ3 from μ̄_ζ through Φ⁻¹ and the logistic, δ2, δ1, and EVERY weight
// gradient of the tile as ONE tensor-core product.
//   P2  z2    = h1·W2           A = h1 (bf16 hi | lo) in TMEM, B = W2 tile in shared memory,   3xBF16
//   P3  δ1pre = δ2·W2ᵀ          A = δ2 in TMEM (same columns as h1 before), B = W2ᵀ tile,      3xBF16
//   P4  G    += [U_hi | U_lo]ᵀ·V    the weight gradients contract over the tile's 128 COLUMNS, so both operands are
//       MN-major shared-memory tiles (bf16 hi/lo planes, [column][64 features] = one 128-byte swizzle row):
//           U = [ h1 (24) | h2 (24) | x, 1 (8) | - ]          V = [ δ2 (24) | δ3 (8) | δ1 (24) | - ]
//       M = 128 = the 64 features of the hi plane and the 64 of the lo plane (descriptor LBO = distance of the
//       planes), N = 56, so the two MMAs per K step (B = V_hi, V_lo) give all four hi/lo products; the halves are
//       added when the accumulator is read, once per CTA.  G holds ∇W2 = h1ᵀδ2, ∇b2 = 1ᵀδ2, ∇W3 = h2ᵀδ3, ∇W1 = xᵀδ1,
//       ∇b1 = 1ᵀδ1 (and cross blocks nobody reads) and stays in TMEM over all tiles of the CTA.
// TMEM (128 columns): G 0..55 | z2 / δ1pre accumulator 64..95 | A hi 96..111 | A lo 112..127 (32 bf16 features each; the
// last 8 stay zero).
// =====================================================================================================
constexpr int TB_NV = 2 * TC_HP + 8;           // features of V = N of the gradient product (56)
constexpr int TB_OX = 2 * TC_HP;               // first feature of [x, 1] in U
constexpr int TB_OD3 = TC_HP, TB_OD1 = TC_HP + 8;
// shared memory: U hi | U lo | V hi | V lo | W hi | W lo (rows n: [W2[k][n], k < 32 | W2[n][k − 32]]) | tables | slots | xw
constexpr int TB_O_U_HI = 0, TB_O_U_LO = TC_PLANE, TB_O_V_HI = 2 * TC_PLANE, TB_O_V_LO = 3 * TC_PLANE, TB_O_W_HI = 4 * TC_PLANE,
              TB_O_W_LO = TB_O_W_HI + 32 * 128, TB_O_TAB = TB_O_W_LO + 32 * 128, TB_O_SLOT = TB_O_TAB + TAB_BYTES,
              TB_O_XW = TB_O_SLOT + 4 * 8 * 4;
__host__ __device__ constexpr int tb_smem_bytes(int nxa) { return TB_O_XW + nxa * 8 + 1024; }
// D = F32, A = B = BF16 (format 1), N >> 3 [17,23), M >> 4 [24,29); bits 15/16: A/B MN-major
constexpr uint32_t TB_IDESC_K = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t TB_IDESC_G = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TB_NV >> 3) << 17) |
                                ((uint32_t)(128 >> 4) << 24);
constexpr int TB_TMEM_COLS = 128, TB_T_ACC = 64, TB_T_AHI = 96, TB_T_ALO = 112;
constexpr int TB_CTAS_PER_SM = 3;

// 24 features → 12 packed hi words and 12 lo words
__device__ __forceinline__ void split24(const float (&x)[TC_HP], uint32_t (&hw)[TC_HP / 2], uint32_t (&lw)[TC_HP / 2]) {
#pragma unroll
  for (int j = 0; j < TC_HP / 2; j++) split_bf2(x[2 * j], x[2 * j + 1], hw[j], lw[j]);
}
__device__ __forceinline__ void a_to_tmem(uint32_t trow, const uint32_t (&hw)[TC_HP / 2], const uint32_t (&lw)[TC_HP / 2]) {
  tmem_st8(trow + TB_T_AHI, hw); tmem_st4(trow + TB_T_AHI + 8, hw + 8);
  tmem_st8(trow + TB_T_ALO, lw); tmem_st4(trow + TB_T_ALO + 8, lw + 8);
}

template <int NI, int H1, int H2, int NOUT>
__global__ void __launch_bounds__(TC_BLOCK, TB_CTAS_PER_SM)
k_mlp3_tc_bwd(const __grid_constant__ Cfg<float> cfg, const float* __restrict__ phi, const float* __restrict__ xM,
              const float* __restrict__ zetaP, const float* __restrict__ mubar, const float* __restrict__ mufwd, int M_units,
              int64_t nb, double* __restrict__ part, double* __restrict__ part_x) {
  using N = MlpTc<NI, H1, H2, NOUT>;
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t bar, barG;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sm = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int Mu = M_units > 0 ? M_units : 1;
  const int nxa = Mu * cfg.npc;
  const bool want_x = part_x != nullptr;
  tc_prologue(&bar, &barG, &tmem_base_s, TB_TMEM_COLS);
  // zero the operand tiles: features nobody writes must be finite
  for (int e = t; e < (4 * TC_PLANE) / 16; e += TC_BLOCK) sts128u(sm + 16 * e, make_uint4(0u, 0u, 0u, 0u));
  for (int e = t; e < 32 * 64; e += TC_BLOCK) {
    const int n = e >> 6, k = e & 63;
    const float w = k < 32 ? N::wfetch(cfg, phi, 1, H1, H2, k, n) : N::wfetch(cfg, phi, 1, H1, H2, n, k - 32);
    store1_bf(sm + TB_O_W_HI, sm + TB_O_W_LO, n, k, w);
  }
  N::build_tables(cfg, phi, sm + TB_O_TAB, t, TC_BLOCK);
  for (int e = t; e < nxa * 2; e += TC_BLOCK) sts32(sm + TB_O_XW + 4 * e, 0.0f);
  cta_sync();   // the TMEM base address is in shared memory
  tc_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  if (warp < TC_THREADS / 32) {   // features 24..31 of the TMEM A operand meet zero weight rows: they must be finite
    const uint32_t zero[4] = {0u, 0u, 0u, 0u};
    tmem_st4(trow + TB_T_AHI + 12, zero);
    tmem_st4(trow + TB_T_ALO + 12, zero);
  }
  tc_publish(true, 0);
  tc_after_sync();
  const uint32_t u_hi = sm + TB_O_U_HI, u_lo = sm + TB_O_U_LO, v_hi = sm + TB_O_V_HI, v_lo = sm + TB_O_V_LO;
  const uint32_t w_hi = sm + TB_O_W_HI, w_lo = sm + TB_O_W_LO, tab = sm + TB_O_TAB;
  const uint32_t tG = tmem, tAcc = tmem + TB_T_ACC;
  // this CTA's contiguous range of units u = site_tile·Mu + m (draws of one site tile are consecutive)
  const int64_t ntot = ((nb + TC_COLS - 1) / TC_COLS) * Mu;
  const int64_t u0 = ntot * blockIdx.x / gridDim.x, u1 = ntot * (blockIdx.x + 1) / gridDim.x;
  if (warp == TC_THREADS / 32) {
    // ===== issuer warp: lane 0 issues the three products of every tile; the descriptors are loop-invariant
    uint64_t d2h[2], d2l[2], d3h[2], d3l[2];
#pragma unroll
    for (int s_ = 0; s_ < 2; s_++) {
      d2h[s_] = make_desc(w_hi + 32 * s_); d2l[s_] = make_desc(w_lo + 32 * s_);
      d3h[s_] = make_desc(w_hi + 64 + 32 * s_); d3l[s_] = make_desc(w_lo + 64 + 32 * s_);
    }
    const uint64_t du = make_desc_mn_(u_hi, TC_PLANE), dvh = make_desc_mn_(v_hi, TC_PLANE), dvl = make_desc_mn_(v_lo, TC_PLANE);
    bool first = true;
    for (int64_t u = u0; u < u1; u++) {
      ready_wait(1);
      if (t == TC_THREADS) {   // P2: z2 = h1·W2
        tc_after_sync();
#pragma unroll
        for (int s_ = 0; s_ < 2; s_++) {
          umma_bf16_ts(tAcc, tmem + TB_T_ALO + 8 * s_, d2h[s_], TB_IDESC_K, s_ > 0 ? 1u : 0u);
          umma_bf16_ts(tAcc, tmem + TB_T_AHI + 8 * s_, d2l[s_], TB_IDESC_K, 1u);
          umma_bf16_ts(tAcc, tmem + TB_T_AHI + 8 * s_, d2h[s_], TB_IDESC_K, 1u);
        }
        umma_commit_(&bar);
      }
      __syncwarp();
      ready_wait(2);
      if (t == TC_THREADS) {   // P3: δ1pre = δ2·W2ᵀ  (the transposed weights sit in features 32..63 of the W rows)
        tc_after_sync();
#pragma unroll
        for (int s_ = 0; s_ < 2; s_++) {
          umma_bf16_ts(tAcc, tmem + TB_T_ALO + 8 * s_, d3h[s_], TB_IDESC_K, s_ > 0 ? 1u : 0u);
          umma_bf16_ts(tAcc, tmem + TB_T_AHI + 8 * s_, d3l[s_], TB_IDESC_K, 1u);
          umma_bf16_ts(tAcc, tmem + TB_T_AHI + 8 * s_, d3h[s_], TB_IDESC_K, 1u);
        }
        umma_commit_(&bar);
      }
      __syncwarp();
      ready_wait(3);
      if (t == TC_THREADS) {   // P4: G += [U_hi | U_lo]ᵀ·(V_hi + V_lo), contraction over the 128 columns of the tile
        tc_after_sync();
#pragma unroll
        for (int s_ = 0; s_ < TC_COLS / 16; s_++) {   // 16 rows = 2048 bytes = 128 descriptor units per K step
          umma_bf16_(tG, du + 128 * s_, dvl + 128 * s_, TB_IDESC_G, (first && s_ == 0) ? 0u : 1u);
          umma_bf16_(tG, du + 128 * s_, dvh + 128 * s_, TB_IDESC_G, 1u);
        }
        if (u + 1 == u1) umma_commit_(&barG);
      }
      __syncwarp();
      first = false;
    }
  } else {
  uint32_t phase = 0;
  float base[TC_HP], xs[8];
  int64_t st_prev = -1;
  for (int64_t u = u0; u < u1; u++) {
    const int64_t st = u / Mu;
    const int m = (int)(u - st * Mu);
    const int64_t site = st * TC_COLS + t;
    const bool in = site < nb;
    // cotangent of this column's outputs and the forward pass's outputs (in flight during layer 1)
    float mb[2] = {0.0f, 0.0f}, muf[2] = {0.0f, 0.0f};
    if (in) {
#pragma unroll
      for (int k = 0; k < NOUT; k++)
        if (k < cfg.nM) {
          const size_t o = M_units > 0 ? ((size_t)m * cfg.nM + k) * nb + site : (size_t)site * cfg.nM + k;
          mb[k] = mubar[o];
          if (mufwd) muf[k] = mufwd[o];
        }
    }
    // ---- E1: layer 1 on the CUDA cores
    if (st != st_prev) { st_prev = st; N::site_base(cfg, tab, xM, site, nb, xs, base); }
    float h1[TC_HP];
    N::layer1(cfg, tab, phi, zetaP, M_units, m, in, base, xs, h1);
    // h1 goes to TMEM now (A operand of P2) and to the U tile only in E2: the gradient product of the previous tile may
    // still be reading U and V, and P2 of this tile is queued behind it on the tensor pipe anyway
    uint32_t h1h[TC_HP / 2], h1l[TC_HP / 2];
    split24(h1, h1h, h1l);
    a_to_tmem(trow, h1h, h1l);
    tc_publish(false, 1);   // P2 reads TMEM and the weight tile only
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_after_sync();
    // ---- E2: h2, outputs, δ3, δ2.  P2 has completed, hence also the gradient product issued before it (tcgen05.commit
    // covers every earlier MMA of the issuing thread): U and V are free
    {
      float z[TC_HP], h2[TC_HP], z3[2], d3[2] = {0.0f, 0.0f};
      tmem_ld24(trow + TB_T_ACC, z);
      {
        uint32_t xh[4], xl[4];
#pragma unroll
        for (int j = 0; j < 4; j++) split_bf2(xs[2 * j], xs[2 * j + 1], xh[j], xl[j]);
        store_words<TC_HP / 2>(u_hi, t, 0, h1h);
        store_words<TC_HP / 2>(u_lo, t, 0, h1l);
        store_words<4>(u_hi, t, TB_OX / 8, xh);
        store_words<4>(u_lo, t, TB_OX / 8, xl);
      }
      N::layer2_epi(tab, z, h2, z3);
      // δ3 = μ̄ · dμ/dp · p(1−p)   (NormalScaling: dμ/dp = σ·√(2π)·exp(q²/2), q = Φ⁻¹(p); src/ModelApplicator.jl:168)
#pragma unroll
      for (int k = 0; k < NOUT; k++) {
        if (k < cfg.nM) {
          const float pk = rcpa(1.0f + ex2a(-1.4426950408889634f * z3[k]));
          float d;
          if (cfg.scale_kind == HVI_SCALE_RANGE) d = mb[k] * cfg.scale_b[k];
          else {
            // q = Φ⁻¹(p): from the forward pass's μ_ζ = a + b·q when the caller kept it (3xTF32 there), else recomputed
            const float qn = mufwd ? (muf[k] - cfg.scale_a[k]) * rcpa(cfg.scale_b[k]) : normcdfinvf(pk);
            d = mb[k] * cfg.scale_b[k] * 2.50662827463100050242f * ex2a(0.72134752044448170368f * qn * qn);
          }
          d3[k] = in ? d * pk * (1.0f - pk) : 0.0f;
        }
      }
      uint32_t hw[TC_HP / 2], lw[TC_HP / 2];
      split24(h2, hw, lw);
      store_words<TC_HP / 2>(u_hi, t, TC_HP / 8, hw);
      store_words<TC_HP / 2>(u_lo, t, TC_HP / 8, lw);
      float d2[TC_HP];
#pragma unroll
      for (int j4 = 0; j4 < TC_HP / 4; j4++) {
        if (4 * j4 < H2) {
          const float4 wa = lds128(tab + 4 * (T_W3 + 8 * j4)), wb = lds128(tab + 4 * (T_W3 + 8 * j4 + 4));
          const float2 D0 = make_float2(d3[0], d3[0]), D1 = make_float2(d3[1], d3[1]);
          const float2 ta = __ffma2_rn(D0, make_float2(wa.x, wa.z), __fmul2_rn(D1, make_float2(wa.y, wa.w)));
          const float2 tb = __ffma2_rn(D0, make_float2(wb.x, wb.z), __fmul2_rn(D1, make_float2(wb.y, wb.w)));
          const float2 ra = times_dtanh2(ta, make_float2(h2[4 * j4], h2[4 * j4 + 1]));
          const float2 rb = times_dtanh2(tb, make_float2(h2[4 * j4 + 2], h2[4 * j4 + 3]));
          d2[4 * j4] = ra.x; d2[4 * j4 + 1] = ra.y; d2[4 * j4 + 2] = rb.x; d2[4 * j4 + 3] = rb.y;
        } else d2[4 * j4] = d2[4 * j4 + 1] = d2[4 * j4 + 2] = d2[4 * j4 + 3] = 0.0f;
      }
      split24(d2, hw, lw);
      a_to_tmem(trow, hw, lw);
      store_words<TC_HP / 2>(v_hi, t, 0, hw);
      store_words<TC_HP / 2>(v_lo, t, 0, lw);
      uint32_t d3h[4] = {0u, 0u, 0u, 0u}, d3l[4] = {0u, 0u, 0u, 0u};
      split_bf2(d3[0], d3[1], d3h[0], d3l[0]);
      store_words<4>(v_hi, t, TB_OD3 / 8, d3h);
      store_words<4>(v_lo, t, TB_OD3 / 8, d3l);
    }
    tc_publish(false, 2);
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_after_sync();
    // ---- E3: δ1 = δ1pre ⊙ (1 − h1²), cotangent of the pbm covariates
    {
      float d1[TC_HP];
      tmem_ld24(trow + TB_T_ACC, d1);
#pragma unroll
      for (int j = 0; j < TC_HP; j += 2) {
        if (j < H1) {
          const float2 v = times_dtanh2(make_float2(d1[j], d1[j + 1]), make_float2(h1[j], h1[j + 1]));
          d1[j] = v.x; d1[j + 1] = v.y;
        } else d1[j] = d1[j + 1] = 0.0f;
      }
      uint32_t hw[TC_HP / 2], lw[TC_HP / 2];
      split24(d1, hw, lw);
      store_words<TC_HP / 2>(v_hi, t, TB_OD1 / 8, hw);
      store_words<TC_HP / 2>(v_lo, t, TB_OD1 / 8, lw);
      if (want_x) {
#pragma unroll
        for (int i = 0; i < NI; i++) {
          if (i >= cfg.nC) {   // x̄[i] = Σ_j W1[i][j] δ1[j], summed over the warp's columns (fixed-order butterfly)
            float sx = 0.0f;
#pragma unroll
            for (int j4 = 0; j4 < H1 / 4; j4++) {
              const float4 w = lds128(tab + 4 * (T_W1 + i * TC_HP + 4 * j4));
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
    tc_publish(true, 3);
    if (want_x) asm volatile("bar.sync 4, %0;" ::"n"(TC_THREADS) : "memory");   // the four warps' slots are written
    if (want_x && t == 32) {   // the four warps' covariate cotangents of this tile, fixed order, into the draw's double sum
      for (int c = 0; c < cfg.npc; c++) {
        const float sx = (lds32(sm + TB_O_SLOT + 4 * c) + lds32(sm + TB_O_SLOT + 4 * (8 + c))) +
                         (lds32(sm + TB_O_SLOT + 4 * (16 + c)) + lds32(sm + TB_O_SLOT + 4 * (24 + c)));
        const uint32_t a = sm + TB_O_XW + 8 * ((M_units > 0 ? m : 0) * cfg.npc + c);
        double acc;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(acc) : "r"(a));
        acc += (double)sx;
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(acc) : "memory");
      }
    }
  }
  }   // epilogue warps
  if (u1 > u0) mbar_wait(&barG, 0);
  tc_after_sync();
  // ---- read G: lane r < 64 holds the hi-plane row of feature r, lane 64 + r the lo-plane row; add, scatter into ϕg's layout
  float* red = reinterpret_cast<float*>(smem_raw) + ((sm - smem_u32(smem_raw)) >> 2);   // reuse of the U planes: [128][TB_NV]
  cta_sync();   // every thread has seen the last gradient product complete before U is overwritten
#pragma unroll 1
  for (int c = 0; c < TB_NV / 8 && warp < TC_THREADS / 32; c++) {
    float g[8];
    if (u1 > u0) { tmem_ld8(trow + 8 * c, g); tmem_wait_ld(); }
    else {
#pragma unroll
      for (int q = 0; q < 8; q++) g[q] = 0.0f;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) red[t * TB_NV + 8 * c + q] = g[q];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  double* out = part + (size_t)blockIdx.x * cfg.nphig;
  for (int e = t; e < 64 * TB_NV; e += TC_BLOCK) {
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
    for (int e = t; e < nxa; e += TC_BLOCK) {
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
int tc_mode() {   // HVI_MLP_TC: 0 off, 1 forward only, 2 (default) forward and backward, 3 the same for every batch size
  const char* e = getenv("HVI_MLP_TC");
  return e ? atoi(e) : 2;
}
// The set-up of a CTA (operand tiles zeroed, weight tiles split and swizzled, TMEM allocation) and the read-out of the gradient
// accumulator cost a few µs per launch: below ≈ 5·10⁵ columns the mma.sync kernels are the faster ones (forward + backward,
// 10⁵ sites: 13 + 20 against 15 + 24 µs; 10⁶: 53 + 92 against 48 + 89; 1.25·10⁶: 64 + 113 against 57 + 105; 10⁷: 451 + 809
// against 388 + 706)
constexpr int64_t TC_MIN_COLUMNS = 500000;
bool tc_worthwhile(int64_t nb, int M_units) { return tc_mode() >= 3 || nb * (int64_t)(M_units > 0 ? M_units : 1) >= TC_MIN_COLUMNS; }

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
  if (tc_mode() < 1 || !tc_worthwhile(nb, M_units)) return cudaErrorNotSupported;
  // HVI_TANH_SHARED=0: one reciprocal per tanh (two MUFU each); 1: one per pair of tanh; 2: one per four
  const char* shr_e = getenv("HVI_TANH_SHARED");
  const int shr = shr_e ? atoi(shr_e) : 1;
#define X(NI, H1, H2, NO_)                                                                                  \
  if (tc_match(cfg, NI, H1, H2, NO_)) {                                                                     \
    auto kern = shr == 2 ? k_mlp3_tc_fwd<NI, H1, H2, NO_, 2> : shr == 1 ? k_mlp3_tc_fwd<NI, H1, H2, NO_, 1> \
                                                                        : k_mlp3_tc_fwd<NI, H1, H2, NO_, 0>; \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM);       \
    if (e != cudaSuccess) return e;                                                                         \
    const int64_t ntot = ((nb + TC_COLS - 1) / TC_COLS) * (M_units > 0 ? M_units : 1);                      \
    int64_t blocks = ntot;                                                                                  \
    if (blocks > (int64_t)L.sm_count * TF_CTAS_PER_SM) blocks = (int64_t)L.sm_count * TF_CTAS_PER_SM;       \
    if (blocks < 1) blocks = 1;                                                                             \
    kern<<<(int)blocks, TC_BLOCK, TF_SMEM, L.stream>>>(cfg, phi, xM, zetaP, M_units, nb, mu);             \
    HVI_TC_CHECK();                                                                                         \
    return cudaSuccess;                                                                                     \
  }
  HVI_TC_CASES(X)
#undef X
  return cudaErrorNotSupported;
}

cudaError_t launch_mlp3_tc_bwd(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM,
                               const float* mubar, int64_t nb, double* part, int* nblk_out, int nblk_max, int M_units,
                               const float* zetaP, double* part_x, const float* mufwd) {
  if (tc_mode() < 2 || !tc_worthwhile(nb, M_units)) return cudaErrorNotSupported;
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
    kern<<<(int)blocks, TC_BLOCK, smem, L.stream>>>(cfg, phi, xM, zetaP, mubar, mufwd, M_units, nb, part, part_x); \
    HVI_TC_CHECK();                                                                                         \
    *nblk_out = (int)blocks;                                                                                \
    return cudaSuccess;                                                                                     \
  }
  HVI_TC_CASES(X)
#undef X
  return cudaErrorNotSupported;
}

}  // namespace hvi
