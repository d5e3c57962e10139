// hvi_mlp_tc.cu -- the small site-batched ML applicator g(xM; ϕg) on the 5th-generation tensor cores (Float32).
//
// g = NormalScaling ∘ MLP(tanh, tanh, logistic) (src/ModelApplicator.jl:163-176,
// ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52) for the 3-layer shapes of DoubleMM (5-20-20-2 and, with one
// pbm covariate, 6-24-24-2), evaluated on sites or -- pass 2 of generate_ζ, src/elbo.jl:508-515 -- on (draw, site) units.
//
// One CTA of 128 threads owns a tile of 128 columns: thread r = TMEM lane r = column r of the tile.
//   * operands live in shared memory as K-major SWIZZLE_128B tiles [row][32 tf32], split hi = tf32(x), lo = x − hi;
//     every product is issued three times (hi·hi + lo·hi + hi·lo, "3xTF32", ≈2^-21 relative) by ONE thread as
//     tcgen05.mma.cta_group::1.kind::tf32, M = 128 (the columns), N = 32 (the features of the layer), K = 8 per instruction;
//   * accumulators live in TMEM (two of 32 columns per CTA); the epilogue of a layer is thread-per-column:
//     tcgen05.ld of the lane's row, bias + tanh (ex2 + rcp), split, and the row goes back to shared memory as the A
//     operand of the next layer.  The layer-1 bias rides in the MMA as the weight row of a constant-1 input; the layer-2
//     bias is folded into the argument scaling of ex2.  The 2-output last layer, logistic and Φ⁻¹ run on the CUDA cores.
//   * the tensor work is asynchronous to the FP32/MUFU epilogues (unlike warp-level mma.sync, which blocks FP32 issue of
//     its scheduler on B200: profiles/microbench/mma_pipes_b200.txt); several CTAs per SM interleave the phases.
// The legacy mma.sync kernels (hvi_mlp_mma.cu) remain the fallback (HVI_MLP_TC=0) and the cross-check in the tests.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "hvi_internal.cuh"
#include "hvi_tc_common.cuh"

namespace hvi {
namespace {

constexpr int TCS_BM = 128;            // columns per tile = TMEM lanes = threads per CTA
constexpr int TCS_N = 32;              // padded feature width of a hidden layer = N of the MMAs
constexpr int TCS_ROW = 128;           // bytes per operand row: 32 tf32 = one swizzle row
constexpr int TCS_A_BYTES = TCS_BM * TCS_ROW;   // one plane of an activation tile (16 KB)
constexpr int TCS_W_BYTES = TCS_N * TCS_ROW;    // one plane of a weight tile (4 KB)

// instruction descriptor: D = F32 [4,6) = 1, A = B = TF32 [7,10) = [10,13) = 2, K-major both, N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t TCS_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TCS_N >> 3) << 17) | ((uint32_t)(TCS_BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

// tf32 hi (round to nearest, integer add + mask) and exact residual lo
__device__ __forceinline__ void split_tf(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}
// shared-memory accesses by 32-bit shared-window address (the tiles are carved out of an aligned-up dynamic buffer,
// which would otherwise make the compiler fall back to generic LD/ST)
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
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
// four consecutive features (one 16-byte chunk) of row `row` of a K-major SWIZZLE_128B tile pair
__device__ __forceinline__ void store_chunk(uint32_t hi_tile, uint32_t lo_tile, int row, int chunk, const float (&v)[4]) {
  float4 h, l;
  split_tf(v[0], h.x, l.x); split_tf(v[1], h.y, l.y); split_tf(v[2], h.z, l.z); split_tf(v[3], h.w, l.w);
  const uint32_t off = row * TCS_ROW + ((chunk ^ (row & 7)) << 4);
  sts128(hi_tile + off, h);
  sts128(lo_tile + off, l);
}
__device__ __forceinline__ void store_elem(uint32_t hi_tile, uint32_t lo_tile, int row, int col, float x) {
  float h, l;
  split_tf(x, h, l);
  const uint32_t off = row * TCS_ROW + ((((col >> 2) ^ (row & 7))) << 4) + ((col & 3) << 2);
  sts32(hi_tile + off, h);
  sts32(lo_tile + off, l);
}

template <int NI, int H1, int H2, int NOUT>
struct MlpTc {
  static_assert(NI + 1 <= 8, "inputs + the bias one fill one K step of 8");
  static_assert(H1 <= TCS_N && H2 <= TCS_N && H1 % 4 == 0 && H2 % 4 == 0, "hidden widths up to 32, whole 16-byte chunks");
  static_assert(NOUT <= 2, "two outputs on the CUDA cores");
  static constexpr int KS2 = (H1 + 7) / 8;      // K steps of the second layer
  static constexpr int NL1 = (H1 + 7) / 8;      // 8-column TMEM loads per accumulator row
  static constexpr int NL2 = (H2 + 7) / 8;

  // shared memory (dynamic, 1024-byte aligned): activation tile hi | lo, W1 hi | lo, W2 hi | lo, then small tables
  static constexpr int O_A_HI = 0, O_A_LO = TCS_A_BYTES, O_W1_HI = 2 * TCS_A_BYTES, O_W1_LO = O_W1_HI + TCS_W_BYTES,
                       O_W2_HI = O_W1_LO + TCS_W_BYTES, O_W2_LO = O_W2_HI + TCS_W_BYTES, O_TAB = O_W2_LO + TCS_W_BYTES;
  static constexpr int TAB_FLOATS = TCS_N + TCS_N * 2;   // prescaled b2 | W3[j][k]
  static constexpr int SMEM_BYTES = O_TAB + TAB_FLOATS * 4 + 1024;

  __device__ static float wfetch(const Cfg<float>& cfg, const float* __restrict__ phi, int l, int nin, int nout, int i, int j) {
    return (i < nin && j < nout) ? phi[cfg.woff[l] + i * nout + j] : 0.0f;
  }
  // B operands [n = output feature][k = input feature]; row NI of W1 is the bias b1 (multiplies the constant-1 input)
  __device__ static void build_weights(const Cfg<float>& cfg, const float* __restrict__ phi, uint32_t sm, int tid, int nthr) {
    for (int e = tid; e < TCS_N * TCS_N; e += nthr) {
      const int j = e >> 5, i = e & 31;
      float w1 = wfetch(cfg, phi, 0, NI, H1, i, j);
      if (i == NI && j < H1) w1 = phi[cfg.boff[0] + j];
      store_elem(sm + O_W1_HI, sm + O_W1_LO, j, i, w1);
      store_elem(sm + O_W2_HI, sm + O_W2_LO, j, i, wfetch(cfg, phi, 1, H1, H2, i, j));
    }
    const uint32_t tab = sm + O_TAB;
    for (int e = tid; e < TCS_N; e += nthr) sts32(tab + 4 * e, e < H2 ? phi[cfg.boff[1] + e] * K2LOG2E : 0.0f);
    for (int e = tid; e < TCS_N * 2; e += nthr) {
      const int j = e >> 1, k = e & 1;
      sts32(tab + 4 * (TCS_N + e), (j < H2 && k < NOUT) ? phi[cfg.woff[2] + j * NOUT + k] : 0.0f);
    }
  }

  // inputs of this thread's column: x[0..NI) then the constant 1; columns beyond the batch are all zero but the 1
  __device__ __forceinline__ static void load_x(const Cfg<float>& cfg, const float* __restrict__ phi, const float* __restrict__ xM,
                                                const float* __restrict__ zetaP, int M_units, int m, int64_t site, int64_t nb,
                                                float (&xv)[8]) {
    const bool in = site < nb;
    const float* row = xM + (in ? site : nb - 1) * cfg.nC;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      float v = 0.0f;
      if (i < NI) {
        if (i < cfg.nC) v = row[i];
        else {
          const int idx = cfg.pbm_covar_idx[i - cfg.nC];
          v = M_units > 0 ? zetaP[(size_t)idx * M_units + m] : phi[cfg.o_muP + idx];
        }
        if (!in) v = 0.0f;
      } else if (i == NI) v = 1.0f;
      xv[i] = v;
    }
  }
  __device__ __forceinline__ static void store_x(uint32_t sm, int row, const float (&xv)[8]) {
    const float a[4] = {xv[0], xv[1], xv[2], xv[3]}, b[4] = {xv[4], xv[5], xv[6], xv[7]};
    store_chunk(sm + O_A_HI, sm + O_A_LO, row, 0, a);
    store_chunk(sm + O_A_HI, sm + O_A_LO, row, 1, b);
  }
  // D(tmem_d) = A·Bᵀ over `ks` K steps of 8, three products per step (issued by one thread)
  __device__ __forceinline__ static void issue_layer(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int ks) {
    for (int s = 0; s < ks; s++) {
      const uint32_t o = s * 32;
      umma_tf32(tmem_d, make_desc(a_lo + o), make_desc(b_hi + o), TCS_IDESC, s > 0 ? 1u : 0u);
      umma_tf32(tmem_d, make_desc(a_hi + o), make_desc(b_lo + o), TCS_IDESC, 1u);
      umma_tf32(tmem_d, make_desc(a_hi + o), make_desc(b_hi + o), TCS_IDESC, 1u);
    }
  }
  // h = tanh(z) of this thread's accumulator row (bias already inside z), features >= H are not computed
  template <int H, int NLD>
  __device__ __forceinline__ static void load_tanh(uint32_t taddr, uint32_t bias_scaled, float (&h)[8 * NLD]) {
    float z[NLD][8];
#pragma unroll
    for (int c = 0; c < NLD; c++) tmem_ld8(taddr + 8 * c, z[c]);
    float b[8 * NLD];
    if (bias_scaled) {
#pragma unroll
      for (int c = 0; c < 2 * NLD; c++) {
        const float4 v = lds128(bias_scaled + 16 * c);
        b[4 * c] = v.x; b[4 * c + 1] = v.y; b[4 * c + 2] = v.z; b[4 * c + 3] = v.w;
      }
    }
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 8 * NLD; j++) {
      if (j < H) {
        const float a = bias_scaled ? fmaf(z[j >> 3][j & 7], K2LOG2E, b[j]) : z[j >> 3][j & 7] * K2LOG2E;
        h[j] = fmaf(-2.0f, rcpa(ex2a(a) + 1.0f), 1.0f);
      } else h[j] = 0.0f;
    }
  }
};

constexpr int TCS_CTAS_PER_SM = 4;

template <int NI, int H1, int H2, int NOUT>
__global__ void __launch_bounds__(TCS_BM, TCS_CTAS_PER_SM)
k_mlp3_tc_fwd(const __grid_constant__ Cfg<float> cfg, const float* __restrict__ phi, const float* __restrict__ xM,
              const float* __restrict__ zetaP, int M_units, int64_t nb, float* __restrict__ mu) {
  using N = MlpTc<NI, H1, H2, NOUT>;
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sm = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  N::build_weights(cfg, phi, sm, t, TCS_BM);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's lane quarter
  const uint32_t a_hi = sm + N::O_A_HI, a_lo = sm + N::O_A_LO;
  const uint32_t w1_hi = sm + N::O_W1_HI, w1_lo = sm + N::O_W1_LO;
  const uint32_t w2_hi = sm + N::O_W2_HI, w2_lo = sm + N::O_W2_LO;
  const uint32_t tab = sm + N::O_TAB;

  const int64_t ntiles = (nb + TCS_BM - 1) / TCS_BM;
  const int64_t ntot = ntiles * (M_units > 0 ? M_units : 1);
  uint32_t phase = 0;
  float xv[8];
  int64_t tile = blockIdx.x;
  int m = 0;
  int64_t site = 0;
  if (tile < ntot) {
    m = M_units > 0 ? (int)(tile / ntiles) : 0;
    site = (tile - (int64_t)m * ntiles) * TCS_BM + t;
    N::load_x(cfg, phi, xM, zetaP, M_units, m, site, nb, xv);
  }
  for (; tile < ntot; tile += gridDim.x) {
    const int m_cur = m;
    const int64_t site_cur = site;
    // ---- layer 1: [x ; 1] · [W1 ; b1]
    N::store_x(sm, t, xv);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      N::issue_layer(tmem, a_hi, a_lo, w1_hi, w1_lo, 1);
      umma_commit_(&bar);
    }
    // the next tile's inputs travel while the tensor core works
    if (tile + gridDim.x < ntot) {
      const int64_t nt = tile + gridDim.x;
      m = M_units > 0 ? (int)(nt / ntiles) : 0;
      site = (nt - (int64_t)m * ntiles) * TCS_BM + t;
      N::load_x(cfg, phi, xM, zetaP, M_units, m, site, nb, xv);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      float h1[8 * N::NL1];
      N::template load_tanh<H1, N::NL1>(trow, 0u, h1);
#pragma unroll
      for (int c = 0; c < 2 * N::KS2; c++) {
        const float v[4] = {h1[4 * c], h1[4 * c + 1], h1[4 * c + 2], h1[4 * c + 3]};
        store_chunk(sm + N::O_A_HI, sm + N::O_A_LO, t, c, v);
      }
    }
    // ---- layer 2
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      N::issue_layer(tmem + TCS_N, a_hi, a_lo, w2_hi, w2_lo, N::KS2);
      umma_commit_(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float h2[8 * N::NL2];
    N::template load_tanh<H2, N::NL2>(trow + TCS_N, tab, h2);
    // ---- layer 3 (NOUT <= 2 outputs), logistic, NormalScaling on the CUDA cores
    float z3[2] = {0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < H2; j++) {
      const float2 w = lds64(tab + 4 * (TCS_N + 2 * j));
      z3[0] = fmaf(h2[j], w.x, z3[0]);
      z3[1] = fmaf(h2[j], w.y, z3[1]);
    }
    if (site_cur < nb) {
#pragma unroll
      for (int k = 0; k < NOUT; k++) {
        if (k < cfg.nM) {
          const float p = rcpa(1.0f + ex2a(-1.4426950408889634f * z3[k]));
          const float v = (cfg.scale_kind == HVI_SCALE_RANGE) ? p * cfg.scale_b[k] + cfg.scale_a[k]
                                                               : cfg.scale_a[k] + cfg.scale_b[k] * normcdfinvf(p);
          if (M_units > 0) mu[((size_t)m_cur * cfg.nM + k) * nb + site_cur] = v;
          else mu[site_cur * cfg.nM + k] = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
}

// =====================================================================================================
// Backward (hand adjoint of g): recompute forward, δ3 from μ̄_ζ through Φ⁻¹ and the logistic, δ2, δ1, and EVERY weight
// gradient of the tile as ONE tensor-core product.
//
// Shared-memory tiles (bf16 hi/lo planes, [column of the tile][64 features] = one 128-byte swizzle row):
//     U = [ h1 (HP) | h2 (HP) | x, 1 (8) ]          V = [ δ2 (HP) | δ3 (8) | δ1 (HP) ]          HP = 24
//   P2  z2    = U[:, h1]·W2          K-major A (first 32 features of U; the weight rows beyond H1 are zero), 3xBF16
//   P3  δ1pre = V[:, δ2]·W2ᵀ         K-major A (first 32 features of V),                                  3xBF16
//   P4  G    += [U_hi | U_lo]ᵀ·V     MN-major operands: the SAME tiles, contraction over the tile's 128 columns.
//       M = 128 = the 64 features of the hi plane and the 64 of the lo plane (descriptor LBO = distance of the
//       planes), so the two MMAs per K step (B = V_hi, V_lo) give all four hi/lo products; the halves are added
//       when the accumulator is read, once per CTA.  G holds ∇W2 = h1ᵀδ2, ∇b2 = 1ᵀδ2, ∇W3 = h2ᵀδ3, ∇W1 = xᵀδ1,
//       ∇b1 = 1ᵀδ1 (and cross blocks nobody reads) and stays in TMEM over all tiles of the CTA.
// Layer 1 runs on the CUDA cores in exact fp32; with pbm covariates the CTA walks the draws of one site tile, so the
// site part of the pre-activation (W1[:nC]ᵀ xM + b1) is computed once per site tile and each draw adds the rank-1 term
// of its covariates (src/elbo.jl:508-515: only ζP[idx] changes between the draws of a site).
// =====================================================================================================
constexpr int TCB_HP = 24;                       // padded hidden width inside the U / V rows (three 16-byte chunks)
constexpr int TCB_NV = 2 * TCB_HP + 8;           // features of V = N of the gradient product (56)
constexpr int TCB_OX = 2 * TCB_HP;               // first feature of [x, 1] in U
constexpr int TCB_OD3 = TCB_HP, TCB_OD1 = TCB_HP + 8;
constexpr int TCB_PLANE = TCS_BM * 128;          // 16 KB
// D = F32, A = B = BF16 (format 1), N >> 3 [17,23), M >> 4 [24,29); bits 15/16: A/B MN-major
constexpr uint32_t TCB_IDESC_K = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t TCB_IDESC_G = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TCB_NV >> 3) << 17) |
                                 ((uint32_t)(128 >> 4) << 24);
constexpr int TCB_TMEM_COLS = 128;               // G (64) | z2 (32) | δ1pre (32)

__device__ __forceinline__ void umma_bf16_(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major SWIZZLE_128B descriptor: atoms of 64 features (128 B) x 8 rows; LBO = distance between 64-feature atoms
// (here: from the hi plane to the lo plane), SBO = distance between 8-row groups (1024 B)
__device__ __forceinline__ uint64_t make_desc_mn_(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void sts128u(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// bf16 hi/lo split of two floats, x0 in the low half
__device__ __forceinline__ void split_bf2(float x0, float x1, uint32_t& h, uint32_t& l) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - h1), "f"(x0 - h0));
}
// eight consecutive features (one 16-byte chunk) of row `row` of a [128][64 bf16] SWIZZLE_128B tile pair
__device__ __forceinline__ void store_chunk_bf(uint32_t hi_tile, uint32_t lo_tile, int row, int chunk, const float (&v)[8]) {
  uint4 h, l;
  split_bf2(v[0], v[1], h.x, l.x); split_bf2(v[2], v[3], h.y, l.y); split_bf2(v[4], v[5], h.z, l.z); split_bf2(v[6], v[7], h.w, l.w);
  const uint32_t off = row * 128 + ((chunk ^ (row & 7)) << 4);
  sts128u(hi_tile + off, h);
  sts128u(lo_tile + off, l);
}
__device__ __forceinline__ void store_elem_bf(uint32_t hi_tile, uint32_t lo_tile, int row, int col, float x) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  const uint32_t off = row * 128 + ((((col >> 3) ^ (row & 7))) << 4) + ((col & 7) << 1);
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(hi_tile + off), "h"(__bfloat16_as_ushort(h)) : "memory");
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(lo_tile + off), "h"(__bfloat16_as_ushort(l)) : "memory");
}

template <int NI, int H1, int H2, int NOUT>
struct MlpTcB {
  static_assert(NI + 1 <= 8 && H1 <= TCB_HP && H2 <= TCB_HP && H1 % 4 == 0 && H2 % 4 == 0 && NOUT <= 2, "shape");
  // shared memory: U hi | U lo | V hi | V lo | W hi | W lo (rows n: [W2[k][n] k<32 | W2ᵀ: W2[n][k] k<32]) | tables | xw | red
  static constexpr int O_U_HI = 0, O_U_LO = TCB_PLANE, O_V_HI = 2 * TCB_PLANE, O_V_LO = 3 * TCB_PLANE, O_W_HI = 4 * TCB_PLANE,
                       O_W_LO = O_W_HI + 32 * 128, O_TAB = O_W_LO + 32 * 128;
  // tables (floats): W1 rows [8][HP] (row NI = b1) | b2 prescaled [HP] | W3 [HP][2]
  static constexpr int T_W1 = 0, T_B2 = 8 * TCB_HP, T_W3 = T_B2 + TCB_HP, TAB_FLOATS = T_W3 + 2 * TCB_HP;
  static constexpr int O_SLOT = O_TAB + ((TAB_FLOATS * 4 + 15) & ~15);  // float slot[4 warps][8]: per-tile covariate cotangents
  static constexpr int O_XW = O_SLOT + 4 * 8 * 4;                       // double xw[nxa]
  __host__ __device__ static constexpr int smem_bytes(int nxa) { return O_XW + nxa * 8 + 1024; }

  __device__ static float wfetch(const Cfg<float>& cfg, const float* __restrict__ phi, int l, int nin, int nout, int i, int j) {
    return (i < nin && j < nout) ? phi[cfg.woff[l] + i * nout + j] : 0.0f;
  }
  __device__ static void build(const Cfg<float>& cfg, const float* __restrict__ phi, uint32_t sm, int tid, int nthr, int nxa) {
    // zero the operand tiles: padding features must be finite (they meet zero weight rows in P2 / P3)
    for (int e = tid; e < (4 * TCB_PLANE) / 16; e += nthr) sts128u(sm + 16 * e, make_uint4(0u, 0u, 0u, 0u));
    for (int e = tid; e < 32 * 64; e += nthr) {
      const int n = e >> 6, k = e & 63;
      const float w = k < 32 ? wfetch(cfg, phi, 1, H1, H2, k, n) : wfetch(cfg, phi, 1, H1, H2, n, k - 32);
      store_elem_bf(sm + O_W_HI, sm + O_W_LO, n, k, w);
    }
    const uint32_t tab = sm + O_TAB;
    for (int e = tid; e < 8 * TCB_HP; e += nthr) {
      const int i = e / TCB_HP, j = e % TCB_HP;
      float w = wfetch(cfg, phi, 0, NI, H1, i, j);
      if (i == NI && j < H1) w = phi[cfg.boff[0] + j];
      sts32(tab + 4 * (T_W1 + e), w);
    }
    for (int e = tid; e < TCB_HP; e += nthr) sts32(tab + 4 * (T_B2 + e), e < H2 ? phi[cfg.boff[1] + e] * K2LOG2E : 0.0f);
    for (int e = tid; e < 2 * TCB_HP; e += nthr) {
      const int j = e >> 1, k = e & 1;
      sts32(tab + 4 * (T_W3 + e), (j < H2 && k < NOUT) ? phi[cfg.woff[2] + j * NOUT + k] : 0.0f);
    }
    for (int e = tid; e < nxa * 2; e += nthr) sts32(sm + O_XW + 4 * e, 0.0f);
  }
};

constexpr int TCB_CTAS_PER_SM = 3;

template <int NI, int H1, int H2, int NOUT>
__global__ void __launch_bounds__(TCS_BM, TCB_CTAS_PER_SM)
k_mlp3_tc_bwd(const __grid_constant__ Cfg<float> cfg, const float* __restrict__ phi, const float* __restrict__ xM,
              const float* __restrict__ zetaP, const float* __restrict__ mubar, int M_units, int64_t nb,
              double* __restrict__ part, double* __restrict__ part_x) {
  using N = MlpTcB<NI, H1, H2, NOUT>;
  constexpr int HP = TCB_HP;
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t bar, barG;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sm = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int Mu = M_units > 0 ? M_units : 1;
  const int nxa = Mu * cfg.npc;
  if (t == 0) {
    mbar_init(&bar, 1);
    mbar_init(&barG, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TCB_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  N::build(cfg, phi, sm, t, TCS_BM, nxa);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tG = tmem, tZ2 = tmem + 64, tD1 = tmem + 96;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
  const uint32_t u_hi = sm + N::O_U_HI, u_lo = sm + N::O_U_LO, v_hi = sm + N::O_V_HI, v_lo = sm + N::O_V_LO;
  const uint32_t w_hi = sm + N::O_W_HI, w_lo = sm + N::O_W_LO, tab = sm + N::O_TAB;

  // this CTA's contiguous range of units u = site_tile·Mu + m (draws of one site tile are consecutive)
  const int64_t ntiles = (nb + TCS_BM - 1) / TCS_BM;
  const int64_t ntot = ntiles * Mu;
  const int64_t u0 = ntot * blockIdx.x / gridDim.x, u1 = ntot * (blockIdx.x + 1) / gridDim.x;
  uint32_t phase = 0, phaseG = 0;
  bool g_pending = false, first = true;
  int64_t st_prev = -1;
  float base[HP];
  float xs[8];
#pragma unroll
  for (int i = 0; i < 8; i++) xs[i] = 0.0f;

  for (int64_t u = u0; u < u1; u++) {
    const int64_t st = u / Mu;
    const int m = (int)(u - st * Mu);
    const int64_t site = st * TCS_BM + t;
    const bool in = site < nb;
    // cotangent of this column's outputs (in flight during E1)
    float mb[2] = {0.0f, 0.0f};
    if (in) {
#pragma unroll
      for (int k = 0; k < NOUT; k++)
        if (k < cfg.nM) mb[k] = M_units > 0 ? mubar[((size_t)m * cfg.nM + k) * nb + site] : mubar[site * cfg.nM + k];
    }
    // ---- E1: layer 1 on the CUDA cores
    if (st != st_prev) {
      st_prev = st;
      const float* row = xM + (in ? site : nb - 1) * cfg.nC;
#pragma unroll
      for (int i = 0; i < 8; i++) xs[i] = (i < NI && i < cfg.nC && in) ? row[i] : (i == NI ? 1.0f : 0.0f);
#pragma unroll
      for (int j = 0; j < HP; j++) base[j] = lds32(tab + 4 * (N::T_W1 + NI * HP + j));
#pragma unroll
      for (int i = 0; i < NI; i++) {
        if (i < cfg.nC) {
#pragma unroll
          for (int j4 = 0; j4 < HP / 4; j4++) {
            const float4 w = lds128(tab + 4 * (N::T_W1 + i * HP + 4 * j4));
            base[4 * j4] = fmaf(xs[i], w.x, base[4 * j4]); base[4 * j4 + 1] = fmaf(xs[i], w.y, base[4 * j4 + 1]);
            base[4 * j4 + 2] = fmaf(xs[i], w.z, base[4 * j4 + 2]); base[4 * j4 + 3] = fmaf(xs[i], w.w, base[4 * j4 + 3]);
          }
        }
      }
    }
    float h1[HP];
#pragma unroll
    for (int j = 0; j < HP; j++) h1[j] = base[j];
#pragma unroll
    for (int i = 0; i < NI; i++) {
      if (i >= cfg.nC) {   // pbm covariates: ζP[idx] of this draw (pass 1: μP[idx])
        const int idx = cfg.pbm_covar_idx[i - cfg.nC];
        const float sv = in ? (M_units > 0 ? zetaP[(size_t)idx * M_units + m] : phi[cfg.o_muP + idx]) : 0.0f;
        xs[i] = sv;
#pragma unroll
        for (int j4 = 0; j4 < HP / 4; j4++) {
          const float4 w = lds128(tab + 4 * (N::T_W1 + i * HP + 4 * j4));
          h1[4 * j4] = fmaf(sv, w.x, h1[4 * j4]); h1[4 * j4 + 1] = fmaf(sv, w.y, h1[4 * j4 + 1]);
          h1[4 * j4 + 2] = fmaf(sv, w.z, h1[4 * j4 + 2]); h1[4 * j4 + 3] = fmaf(sv, w.w, h1[4 * j4 + 3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < HP; j++) h1[j] = j < H1 ? fmaf(-2.0f, rcpa(ex2a(h1[j] * K2LOG2E) + 1.0f), 1.0f) : 0.0f;
    // U is free once the gradient product of the previous tile has read it
    if (g_pending) { mbar_wait(&barG, phaseG); phaseG ^= 1; g_pending = false; }
#pragma unroll
    for (int c = 0; c < HP / 8; c++) {
      const float v[8] = {h1[8 * c], h1[8 * c + 1], h1[8 * c + 2], h1[8 * c + 3], h1[8 * c + 4], h1[8 * c + 5], h1[8 * c + 6], h1[8 * c + 7]};
      store_chunk_bf(u_hi, u_lo, t, c, v);
    }
    store_chunk_bf(u_hi, u_lo, t, TCB_OX / 8, xs);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {   // P2: z2 = h1·W2
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const uint32_t o = s * 32;
        umma_bf16_(tZ2, make_desc(u_lo + o), make_desc(w_hi + o), TCB_IDESC_K, s > 0 ? 1u : 0u);
        umma_bf16_(tZ2, make_desc(u_hi + o), make_desc(w_lo + o), TCB_IDESC_K, 1u);
        umma_bf16_(tZ2, make_desc(u_hi + o), make_desc(w_hi + o), TCB_IDESC_K, 1u);
      }
      umma_commit_(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- E2: h2, outputs, δ3, δ2
    float d2[HP];
    float d3[2] = {0.0f, 0.0f};
    {
      float z[HP];
#pragma unroll
      for (int c = 0; c < HP / 8; c++) {
        float v[8];
        tmem_ld8(tZ2 + lane_sel + 8 * c, v);
#pragma unroll
        for (int q = 0; q < 8; q++) z[8 * c + q] = v[q];
      }
      tmem_wait_ld();
      float h2[HP];
      float z3[2] = {0.0f, 0.0f};
#pragma unroll
      for (int j = 0; j < HP; j++) {
        if (j < H2) {
          h2[j] = fmaf(-2.0f, rcpa(ex2a(fmaf(z[j], K2LOG2E, lds32(tab + 4 * (N::T_B2 + j)))) + 1.0f), 1.0f);
          const float2 w = lds64(tab + 4 * (N::T_W3 + 2 * j));
          z3[0] = fmaf(h2[j], w.x, z3[0]);
          z3[1] = fmaf(h2[j], w.y, z3[1]);
        } else h2[j] = 0.0f;
      }
      // δ3 = μ̄ · dμ/dp · p(1−p)   (NormalScaling: dμ/dp = σ·√(2π)·exp(q²/2), q = Φ⁻¹(p); src/ModelApplicator.jl:168)
#pragma unroll
      for (int k = 0; k < NOUT; k++) {
        if (k < cfg.nM) {
          const float pk = rcpa(1.0f + ex2a(-1.4426950408889634f * z3[k]));
          float d;
          if (cfg.scale_kind == HVI_SCALE_RANGE) d = mb[k] * cfg.scale_b[k];
          else {
            const float q = normcdfinvf(pk);
            d = mb[k] * cfg.scale_b[k] * 2.50662827463100050242f * ex2a(0.72134752044448170368f * q * q);
          }
          d3[k] = in ? d * pk * (1.0f - pk) : 0.0f;
        }
      }
#pragma unroll
      for (int j = 0; j < HP; j++) {
        if (j < H2) {
          const float2 w = lds64(tab + 4 * (N::T_W3 + 2 * j));
          d2[j] = fmaf(d3[0], w.x, d3[1] * w.y) * fmaf(-h2[j], h2[j], 1.0f);
        } else d2[j] = 0.0f;
      }
#pragma unroll
      for (int c = 0; c < HP / 8; c++) {
        const float v[8] = {h2[8 * c], h2[8 * c + 1], h2[8 * c + 2], h2[8 * c + 3], h2[8 * c + 4], h2[8 * c + 5], h2[8 * c + 6], h2[8 * c + 7]};
        store_chunk_bf(u_hi, u_lo, t, HP / 8 + c, v);
        const float w[8] = {d2[8 * c], d2[8 * c + 1], d2[8 * c + 2], d2[8 * c + 3], d2[8 * c + 4], d2[8 * c + 5], d2[8 * c + 6], d2[8 * c + 7]};
        store_chunk_bf(v_hi, v_lo, t, c, w);
      }
      const float v3[8] = {d3[0], d3[1], 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
      store_chunk_bf(v_hi, v_lo, t, TCB_OD3 / 8, v3);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {   // P3: δ1pre = δ2·W2ᵀ  (the transposed weights sit in features 32..63 of the W rows)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const uint32_t o = s * 32;
        umma_bf16_(tD1, make_desc(v_lo + o), make_desc(w_hi + 64 + o), TCB_IDESC_K, s > 0 ? 1u : 0u);
        umma_bf16_(tD1, make_desc(v_hi + o), make_desc(w_lo + 64 + o), TCB_IDESC_K, 1u);
        umma_bf16_(tD1, make_desc(v_hi + o), make_desc(w_hi + 64 + o), TCB_IDESC_K, 1u);
      }
      umma_commit_(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- E3: δ1 = δ1pre ⊙ (1 − h1²), cotangent of the pbm covariates
    {
      float d1[HP];
#pragma unroll
      for (int c = 0; c < HP / 8; c++) {
        float v[8];
        tmem_ld8(tD1 + lane_sel + 8 * c, v);
#pragma unroll
        for (int q = 0; q < 8; q++) d1[8 * c + q] = v[q];
      }
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < HP; j++) d1[j] = j < H1 ? d1[j] * fmaf(-h1[j], h1[j], 1.0f) : 0.0f;
#pragma unroll
      for (int c = 0; c < HP / 8; c++) {
        const float v[8] = {d1[8 * c], d1[8 * c + 1], d1[8 * c + 2], d1[8 * c + 3], d1[8 * c + 4], d1[8 * c + 5], d1[8 * c + 6], d1[8 * c + 7]};
        store_chunk_bf(v_hi, v_lo, t, TCB_OD1 / 8 + c, v);
      }
      if (part_x) {
#pragma unroll
        for (int i = 0; i < NI; i++) {
          if (i >= cfg.nC) {   // x̄[i] = Σ_j W1[i][j] δ1[j], summed over the warp's columns (fixed-order butterfly)
            float sx = 0.0f;
#pragma unroll
            for (int j = 0; j < H1; j++) sx = fmaf(lds32(tab + 4 * (N::T_W1 + i * HP + j)), d1[j], sx);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o);
            if (lane == 0) sts32(sm + N::O_SLOT + 4 * (warp * 8 + (i - cfg.nC)), sx);
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {   // P4: G += [U_hi | U_lo]ᵀ·(V_hi + V_lo), contraction over the 128 columns of the tile
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < TCS_BM / 16; s++) {
        const uint32_t o = s * 2048;
        umma_bf16_(tG, make_desc_mn_(u_hi + o, TCB_PLANE), make_desc_mn_(v_lo + o, TCB_PLANE), TCB_IDESC_G, (first && s == 0) ? 0u : 1u);
        umma_bf16_(tG, make_desc_mn_(u_hi + o, TCB_PLANE), make_desc_mn_(v_hi + o, TCB_PLANE), TCB_IDESC_G, 1u);
      }
      umma_commit_(&barG);
    }
    if (part_x && t == 32) {   // the four warps' covariate cotangents of this tile, fixed order, into the draw's double sum
      for (int c = 0; c < cfg.npc; c++) {
        const float sx = (lds32(sm + N::O_SLOT + 4 * c) + lds32(sm + N::O_SLOT + 4 * (8 + c))) +
                         (lds32(sm + N::O_SLOT + 4 * (16 + c)) + lds32(sm + N::O_SLOT + 4 * (24 + c)));
        const uint32_t a = sm + N::O_XW + 8 * ((M_units > 0 ? m : 0) * cfg.npc + c);
        double acc;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(acc) : "r"(a));
        acc += (double)sx;
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(acc) : "memory");
      }
    }
    first = false;
    g_pending = true;
  }
  if (g_pending) { mbar_wait(&barG, phaseG); phaseG ^= 1; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- read G: lane r < 64 holds the hi-plane row of feature r, lane 64 + r the lo-plane row; add, scatter into ϕg's layout
  float* red = reinterpret_cast<float*>(smem_raw) + ((sm - smem_u32(smem_raw)) >> 2);   // reuse of the U planes: [64][TCB_NV]
  float g[TCB_NV];
  if (u1 > u0) {
#pragma unroll
    for (int c = 0; c < TCB_NV / 8; c++) {
      float v[8];
      tmem_ld8(tG + lane_sel + 8 * c, v);
#pragma unroll
      for (int q = 0; q < 8; q++) g[8 * c + q] = v[q];
    }
    tmem_wait_ld();
  } else {
#pragma unroll
    for (int q = 0; q < TCB_NV; q++) g[q] = 0.0f;
  }
  if (t >= 64) {
#pragma unroll
    for (int q = 0; q < TCB_NV; q++) red[(t - 64) * TCB_NV + q] = g[q];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  double* out = part + (size_t)blockIdx.x * cfg.nphig;
  if (t < 64) {
#pragma unroll
    for (int q = 0; q < TCB_NV; q++) g[q] += red[t * TCB_NV + q];
    if (t < H1) {                       // row h1[t]: ∇W2[t][j]
#pragma unroll
      for (int j = 0; j < H2; j++) out[cfg.woff[1] + t * H2 + j] = (double)g[j];
    } else if (t >= HP && t < HP + H2) {   // row h2[t − HP]: ∇W3[·][k]
#pragma unroll
      for (int k = 0; k < NOUT; k++) out[cfg.woff[2] + (t - HP) * NOUT + k] = (double)g[TCB_OD3 + k];
    } else if (t >= TCB_OX && t < TCB_OX + NI) {   // row x[i]: ∇W1[i][j]
#pragma unroll
      for (int j = 0; j < H1; j++) out[cfg.woff[0] + (t - TCB_OX) * H1 + j] = (double)g[TCB_OD1 + j];
    } else if (t == TCB_OX + NI) {       // row of the constant 1: ∇b1 = 1ᵀδ1, ∇b2 = 1ᵀδ2
#pragma unroll
      for (int j = 0; j < H1; j++) out[cfg.boff[0] + j] = (double)g[TCB_OD1 + j];
#pragma unroll
      for (int j = 0; j < H2; j++) out[cfg.boff[1] + j] = (double)g[j];
    }
  }
  if (part_x) {
    for (int e = t; e < nxa; e += TCS_BM) {
      double v;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sm + N::O_XW + 8 * (uint32_t)e));
      part_x[(size_t)blockIdx.x * nxa + e] = v;
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCB_TMEM_COLS));
}

bool tc_match(const Cfg<float>& c, int ni, int h1, int h2, int no) {
  return c.nL == 3 && c.l_in[0] == ni && c.l_out[0] == h1 && c.l_out[1] == h2 && c.l_out[2] == no &&
         c.l_act[0] == HVI_ACT_TANH && c.l_act[1] == HVI_ACT_TANH && c.l_act[2] == HVI_ACT_LOGISTIC && c.l_bias[0] &&
         c.l_bias[1] && !c.l_bias[2] && c.nM <= no && c.nC <= ni;
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
    using N = MlpTc<NI, H1, H2, NO_>;                                                                       \
    auto kern = k_mlp3_tc_fwd<NI, H1, H2, NO_>;                                                             \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, N::SMEM_BYTES); \
    if (e != cudaSuccess) return e;                                                                         \
    const int64_t ntot = ((nb + TCS_BM - 1) / TCS_BM) * (M_units > 0 ? M_units : 1);                        \
    int64_t blocks = ntot;                                                                                  \
    if (blocks > (int64_t)L.sm_count * TCS_CTAS_PER_SM) blocks = (int64_t)L.sm_count * TCS_CTAS_PER_SM;     \
    if (blocks < 1) blocks = 1;                                                                             \
    kern<<<(int)blocks, TCS_BM, N::SMEM_BYTES, L.stream>>>(cfg, phi, xM, zetaP, M_units, nb, mu);           \
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
    using N = MlpTcB<NI, H1, H2, NO_>;                                                                      \
    const int nxa = (M_units > 0 ? M_units : 1) * cfg.npc;                                                  \
    const int smem = N::smem_bytes(nxa);                                                                    \
    if (smem > 200 * 1024) return cudaErrorNotSupported;                                                    \
    auto kern = k_mlp3_tc_bwd<NI, H1, H2, NO_>;                                                             \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
    if (e != cudaSuccess) return e;                                                                         \
    const int64_t ntot = ((nb + TCS_BM - 1) / TCS_BM) * (M_units > 0 ? M_units : 1);                        \
    int64_t blocks = ntot;                                                                                  \
    const int64_t cap = nblk_max < L.sm_count * TCB_CTAS_PER_SM ? nblk_max : L.sm_count * TCB_CTAS_PER_SM;  \
    if (blocks > cap) blocks = cap;                                                                         \
    if (blocks < 1) blocks = 1;                                                                             \
    kern<<<(int)blocks, TCS_BM, smem, L.stream>>>(cfg, phi, xM, zetaP, mubar, M_units, nb, part, part_x);   \
    HVI_TC_CHECK();                                                                                         \
    *nblk_out = (int)blocks;                                                                                \
    return cudaSuccess;                                                                                     \
  }
  HVI_TC_CASES(X)
#undef X
  return cudaErrorNotSupported;
}

}  // namespace hvi
