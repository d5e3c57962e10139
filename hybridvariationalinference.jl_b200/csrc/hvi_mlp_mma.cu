// hvi_mlp_mma.cu -- the small site-batched ML applicator g(xM; ϕg) on the tensor cores (Float32).
//
// g = NormalScaling ∘ MLP(tanh, tanh, logistic) (src/ModelApplicator.jl:163-176,
// ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52) for the 3-layer shapes of DoubleMM
// (5-20-20-2 and, with one pbm covariate, 6-24-24-2).  One warp owns a tile of 16 columns (sites, or
// (draw, site) units in pass 2 of generate_ζ, src/elbo.jl:508-515) and keeps the whole layer chain in
// registers: the two hidden layers are warp-level mma.sync instructions whose operands are split hi + lo
// and multiplied three times (hi·hi + hi·lo + lo·hi).  The accumulator fragment of one layer (row =
// site, columns 2t, 2t+1) is, after tanh, the A fragment of the next layer -- directly for m16n8k16 (bf16),
// through a permutation of the K index for m16n8k8 (tf32) -- so nothing goes through shared memory between
// layers.  The last layer (2 outputs) runs on the CUDA cores in exact fp32.
//
// Forward kernel: 3xTF32 (operands carry ≈21 bits): μ_ζ feeds the streaming kernel, whose σ- and ρ-gradients
// amplify its rounding (a 3xBF16 forward missed the 1e-4 tolerance on ρsM).
// Backward kernel (hand adjoint): recompute forward in 3xBF16 (≈2^-17), δ3 from μ̄_ζ through logistic and Φ⁻¹,
// ∇W3 and δ2 = (δ3 W3ᵀ)⊙(1−h2²) on the CUDA cores, δ1 = (δ2 W2ᵀ)⊙(1−h1²) as MMAs against pre-split transposed
// weight fragments, and the weight gradients ∇W_l = [a_{l-1}; 1]ᵀ δ_l (l = 1, 2) as MMAs whose K dimension is the
// tile's 16 sites: both operands are the packed register fragments transposed in place with movmatrix (no
// shared memory).  The bias gradient rides along as the row of the constant 1 appended to each activation.
// Accumulators stay in fp32 registers over the warp's tiles and are reduced over warps in a fixed order into one
// row of double partial sums per block (same hand-off as k_mlp3_bwd: k_reduce sums the rows).
//
// Measured on B200 (profiles/microbench/mma_pipes.cu): HMMA.16816 / HMMA.1688 issue at 0.457 warp-inst/clk/SM and
// FP32 instructions of the same scheduler do not overlap them, so the cost of a tile is ≈ 8.75 clk x HMMA + the
// other instructions; that is why the thin last layer is cheaper on the CUDA cores.
#include <math.h>
#include <stdlib.h>

#include "hvi_internal.cuh"

namespace hvi {
namespace {

__device__ __forceinline__ float rcpf_(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2f_(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// tanh(x) = 1 − 2/(1 + e^{2x}): absolute error ≈ 1.5e-7 (2 MUFU + 3 FP32); saturates correctly at ±∞
__device__ __forceinline__ float tanh5(float x) {
  const float e = ex2f_(x * 2.885390081777927f);
  return fmaf(-2.0f, rcpf_(e + 1.0f), 1.0f);
}
__device__ __forceinline__ float logisticf_(float z) { return rcpf_(1.0f + ex2f_(-1.4426950408889634f * z)); }

// bf16 hi/lo split of two floats: h = {bf16(x1), bf16(x0)}, l = {bf16(x1 − hi1), bf16(x0 − hi0)}; x0 in the low half
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& h, uint32_t& l) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ uint32_t tr8(uint32_t v) {   // transpose of an 8×8 b16 fragment across the warp
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// c += A·B with A = Ah + Al, B = Bh + Bl (the lo·lo term, ≈2^-18 relative, is dropped)
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                     uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma16816(c, al[0], al[1], al[2], al[3], bh0, bh1);
  mma16816(c, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
  mma16816(c, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}

// tf32 hi/lo split: hi = x rounded to 10 mantissa bits (integer add + mask: finite inputs only; cvt.rna.tf32 costs
// four instructions on sm_100), lo = x − hi exactly, of which the tensor core reads the top 10 bits: x = hi + lo to ≈2^-21
__device__ __forceinline__ void split_tf32(float x, uint32_t& h, uint32_t& l) {
  h = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  l = __float_as_uint(x - __uint_as_float(h));
}
__device__ __forceinline__ void mma1688(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }

template <int NI, int H1, int H2, int NOUT>
struct MlpMma {
  static_assert(NI + 1 <= 8, "inputs + the bias one must fit the first 8 columns of the A fragment");
  static_assert(H1 <= 24 && H2 <= 24 && H1 % 2 == 0 && H2 % 2 == 0, "hidden widths up to 24 (two k-steps incl. the bias one)");
  static_assert(NOUT <= 2, "epilogue redistributes 16 sites × 2 outputs over the 32 lanes");
  // computed n-tiles of each hidden layer and tiles including the column of the constant 1
  static constexpr int NTH1 = cdiv(H1, 8), NTH2 = cdiv(H2, 8), NT1 = H1 / 8 + 1;
  // weight fragment tables, one uint4 {b0 hi, b1 hi, b0 lo, b1 lo} per lane: index (ks·ntiles + nt)
  static constexpr int F_W1 = 0;                    // B[k=i][n=j] = W1[i][j], 1 k-step × NTH1
  static constexpr int F_W2 = F_W1 + NTH1;          // W2, 2 × NTH2
  static constexpr int F_W2T = F_W2 + 2 * NTH2;     // B[k=j][n=i] = W2[i][j], 2 × NTH1
  static constexpr int F_W1T = F_W2T + 2 * NTH1;    // B[k=j][n=i] = W1[i][j], 2 × 1 (cotangent of the inputs)
  static constexpr int NFRAG = F_W1T + 2;
  static constexpr int NB = 8 * (NTH1 + NTH2);      // padded biases b1 | b2
  // gradient accumulators: fp32 C fragments ∇W1 (1×NTH1 tiles) and ∇W2 (2×NTH2), then this lane's partial sums of
  // ∇W3 (rows 8nt+2t+c of W3, both outputs), NRED floats per lane in all
  static constexpr int A_W1 = 0, A_W2 = A_W1 + NTH1, NACCT = A_W2 + 2 * NTH2, NACC3 = NTH2 * 2 * NOUT,
                       NRED = NACCT * 4 + NACC3;

  __device__ static float wfetch(const Cfg<float>& cfg, const float* __restrict__ phi, int l, int nin, int nout, int i, int j) {
    return (i < nin && j < nout) ? phi[cfg.woff[l] + i * nout + j] : 0.0f;
  }

  __device__ static void build_frags(const Cfg<float>& cfg, const float* __restrict__ phi, uint4* __restrict__ wf,
                                     float* __restrict__ sb, int nfrag, int tid, int nthr) {
    for (int e = tid; e < nfrag * 32; e += nthr) {
      const int f = e >> 5, lane = e & 31, g = lane >> 2, t = lane & 3;
      int l, nin, nout, ks, nt;
      bool tr;
      if (f < F_W2) { l = 0; nin = NI; nout = H1; tr = false; ks = 0; nt = f - F_W1; }
      else if (f < F_W2T) { l = 1; nin = H1; nout = H2; tr = false; ks = (f - F_W2) / NTH2; nt = (f - F_W2) % NTH2; }
      else if (f < F_W1T) { l = 1; nin = H1; nout = H2; tr = true; ks = (f - F_W2T) / NTH1; nt = (f - F_W2T) % NTH1; }
      else { l = 0; nin = NI; nout = H1; tr = true; ks = f - F_W1T; nt = 0; }
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int k = 16 * ks + 2 * t + (q & 1) + (q >> 1) * 8, n = 8 * nt + g;
        v[q] = tr ? wfetch(cfg, phi, l, nin, nout, n, k) : wfetch(cfg, phi, l, nin, nout, k, n);
      }
      uint4 w;
      split2(v[0], v[1], w.x, w.z);
      split2(v[2], v[3], w.y, w.w);
      wf[e] = w;
    }
    for (int e = tid; e < NB; e += nthr) {
      const bool first = e < 8 * NTH1;
      const int j = first ? e : e - 8 * NTH1;
      sb[e] = first ? (j < H1 ? phi[cfg.boff[0] + j] : 0.0f) : (j < H2 ? phi[cfg.boff[1] + j] : 0.0f);
    }
  }

  // A fragments (two k-steps) of an activation held as packed tiles p[tile][0: sites g, 1: sites g+8]
  template <int NT>
  __device__ __forceinline__ static void afrag(const uint32_t (&p)[NT][2], int ks, uint32_t (&a)[4]) {
    a[0] = p[2 * ks][0]; a[1] = p[2 * ks][1];
    if (2 * ks + 1 < NT) { a[2] = p[2 * ks + 1][0]; a[3] = p[2 * ks + 1][1]; }
    else { a[2] = 0u; a[3] = 0u; }
  }

  // one dense layer: acc[nt] (+)= A·W, A given as packed hi/lo tiles, NK k-steps, NTN n-tiles
  template <int NTA, int NK, int NTN>
  __device__ __forceinline__ static void dense(const uint4* __restrict__ wf, int lane, const uint32_t (&ph)[NTA][2],
                                               const uint32_t (&pl)[NTA][2], float (&acc)[NTN][4]) {
#pragma unroll
    for (int ks = 0; ks < NK; ks++) {
      if (2 * ks >= NTA) break;
      uint32_t ah[4], al[4];
      afrag<NTA>(ph, ks, ah);
      afrag<NTA>(pl, ks, al);
      uint4 w[NTN];
#pragma unroll
      for (int nt = 0; nt < NTN; nt++) w[nt] = wf[(ks * NTN + nt) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < NTN; nt++) mma16816(acc[nt], al[0], al[1], al[2], al[3], w[nt].x, w[nt].y);
#pragma unroll
      for (int nt = 0; nt < NTN; nt++) mma16816(acc[nt], ah[0], ah[1], ah[2], ah[3], w[nt].z, w[nt].w);
#pragma unroll
      for (int nt = 0; nt < NTN; nt++) mma16816(acc[nt], ah[0], ah[1], ah[2], ah[3], w[nt].x, w[nt].y);
    }
  }

  // tanh of the computed tiles, the constant 1 in column H, pack + split
  template <int H, int NTH, int NT>
  __device__ __forceinline__ static void activate(float (&a)[NTH][4], int t, uint32_t (&ph)[NT][2], uint32_t (&pl)[NT][2]) {
#pragma unroll
    for (int nt = 0; nt < NTH; nt++) {
#pragma unroll
      for (int r = 0; r < 4; r++) a[nt][r] = tanh5(a[nt][r]);
    }
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
      if (nt < NTH) { c0 = a[nt < NTH ? nt : 0][0]; c1 = a[nt < NTH ? nt : 0][1]; c2 = a[nt < NTH ? nt : 0][2]; c3 = a[nt < NTH ? nt : 0][3]; }
      if (nt == H / 8) {   // column H (even) of this tile: lanes t == (H%8)/2, low element
        const bool me = (t == (H % 8) / 2);
        c0 = me ? 1.0f : c0; c2 = me ? 1.0f : c2;
      }
      if (nt < NTH) {
        split2(c0, c1, ph[nt][0], pl[nt][0]);
        split2(c2, c3, ph[nt][1], pl[nt][1]);
      } else {   // tile holding nothing but the 1
        ph[nt][0] = ph[nt][1] = (t == 0) ? 0x00003f80u : 0u;
        pl[nt][0] = pl[nt][1] = 0u;
      }
    }
  }

  // inputs of a tile: this lane's elements x[g][2t], x[g][2t+1], x[g+8][2t], x[g+8][2t+1] incl. the 1 in column NI.
  // kind of the lane's two columns (fixed for the whole kernel): 0 empty, 1 covariate column of xM, 2 pbm covariate, 3 the 1
  struct XCols {
    int kind[2], idx[2];
  };
  __device__ __forceinline__ static XCols x_cols(const Cfg<float>& cfg, int t) {
    XCols c;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int i = 2 * t + q;
      c.kind[q] = i == NI ? 3 : (i > NI ? 0 : (i < cfg.nC ? 1 : 2));
      c.idx[q] = c.kind[q] == 2 ? cfg.pbm_covar_idx[i - cfg.nC] : i;
    }
    return c;
  }
  __device__ __forceinline__ static void load_x(const Cfg<float>& cfg, const XCols& xc, const float* __restrict__ phi,
                                                const float* __restrict__ xM, const float* __restrict__ zetaP,
                                                int M_units, int m, int64_t site0, int64_t nb, int g, float (&xv)[4]) {
    float cov[2];
#pragma unroll
    for (int q = 0; q < 2; q++)
      cov[q] = xc.kind[q] == 2 ? (M_units > 0 ? zetaP[(size_t)xc.idx[q] * M_units + m] : phi[cfg.o_muP + xc.idx[q]])
                               : (xc.kind[q] == 3 ? 1.0f : 0.0f);
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int64_t s = site0 + g + 8 * r;
      const bool in = s < nb;
      const float* row = xM + (in ? s : nb - 1) * cfg.nC;
#pragma unroll
      for (int q = 0; q < 2; q++) xv[2 * r + q] = xc.kind[q] == 1 ? row[xc.idx[q]] : cov[q];
    }
  }
  // columns beyond the batch are zeroed when the tile is consumed (not at the load: the select would wait for it)
  __device__ __forceinline__ static void mask_x(const XCols& xc, int64_t site0, int64_t nb, int g, float (&xv)[4]) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const bool in = site0 + g + 8 * r < nb;
#pragma unroll
      for (int q = 0; q < 2; q++)
        if (!in && xc.kind[q] != 3) xv[2 * r + q] = 0.0f;
    }
  }
  // tile index → (draw, first site); no division in site mode
  __device__ __forceinline__ static void tile_pos(int64_t tile, int64_t ntiles, int M_units, int& m, int64_t& site0) {
    if (M_units > 0) { m = (int)(tile / ntiles); site0 = (tile - (int64_t)m * ntiles) * 16; }
    else { m = 0; site0 = tile * 16; }
  }

  // forward of one tile up to the activations of the second hidden layer (3xBF16; h1 stays packed for ∇W2)
  __device__ __forceinline__ static void forward_h2(const uint4* __restrict__ wf, const float* __restrict__ sb, int lane, int t,
                                                    const uint32_t (&xh)[1][2], const uint32_t (&xl)[1][2], float (&h1)[NTH1][4],
                                                    uint32_t (&h1h)[NT1][2], uint32_t (&h1l)[NT1][2], float (&h2)[NTH2][4]) {
#pragma unroll
    for (int nt = 0; nt < NTH1; nt++) {
      const float2 b = *reinterpret_cast<const float2*>(sb + 8 * nt + 2 * t);
      h1[nt][0] = h1[nt][2] = b.x; h1[nt][1] = h1[nt][3] = b.y;
    }
    dense<1, 1, NTH1>(wf + F_W1 * 32, lane, xh, xl, h1);
    activate<H1, NTH1, NT1>(h1, t, h1h, h1l);
#pragma unroll
    for (int nt = 0; nt < NTH2; nt++) {
      const float2 b = *reinterpret_cast<const float2*>(sb + 8 * NTH1 + 8 * nt + 2 * t);
      h2[nt][0] = h2[nt][2] = b.x; h2[nt][1] = h2[nt][3] = b.y;
    }
    dense<NT1, 2, NTH2>(wf + F_W2 * 32, lane, h1h, h1l, h2);
#pragma unroll
    for (int nt = 0; nt < NTH2; nt++)
#pragma unroll
      for (int r = 0; r < 4; r++) h2[nt][r] = tanh5(h2[nt][r]);
  }

  // The last layer has NOUT <= 2 outputs: an MMA tile would spend 8 columns (and, on this path, issue slots the
  // FP32 pipe cannot use meanwhile) on them, so it runs on the CUDA cores in exact fp32.  This lane's rows of W3:
  // w3[nt][c][k] = W3[8nt + 2t + c][k].
  __device__ __forceinline__ static void load_w3(const Cfg<float>& cfg, const float* __restrict__ phi, int t,
                                                 float (&w3)[NTH2][2][NOUT]) {
#pragma unroll
    for (int nt = 0; nt < NTH2; nt++)
#pragma unroll
      for (int c = 0; c < 2; c++)
#pragma unroll
        for (int k = 0; k < NOUT; k++) w3[nt][c][k] = wfetch(cfg, phi, 2, H2, NOUT, 8 * nt + 2 * t + c, k);
  }
  // z3 of rows g (pz[0]) and g+8 (pz[1]); after the two butterfly steps every lane of a quad holds all of them
  __device__ __forceinline__ static void last_layer(const float (&h2)[NTH2][4], const float (&w3)[NTH2][2][NOUT],
                                                    float (&pz)[2][NOUT]) {
#pragma unroll
    for (int hb = 0; hb < 2; hb++)
#pragma unroll
      for (int k = 0; k < NOUT; k++) {
        float s = 0.0f;
#pragma unroll
        for (int nt = 0; nt < NTH2; nt++)
#pragma unroll
          for (int c = 0; c < 2; c++) s = fmaf(h2[nt][2 * hb + c], w3[nt][c][k], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        pz[hb][k] = s;
      }
  }
  // lane (g, t) of the epilogue owns row g + 8·(t >> 1), output t & 1
  __device__ __forceinline__ static float pick(const float (&pz)[2][NOUT], int t) {
    const float a = (t & 1) && NOUT > 1 ? pz[0][NOUT > 1 ? 1 : 0] : pz[0][0];
    const float b = (t & 1) && NOUT > 1 ? pz[1][NOUT > 1 ? 1 : 0] : pz[1][0];
    return (t >> 1) ? b : a;
  }

  // ---- forward-only variant in 3xTF32 (m16n8k8; operands carry 22 bits): μ_ζ feeds the whole streaming
  // kernel, whose σ- and ρ-gradients amplify its rounding, so the value path gets the tighter split.
  // The A fragment of m16n8k8 wants columns t and t+4 while the accumulator holds 2t and 2t+1: the K index
  // is permuted (logical k = t ↔ column 2t, k = t+4 ↔ column 2t+1) and the weight rows with it.
  static constexpr int G_W1 = 0, G_W2 = G_W1 + NTH1, NFRAG32 = G_W2 + NTH1 * NTH2;

  __device__ static void build_frags32(const Cfg<float>& cfg, const float* __restrict__ phi, uint4* __restrict__ wf,
                                       float* __restrict__ sb, int tid, int nthr) {
    for (int e = tid; e < NFRAG32 * 32; e += nthr) {
      const int f = e >> 5, lane = e & 31, g = lane >> 2, t = lane & 3;
      int l, nin, nout, ks, nt;
      if (f < G_W2) { l = 0; nin = NI; nout = H1; ks = 0; nt = f - G_W1; }
      else { l = 1; nin = H1; nout = H2; ks = (f - G_W2) / NTH2; nt = (f - G_W2) % NTH2; }
      uint4 w;
      split_tf32(wfetch(cfg, phi, l, nin, nout, 8 * ks + 2 * t, 8 * nt + g), w.x, w.z);
      split_tf32(wfetch(cfg, phi, l, nin, nout, 8 * ks + 2 * t + 1, 8 * nt + g), w.y, w.w);
      wf[e] = w;
    }
    for (int e = tid; e < NB; e += nthr) {
      const bool first = e < 8 * NTH1;
      const int j = first ? e : e - 8 * NTH1;
      sb[e] = first ? (j < H1 ? phi[cfg.boff[0] + j] : 0.0f) : (j < H2 ? phi[cfg.boff[1] + j] : 0.0f);
    }
  }

  // acc[nt] += A·W for NK k-steps of 8; the A fragment of k-step ks is the accumulator tile ks of the previous layer
  template <int NK, int NTN>
  __device__ __forceinline__ static void dense32(const uint4* __restrict__ wf, int lane, const float (&a)[NK][4],
                                                 float (&acc)[NTN][4]) {
#pragma unroll
    for (int ks = 0; ks < NK; ks++) {
      uint32_t ah[4], al[4];
      split_tf32(a[ks][0], ah[0], al[0]);   // (row g,   column 2t)
      split_tf32(a[ks][2], ah[1], al[1]);   // (row g+8, column 2t)
      split_tf32(a[ks][1], ah[2], al[2]);   // (row g,   column 2t+1)
      split_tf32(a[ks][3], ah[3], al[3]);   // (row g+8, column 2t+1)
      // consecutive MMAs go to different accumulators (the n-tiles), so the dependent-issue latency of HMMA is not
      // serialised inside a warp
      uint4 w[NTN];
#pragma unroll
      for (int nt = 0; nt < NTN; nt++) w[nt] = wf[(ks * NTN + nt) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < NTN; nt++) mma1688(acc[nt], al, w[nt].x, w[nt].y);
#pragma unroll
      for (int nt = 0; nt < NTN; nt++) mma1688(acc[nt], ah, w[nt].z, w[nt].w);
#pragma unroll
      for (int nt = 0; nt < NTN; nt++) mma1688(acc[nt], ah, w[nt].x, w[nt].y);
    }
  }

  __device__ __forceinline__ static void forward32(const uint4* __restrict__ wf, const float* __restrict__ sb, int lane, int t,
                                                   const float (&xv)[4], float (&h2)[NTH2][4]) {
    float x[1][4] = {{xv[0], xv[1], xv[2], xv[3]}};
    float h1[NTH1][4];
#pragma unroll
    for (int nt = 0; nt < NTH1; nt++) {
      const float2 b = *reinterpret_cast<const float2*>(sb + 8 * nt + 2 * t);
      h1[nt][0] = h1[nt][2] = b.x; h1[nt][1] = h1[nt][3] = b.y;
    }
    dense32<1, NTH1>(wf + G_W1 * 32, lane, x, h1);
#pragma unroll
    for (int nt = 0; nt < NTH1; nt++)
#pragma unroll
      for (int r = 0; r < 4; r++) h1[nt][r] = tanh5(h1[nt][r]);
#pragma unroll
    for (int nt = 0; nt < NTH2; nt++) {
      const float2 b = *reinterpret_cast<const float2*>(sb + 8 * NTH1 + 8 * nt + 2 * t);
      h2[nt][0] = h2[nt][2] = b.x; h2[nt][1] = h2[nt][3] = b.y;
    }
    dense32<NTH1, NTH2>(wf + G_W2 * 32, lane, h1, h2);
#pragma unroll
    for (int nt = 0; nt < NTH2; nt++)
#pragma unroll
      for (int r = 0; r < 4; r++) h2[nt][r] = tanh5(h2[nt][r]);
  }
};

constexpr int MMA_BS = 256, MMA_NW = MMA_BS / 32;      // forward: 3 blocks per SM
constexpr int MMA_BSB = 128, MMA_NWB = MMA_BSB / 32;   // backward (168 registers): 3 blocks per SM

template <int NI, int H1, int H2, int NOUT>
__global__ void __launch_bounds__(MMA_BS, 4) k_mlp3_mma_fwd(const __grid_constant__ Cfg<float> cfg, const float* __restrict__ phi,
                                                             const float* __restrict__ xM, const float* __restrict__ zetaP,
                                                             int M_units, int64_t nb, float* __restrict__ mu) {
  using N = MlpMma<NI, H1, H2, NOUT>;
  __shared__ __align__(16) uint4 wf[N::NFRAG32 * 32];
  __shared__ __align__(16) float sb[N::NB];
  N::build_frags32(cfg, phi, wf, sb, threadIdx.x, MMA_BS);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t ntiles = (nb + 15) / 16;
  const int64_t ntot = ntiles * (M_units > 0 ? M_units : 1);
  const int64_t stride = (int64_t)gridDim.x * MMA_NW;
  int64_t tile = (int64_t)blockIdx.x * MMA_NW + warp;
  const typename N::XCols xcols = N::x_cols(cfg, t);
  float w3[N::NTH2][2][NOUT];
  N::load_w3(cfg, phi, t, w3);
  float xv[4];
  int m = 0, m_next = 0;
  int64_t site0 = 0, site0_next = 0;
  if (tile < ntot) {
    N::tile_pos(tile, ntiles, M_units, m_next, site0_next);
    N::load_x(cfg, xcols, phi, xM, zetaP, M_units, m_next, site0_next, nb, g, xv);
  }
  for (; tile < ntot; tile += stride) {
    m = m_next; site0 = site0_next;
    float xc[4] = {xv[0], xv[1], xv[2], xv[3]};
    N::mask_x(xcols, site0, nb, g, xc);
    if (tile + stride < ntot) {   // next tile's inputs are in flight during this tile's arithmetic
      N::tile_pos(tile + stride, ntiles, M_units, m_next, site0_next);
      N::load_x(cfg, xcols, phi, xM, zetaP, M_units, m_next, site0_next, nb, g, xv);
    }
    float h2[N::NTH2][4], pz[2][NOUT];
    N::forward32(wf, sb, lane, t, xc, h2);
    N::last_layer(h2, w3, pz);
    const float z = N::pick(pz, t);
    const int k = t & 1;
    const int64_t site = site0 + g + 8 * (t >> 1);
    if (k < NOUT && k < cfg.nM && site < nb) {
      const float p = logisticf_(z);
      const float v = (cfg.scale_kind == HVI_SCALE_RANGE) ? p * cfg.scale_b[k] + cfg.scale_a[k]
                                                           : cfg.scale_a[k] + cfg.scale_b[k] * normcdfinvf(p);
      if (M_units > 0) mu[((size_t)m * cfg.nM + k) * nb + site] = v;
      else mu[site * cfg.nM + k] = v;
    }
  }
}

template <int NI, int H1, int H2, int NOUT>
__global__ void __launch_bounds__(MMA_BSB, 3) k_mlp3_mma_bwd(const __grid_constant__ Cfg<float> cfg, const float* __restrict__ phi,
                                                             const float* __restrict__ xM, const float* __restrict__ zetaP,
                                                             const float* __restrict__ mubar, int M_units, int64_t nb,
                                                             double* __restrict__ part, double* __restrict__ part_x) {
  using N = MlpMma<NI, H1, H2, NOUT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* wf = reinterpret_cast<uint4*>(smem_raw);                                   // [NFRAG][32]
  float* sb = reinterpret_cast<float*>(wf + N::NFRAG * 32);                         // [NB]
  float* red = sb + ((N::NB + 3) & ~3);                                             // [MMA_NWB][NRED][32]
  double* xw = reinterpret_cast<double*>(red + (size_t)MMA_NWB * N::NRED * 32);      // [MMA_NWB][nxa]
  const int nxa = (M_units > 0 ? M_units : 1) * cfg.npc;
  N::build_frags(cfg, phi, wf, sb, N::NFRAG, threadIdx.x, MMA_BSB);
  for (int e = threadIdx.x; e < MMA_NWB * nxa; e += MMA_BSB) xw[e] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t ntiles = (nb + 15) / 16;
  const int64_t ntot = ntiles * (M_units > 0 ? M_units : 1);
  const int64_t stride = (int64_t)gridDim.x * MMA_NWB;
  float acc[N::NACCT][4];
#pragma unroll
  for (int a = 0; a < N::NACCT; a++) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.0f;
  const bool has_x = cfg.npc > 0;

  int64_t tile = (int64_t)blockIdx.x * MMA_NWB + warp;
  const typename N::XCols xcols = N::x_cols(cfg, t);
  float xv[4];
  int m = 0, m_next = 0;
  int64_t site0 = 0, site0_next = 0;
  float w3[N::NTH2][2][NOUT], acc3[N::NTH2][2][NOUT];
  N::load_w3(cfg, phi, t, w3);
#pragma unroll
  for (int nt = 0; nt < N::NTH2; nt++)
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
      for (int k = 0; k < NOUT; k++) acc3[nt][c][k] = 0.0f;
  // cotangent of this lane's epilogue element (row g + 8·(t >> 1) of the tile, output t & 1), clamped address
  const int ke = t & 1, re = g + 8 * (t >> 1);
  const bool k_ok = ke < NOUT && ke < cfg.nM;
  auto load_mb = [&](int mm, int64_t s0) -> float {
    int64_t se = s0 + re;
    if (se >= nb) se = nb - 1;
    const int kk = k_ok ? ke : 0;
    return M_units > 0 ? mubar[((size_t)mm * cfg.nM + kk) * nb + se] : mubar[se * cfg.nM + kk];
  };
  float mb_next = 0.0f;
  if (tile < ntot) {
    N::tile_pos(tile, ntiles, M_units, m_next, site0_next);
    N::load_x(cfg, xcols, phi, xM, zetaP, M_units, m_next, site0_next, nb, g, xv);
    mb_next = load_mb(m_next, site0_next);
  }
  for (; tile < ntot; tile += stride) {
    m = m_next; site0 = site0_next;
    const int64_t site_e = site0 + re;
    const bool act_e = k_ok && site_e < nb;
    const float mb = mb_next;

    uint32_t xh[1][2], xl[1][2];
    {
      float xc[4] = {xv[0], xv[1], xv[2], xv[3]};
      N::mask_x(xcols, site0, nb, g, xc);
      split2(xc[0], xc[1], xh[0][0], xl[0][0]);
      split2(xc[2], xc[3], xh[0][1], xl[0][1]);
    }
    if (tile + stride < ntot) {
      N::tile_pos(tile + stride, ntiles, M_units, m_next, site0_next);
      N::load_x(cfg, xcols, phi, xM, zetaP, M_units, m_next, site0_next, nb, g, xv);
      mb_next = load_mb(m_next, site0_next);
    }
    float h1[N::NTH1][4], h2[N::NTH2][4], pz[2][NOUT];
    uint32_t h1h[N::NT1][2], h1l[N::NT1][2];
    N::forward_h2(wf, sb, lane, t, xh, xl, h1, h1h, h1l, h2);
    N::last_layer(h2, w3, pz);

    // δ3 = μ̄ · dμ/dp · p(1−p)   (NormalScaling: dμ/dp = σ·√(2π)·exp(q²/2), q = Φ⁻¹(p); src/ModelApplicator.jl:168)
    float d = 0.0f;
    if (act_e) {
      const float pk = logisticf_(N::pick(pz, t));
      if (cfg.scale_kind == HVI_SCALE_RANGE) d = mb * cfg.scale_b[ke];
      else {
        const float q = normcdfinvf(pk);
        d = mb * cfg.scale_b[ke] * 2.50662827463100050242f * ex2f_(0.72134752044448170368f * q * q);
      }
      d *= pk * (1.0f - pk);
    }
    // every lane of the quad needs δ3 of rows g and g+8, both outputs: dq[hb][k] sits in lane 2·hb + k of the quad
    float dq[2][NOUT];
#pragma unroll
    for (int hb = 0; hb < 2; hb++)
#pragma unroll
      for (int k = 0; k < NOUT; k++) dq[hb][k] = __shfl_sync(0xffffffffu, d, (lane & ~3) + 2 * hb + k);
    // ∇W3 += h2ᵀ δ3 (this lane's rows of W3) and δ2 = (δ3 W3ᵀ) ⊙ (1 − h2²), both on the CUDA cores
    uint32_t d2h[N::NTH2][2], d2l[N::NTH2][2];
#pragma unroll
    for (int nt = 0; nt < N::NTH2; nt++) {
      float p[4];
#pragma unroll
      for (int c = 0; c < 2; c++) {
#pragma unroll
        for (int k = 0; k < NOUT; k++)
          acc3[nt][c][k] = fmaf(h2[nt][c], dq[0][k], fmaf(h2[nt][2 + c], dq[1][k], acc3[nt][c][k]));
#pragma unroll
        for (int hb = 0; hb < 2; hb++) {
          float v = 0.0f;
#pragma unroll
          for (int k = 0; k < NOUT; k++) v = fmaf(dq[hb][k], w3[nt][c][k], v);
          p[2 * hb + c] = v * fmaf(-h2[nt][2 * hb + c], h2[nt][2 * hb + c], 1.0f);
        }
      }
      split2(p[0], p[1], d2h[nt][0], d2l[nt][0]);
      split2(p[2], p[3], d2h[nt][1], d2l[nt][1]);
    }
    {  // ∇W2 += [h1; 1]ᵀ δ2
      uint32_t bh[N::NTH2][2], bl[N::NTH2][2];
#pragma unroll
      for (int nt = 0; nt < N::NTH2; nt++) {
        bh[nt][0] = tr8(d2h[nt][0]); bh[nt][1] = tr8(d2h[nt][1]); bl[nt][0] = tr8(d2l[nt][0]); bl[nt][1] = tr8(d2l[nt][1]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        if (2 * mt >= N::NT1) break;
        uint32_t ah[4], al[4];
        ah[0] = tr8(h1h[2 * mt][0]); ah[2] = tr8(h1h[2 * mt][1]); al[0] = tr8(h1l[2 * mt][0]); al[2] = tr8(h1l[2 * mt][1]);
        if (2 * mt + 1 < N::NT1) {
          ah[1] = tr8(h1h[2 * mt + 1][0]); ah[3] = tr8(h1h[2 * mt + 1][1]);
          al[1] = tr8(h1l[2 * mt + 1][0]); al[3] = tr8(h1l[2 * mt + 1][1]);
        } else { ah[1] = ah[3] = al[1] = al[3] = 0u; }
#pragma unroll
        for (int nt = 0; nt < N::NTH2; nt++) mma3(acc[N::A_W2 + mt * N::NTH2 + nt], ah, al, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
      }
    }
    // δ1 = (δ2 W2ᵀ) ⊙ (1 − h1²)
    uint32_t d1h[N::NTH1][2], d1l[N::NTH1][2];
    {
      float p[N::NTH1][4];
#pragma unroll
      for (int nt = 0; nt < N::NTH1; nt++) p[nt][0] = p[nt][1] = p[nt][2] = p[nt][3] = 0.0f;
      N::template dense<N::NTH2, 2, N::NTH1>(wf + N::F_W2T * 32, lane, d2h, d2l, p);
#pragma unroll
      for (int nt = 0; nt < N::NTH1; nt++) {
#pragma unroll
        for (int r = 0; r < 4; r++) p[nt][r] *= fmaf(-h1[nt][r], h1[nt][r], 1.0f);
        split2(p[nt][0], p[nt][1], d1h[nt][0], d1l[nt][0]);
        split2(p[nt][2], p[nt][3], d1h[nt][1], d1l[nt][1]);
      }
    }
    {  // ∇W1 += [x; 1]ᵀ δ1   (rows 8..15 of the A tile are empty)
      uint32_t ah[4], al[4];
      ah[0] = tr8(xh[0][0]); ah[2] = tr8(xh[0][1]); al[0] = tr8(xl[0][0]); al[2] = tr8(xl[0][1]);
      ah[1] = ah[3] = al[1] = al[3] = 0u;
#pragma unroll
      for (int nt = 0; nt < N::NTH1; nt++)
        mma3(acc[N::A_W1 + nt], ah, al, tr8(d1h[nt][0]), tr8(d1h[nt][1]), tr8(d1l[nt][0]), tr8(d1l[nt][1]));
    }
    if (has_x) {
      // cotangent of the pbm covariate inputs x̄[i] = Σ_j W1[i][j] δ1[j], summed over the tile's sites
      // (shuffle tree, fixed order) and added to this warp's own double accumulator of draw m
      float xb[1][4];
      xb[0][0] = xb[0][1] = xb[0][2] = xb[0][3] = 0.0f;
      N::template dense<N::NTH1, 2, 1>(wf + N::F_W1T * 32, lane, d1h, d1l, xb);
      float s0 = xb[0][0] + xb[0][2], s1 = xb[0][1] + xb[0][3];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
      if (g == 0) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int c = 2 * t + q - cfg.nC;
          if (c >= 0 && c < cfg.npc) xw[(size_t)warp * nxa + (M_units > 0 ? m : 0) * cfg.npc + c] += (double)(q ? s1 : s0);
        }
      }
    }
  }

  // block reduction over the warps (fixed order), scatter into the layout of ϕg
#pragma unroll
  for (int a = 0; a < N::NACCT; a++)
#pragma unroll
    for (int r = 0; r < 4; r++) red[((size_t)warp * N::NRED + a * 4 + r) * 32 + lane] = acc[a][r];
#pragma unroll
  for (int nt = 0; nt < N::NTH2; nt++)
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
      for (int k = 0; k < NOUT; k++)
        red[((size_t)warp * N::NRED + N::NACCT * 4 + (nt * 2 + c) * NOUT + k) * 32 + lane] = acc3[nt][c][k];
  __syncthreads();
  double* out = part + (size_t)blockIdx.x * cfg.nphig;
  for (int e = threadIdx.x; e < N::NACCT * 4 * 32; e += MMA_BSB) {
    const int ln = e & 31, r = (e >> 5) & 3, a = e >> 7, gg = ln >> 2, tt = ln & 3;
    double s = 0.0;
    for (int w = 0; w < MMA_NWB; w++) s += (double)red[((size_t)w * N::NRED + a * 4 + r) * 32 + ln];
    int l, nin, nout, mt, nt;
    if (a < N::A_W2) { l = 0; nin = NI; nout = H1; mt = 0; nt = a - N::A_W1; }
    else { l = 1; nin = H1; nout = H2; mt = (a - N::A_W2) / N::NTH2; nt = (a - N::A_W2) % N::NTH2; }
    const int i = 16 * mt + gg + (r >> 1) * 8, j = 8 * nt + 2 * tt + (r & 1);
    if (j < nout) {
      if (i < nin) out[cfg.woff[l] + i * nout + j] = s;
      else if (i == nin && cfg.l_bias[l]) out[cfg.boff[l] + j] = s;
    }
  }
  // ∇W3[i][k]: the lanes with the same t hold partial sums over their rows (8 values of g per warp)
  for (int e = threadIdx.x; e < H2 * NOUT; e += MMA_BSB) {
    const int i = e / NOUT, k = e % NOUT, nt = i >> 3, tt = (i & 7) >> 1, c = i & 1;
    double s = 0.0;
    for (int w = 0; w < MMA_NWB; w++)
      for (int gg = 0; gg < 8; gg++)
        s += (double)red[((size_t)w * N::NRED + N::NACCT * 4 + (nt * 2 + c) * NOUT + k) * 32 + 4 * gg + tt];
    out[cfg.woff[2] + i * NOUT + k] = s;
  }
  if (part_x)
    for (int e = threadIdx.x; e < nxa; e += MMA_BSB) {
      double s = 0.0;
      for (int w = 0; w < MMA_NWB; w++) s += xw[(size_t)w * nxa + e];
      part_x[(size_t)blockIdx.x * nxa + e] = s;
    }
}

bool mma_match(const Cfg<float>& c, int ni, int h1, int h2, int no) {
  return c.nL == 3 && c.l_in[0] == ni && c.l_out[0] == h1 && c.l_out[1] == h2 && c.l_out[2] == no &&
         c.l_act[0] == HVI_ACT_TANH && c.l_act[1] == HVI_ACT_TANH && c.l_act[2] == HVI_ACT_LOGISTIC && c.l_bias[0] &&
         c.l_bias[1] && !c.l_bias[2] && c.nM <= no;
}

bool mma_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HVI_MLP_MMA"); v = e ? atoi(e) : 1; }
  return v != 0;
}

#define HVI_MMA_CASES(X) X(5, 20, 20, 2) X(6, 24, 24, 2)

}  // namespace

#define HVI_MMA_CHECK()                            \
  do {                                             \
    cudaError_t e_ = cudaGetLastError();           \
    if (e_ != cudaSuccess) return e_;              \
    if (L.counter) ++*L.counter;                   \
  } while (0)

// returns cudaErrorNotSupported when the shape is not one of the tensor-core cases (caller falls back to the
// CUDA-core kernels of the same layout)
cudaError_t launch_mlp3_mma_fwd(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM, int64_t nb,
                                float* mu, int M_units, const float* zetaP) {
  if (!mma_enabled()) return cudaErrorNotSupported;
#define X(NI, H1, H2, NO_)                                                                                  \
  if (mma_match(cfg, NI, H1, H2, NO_)) {                                                                    \
    const int64_t ntot = ((nb + 15) / 16) * (M_units > 0 ? M_units : 1);                                    \
    int64_t blocks = (ntot + MMA_NW - 1) / MMA_NW;                                                          \
    if (blocks > (int64_t)L.sm_count * 4) blocks = (int64_t)L.sm_count * 4;                                 \
    k_mlp3_mma_fwd<NI, H1, H2, NO_><<<(int)blocks, MMA_BS, 0, L.stream>>>(cfg, phi, xM, zetaP, M_units, nb, mu); \
    HVI_MMA_CHECK();                                                                                        \
    return cudaSuccess;                                                                                     \
  }
  HVI_MMA_CASES(X)
#undef X
  return cudaErrorNotSupported;
}

cudaError_t launch_mlp3_mma_bwd(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM,
                                const float* mubar, int64_t nb, double* part, int* nblk_out, int nblk_max, int M_units,
                                const float* zetaP, double* part_x) {
  if (!mma_enabled()) return cudaErrorNotSupported;
#define X(NI, H1, H2, NO_)                                                                                  \
  if (mma_match(cfg, NI, H1, H2, NO_)) {                                                                    \
    using N = MlpMma<NI, H1, H2, NO_>;                                                                      \
    const size_t nxa = (size_t)(M_units > 0 ? M_units : 1) * cfg.npc;                                       \
    const size_t smem = sizeof(uint4) * N::NFRAG * 32 + sizeof(float) * ((N::NB + 3) & ~3) +                \
                        sizeof(float) * (size_t)MMA_NWB * N::NRED * 32 + sizeof(double) * MMA_NWB * nxa + 16;    \
    if (smem > 200 * 1024) return cudaErrorNotSupported;                                                    \
    auto kern = k_mlp3_mma_bwd<NI, H1, H2, NO_>;                                                            \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    if (e != cudaSuccess) return e;                                                                         \
    const int64_t ntot = ((nb + 15) / 16) * (M_units > 0 ? M_units : 1);                                    \
    int64_t blocks = (ntot + MMA_NWB - 1) / MMA_NWB;                                                          \
    const int64_t cap = nblk_max < L.sm_count * 3 ? nblk_max : L.sm_count * 3;                              \
    if (blocks > cap) blocks = cap;                                                                         \
    if (blocks < 1) blocks = 1;                                                                             \
    kern<<<(int)blocks, MMA_BSB, smem, L.stream>>>(cfg, phi, xM, zetaP, mubar, M_units, nb, part, part_x);   \
    HVI_MMA_CHECK();                                                                                        \
    *nblk_out = (int)blocks;                                                                                \
    return cudaSuccess;                                                                                     \
  }
  HVI_MMA_CASES(X)
#undef X
  return cudaErrorNotSupported;
}

}  // namespace hvi
