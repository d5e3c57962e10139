// hvi_dense_tc.cu -- tensor-core dense layer of the wide applicator (BASELINE config 5: 4x512 hidden):
//   C[Mc x N] = act(A[Mc x K] * W[N x K]^T + b),  fp32 in / fp32 out, error-compensated 3xTF32
// on the 5th-generation tensor cores: tcgen05.mma kind::tf32 issued by one thread, fp32 accumulators in
// TMEM, operands in 128-byte-swizzled K-major shared-memory tiles, epilogue via tcgen05.ld.
//
// Every fp32 operand x is split on the way into shared memory into hi = tf32(x) and lo = x - hi
// (exact), and each K step issues hi*hi + hi*lo + lo*hi into the same accumulator: relative error
// ~1e-6 instead of tf32's 5e-4, which is what the 1e-4 gradient tolerance of the path needs
// (SURVEY section 7 "Tolerance vs tensor cores").
//
// Reference: the dense layers of g (ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52,
// ext/HybridVariationalInferenceFluxExt.jl:26-31) -- cuBLAS/SimpleChains GEMMs there.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "hvi_internal.cuh"
#include "hvi_tc_common.cuh"

namespace hvi {

constexpr int TC_BM = 128;      // rows per CTA = TMEM lanes
constexpr int TC_BN = 128;      // output columns per CTA = TMEM columns
constexpr int TC_BK = 32;       // K per stage: 32 tf32 = one 128-byte swizzle row
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 256;     // 8 warps: two per scheduler for the split/store phase; warps 0-3 and 4-7 share the epilogue
constexpr int TC_LD = 1024 / TC_THREADS;   // float4 per thread, operand and stage (= 4)
constexpr int TC_BLOCK = TC_THREADS + 32;   // 8 producer/epilogue warps + 1 warp whose lane 0 issues the MMAs
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;              // 16 KB per operand half
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;             // A hi, A lo, B hi, B lo

// instruction descriptor (InstrDescriptor): D = F32 [4,6)=1, A = TF32 [7,10)=2, B = TF32 [10,13)=2,
// both K-major, N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

// the same with A = B = BF16 (format 1) for kind::f16
constexpr uint32_t TC_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(TC_IDESC_BF16), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate)
      : "memory");
}

// round to nearest tf32 (10-bit mantissa), result as an fp32 bit pattern
// (integer add + mask on the ALU pipe: the cvt.rna.tf32 instruction runs at the 16/clk conversion rate)
__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// split 4 fp32 values into tf32 hi / residual lo and store both as one 16-byte chunk each at the
// swizzled position of (row, chunk) inside a [rows][32 floats] tile
__device__ __forceinline__ void split_store(unsigned char* tile_hi, unsigned char* tile_lo, int row, int chunk, float4 v) {
  float4 hi, lo;
  hi.x = to_tf32(v.x); lo.x = v.x - hi.x;
  hi.y = to_tf32(v.y); lo.y = v.y - hi.y;
  hi.z = to_tf32(v.z); lo.z = v.z - hi.z;
  hi.w = to_tf32(v.w); lo.w = v.w - hi.w;
  const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
  *reinterpret_cast<float4*>(tile_hi + off) = hi;
  *reinterpret_cast<float4*>(tile_lo + off) = lo;
}

// scalar variant of split_store for the transposing loader of the weight-gradient GEMM: element
// (row, col) of a [rows][32] K-major tile
__device__ __forceinline__ void split_store1(unsigned char* tile_hi, unsigned char* tile_lo, int row, int col, float x) {
  const float hi = to_tf32(x);
  const int off = row * 128 + ((((col >> 2) ^ (row & 7))) << 4) + ((col & 3) << 2);
  *reinterpret_cast<float*>(tile_hi + off) = hi;
  *reinterpret_cast<float*>(tile_lo + off) = x - hi;
}

// eight consecutive K values (two float4) → one 16-byte chunk of bf16 hi and one of lo at (row, chunk)
__device__ __forceinline__ void split_store_bf16(unsigned char* tile_hi, unsigned char* tile_lo, int row, int chunk, float4 v0, float4 v1) {
  unsigned short h[8], l[8];
  bf16_split(v0.x, h[0], l[0]); bf16_split(v0.y, h[1], l[1]); bf16_split(v0.z, h[2], l[2]); bf16_split(v0.w, h[3], l[3]);
  bf16_split(v1.x, h[4], l[4]); bf16_split(v1.y, h[5], l[5]); bf16_split(v1.z, h[6], l[6]); bf16_split(v1.w, h[7], l[7]);
  const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
  *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16),
                                                        h[4] | ((uint32_t)h[5] << 16), h[6] | ((uint32_t)h[7] << 16));
  *reinterpret_cast<uint4*>(tile_lo + off) = make_uint4(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16),
                                                        l[4] | ((uint32_t)l[5] << 16), l[6] | ((uint32_t)l[7] << 16));
}
// element (row, col) of a [rows][64 bf16] K-major tile
__device__ __forceinline__ void split_store1_bf16(unsigned char* tile_hi, unsigned char* tile_lo, int row, int col, float x) {
  unsigned short h, l;
  bf16_split(x, h, l);
  const int off = row * 128 + ((((col >> 3) ^ (row & 7))) << 4) + ((col & 7) << 1);
  *reinterpret_cast<unsigned short*>(tile_hi + off) = h;
  *reinterpret_cast<unsigned short*>(tile_lo + off) = l;
}

// MODE 0 (layer / data gradient):  C[Mc x N] = epi(A[Mc x K] · W[N x K]ᵀ),
//     epi(z) = act(z + bias)                        if Hd == NULL   (forward layer)
//     epi(z) = z · act'(Hd[row][col])               otherwise       (δ_{l-1} = (δ_l W_l) ⊙ act'(H_{l-1}))
// MODE 1 (weight gradient, split over the rows): P[split][Md x Nd] = Σ_{r in split} X[r][m]·Y[r][n]
//     with A ≙ X (Mc x Md), W ≙ Y (Mc x Nd), N ≙ Nd, K ≙ Md; operands are transposed on the way into
//     shared memory (lane ↔ row r, so the scalar stores of a warp hit 32 different banks)
// PREC 0: 3xTF32 (K = 32 per stage);  PREC 1: 3xBF16 (K = 64 per stage)
template <int MODE, int PREC>
__global__ void __launch_bounds__(TC_BLOCK, 1)
k_dense_tc(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ C,
           int64_t Mc, int N, int K, int act, const float* __restrict__ Hd, int64_t rows_per_split) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar_full[TC_STAGES];  // "all 8 producer warps have stored stage s" (count 8)
  __shared__ uint64_t bar_mma[TC_STAGES];   // "the MMAs that read stage s have completed" (tcgen05.commit)
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  unsigned char* tiles = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // MODE 0: blockIdx.x = column tile (fastest: the CTAs that share a row tile of A run together and
  // re-read it from L2), blockIdx.y = row tile.  MODE 1: blockIdx.x = m tile, blockIdx.y = n tile.
  const int64_t row0 = (int64_t)(MODE == 0 ? blockIdx.y : blockIdx.x) * TC_BM;
  const int col0 = (MODE == 0 ? blockIdx.x : blockIdx.y) * TC_BN;
  if (t == 0) {
    for (int s = 0; s < TC_STAGES; s++) { mbar_init(&bar_mma[s], 1); mbar_init(&bar_full[s], TC_THREADS / 32); }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TC_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  // MODE 1: this CTA's row range
  int64_t r_begin = 0, r_end = 0;
  if (MODE == 1) {
    r_begin = (int64_t)blockIdx.z * rows_per_split;
    r_end = r_begin + rows_per_split < Mc ? r_begin + rows_per_split : Mc;
    if (r_end < r_begin) r_end = r_begin;
  }
  constexpr int BKV = PREC == 0 ? TC_BK : 2 * TC_BK;       // K values per stage
  constexpr int NV = PREC == 0 ? TC_LD : 2 * TC_LD;        // float4 per thread, operand and stage
  const int KT = MODE == 0 ? K / BKV : (int)((r_end - r_begin + BKV - 1) / BKV);
  // two register sets: the global loads of tile kt+1 are issued before tile kt is split and stored,
  // so a full iteration (stores, barrier, MMA issue) hides their latency
  float4 va[2][NV], vb[2][NV];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_tile = [&](int kt_, float4 (&xa)[NV], float4 (&xb)[NV]) {
    if (MODE == 0) {   // element e = i*TC_THREADS + t → row e/8, 16-byte chunk e%8 of the K-major tile
#pragma unroll
      for (int i = 0; i < TC_LD; i++) {
        const int e = i * TC_THREADS + t, r = e >> 3, c = e & 7;
        const int64_t gr = row0 + r;
        if (PREC == 0) {
          xa[i] = gr < Mc ? __ldg(reinterpret_cast<const float4*>(A + gr * K + (size_t)kt_ * BKV + c * 4)) : zero4;
          xb[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)(col0 + r) * K + (size_t)kt_ * BKV + c * 4));
        } else {       // a chunk holds 8 bf16 = 8 K values = two float4
          const float4* pa = reinterpret_cast<const float4*>(A + gr * K + (size_t)kt_ * BKV + c * 8);
          const float4* pb = reinterpret_cast<const float4*>(W + (size_t)(col0 + r) * K + (size_t)kt_ * BKV + c * 8);
          xa[2 * i] = gr < Mc ? __ldg(pa) : zero4;
          xa[2 * i + 1] = gr < Mc ? __ldg(pa + 1) : zero4;
          xb[2 * i] = __ldg(pb);
          xb[2 * i + 1] = __ldg(pb + 1);
        }
      }
    } else {           // lane ↔ row r of the step, warp w covers m (n) = 16w … 16w+15, four at a time
#pragma unroll
      for (int h = 0; h < (PREC == 0 ? 1 : 2); h++) {
        const int64_t gr = r_begin + (int64_t)kt_ * BKV + h * 32 + lane;
#pragma unroll
        for (int i = 0; i < TC_LD; i++) {
          const int m4 = warp * TC_LD + i;
          xa[h * TC_LD + i] = gr < r_end ? __ldg(reinterpret_cast<const float4*>(A + gr * K + row0 + 4 * m4)) : zero4;
          xb[h * TC_LD + i] = gr < r_end ? __ldg(reinterpret_cast<const float4*>(W + gr * N + col0 + 4 * m4)) : zero4;
        }
      }
    }
  };
  auto store_tile = [&](unsigned char* st, const float4 (&xa)[NV], const float4 (&xb)[NV]) {
    unsigned char *ahi = st, *alo = st + TC_TILE_BYTES, *bhi = st + 2 * TC_TILE_BYTES, *blo = st + 3 * TC_TILE_BYTES;
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < TC_LD; i++) {
        const int e = i * TC_THREADS + t, r = e >> 3, c = e & 7;
        if (PREC == 0) {
          split_store(ahi, alo, r, c, xa[i]);
          split_store(bhi, blo, r, c, xb[i]);
        } else {
          split_store_bf16(ahi, alo, r, c, xa[2 * i], xa[2 * i + 1]);
          split_store_bf16(bhi, blo, r, c, xb[2 * i], xb[2 * i + 1]);
        }
      }
    } else {
#pragma unroll
      for (int h = 0; h < (PREC == 0 ? 1 : 2); h++) {
        const int col = h * 32 + lane;
#pragma unroll
        for (int i = 0; i < TC_LD; i++) {
          const int m = 4 * (warp * TC_LD + i);
          const float4 a4 = xa[h * TC_LD + i], b4 = xb[h * TC_LD + i];
          if (PREC == 0) {
            split_store1(ahi, alo, m + 0, col, a4.x); split_store1(ahi, alo, m + 1, col, a4.y);
            split_store1(ahi, alo, m + 2, col, a4.z); split_store1(ahi, alo, m + 3, col, a4.w);
            split_store1(bhi, blo, m + 0, col, b4.x); split_store1(bhi, blo, m + 1, col, b4.y);
            split_store1(bhi, blo, m + 2, col, b4.z); split_store1(bhi, blo, m + 3, col, b4.w);
          } else {
            split_store1_bf16(ahi, alo, m + 0, col, a4.x); split_store1_bf16(ahi, alo, m + 1, col, a4.y);
            split_store1_bf16(ahi, alo, m + 2, col, a4.z); split_store1_bf16(ahi, alo, m + 3, col, a4.w);
            split_store1_bf16(bhi, blo, m + 0, col, b4.x); split_store1_bf16(bhi, blo, m + 1, col, b4.y);
            split_store1_bf16(bhi, blo, m + 2, col, b4.z); split_store1_bf16(bhi, blo, m + 3, col, b4.w);
          }
        }
      }
    }
  };
  // warp-specialised main loop, no block-wide barrier: warps 0-7 produce stages (global → registers →
  // tf32/bf16 split → swizzled shared memory) and arrive on bar_full; lane 0 of warp 8 waits for a full
  // stage, issues its MMAs and lets tcgen05.commit release the stage (bar_mma) when they have completed
  if (warp < TC_THREADS / 32) {
    if (KT > 0) load_tile(0, va[0], vb[0]);
#pragma unroll 2
    for (int kt = 0; kt < KT; kt++) {
      const int s = kt % TC_STAGES;
      if (kt + 1 < KT) {
        if (kt & 1) load_tile(kt + 1, va[0], vb[0]); else load_tile(kt + 1, va[1], vb[1]);
      }
      // the stage is free once the MMAs issued TC_STAGES iterations ago have completed
      if (kt >= TC_STAGES) mbar_wait(&bar_mma[s], ((kt / TC_STAGES) - 1) & 1);
      unsigned char* st = tiles + (size_t)s * TC_STAGE_BYTES;
      if (kt & 1) store_tile(st, va[1], vb[1]); else store_tile(st, va[0], vb[0]);
      // generic-proxy writes → visible to the tensor core's async proxy, then one arrival per warp
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[s]);
    }
  } else if (lane == 0) {
    for (int kt = 0; kt < KT; kt++) {
      const int s = kt % TC_STAGES;
      mbar_wait(&bar_full[s], (kt / TC_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(tiles + (size_t)s * TC_STAGE_BYTES), a_lo = a_hi + TC_TILE_BYTES,
                     b_hi = a_hi + 2 * TC_TILE_BYTES, b_lo = a_hi + 3 * TC_TILE_BYTES;
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {   // UMMA_K = 8 tf32 or 16 bf16: 32 bytes along the swizzled 128-byte row
        const uint32_t o = ks * 32;
        if (PREC == 0) {
          mma_tf32(tmem_d, make_desc(a_hi + o), make_desc(b_hi + o), (kt > 0 || ks > 0) ? 1u : 0u);
          mma_tf32(tmem_d, make_desc(a_hi + o), make_desc(b_lo + o), 1u);
          mma_tf32(tmem_d, make_desc(a_lo + o), make_desc(b_hi + o), 1u);
        } else {
          mma_bf16(tmem_d, make_desc(a_hi + o), make_desc(b_hi + o), (kt > 0 || ks > 0) ? 1u : 0u);
          mma_bf16(tmem_d, make_desc(a_hi + o), make_desc(b_lo + o), 1u);
          mma_bf16(tmem_d, make_desc(a_lo + o), make_desc(b_hi + o), 1u);
        }
      }
      // arrives on the barrier when all MMAs issued so far have completed (implies before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma[s])) : "memory");
      if (kt == KT - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
    }
  }
  // epilogue: TMEM → registers → global.  Warp w owns TMEM lanes 32w … 32w+31.
  if (KT > 0 && warp < TC_THREADS / 32) mbar_wait(&bar_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w may only touch TMEM lanes 32·(w mod 4) …; warps 0-3 take the first half of the columns, 4-7 the second
  const int wq = warp & 3;
  const int64_t grow = row0 + wq * 32 + lane;
#pragma unroll 1
  for (int cb = (warp >> 2) * (TC_BN / 2); cb < ((warp >> 2) + 1) * (TC_BN / 2) && warp < TC_THREADS / 32; cb += 32) {
    uint32_t v[32];
    if (KT > 0) {
      const uint32_t taddr = tmem_d + ((uint32_t)(wq * 32) << 16) + cb;
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
    } else {
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = 0u;
    }
    if (MODE == 1) {
      float* out = C + (size_t)blockIdx.z * (size_t)K * N + (size_t)grow * N + col0 + cb;   // K ≙ Md
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(out + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                          __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    } else if (grow < Mc) {
      float* out = C + grow * N + col0 + cb;
      const float* hd = Hd ? Hd + grow * N + col0 + cb : nullptr;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o;
        float* po = &o.x;
        float4 h4 = hd ? *reinterpret_cast<const float4*>(hd + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float* ph = &h4.x;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          float z = __uint_as_float(v[j + c]);
          if (hd) z *= act_deriv_f(act, ph[c]);
          else {
            z += bias ? bias[col0 + cb + j + c] : 0.f;
            if (act == HVI_ACT_TANH) z = tanhf(z);
            else if (act == HVI_ACT_LOGISTIC) z = 1.f / (1.f + expf(-z));
          }
          po[c] = z;
        }
        *reinterpret_cast<float4*>(out + j) = o;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TC_BN));
}

// operand precision of the tensor-core GEMMs: HVI_TC_PREC=0 → 3xTF32, 1 → 3xBF16 (default)
static int tc_prec() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HVI_TC_PREC"); v = e ? atoi(e) : 0; }
  return v;
}

// C = epi(A·Wᵀ); device pointers; requires K % 64 == 0 (32 with TF32), N % 128 == 0, 16-byte aligned rows
cudaError_t launch_dense_tc(cudaStream_t stream, const float* A, const float* W, const float* bias, float* C, int64_t Mc,
                            int N, int K, int act, const float* Hd = nullptr) {
  if (K % TC_BK != 0 || N % TC_BN != 0 || Mc < 1) return cudaErrorInvalidValue;
  const bool bf = tc_prec() == 1 && K % (2 * TC_BK) == 0;
  const size_t smem = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024;
  auto kern = bf ? k_dense_tc<0, 1> : k_dense_tc<0, 0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)(N / TC_BN), (unsigned)((Mc + TC_BM - 1) / TC_BM));
  kern<<<grid, TC_BLOCK, smem, stream>>>(A, W, bias, C, Mc, N, K, act, Hd, 0);
  return cudaGetLastError();
}

// weight gradient P[split][Md x Nd] = Σ_rows X[r][:]ᵀ Y[r][:] ; Md, Nd multiples of 128
cudaError_t launch_wgrad_tc(cudaStream_t stream, const float* X, const float* Y, float* P, int64_t rows, int Md, int Nd,
                            int nsplit) {
  if (Md % TC_BM != 0 || Nd % TC_BN != 0 || rows < 1 || nsplit < 1) return cudaErrorInvalidValue;
  const bool bf = tc_prec() == 1;
  const int bkv = bf ? 2 * TC_BK : TC_BK;
  const size_t smem = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024;
  auto kern = bf ? k_dense_tc<1, 1> : k_dense_tc<1, 0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int64_t per = (rows + nsplit - 1) / nsplit;
  per = (per + bkv - 1) / bkv * bkv;
  dim3 grid((unsigned)(Md / TC_BM), (unsigned)(Nd / TC_BN), (unsigned)nsplit);
  kern<<<grid, TC_BLOCK, smem, stream>>>(X, Y, nullptr, P, rows, Nd, Md, 0, nullptr, per);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// wide applicator, forward: g for networks with layers wider than HVI_MAX_WIDTH (BASELINE config 5:
// 5-512-512-512-512-2).  Columns (sites, or draw x site with pbm covariates) are processed in chunks;
// activations of a chunk live in two ping-pong HBM buffers [chunk x width]; hidden layers whose shape
// fits (K % 32 == 0, N % 128 == 0, Float32) run on the tensor cores (k_dense_tc), the first layer
// (K = n_covar) and the last (N = n_θM, fused logistic + Φ⁻¹ scaling) on CUDA cores.
// ---------------------------------------------------------------------------------------
// W_ref: the reference's layout vec(W[out x in]) column-major = [in][out]; Wt: [out][in] row-major
template <typename T>
__global__ void k_transpose_w(const T* __restrict__ Wref, T* __restrict__ Wt, int n_in, int n_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_in * n_out) { const int i = e / n_out, j = e % n_out; Wt[(size_t)j * n_in + i] = Wref[e]; }
}

// generic layer on CUDA cores: C[r][j] = act(Σ_i in(r,i)·W[i][j] + b[j]); one thread per (r, 4 consecutive j)
// first-layer mode (xM != NULL): in(r, i) = xM[site][i] for i < nC, else the pbm covariate value extra[i − nC]
template <typename T>
__global__ void __launch_bounds__(256) k_dense_simple(const T* __restrict__ A, const T* __restrict__ xM, int nC,
                                                      const T* __restrict__ extra, int64_t site0,
                                                      const T* __restrict__ Wref, const T* __restrict__ bias, int act,
                                                      T* __restrict__ C, int64_t rows, int N, int K) {
  const int N4 = (N + 3) / 4;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * N4) return;
  const int64_t r = e / N4;
  const int j0 = (int)(e % N4) * 4;
  T z[4];
#pragma unroll
  for (int c = 0; c < 4; c++) z[c] = (bias && j0 + c < N) ? bias[j0 + c] : T(0);
  for (int i = 0; i < K; i++) {
    const T x = xM ? (i < nC ? xM[(site0 + r) * nC + i] : extra[i - nC]) : A[r * K + i];
#pragma unroll
    for (int c = 0; c < 4; c++)
      if (j0 + c < N) z[c] = fma(x, Wref[(size_t)i * N + j0 + c], z[c]);
  }
#pragma unroll
  for (int c = 0; c < 4; c++)
    if (j0 + c < N) C[r * N + j0 + c] = act_apply<T>(act, z[c]);
}

// last layer + output scaling: one warp per row, lanes over K, outputs n_out (≤ 8)
template <typename T>
__global__ void __launch_bounds__(256) k_wide_last(const Cfg<T> cfg, const T* __restrict__ A, const T* __restrict__ Wref,
                                                   int act, int64_t rows, int K, int n_out, T* __restrict__ mu,
                                                   int64_t site0, int m, int M_units, int64_t nb) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  T acc[MAXP];
  for (int k = 0; k < n_out; k++) acc[k] = T(0);
  for (int i = lane; i < K; i += 32) {
    const T x = A[r * K + i];
    for (int k = 0; k < n_out; k++) acc[k] = fma(x, Wref[(size_t)i * n_out + k], acc[k]);
  }
  for (int k = 0; k < n_out; k++) {
    T s = acc[k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && k < cfg.nM) {
      const T p = act_apply<T>(act, s);
      const T v = (cfg.scale_kind == HVI_SCALE_RANGE) ? p * cfg.scale_b[k] + cfg.scale_a[k]
                                                      : cfg.scale_a[k] + cfg.scale_b[k] * (T)normcdfinv((double)p);
      const int64_t site = site0 + r;
      if (M_units > 0) mu[((size_t)m * cfg.nM + k) * nb + site] = v;
      else mu[site * cfg.nM + k] = v;
    }
  }
}

template <typename T>
cudaError_t launch_mlp_fwd_wide(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* xM, int64_t nb, T* mu,
                                int M_units, const T* zetaP, T* work /*wide_work_elems(cfg)*/, WideCache cache) {
  if (cfg.l_bias[cfg.nL - 1] || cfg.l_out[cfg.nL - 1] > MAXP) return cudaErrorInvalidConfiguration;
  if constexpr (sizeof(T) == 4) {   // TMA-fed 3xBF16 chain (hvi_dense_tma.cu) when every hidden width is a multiple of 128
    if (bf3_eligible(cfg)) return launch_mlp_fwd_bf3(L, cfg, phi, xM, nb, mu, M_units, zetaP, work, cache);
  }
  const int64_t CHUNK = 65536;
  const int maxw = cfg.max_width;
  T* buf0 = work;
  T* buf1 = work + (size_t)CHUNK * maxw;
  T* wt = buf1 + (size_t)CHUNK * maxw;          // transposed hidden-layer weights, [out][in] each
  T* extra_h = wt + (size_t)cfg.nphig;          // pbm covariate values of the current draw (device)
  // transposed copies of the hidden-layer weights for the tensor-core layers
  for (int l = 1; l < cfg.nL - 1; l++) {
    const int n = cfg.l_in[l] * cfg.l_out[l];
    k_transpose_w<T><<<(n + 255) / 256, 256, 0, L.stream>>>(phi + cfg.woff[l], wt + cfg.woff[l], cfg.l_in[l], cfg.l_out[l]);
    if (L.counter) ++*L.counter;
  }
  const int nm = M_units > 0 ? M_units : 1;
  for (int m = 0; m < nm; m++) {
    // pbm covariates of this draw: ζP[idx, m] (unit mode) or μP[idx] (site mode)
    for (int c = 0; c < cfg.npc; c++) {
      const T* src = M_units > 0 ? zetaP + (size_t)cfg.pbm_covar_idx[c] * M_units + m : phi + cfg.o_muP + cfg.pbm_covar_idx[c];
      cudaError_t e = cudaMemcpyAsync(extra_h + c, src, sizeof(T), cudaMemcpyDeviceToDevice, L.stream);
      if (e != cudaSuccess) return e;
    }
    for (int64_t s0 = 0; s0 < nb; s0 += CHUNK) {
      const int64_t rows = nb - s0 < CHUNK ? nb - s0 : CHUNK;
      T *in = buf0, *out = buf1;
      {  // first layer
        const int N = cfg.l_out[0], K = cfg.l_in[0];
        const int64_t n = rows * ((N + 3) / 4);
        k_dense_simple<T><<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(
            nullptr, xM, cfg.nC, extra_h, s0, phi + cfg.woff[0], cfg.l_bias[0] ? phi + cfg.boff[0] : nullptr, cfg.l_act[0],
            in, rows, N, K);
        if (L.counter) ++*L.counter;
      }
      for (int l = 1; l < cfg.nL - 1; l++) {
        const int N = cfg.l_out[l], K = cfg.l_in[l];
        const T* b = cfg.l_bias[l] ? phi + cfg.boff[l] : nullptr;
        bool tc = false;
        if constexpr (sizeof(T) == 4) {
          if (K % TC_BK == 0 && N % TC_BN == 0) {
            cudaError_t e = launch_dense_tc(L.stream, (const float*)in, (const float*)(wt + cfg.woff[l]), (const float*)b,
                                            (float*)out, rows, N, K, cfg.l_act[l]);
            if (e != cudaSuccess) return e;
            tc = true;
          }
        }
        if (!tc) {
          const int64_t n = rows * ((N + 3) / 4);
          k_dense_simple<T><<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(in, nullptr, 0, nullptr, 0, phi + cfg.woff[l], b,
                                                                           cfg.l_act[l], out, rows, N, K);
        }
        if (L.counter) ++*L.counter;
        T* tmp = in; in = out; out = tmp;
      }
      {  // last layer + scaling
        const int l = cfg.nL - 1;
        k_wide_last<T><<<(unsigned)((rows * 32 + 255) / 256), 256, 0, L.stream>>>(cfg, in, phi + cfg.woff[l], cfg.l_act[l], rows,
                                                                              cfg.l_in[l], cfg.l_out[l], mu, s0, m, M_units, nb);
        if (L.counter) ++*L.counter;
      }
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------
// wide applicator, backward: ∇ϕg (and the cotangent of the pbm covariate inputs) from μ̄_ζ.
// Per chunk of columns: forward keeping every activation, δ of the last layer on CUDA cores, then per
// hidden layer  δ_{l-1} = (δ_l W_l) ⊙ act'(H_{l-1})  (k_dense_tc<0>, W_l in the reference's own
// [in][out] layout is already K-major for this product) and  ∇W_l = H_{l-1}ᵀ δ_l  (k_dense_tc<1>, split
// over the rows, fp32 partial tiles summed in a fixed order into a double accumulator).  The first
// layer (K = n_covar), the last (N = n_θM) and all biases use the "skinny" reduction kernels.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_skinny(const T* __restrict__ S, int ls, int ns, const T* __restrict__ B, int lb,
                                                int nbig, int64_t rows, float* __restrict__ part) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * SKINNY_ROWS;
  const int64_t r1 = r0 + SKINNY_ROWS < rows ? r0 + SKINNY_ROWS : rows;
  if (b >= nbig) return;
  float acc[8];
  for (int s = 0; s < 8; s++) acc[s] = 0.f;
  int64_t r = r0;
  for (; r + 4 <= r1; r += 4) {   // four independent loads in flight per thread
    const float y0 = (float)B[r * lb + b], y1 = (float)B[(r + 1) * lb + b], y2 = (float)B[(r + 2) * lb + b],
                y3 = (float)B[(r + 3) * lb + b];
    for (int s = 0; s < ns; s++) {
      acc[s] = fmaf((float)S[r * ls + s], y0, acc[s]);
      acc[s] = fmaf((float)S[(r + 1) * ls + s], y1, acc[s]);
      acc[s] = fmaf((float)S[(r + 2) * ls + s], y2, acc[s]);
      acc[s] = fmaf((float)S[(r + 3) * ls + s], y3, acc[s]);
    }
  }
  for (; r < r1; r++) {
    const float y = (float)B[r * lb + b];
    for (int s = 0; s < ns; s++) acc[s] = fmaf((float)S[r * ls + s], y, acc[s]);
  }
  for (int s = 0; s < ns; s++) part[((size_t)blockIdx.y * ns + s) * nbig + b] = acc[s];
}
template <>
__global__ void __launch_bounds__(256) k_skinny<double>(const double* __restrict__ S, int ls, int ns, const double* __restrict__ B,
                                                        int lb, int nbig, int64_t rows, float* __restrict__ part_f) {
  double* part = reinterpret_cast<double*>(part_f);
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * SKINNY_ROWS;
  const int64_t r1 = r0 + SKINNY_ROWS < rows ? r0 + SKINNY_ROWS : rows;
  if (b >= nbig) return;
  double acc[8];
  for (int s = 0; s < 8; s++) acc[s] = 0.0;
  for (int64_t r = r0; r < r1; r++) {
    const double y = B[r * lb + b];
    for (int s = 0; s < ns; s++) acc[s] = fma(S[r * ls + s], y, acc[s]);
  }
  for (int s = 0; s < ns; s++) part[((size_t)blockIdx.y * ns + s) * nbig + b] = acc[s];
}

// last layer backward: one warp per row.  z_k = Σ_i H[r][i] W[i][k]; p = act(z); δ_L[k] = μ̄·scale'·act'(p);
// writes δ_L (rows x 8) and δ_{L-1}[r][i] = (Σ_k W[i][k] δ_L[k])·act'_{L-1}(H[r][i])
template <typename T>
__global__ void __launch_bounds__(256) k_wide_last_bwd(const Cfg<T> cfg, const T* __restrict__ H, const T* __restrict__ Wref,
                                                       int act, int act_prev, int64_t rows, int K, int n_out,
                                                       const T* __restrict__ mubar, int64_t site0, int m, int M_units,
                                                       int64_t nb, T* __restrict__ dlast, T* __restrict__ Dprev) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  T acc[MAXP], dl[MAXP];
  for (int k = 0; k < n_out; k++) acc[k] = T(0);
  for (int i = lane; i < K; i += 32) {
    const T x = H[r * K + i];
    for (int k = 0; k < n_out; k++) acc[k] = fma(x, Wref[(size_t)i * n_out + k], acc[k]);
  }
  const int64_t site = site0 + r;
  for (int k = 0; k < n_out; k++) {
    T s = acc[k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    T d = T(0);
    if (k < cfg.nM) {
      const T p = act_apply<T>(act, s);
      const T mb = M_units > 0 ? mubar[((size_t)m * cfg.nM + k) * nb + site] : mubar[site * cfg.nM + k];
      if (cfg.scale_kind == HVI_SCALE_RANGE) d = mb * cfg.scale_b[k];
      else { const T q = (T)normcdfinv((double)p); d = mb * cfg.scale_b[k] * T(2.50662827463100050242) * exp(T(0.5) * q * q); }
      d *= (act == HVI_ACT_LOGISTIC) ? p * (T(1) - p) : (act == HVI_ACT_TANH) ? T(1) - p * p : T(1);
    }
    dl[k] = d;
    if (lane == 0) dlast[r * 8 + k] = d;
  }
  if (lane == 0) for (int k = n_out; k < 8; k++) dlast[r * 8 + k] = T(0);
  for (int i = lane; i < K; i += 32) {
    T s = T(0);
    for (int k = 0; k < n_out; k++) s = fma(Wref[(size_t)i * n_out + k], dl[k], s);
    const T h = H[r * K + i];
    const T dv = (act_prev == HVI_ACT_TANH) ? T(1) - h * h : (act_prev == HVI_ACT_LOGISTIC) ? h * (T(1) - h) : T(1);
    Dprev[r * K + i] = s * dv;
  }
}

// CUDA-core fallbacks (Float64, or shapes the tensor-core kernel does not take)
template <typename T>
__global__ void __launch_bounds__(256) k_dgrad_simple(const T* __restrict__ D, const T* __restrict__ Wref, const T* __restrict__ Hprev,
                                                      int act_prev, T* __restrict__ C, int64_t rows, int n_in, int n_out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * n_in) return;
  const int64_t r = e / n_in;
  const int i = (int)(e % n_in);
  T s = T(0);
  for (int j = 0; j < n_out; j++) s = fma(D[r * n_out + j], Wref[(size_t)i * n_out + j], s);
  const T h = Hprev[r * n_in + i];
  const T dv = (act_prev == HVI_ACT_TANH) ? T(1) - h * h : (act_prev == HVI_ACT_LOGISTIC) ? h * (T(1) - h) : T(1);
  C[r * n_in + i] = s * dv;
}
template <typename T>
__global__ void __launch_bounds__(256) k_wgrad_simple(const T* __restrict__ Hprev, const T* __restrict__ D, int64_t rows, int n_in,
                                                      int n_out, double* __restrict__ acc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_in * n_out) return;
  const int i = e / n_out, j = e % n_out;
  double s = 0.0;
  for (int64_t r = 0; r < rows; r++) s += (double)Hprev[r * n_in + i] * (double)D[r * n_out + j];
  acc[e] += s;
}

size_t wide_work_elems(int max_width, int nphig) {
  const size_t a = (size_t)2 * 65536 * max_width + (size_t)nphig + 16, b = (bf3_fwd_work_bytes(max_width, nphig) + 3) / 4;
  return a > b ? a : b;
}
// elements of T (treated as 8 bytes each for the double/float mix) of the backward work space
size_t wide_bwd_work_bytes(int nL, int max_width, int nphig) {
  const size_t cw = (size_t)WIDE_CHUNK * max_width;
  const size_t a = 8 * ((size_t)(nL + 1) * cw + (size_t)nphig + (size_t)WIDE_SPLIT * max_width * max_width +
                        (size_t)WIDE_CHUNK * (8 + XS_LD) + (size_t)(WIDE_CHUNK / SKINNY_ROWS) * 16 * max_width + 4 * (size_t)max_width + 1024);
  const size_t b = bf3_bwd_work_bytes(nL, max_width, nphig);
  return a > b ? a : b;
}

template <typename T>
cudaError_t launch_mlp_bwd_wide(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* xM, int64_t nb, const T* mubar,
                                int M_units, const T* zetaP, void* work, double* grad_acc, double* xacc, WideCache cache) {
  if (cfg.l_bias[cfg.nL - 1] || cfg.l_out[cfg.nL - 1] > 8 || cfg.nC + cfg.npc + 1 > 8) return cudaErrorInvalidConfiguration;
  if constexpr (sizeof(T) == 4) {
    if (bf3_eligible(cfg)) return launch_mlp_bwd_bf3(L, cfg, phi, xM, nb, mubar, M_units, zetaP, work, grad_acc, xacc, cache);
  }
  const int nL = cfg.nL, maxw = cfg.max_width;
  const size_t cw = (size_t)WIDE_CHUNK * maxw;
  // carve the work space
  T* Hbuf = reinterpret_cast<T*>(work);                  // H_0 … H_{nL-2}: nL-1 buffers
  T* D0 = Hbuf + (size_t)(nL - 1) * cw;
  T* D1 = D0 + cw;
  T* wt = D1 + cw;                                       // transposed hidden weights (forward GEMMs)
  T* extra_h = wt + cfg.nphig;
  T* Xs = extra_h + 16;                                  // [chunk][XS_LD]
  T* dlast = Xs + (size_t)WIDE_CHUNK * XS_LD;            // [chunk][8]
  float* P = reinterpret_cast<float*>(dlast + (size_t)WIDE_CHUNK * 8);   // wgrad partial tiles / skinny partials
  double* db0 = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(P) +
                                          8 * ((size_t)WIDE_SPLIT * maxw * maxw + (size_t)(WIDE_CHUNK / SKINNY_ROWS) * 16 * maxw));
  auto count = [&]() { if (L.counter) ++*L.counter; };
  for (int l = 1; l < nL - 1; l++) {
    const int n = cfg.l_in[l] * cfg.l_out[l];
    k_transpose_w<T><<<(n + 255) / 256, 256, 0, L.stream>>>(phi + cfg.woff[l], wt + cfg.woff[l], cfg.l_in[l], cfg.l_out[l]);
    count();
  }
  const int nm = M_units > 0 ? M_units : 1;
  for (int m = 0; m < nm; m++) {
    for (int c = 0; c < cfg.npc; c++) {
      const T* src = M_units > 0 ? zetaP + (size_t)cfg.pbm_covar_idx[c] * M_units + m : phi + cfg.o_muP + cfg.pbm_covar_idx[c];
      cudaError_t e = cudaMemcpyAsync(extra_h + c, src, sizeof(T), cudaMemcpyDeviceToDevice, L.stream);
      if (e != cudaSuccess) return e;
    }
    for (int64_t s0 = 0; s0 < nb; s0 += WIDE_CHUNK) {
      const int64_t rows = nb - s0 < WIDE_CHUNK ? nb - s0 : WIDE_CHUNK;
      const int nslice = (int)((rows + SKINNY_ROWS - 1) / SKINNY_ROWS);
      auto H = [&](int l) { return Hbuf + (size_t)l * cw; };   // output of layer l
      // ---- forward, keeping all activations
      k_build_xs<T><<<(unsigned)((rows + 255) / 256), 256, 0, L.stream>>>(xM, cfg.nC, extra_h, cfg.npc, s0, rows, Xs);
      count();
      {
        const int N = cfg.l_out[0], K = cfg.l_in[0];
        const int64_t n = rows * ((N + 3) / 4);
        k_dense_simple<T><<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(nullptr, xM, cfg.nC, extra_h, s0, phi + cfg.woff[0],
                                                                         cfg.l_bias[0] ? phi + cfg.boff[0] : nullptr,
                                                                         cfg.l_act[0], H(0), rows, N, K);
        count();
      }
      for (int l = 1; l < nL - 1; l++) {
        const int N = cfg.l_out[l], K = cfg.l_in[l];
        const T* b = cfg.l_bias[l] ? phi + cfg.boff[l] : nullptr;
        bool tc = false;
        if constexpr (sizeof(T) == 4) {
          if (K % TC_BK == 0 && N % TC_BN == 0) {
            cudaError_t e = launch_dense_tc(L.stream, (const float*)H(l - 1), (const float*)(wt + cfg.woff[l]), (const float*)b,
                                            (float*)H(l), rows, N, K, cfg.l_act[l]);
            if (e != cudaSuccess) return e;
            tc = true;
          }
        }
        if (!tc) {
          const int64_t n = rows * ((N + 3) / 4);
          k_dense_simple<T><<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(H(l - 1), nullptr, 0, nullptr, 0, phi + cfg.woff[l], b,
                                                                           cfg.l_act[l], H(l), rows, N, K);
        }
        count();
      }
      // ---- last layer: δ_L and δ_{L-1}, ∇W_L
      T* Dcur = D0;   // δ of layer l's output, [rows][l_out[l]]
      T* Dnxt = D1;
      {
        const int l = nL - 1, K = cfg.l_in[l], no = cfg.l_out[l];
        k_wide_last_bwd<T><<<(unsigned)((rows * 32 + 255) / 256), 256, 0, L.stream>>>(
            cfg, H(l - 1), phi + cfg.woff[l], cfg.l_act[l], cfg.l_act[l - 1], rows, K, no, mubar, s0, m, M_units, nb, dlast, Dcur);
        count();
        k_skinny<T><<<dim3((K + 255) / 256, nslice), 256, 0, L.stream>>>(dlast, 8, no, H(l - 1), K, K, rows, P);
        count();
        if constexpr (sizeof(T) == 4)
          k_accum<float><<<(no * K + 255) / 256, 256, 0, L.stream>>>(P, nslice, (size_t)no * K, no, K, grad_acc, cfg.woff[l], 1, no, nullptr);
        else
          k_accum<double><<<(no * K + 255) / 256, 256, 0, L.stream>>>((const double*)P, nslice, (size_t)no * K, no, K, grad_acc, cfg.woff[l], 1, no, nullptr);
        count();
      }
      // ---- hidden layers l = nL-2 … 1: bias, weight gradient, then δ of the layer below
      for (int l = nL - 2; l >= 1; l--) {
        const int N = cfg.l_out[l], K = cfg.l_in[l];   // Dcur: [rows][N]
        if (cfg.l_bias[l]) {
          k_skinny<T><<<dim3((N + 255) / 256, nslice), 256, 0, L.stream>>>(Xs + cfg.nC + cfg.npc, XS_LD, 1, Dcur, N, N, rows, P);
          count();
          if constexpr (sizeof(T) == 4)
            k_accum<float><<<(N + 255) / 256, 256, 0, L.stream>>>(P, nslice, (size_t)N, 1, N, grad_acc, cfg.boff[l], 0, 1, nullptr);
          else
            k_accum<double><<<(N + 255) / 256, 256, 0, L.stream>>>((const double*)P, nslice, (size_t)N, 1, N, grad_acc, cfg.boff[l], 0, 1, nullptr);
          count();
        }
        bool tc = false;
        if constexpr (sizeof(T) == 4) {
          if (K % TC_BM == 0 && N % TC_BN == 0 && K % TC_BK == 0) {
            // ∇W_l[in][out] = H_{l-1}ᵀ δ_l
            cudaError_t e = launch_wgrad_tc(L.stream, (const float*)H(l - 1), (const float*)Dcur, P, rows, K, N, WIDE_SPLIT);
            if (e != cudaSuccess) return e;
            count();
            k_accum<float><<<(K * N + 255) / 256, 256, 0, L.stream>>>(P, WIDE_SPLIT, (size_t)K * N, 1, K * N, grad_acc, cfg.woff[l], 0, 1, nullptr);
            count();
            // δ_{l-1} = (δ_l W_l) ⊙ act'(H_{l-1}):  A = δ_l [rows x N], "W" = W_l as [in][out] = [K x N] row-major
            e = launch_dense_tc(L.stream, (const float*)Dcur, (const float*)(phi + cfg.woff[l]), nullptr, (float*)Dnxt, rows, K, N,
                                cfg.l_act[l - 1], (const float*)H(l - 1));
            if (e != cudaSuccess) return e;
            count();
            tc = true;
          }
        }
        if (!tc) {
          k_wgrad_simple<T><<<(K * N + 255) / 256, 256, 0, L.stream>>>(H(l - 1), Dcur, rows, K, N, grad_acc + cfg.woff[l]);
          count();
          const int64_t n = rows * K;
          k_dgrad_simple<T><<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(Dcur, phi + cfg.woff[l], H(l - 1), cfg.l_act[l - 1], Dnxt,
                                                                           rows, K, N);
          count();
        }
        T* tmp = Dcur; Dcur = Dnxt; Dnxt = tmp;
      }
      // ---- first layer: ∇b_0, ∇W_0 = [x; covariates]ᵀ δ_0, and the cotangent of the covariate inputs
      {
        const int N = cfg.l_out[0], K = cfg.l_in[0];
        const int ns = K + 1;   // inputs + the constant 1
        k_skinny<T><<<dim3((N + 255) / 256, nslice), 256, 0, L.stream>>>(Xs, XS_LD, ns, Dcur, N, N, rows, P);
        count();
        // the skinny partial is [slice][K+1][N]: rows 0..K-1 → W_0[i][j]; row K → b_0[j] (its chunk sum Σ_r δ_0 is kept for x̄)
        const size_t sst = (size_t)ns * N;
        double* acc_b = cfg.l_bias[0] ? grad_acc : nullptr;
        if constexpr (sizeof(T) == 4) {
          k_accum<float><<<(K * N + 255) / 256, 256, 0, L.stream>>>(P, nslice, sst, K, N, grad_acc, cfg.woff[0], N, 1, nullptr);
          k_accum<float><<<(N + 255) / 256, 256, 0, L.stream>>>(P + (size_t)K * N, nslice, sst, 1, N, acc_b, cfg.boff[0], 0, 1, db0);
        } else {
          k_accum<double><<<(K * N + 255) / 256, 256, 0, L.stream>>>((const double*)P, nslice, sst, K, N, grad_acc, cfg.woff[0], N, 1, nullptr);
          k_accum<double><<<(N + 255) / 256, 256, 0, L.stream>>>((const double*)P + (size_t)K * N, nslice, sst, 1, N, acc_b, cfg.boff[0], 0, 1, db0);
        }
        count(); count();
        if (cfg.npc > 0 && xacc) {
          k_xbar<T><<<1, 32 * (cfg.npc > 0 ? cfg.npc : 1), 0, L.stream>>>(phi + cfg.woff[0], cfg.nC, cfg.npc, N, db0, xacc + (size_t)(M_units > 0 ? m : 0) * cfg.npc);
          count();
        }
      }
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

template cudaError_t launch_mlp_bwd_wide<float>(const Launch&, const Cfg<float>&, const float*, const float*, int64_t, const float*,
                                                int, const float*, void*, double*, double*, WideCache);
template cudaError_t launch_mlp_bwd_wide<double>(const Launch&, const Cfg<double>&, const double*, const double*, int64_t,
                                                 const double*, int, const double*, void*, double*, double*, WideCache);

template cudaError_t launch_mlp_fwd_wide<float>(const Launch&, const Cfg<float>&, const float*, const float*, int64_t, float*,
                                                int, const float*, float*, WideCache);
template cudaError_t launch_mlp_fwd_wide<double>(const Launch&, const Cfg<double>&, const double*, const double*, int64_t,
                                                 double*, int, const double*, double*, WideCache);

}  // namespace hvi
