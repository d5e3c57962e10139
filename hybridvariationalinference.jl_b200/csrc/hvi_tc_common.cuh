// hvi_tc_common.cuh -- PTX helpers shared by the tcgen05 kernels of the wide applicator
// (hvi_dense_tc.cu: register-staged 3xTF32 GEMM; hvi_dense_tma.cu: TMA-fed persistent 3xBF16 GEMM).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hvi_internal.cuh"

namespace hvi {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start address >> 4 [0,14) | LBO (unused for swizzled K-major) = 1 [16,30) | SBO = 1024 B >> 4 [32,46)
// | version = 1 [46,48) | layout type SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// bf16 variant of the split: x ≈ hi + lo with hi = bf16(x), lo = bf16(x − hi) (16 mantissa bits together);
// three products hi·hi + hi·lo + lo·hi at twice the tf32 rate and half the shared-memory bytes
__device__ __forceinline__ void bf16_split(float x, unsigned short& hi, unsigned short& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(l);
}
__device__ __forceinline__ float act_deriv_f(int act, float h) {
  if (act == HVI_ACT_TANH) return 1.f - h * h;
  if (act == HVI_ACT_LOGISTIC) return h * (1.f - h);
  return 1.f;
}


constexpr int64_t WIDE_CHUNK = 262144;   // columns per chunk of the wide applicator (65 536 in round 1: 4x the launches, C5 1.74 s instead of 1.47 s)
constexpr int WIDE_SPLIT = 8;        // row splits of the weight-gradient GEMM
constexpr int SKINNY_ROWS = 128;     // rows per block of the skinny reductions (many blocks: the kernel is latency bound)
constexpr int XS_LD = 16;            // leading dimension of the small input matrix [x ; pbm covariates ; 1]

// G[s][b] = Σ_r S[r·ls + s]·B[r·lb + b] over this block's row slice; one thread per b

template <typename T>
__device__ __forceinline__ T act_apply(int act, T z) {
  if (act == HVI_ACT_TANH) return tanh(z);
  if (act == HVI_ACT_LOGISTIC) return T(1) / (T(1) + exp(-z));
  return z;
}

template <typename T>
__global__ void k_build_xs(const T* __restrict__ xM, int nC, const T* __restrict__ extra, int npc, int64_t site0,
                           int64_t rows, T* __restrict__ Xs) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  for (int i = 0; i < XS_LD; i++) {
    T v = T(0);
    if (i < nC) v = xM[(site0 + r) * nC + i];
    else if (i < nC + npc) v = extra[i - nC];
    else if (i == nC + npc) v = T(1);
    Xs[r * XS_LD + i] = v;
  }
}

// acc[off + s·stride_s + b·stride_b] += Σ_slices part[slice·slice_stride + s·nbig + b]  (fixed order);
// optional copy of the sums; acc == NULL only fills chunk_sum
template <typename PT>
__global__ void k_accum(const PT* __restrict__ part, int nslice, size_t slice_stride, int ns, int nbig,
                        double* __restrict__ acc, int off, int stride_s, int stride_b, double* __restrict__ chunk_sum) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ns * nbig) return;
  const int s = e / nbig, b = e % nbig;
  double t = 0.0;
  for (int sl = 0; sl < nslice; sl++) t += (double)part[(size_t)sl * slice_stride + (size_t)s * nbig + b];
  if (acc) acc[off + s * stride_s + b * stride_b] += t;
  if (chunk_sum) chunk_sum[e] = t;
}

// x̄[m][c] += Σ_j W_0[nC + c][j]·(Σ_r δ_0[r][j])  for this chunk
template <typename T>
__global__ void k_xbar(const T* __restrict__ W0, int nC, int npc, int N, const double* __restrict__ db0_chunk,
                       double* __restrict__ xacc_m) {
  // one warp per covariate, lanes stride over the features, fixed-order butterfly
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (c >= npc) return;
  double s = 0.0;
  for (int j = lane; j < N; j += 32) s += (double)W0[(size_t)(nC + c) * N + j] * db0_chunk[j];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) xacc_m[c] += s;
}


}  // namespace hvi
