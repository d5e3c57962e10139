// umma_small.cu -- cost of small tcgen05.mma instructions on B200 (one CTA, one issuing thread, unrolled issue loop with
// loop-invariant descriptors): cycles from the first issue to the arrival of tcgen05.commit for chains of NMMA instructions,
// and the cycles the issuing thread itself spends in the chain (issue only).
//   F0 ss_k      A, B K-major SWIZZLE_128B tiles in shared memory      M = 128, N = 32,  K = 16 bf16, one accumulator
//   F1 ts_k      A in tensor memory, B K-major in shared memory        M = 128, N = 32
//   F2 ss_mn56   A, B MN-major tiles in shared memory                  M = 128, N = 56
//   F3 ss_mn112  the same with N = 112
//   F4 ts_k x2   F1 alternating between two accumulators
//   F5 ss_mn56x2 F2 alternating between two accumulators
//   F6 ss_k256   M = 128, N = 256 (the GEMM shape: 128 cycles of math per instruction)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_small umma_small.cu && ./umma_small
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k(uint32_t a) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t a, uint32_t lbo) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__host__ __device__ constexpr uint32_t idesc(int n, bool mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (mn ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
constexpr int NFL = 7;

template <int F, int NMMA>
__device__ __forceinline__ void chain(uint32_t tm, uint32_t sm, uint64_t* bar, uint32_t& ph, long long* out) {
  // loop-invariant descriptors (8 K positions)
  const uint64_t ak = desc_k(sm), bk = desc_k(sm + 32768), amn = desc_mn(sm, 16384), bmn = desc_mn(sm + 32768, 16384);
  for (int rep = 0; rep < 3; rep++) {
    const long long c0 = clock64();
#pragma unroll
    for (int i = 0; i < NMMA; i++) {
      const uint32_t o = i & 7;
      if (F == 0) mma_ss(tm, ak + 2 * (o & 3), bk + 2 * (o & 3), idesc(32, false), i > 0);
      if (F == 1) mma_ts(tm, tm + 256 + 8 * (o & 1), bk + 2 * (o & 3), idesc(32, false), i > 0);
      if (F == 2) mma_ss(tm, amn + 128 * o, bmn + 128 * o, idesc(56, true), i > 0);
      if (F == 3) mma_ss(tm, amn + 128 * o, bmn + 128 * o, idesc(112, true), i > 0);
      if (F == 4) mma_ts(tm + 32 * (i & 1), tm + 256 + 8 * (o & 1), bk + 2 * (o & 3), idesc(32, false), i > 1);
      if (F == 5) mma_ss(tm + 64 * (i & 1), amn + 128 * o, bmn + 128 * o, idesc(56, true), i > 1);
      if (F == 6) mma_ss(tm, ak + 2 * (o & 3), bk + 2 * (o & 3), idesc(256, false), i > 0);
    }
    const long long c1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    mbar_wait(bar, ph);
    ph ^= 1;
    const long long c2 = clock64();
    out[(F * 3 + rep) * 2] = c2 - c0;
    out[(F * 3 + rep) * 2 + 1] = c1 - c0;
  }
}

template <int NMMA>
__global__ void k(long long* out) {
  extern __shared__ __align__(1024) unsigned char raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const uint32_t sm = (smem_u32(raw) + 1023u) & ~1023u;
  for (int e = threadIdx.x; e < 100 * 1024 / 4; e += blockDim.x) ((uint32_t*)raw)[e] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tbase;
  if (threadIdx.x == 0) {
    uint32_t ph = 0;
    chain<0, NMMA>(tm, sm, &bar, ph, out);
    chain<1, NMMA>(tm, sm, &bar, ph, out);
    chain<2, NMMA>(tm, sm, &bar, ph, out);
    chain<3, NMMA>(tm, sm, &bar, ph, out);
    chain<4, NMMA>(tm, sm, &bar, ph, out);
    chain<5, NMMA>(tm, sm, &bar, ph, out);
    chain<6, NMMA>(tm, sm, &bar, ph, out);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int NMMA>
int run(long long* d) {
  const char* names[NFL] = {"ss_k   N32  (A,B K-major smem)", "ts_k   N32  (A tmem)", "ss_mn  N56  (A,B MN-major smem)", "ss_mn  N112 (A,B MN-major smem)",
                            "ts_k   N32, two accumulators", "ss_mn  N56, two accumulators", "ss_k   N256 (GEMM shape)"};
  cudaFuncSetAttribute(k<NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 101 * 1024);
  k<NMMA><<<1, 128, 101 * 1024>>>(d);
  long long h[NFL * 6];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  for (int f = 0; f < NFL; f++)
    printf("%-34s chain of %2d: %6lld cycles to commit arrival (best of 3: %6lld), %5lld in the issue loop = %5.1f per MMA\n", names[f], NMMA, h[f * 6],
           h[f * 6 + 4] < h[f * 6 + 2] ? h[f * 6 + 4] : h[f * 6 + 2], h[f * 6 + 5], (double)h[f * 6 + 5] / NMMA);
  return 0;
}

int main() {
  long long* d;
  cudaMalloc(&d, NFL * 6 * sizeof(long long));
  if (run<1>(d) || run<2>(d) || run<6>(d) || run<16>(d) || run<64>(d)) return 1;
  return 0;
}
