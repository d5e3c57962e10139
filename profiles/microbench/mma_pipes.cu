// Rates of the warp-level (legacy) tensor path and of the helper instructions a register-resident
// small-MLP chain needs, on B200 (sm_100a): warp-instructions per clock per SM.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_pipes mma_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
  float c[ILP][4];
  float x[ILP];
  double d[ILP];
  uint32_t u[ILP];
  uint32_t A[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
  uint32_t B[2] = {threadIdx.x * 11u, threadIdx.x * 13u};
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    for (int j = 0; j < 4; j++) c[i][j] = a * i + j;
    x[i] = a + i + threadIdx.x; d[i] = a + i; u[i] = threadIdx.x + i;
  }
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  float2 v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) v[i] = make_float2(a + i, b + threadIdx.x);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (KIND == 0) mma_bf16(c[i], A, B);
      if (KIND == 1) mma_tf32(c[i], A, B);
      if (KIND == 2) mma_f16(c[i], A, B);
      if (KIND == 3) { mma_bf16(c[i], A, B); x[i] = fmaf(x[i], a, b); }                       // HMMA + FFMA
      if (KIND == 4) { mma_bf16(c[i], A, B); x[i] = fmaf(x[i], a, b); v[i] = __ffma2_rn(v[i], aa, bb); }   // HMMA + FFMA + FFMA2
      if (KIND == 5) d[i] = fma(d[i], (double)a, (double)b);                                  // DFMA
      if (KIND == 6) { d[i] = fma(d[i], (double)a, (double)b); v[i] = __ffma2_rn(v[i], aa, bb); }   // DFMA + FFMA2
      if (KIND == 7) asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %0;" : "+r"(u[i]));
      if (KIND == 8) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(x[(i + 1) % ILP]));
      if (KIND == 9) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      if (KIND == 10) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(x[i])); x[i] += a; }   // F2F + FADD
      if (KIND == 11) { mma_bf16(c[i], A, B); mma_bf16(c[i], A, B); mma_bf16(c[i], A, B);            // 3 HMMA + 10 FFMA
#pragma unroll
                        for (int r = 0; r < 10; r++) x[i] = fmaf(x[i], a, b); }
    }
    if (KIND == 8) {
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] += __uint_as_float(u[i] & 0x3f800000u);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i] + c[i][0] + c[i][1] + c[i][2] + c[i][3] + (float)d[i] + (float)u[i] + v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, double per_iter) {
  int dev = 0, sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  float* out;
  const int blocks = sms * 4, threads = 256;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<blocks, threads>>>(out, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<KIND><<<blocks, threads>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double winst = (double)blocks * (threads / 32) * ITERS * ILP * per_iter;
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-28s %8.3f ms  %.3f warp-inst/clk/SM (counted inst per iteration: %.0f; max clock %d MHz)\n", name, ms,
         winst / cycles / sms, per_iter, clk / 1000);
  cudaFree(out);
}

int main() {
  run<0>("HMMA.16816 bf16", 1);
  run<1>("HMMA.1688 tf32", 1);
  run<2>("HMMA.16816 f16", 1);
  run<3>("HMMA bf16 + FFMA", 2);
  run<4>("HMMA bf16 + FFMA + FFMA2", 3);
  run<5>("DFMA", 1);
  run<6>("DFMA + FFMA2", 2);
  run<7>("MOVMATRIX", 1);
  run<8>("F2FP.BF16 pack (+1/8 LOP,FADD)", 1);
  run<9>("MUFU.TANH", 1);
  run<10>("F2F.F64.F32 + FADD", 2);
  run<11>("3 HMMA + 10 FFMA", 13);
  return 0;
}
