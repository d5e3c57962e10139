// Pipe-rate microbenchmark for B200 (sm_100a): warp-instructions per clock per SM of the
// instruction kinds the fused neg-ELBO kernel is made of.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int KIND>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
  float x[ILP];
  float2 v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { x[i] = a + i + threadIdx.x; v[i] = make_float2(a + i, b + threadIdx.x); }
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (KIND == 0) x[i] = fmaf(x[i], a, b);                          // FFMA
      if (KIND == 1) v[i] = __ffma2_rn(v[i], aa, bb);                  // FFMA2
      if (KIND == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));   // MUFU.RCP
      if (KIND == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));   // MUFU.EX2
      if (KIND == 4) x[i] = x[i] + a;                                  // FADD
      if (KIND == 5) v[i] = __fadd2_rn(v[i], aa);                      // FADD2
      if (KIND == 6) v[i] = __fmul2_rn(v[i], aa);                      // FMUL2
      if (KIND == 7) x[i] = fmaxf(x[i], a) + 0.f * b;                  // FMNMX (+FADD folded away?)
      if (KIND == 8) { x[i] = fmaf(x[i], a, b); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[(i + 1) % ILP])); }  // 1 FFMA + 1 MUFU
      if (KIND == 9) { v[i] = __ffma2_rn(v[i], aa, bb); x[i] = fmaf(x[i], a, b); }   // FFMA2 + FFMA mix
    }
    if (KIND == 10) {   // the same mix in runs of 8: all packed, then all scalar
#pragma unroll
      for (int i = 0; i < ILP; i++) v[i] = __ffma2_rn(v[i], aa, bb);
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
    }
    if (KIND == 11) {   // 2 packed : 1 scalar, interleaved (the ratio of the streaming kernel's hot loop is 56 : 35)
#pragma unroll
      for (int i = 0; i < ILP; i++) { v[i] = __ffma2_rn(v[i], aa, bb); if (i & 1) x[i] = fmaf(x[i], a, b); }
    }
    if (KIND == 12) {   // FFMA2 with three distinct register operands each
#pragma unroll
      for (int i = 0; i < ILP; i++) v[i] = __ffma2_rn(v[i], v[(i + 3) % ILP], v[(i + 5) % ILP]);
    }
    if (KIND == 14) {   // scalar FFMA with three distinct register operands
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], x[(i + 3) % ILP], x[(i + 5) % ILP]);
    }
    if (KIND == 15) {   // FADD2 / FMUL2 with two distinct register operands
#pragma unroll
      for (int i = 0; i < ILP; i++) v[i] = __fadd2_rn(v[i], v[(i + 3) % ILP]);
    }
    if (KIND == 16) {
#pragma unroll
      for (int i = 0; i < ILP; i++) v[i] = __fmul2_rn(v[i], v[(i + 3) % ILP]);
    }
    if (KIND == 17) {   // FFMA2 with two distinct operands + one reused
#pragma unroll
      for (int i = 0; i < ILP; i++) v[i] = __ffma2_rn(v[i], v[(i + 3) % ILP], bb);
    }
    if (KIND == 18) {   // two scalar FFMA doing the work of one distinct-operand FFMA2
#pragma unroll
      for (int i = 0; i < ILP; i++) { v[i].x = fmaf(v[i].x, v[(i + 3) % ILP].x, v[(i + 5) % ILP].x); v[i].y = fmaf(v[i].y, v[(i + 3) % ILP].y, v[(i + 5) % ILP].y); }
    }
    if (KIND == 13) {   // FFMA2 + FMNMX (ALU pipe) interleaved
#pragma unroll
      for (int i = 0; i < ILP; i++) { v[i] = __ffma2_rn(v[i], aa, bb); x[i] = fmaxf(x[i], a); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i] + v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, double per_iter) {
  int dev = 0, sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  float* out;
  const int blocks = sms * 8, threads = 256;
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
  printf("%-14s %8.3f ms  %.3f warp-inst/clk/SM (at max clock %d MHz)\n", name, ms, winst / cycles / sms, clk / 1000);
  cudaFree(out);
}

int main() {
  run<0>("FFMA", 1);
  run<1>("FFMA2", 1);
  run<2>("MUFU.RCP", 1);
  run<3>("MUFU.EX2", 1);
  run<4>("FADD", 1);
  run<5>("FADD2", 1);
  run<6>("FMUL2", 1);
  run<7>("FMNMX", 1);
  run<8>("FFMA+RCP", 2);
  run<9>("FFMA2+FFMA", 2);
  run<10>("FFMA2 x8 then FFMA x8", 2);
  run<11>("FFMA2 + FFMA 2:1", 1.5);
  run<12>("FFMA2 distinct operands", 1);
  run<13>("FFMA2 + FMNMX", 2);
  run<14>("FFMA distinct operands", 1);
  run<15>("FADD2 distinct operands", 1);
  run<16>("FMUL2 distinct operands", 1);
  run<17>("FFMA2 2 distinct + 1 reused", 1);
  run<18>("2 x FFMA distinct (= 1 FFMA2)", 2);
  return 0;
}
