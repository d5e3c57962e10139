// Can the FP64 pipe take part of the observation loop of k_stream off the FP32 pipe?
// The loop body is the one of obs_loop (hvi_kernels.cu): per observation 14 FP32 operations + 1 reciprocal, two
// observations per packed instruction.  MODE 0: all 8 observations packed FP32 (what k_stream does).
// MODE 1: observations 0-5 packed FP32, 6-7 in double (needs F2F conversions of the parameters, of the
// reciprocal and of the four partial sums per draw).  MODE 2: 0-5 packed, 6 scalar FP32, 7 double.
// One block of 256 threads per SM (2 warps per scheduler), as k_stream runs.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o obs_mix obs_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

#define DRAWS 4096

__device__ __forceinline__ float rcpf(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const float* __restrict__ in, float* __restrict__ out, float eps) {
  constexpr int NPK = MODE == 0 ? 4 : 3;   // packed pairs
  float2 S1[4], S2[4], S12[4], nyo[4], w[4];
  const float* p = in + threadIdx.x * 32;
#pragma unroll
  for (int o = 0; o < 4; o++) {
    S1[o] = make_float2(p[2 * o], p[2 * o + 1]); S2[o] = make_float2(p[8 + 2 * o], p[9 + 2 * o]);
    nyo[o] = make_float2(p[16 + 2 * o], p[17 + 2 * o]); w[o] = make_float2(p[24 + 2 * o], p[25 + 2 * o]);
    S12[o] = __fmul2_rn(S1[o], S2[o]);
  }
  // double copies of observations 6, 7
  const double dS1[2] = {S1[3].x, S1[3].y}, dS2[2] = {S2[3].x, S2[3].y}, dS12[2] = {(double)S1[3].x * S2[3].x, (double)S1[3].y * S2[3].y};
  const double dny[2] = {nyo[3].x, nyo[3].y}, dw[2] = {w[3].x, w[3].y};
  float2 nly2 = make_float2(0.f, 0.f);
  double dnly = 0.0;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  float r0 = 0.3f, r1 = 0.5f, k1 = 0.2f, k2 = 2.0f;
#pragma unroll 4
  for (int m = 0; m < DRAWS; m++) {
    r0 += eps; r1 += eps; k1 += eps; k2 += eps;   // new parameters every draw
    const float2 R0 = make_float2(r0, r0), R1 = make_float2(r1, r1), K1 = make_float2(k1, k1), K2 = make_float2(k2, k2);
    float2 g0 = make_float2(0.f, 0.f), g1 = g0, gA = g0, gB = g0;
#pragma unroll
    for (int o = 0; o < NPK; o++) {
      const float2 d1 = __fadd2_rn(K1, S1[o]), d2 = __fadd2_rn(K2, S2[o]);
      const float2 pp = __fmul2_rn(d1, d2);
      const float2 rp = make_float2(rcpf(pp.x), rcpf(pp.y));
      const float2 ab = __fmul2_rn(S12[o], rp);
      const float2 y = __ffma2_rn(R1, ab, R0);
      const float2 res = __fadd2_rn(y, nyo[o]);
      const float2 dy = __fmul2_rn(res, w[o]);
      nly2 = __ffma2_rn(res, dy, nly2);
      g0 = __fadd2_rn(g0, dy);
      const float2 t = __fmul2_rn(dy, ab);
      g1 = __fadd2_rn(g1, t);
      const float2 u = __fmul2_rn(t, rp);
      gA = __ffma2_rn(u, d2, gA);
      gB = __ffma2_rn(u, d1, gB);
    }
    float h0 = g0.x + g0.y, h1 = g1.x + g1.y, hA = gA.x + gA.y, hB = gB.x + gB.y;
    if (MODE == 2) {   // observation 6 in scalar FP32
      const float d1 = k1 + S1[3].x, d2 = k2 + S2[3].x, pp = d1 * d2, rp = rcpf(pp), ab = S12[3].x * rp;
      const float res = fmaf(r1, ab, r0) + nyo[3].x, dy = res * w[3].x;
      nly2.x = fmaf(res, dy, nly2.x);
      h0 += dy;
      const float t = dy * ab; h1 += t;
      const float u = t * rp; hA = fmaf(u, d2, hA); hB = fmaf(u, d1, hB);
    }
    if (MODE >= 1) {   // observations (6,) 7 in double
      const double dr0 = r0, dr1 = r1, dk1 = k1, dk2 = k2;
      double e0 = 0, e1 = 0, eA = 0, eB = 0;
#pragma unroll
      for (int q = (MODE == 2 ? 1 : 0); q < 2; q++) {
        const double d1 = dk1 + dS1[q], d2 = dk2 + dS2[q], pp = d1 * d2;
        const double rp = (double)rcpf((float)pp);
        const double ab = dS12[q] * rp;
        const double res = fma(dr1, ab, dr0) + dny[q], dy = res * dw[q];
        dnly = fma(res, dy, dnly);
        e0 += dy;
        const double t = dy * ab; e1 += t;
        const double u = t * rp; eA = fma(u, d2, eA); eB = fma(u, d1, eB);
      }
      h0 += (float)e0; h1 += (float)e1; hA += (float)eA; hB += (float)eB;
    }
    // stand-in for the rest of the draw: the four sums feed accumulators
    acc0 = fmaf(h0, r0, acc0); acc1 = fmaf(h1, r1, acc1); acc2 = fmaf(-r1 * hA, k1, acc2); acc3 = fmaf(-r1 * hB, k2, acc3);
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc0 + acc1 + acc2 + acc3 + nly2.x + nly2.y + (float)dnly;
}

template <int MODE>
void run(const char* name, const float* in, float* out) {
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, 256>>>(in, out, 1e-6f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<sms, 256>>>(in, out, 1e-6f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * clk * 1e3;
  printf("%-40s %8.3f ms  %.1f clk per warp-draw per scheduler (2 warps each)\n", name, ms, cyc / DRAWS / 2);
}

int main() {
  float *in, *out;
  cudaMalloc(&in, sizeof(float) * 256 * 32);
  cudaMalloc(&out, sizeof(float) * 256 * 256);
  float h[256 * 32];
  for (int i = 0; i < 256 * 32; i++) h[i] = 0.1f + 0.001f * (i % 977);
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  run<0>("8 obs packed FP32", in, out);
  run<1>("6 obs packed FP32 + 2 obs FP64", in, out);
  run<2>("6 packed + 1 scalar FP32 + 1 FP64", in, out);
  return 0;
}
