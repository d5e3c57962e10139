// hvi_dev.cuh -- small device helpers shared by the kernel translation units
#pragma once
#include <math.h>
#include "hvi_internal.cuh"

// after every kernel launch of a launcher whose Launch argument is called L
#define HVI_LAUNCH_CHECK()                         \
  do {                                             \
    cudaError_t e_ = cudaGetLastError();           \
    if (e_ != cudaSuccess) return e_;              \
    if (L.counter) ++*L.counter;                   \
  } while (0)

namespace hvi {

// ---------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ double rcp_fast(double x) { return 1.0 / x; }
// ex2.approx of x·log2(e): relative error ≈ 2^-22 plus |x|·6e-8 from the argument rounding
__device__ __forceinline__ float exp_t(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}
__device__ __forceinline__ double exp_t(double x) { return exp(x); }
__device__ __forceinline__ float log_t(float x) { return logf(x); }
__device__ __forceinline__ double log_t(double x) { return log(x); }
__device__ __forceinline__ float tanh_t(float x) { return tanhf(x); }
__device__ __forceinline__ double tanh_t(double x) { return tanh(x); }
__device__ __forceinline__ float ndtri_t(float p) { return normcdfinvf(p); }
__device__ __forceinline__ double ndtri_t(double p) { return normcdfinv(p); }
__device__ __forceinline__ float log1p_t(float x) { return log1pf(x); }
__device__ __forceinline__ double log1p_t(double x) { return log1p(x); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// two-wide values: float2 maps to the packed FP32 instructions of sm_100 (FADD2/FMUL2/FFMA2 --
// same FLOP rate as the scalar forms on B200, half the issue slots), double2 is plain pairs
template <typename T> struct PairT;
template <> struct PairT<float> { using type = float2; };
template <> struct PairT<double> { using type = double2; };
__device__ __forceinline__ float2 mk2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ double2 mk2(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ double2 padd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 pmul(double2 a, double2 b) { return make_double2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ double2 pfma(double2 a, double2 b, double2 c) { return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }

template <typename T>
__device__ __forceinline__ T act_f(int act, T z) {
  if (act == HVI_ACT_TANH) return tanh_t(z);
  if (act == HVI_ACT_LOGISTIC) return T(1) / (T(1) + exp_t(-z));
  return z;
}
template <typename T>
__device__ __forceinline__ T act_d(int act, T a) {
  if (act == HVI_ACT_TANH) return T(1) - a * a;
  if (act == HVI_ACT_LOGISTIC) return a * (T(1) - a);
  return T(1);
}

// θ = T(ζ) with the reference's clamp (src/elbo.jl:783-808; bijectors_utils.jl:10-23,38-52) fused
// with the prior on θ (src/elbo.jl:203-260).  Outputs: θ, dθ/dζ (0 when clamped), the variable part
// of −logpdf(prior, θ) and its derivative w.r.t. ζ, the log-Jacobian term and its derivative.
template <typename T>
__device__ __forceinline__ void transform_prior(int tr, int pk, T a, T ib, T lo, T hi, T zeta, T& th, T& dth, T& pv,
                                                T& dz, T& lj) {
  T raw, d, dlj, dpz;
  if (tr == HVI_TR_EXP) {
    raw = exp_t(zeta); d = raw; lj = zeta; dlj = T(1);
  } else if (tr == HVI_TR_LOGISTIC) {
    raw = T(1) / (T(1) + exp_t(-zeta)); d = raw * (T(1) - raw);
    lj = -(log1p_t(exp_t(-zeta)) + log1p_t(exp_t(zeta))); dlj = T(1) - T(2) * raw;
  } else {
    raw = zeta; d = T(1); lj = T(0); dlj = T(0);
  }
  th = fmin(hi, raw);
  th = fmax(lo, th);
  const bool uncl = (th == raw);
  dth = uncl ? d : T(0);
  if (pk == HVI_PR_LOGNORMAL) {
    const T lt = (tr == HVI_TR_EXP && uncl) ? zeta : log_t(th);
    const T u = (lt - a) * ib;
    pv = lt + T(0.5) * u * u;
    // d/dθ = (1 + u·ib)/θ ; times dθ/dζ
    if (tr == HVI_TR_EXP) { dz = uncl ? u * ib : T(-1); return; }   // (1 + u·ib) − 1 when unclamped
    dpz = (T(1) + u * ib) / th * dth;
  } else if (pk == HVI_PR_NORMAL) {
    const T u = (th - a) * ib;
    pv = T(0.5) * u * u;
    dpz = u * ib * dth;
  } else {
    pv = T(0); dpz = T(0);
  }
  dz = dpz - dlj;
}
}  // namespace hvi
