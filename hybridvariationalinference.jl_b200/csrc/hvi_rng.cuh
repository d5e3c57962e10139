// hvi_rng.cuh -- the library's counter-based standard-normal generator (Philox4x32-10 + Box-Muller), shared by the
// streaming kernel, hvi_randn and the prediction kernel: draws are keyed by (seed, global column, draw / 4), so results
// do not depend on the sharding.  Stands in for CUDA.randn (ext/HybridVariationalInferenceCUDAExt.jl:82-86).
#pragma once
#include <stdint.h>

namespace hvi {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}


// four standard normals of column gcol (= global site·n_θM + k), draws 4·m4 … 4·m4+3:
// Philox4x32-10 keyed by the seed, Box-Muller on the MUFU pipe
__device__ __forceinline__ void normal4(uint64_t seed, uint64_t gcol, int m4, float (&n4)[4]) {
  uint32_t r[4];
  philox4x32_10((uint32_t)gcol, (uint32_t)(gcol >> 32), (uint32_t)m4, 0x5eedu, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
  for (int h = 0; h < 2; h++) {
    // u1 = (k + 1)·2⁻²⁴ in (0, 1], angle = 2π·k'·2⁻²⁴ − π in [−π, π): one FFMA each after the conversion (the same values
    // as the two-step forms, the scalings are powers of two); sqrt on the MUFU pipe (sqrt.approx: 1 ulp, a draw does not
    // need the IEEE square root's fix-up sequence)
    const float u1 = fmaf((float)(r[2 * h] >> 8), 5.9604644775390625e-08f, 5.9604644775390625e-08f);
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-1.3862943611198906f * __log2f(u1)));   // sqrt(−2 ln u1)
    float sn, cs;
    __sincosf(fmaf((float)(r[2 * h + 1] >> 8), 6.283185307179586f * 5.9604644775390625e-08f, -3.141592653589793f), &sn, &cs);
    n4[2 * h] = rad * cs; n4[2 * h + 1] = rad * sn;
  }
}

}  // namespace hvi
