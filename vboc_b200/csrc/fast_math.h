// fast_math.h -- FP64 reciprocal and reciprocal square root without the special-case handling of the CUDA
// library routines: the hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~20 good bits) and two Newton steps.
// Valid for positive normal arguments (the solver divides by slacks t >= 1e-16 and takes rsqrt of pivots
// it has already tested for > 0); relative error <= 2 ulp (tools/fastmath_test.cu measures it on the GPU).
// The IEEE-exact sequences they replace are ~25-30 instructions each with a slow-path branch; in the solve
// kernel they sit in the constraint pass (per constraint) and in the Riccati factorisation (per stage).
#pragma once
#if defined(__CUDACC__) && !defined(VBOC_EMU)
#include <math.h>
__host__ __device__ __forceinline__ double vb_rcp_pos(double x) {
#if !defined(__CUDA_ARCH__)
    return 1.0 / x;
#else
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
#endif
}
__host__ __device__ __forceinline__ double vb_rsqrt_pos(double x) {
#if !defined(__CUDA_ARCH__)
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);  // 1 - x y^2
    y = fma(0.5 * y, e, y);
    e = fma(-x * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    return y;
#endif
}
#else
#include <cmath>
inline double vb_rcp_pos(double x) { return 1.0 / x; }
inline double vb_rsqrt_pos(double x) { return 1.0 / std::sqrt(x); }
#endif
