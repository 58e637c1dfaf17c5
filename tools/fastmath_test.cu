// Accuracy of vb_rcp_pos / vb_rsqrt_pos against the IEEE results, on the GPU:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/fmt tools/fastmath_test.cu && /tmp/fmt
#include <cstdio>
#include <cmath>
#include <cstdint>
#include "../vboc_b200/csrc/fast_math.h"
__global__ void k(int n, double *maxrel) {
    unsigned long long s = 88172645463325252ull + threadIdx.x + blockIdx.x * 1315423911ull;
    double m1 = 0, m2 = 0;
    for (int i = 0; i < n; ++i) {
        s ^= s << 13, s ^= s >> 7, s ^= s << 17;
        double mant = 1.0 + (double)(s >> 11) * (1.0 / 9007199254740992.0);
        int ex = (int)((s >> 3) % 121) - 60;  // 2^-60 .. 2^60
        double x = ldexp(mant, ex);
        double r0 = 1.0 / x, r1 = vb_rcp_pos(x);
        double q0 = 1.0 / sqrt(x), q1 = vb_rsqrt_pos(x);
        m1 = fmax(m1, fabs(r1 - r0) / r0), m2 = fmax(m2, fabs(q1 - q0) / q0);
    }
    atomicMax((unsigned long long *)maxrel, __double_as_longlong(m1));
    atomicMax((unsigned long long *)maxrel + 1, __double_as_longlong(m2));
}
int main() {
    double *d, h[2] = {0, 0};
    cudaMalloc(&d, 16), cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
    k<<<296, 256>>>(2000, d);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("max relative error vs IEEE: rcp %.3e (%.2f ulp)  rsqrt %.3e (%.2f ulp)\n", h[0], h[0] / 1.11e-16, h[1], h[1] / 1.11e-16);
    return 0;
}
