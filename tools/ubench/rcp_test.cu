// Accuracy of 1/x from rcp.approx.ftz.f64 + (a) two Newton steps, (b) one third-order step.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double rcp_newton2(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ double rcp_cubic(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}
__device__ double rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__global__ void k(double *out, int n) {
    double m0 = 0, m1 = 0, m2 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // x over many binades and mantissas
        const double x = ldexp(1.0 + (double)((i * 2654435761u) & 0xffffff) / 16777216.0 + 1e-9 * i, (i % 41) - 20);
        const double ref = 1.0 / x;
        m0 = fmax(m0, fabs(rcp_seed(x) - ref) / ref);
        m1 = fmax(m1, fabs(rcp_newton2(x) - ref) / ref);
        m2 = fmax(m2, fabs(rcp_cubic(x) - ref) / ref);
    }
    atomicMax((unsigned long long *)out + 0, __double_as_longlong(m0));
    atomicMax((unsigned long long *)out + 1, __double_as_longlong(m1));
    atomicMax((unsigned long long *)out + 2, __double_as_longlong(m2));
}
int main() {
    double *d, h[3] = {0, 0, 0};
    cudaMalloc(&d, 24);
    cudaMemcpy(d, h, 24, cudaMemcpyHostToDevice);
    k<<<296, 256>>>(d, 1 << 26);
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("{\"seed_rel_err\": %.3e, \"newton2_rel_err\": %.3e, \"cubic_rel_err\": %.3e, \"eps\": %.3e}\n", h[0], h[1], h[2], 2.220446e-16);
    return 0;
}
