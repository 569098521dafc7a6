// fp64_peak.cu — measures the FP64 peaks the FP64-bound rooflines use (SURVEY §8d says "to be measured"):
//   DFMA (vector pipe) and DMMA (mma.sync f64 tensor shapes) throughput on the whole GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu ; prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int CH>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 : 256 FMA per warp instruction
template <int CH>
__global__ void __launch_bounds__(256) dmma884_kernel(double *out, int iters, double a, double b) {
    double c0[CH], c1[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

// m16n8k16 f64: A 16x16 (8 regs), B 16x8 (4 regs), C 16x8 (4 regs): 2048 FMA per warp instruction
template <int CH>
__global__ void __launch_bounds__(256) dmma16816_kernel(double *out, int iters, double a, double b) {
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) dmma1684_kernel(double *out, int iters, double a, double b) {
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, 8));
    const int iters = 20000, blocks = sms * 8, threads = 256;
    const double warps = (double)blocks * threads / 32;
    float t;
    t = time_ms([&] { dfma_kernel<16><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
    const double dfma = 2.0 * blocks * threads * 16.0 * iters / (t * 1e-3) / 1e12;
    t = time_ms([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
    const double d884 = 2.0 * warps * 8 * 256.0 * iters / (t * 1e-3) / 1e12;
    t = time_ms([&] { dmma1684_kernel<4><<<blocks, threads>>>(out, iters, 0.999, 1e-3); });
    const double d1684 = 2.0 * warps * 4 * 512.0 * iters / (t * 1e-3) / 1e12;
    t = time_ms([&] { dmma16816_kernel<4><<<blocks, threads>>>(out, iters / 4, 0.999, 1e-3); });
    const double d16816 = 2.0 * warps * 4 * 2048.0 * (iters / 4) / (t * 1e-3) / 1e12;
    CK(cudaGetLastError());
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"dmma_m16n8k4_tflops\": %.2f, \"dmma_m16n8k16_tflops\": %.2f}\n",
           prop.name, sms, dfma, d884, d1684, d16816);
    return 0;
}
