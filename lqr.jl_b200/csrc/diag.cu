// diag.cu — diagnostics behind the C ABI: the FP64 peaks the FP64-bound rooflines are quoted against, measured on
// the device the handle owns (DFMA on the vector pipe, DMMA = mma.sync.m8n8k4.f64 on the tensor pipe; there is no
// tcgen05 kind for FP64).  bench.py runs these inside its own clock-sampled region so that the denominators of the
// FP64 fractions carry the same clock record as the kernels they are compared with.
#include "common.cuh"

template <int CH>
__global__ void __launch_bounds__(256) diag_dfma_kernel(double *out, int iters, double a, double b) {
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

// one m8n8k4 instruction = 8 x 8 x 4 = 256 FMAs per warp
template <int CH>
__global__ void __launch_bounds__(256) diag_dmma_kernel(double *out, int iters, double a, double b) {
    double c0[CH], c1[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        c0[i] = threadIdx.x * 1e-3;
        c1[i] = i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

extern "C" int32_t lqrb_fp64_peak_f64(lqrb_handle_t h, int32_t kind, double seconds, double *tflops) {
    if (!h) return -1;
    if (kind != 0 && kind != 1) return lqrb_fail(h, -2, "kind must be 0 (DFMA) or 1 (DMMA m8n8k4)");
    if (!(seconds > 0.0) || seconds > 10.0) return lqrb_fail(h, -3, "seconds out of range (0, 10]");
    if (!tflops) return -4;
    LQRB_CUDA(h, cudaSetDevice(h->device));
    double *out = (double *)lqrb_scratch(h, SCR_MISC, 64);
    if (!out) return 1000 + (int)cudaErrorMemoryAllocation;
    const int blocks = h->sm_count * 8, threads = 256, iters = 20000;
    constexpr int CH_F = 16, CH_M = 8;
    cudaStream_t s = h->stream;
    auto launch = [&] {
        if (kind == 0)
            diag_dfma_kernel<CH_F><<<blocks, threads, 0, s>>>(out, iters, 0.999, 1e-3);
        else
            diag_dmma_kernel<CH_M><<<blocks, threads, 0, s>>>(out, iters, 0.999, 1e-3);
        h->launches++;
    };
    const double flops = kind == 0 ? 2.0 * blocks * threads * CH_F * (double)iters
                                   : 2.0 * (blocks * threads / 32.0) * CH_M * 256.0 * (double)iters;
    cudaEvent_t e0, e1;
    LQRB_CUDA(h, cudaEventCreate(&e0));
    LQRB_CUDA(h, cudaEventCreate(&e1));
    launch();
    launch();
    LQRB_CUDA(h, cudaStreamSynchronize(s));
    double best = 0.0, spent = 0.0;
    int reps = 0;
    while (spent < seconds * 1e3 && reps < 1000) {
        cudaEventRecord(e0, s);
        launch();
        cudaEventRecord(e1, s);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            return lqrb_cuda_fail(h, e, "fp64 peak kernel");
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        spent += ms;
        ++reps;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    h->kernel_name = kind == 0 ? "diag_dfma" : "diag_dmma_m8n8k4";
    *tflops = best;
    return 0;
}
