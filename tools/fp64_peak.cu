// Measures the FP64 FMA issue rate of the device (the roofline denominator SURVEY.md 8d asks for:
// it is not in MEASURED_PEAKS.json).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak.bin tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, int iters, double a, double b)
{
    double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    for (int i = 0; i < iters; i++) {
        v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
        v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int threads = 512, iters = 20000;
    for (int per_sm = 1; per_sm <= 4; per_sm *= 2) {
        const int blocks = p.multiProcessorCount * per_sm;
        double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<blocks, threads>>>(out, 100, 0.999999, 1e-9);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0); k<<<blocks, threads>>>(out, iters, 0.999999, 1e-9); cudaEventRecord(e1);
            cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double fma = (double)blocks * threads * iters * 8.0;
        printf("{\"sms\": %d, \"ctas_per_sm\": %d, \"threads\": %d, \"ms\": %.3f, \"fp64_tfma_per_s\": %.3f, \"fp64_tflops\": %.3f}\n",
               p.multiProcessorCount, per_sm, threads, best, fma / best / 1e9, 2.0 * fma / best / 1e9);
        cudaFree(out);
    }
    return 0;
}
