// Does the FP64 tensor instruction (mma.sync.m8n8k4.f64, SASS DMMA) add throughput next to the FP64 FMA pipe on this part,
// or do both share one unit?  Three kernels: DFMA only, DMMA only, both interleaved; FMA-equivalents per second each.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe.bin tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>   // 0: DFMA, 1: DMMA, 2: both
__global__ void k(double *out, int iters, double a, double b)
{
    double v[8], c[8][2];
    for (int i = 0; i < 8; i++) { v[i] = threadIdx.x + i; c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE != 1) v[i] = fma(v[i], a, b);
            if (MODE != 0) dmma(c[i][0], c[i][1], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += v[i] + c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> static void run(const char *name, int sms)
{
    const int threads = 512, iters = 4000, blocks = sms;
    double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 10, 0.999999, 1e-9); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters, 0.999999, 1e-9); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // per warp-instruction: DFMA = 32 FMA, DMMA m8n8k4 = 8*8*4 = 256 FMA
    const double warps = (double)blocks * threads / 32.0, n = warps * iters * 8.0;
    const double fma = (MODE != 1 ? n * 32.0 : 0.0) + (MODE != 0 ? n * 256.0 : 0.0);
    printf("{\"kernel\": \"%s\", \"ms\": %.3f, \"tfma_per_s\": %.3f, \"dfma_warp_instr_per_clk_per_sm\": %.3f}\n", name, best, fma / best / 1e9,
           (MODE != 1 ? n : 0.0) / (best * 1e-3) / 1.965e9 / sms);
    cudaFree(out);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    run<0>("dfma", p.multiProcessorCount); run<1>("dmma", p.multiProcessorCount); run<2>("dfma+dmma", p.multiProcessorCount);
    return 0;
}
