// FP64 dependent-issue latency and throughput vs. (warps per scheduler, independent chains per warp) on one SM-full grid.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_lat.bin tools/fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b, long long *cyc)
{
    double v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) v[j] = threadIdx.x + j;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) v[j] = fma(v[j], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int threads, int sms)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * sms * threads); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<ILP><<<sms, threads>>>(out, 16, 0.999999, 1e-9, cyc);
    k<ILP><<<sms, threads>>>(out, iters, 0.999999, 1e-9, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_iter = (double)h / iters;
    printf("{\"warps_per_scheduler\": %d, \"ilp\": %d, \"cycles_per_dependent_step\": %.2f, \"dfma_per_cycle_per_sm\": %.1f}\n",
           threads / 128, ILP, per_iter, (double)threads * ILP / per_iter);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int threads = 128; threads <= 1024; threads *= 2) { run<1>(threads, p.multiProcessorCount); run<2>(threads, p.multiProcessorCount); run<4>(threads, p.multiProcessorCount); run<8>(threads, p.multiProcessorCount); }
    return 0;
}
