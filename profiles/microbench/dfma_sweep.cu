// DFMA issue-rate sweep: achieved FP64 TFLOP/s versus warps per SM sub-partition and independent chains per thread.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_sweep dfma_sweep.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 1.2345e-300) out[0] = s;
}
template <int ILP>
void run(int sms, int threads, double *d)
{
    const int iters = 2048;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        k<ILP><<<sms, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    double flops = 2.0 * 16 * ILP * iters * (double)threads * sms;
    printf("warps/SMSP %2d  ILP %2d : %7.2f TFLOP/s\n", threads / 128, ILP, flops / (best * 1e-3) / 1e12);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *d; cudaMalloc(&d, 64);
    for (int threads : {128, 256, 512, 1024}) {
        run<1>(p.multiProcessorCount, threads, d);
        run<2>(p.multiProcessorCount, threads, d);
        run<4>(p.multiProcessorCount, threads, d);
        run<8>(p.multiProcessorCount, threads, d);
        run<16>(p.multiProcessorCount, threads, d);
    }
    return 0;
}
