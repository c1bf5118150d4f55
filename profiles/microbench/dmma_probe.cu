// Is mma.sync f64 (DMMA) bit-identical to a k-ascending chain of IEEE fma?  And how fast is it on B200?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// D[8x8] = sum over K (multiple of 4) of A[8xK] * B[Kx8], A row-major [8][K], B given as Bt[8][K] (col-major B)
__global__ void k_check(const double *A, const double *Bt, int K, double *D)
{
    const int lane = threadIdx.x;
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 0; k0 < K; k0 += 4) {
        const double a = A[(lane >> 2) * K + k0 + (lane & 3)];
        const double b = Bt[(lane >> 2) * K + k0 + (lane & 3)];
        dmma884(c0, c1, a, b, c0, c1);
    }
    D[(lane >> 2) * 8 + (lane & 3) * 2] = c0;
    D[(lane >> 2) * 8 + (lane & 3) * 2 + 1] = c1;
}

template <int ILP>
__global__ void k_rate(double *out, int iters, double a, double b)
{
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b, c[i][0], c[i][1]);
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345e-300) out[0] = s;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int K = 64;
    double hA[8 * K], hB[8 * K], hD[64];
    double *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
    srand(7);
    long mism_seq = 0, mism_any = 0, total = 0;
    for (int trial = 0; trial < 2000; ++trial) {
        for (int i = 0; i < 8 * K; ++i) {
            // wide dynamic range so that the accumulation order matters
            hA[i] = ((rand() / (double)RAND_MAX) - 0.5) * exp2((double)(rand() % 40 - 20));
            hB[i] = ((rand() / (double)RAND_MAX) - 0.5) * exp2((double)(rand() % 40 - 20));
        }
        cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
        k_check<<<1, 32>>>(dA, dB, K, dD);
        cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
        for (int r = 0; r < 8; ++r)
            for (int c = 0; c < 8; ++c) {
                double seq = 0.0;
                for (int k = 0; k < K; ++k) seq = fma(hA[r * K + k], hB[c * K + k], seq);
                ++total;
                if (seq != hD[r * 8 + c]) ++mism_seq;
                // alternative hypothesis: unfused products summed
                double alt = 0.0;
                for (int k = 0; k < K; ++k) alt += hA[r * K + k] * hB[c * K + k];
                if (alt != hD[r * 8 + c]) ++mism_any;
            }
    }
    printf("DMMA m8n8k4 vs sequential k-ascending fma chain: %ld mismatches of %ld\n", mism_seq, total);
    printf("DMMA m8n8k4 vs unfused mul+add chain           : %ld mismatches of %ld\n", mism_any, total);
    double *d; cudaMalloc(&d, 64);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int threads : {128, 256, 512, 1024}) {
        const int iters = 4096;
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(a);
            k_rate<8><<<p.multiProcessorCount, threads>>>(d, iters, 0.999999, 1e-9);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        const double flops = 2.0 * 256 * 8 * 8 * iters * (threads / 32.0) * p.multiProcessorCount;
        printf("DMMA rate, %2d warps/SMSP, 8 independent accumulators: %7.2f TFLOP/s\n", threads / 128, flops / (best * 1e-3) / 1e12);
    }
    return 0;
}
