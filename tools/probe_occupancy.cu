#include <cstdio>
#include <cuda_runtime.h>
template <int NF>
__global__ void __launch_bounds__(1024) k(double *sink, int iters, double a, double b)
{
    double f[NF];
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
    double ra = a + threadIdx.x, rb = b - threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = fma(f[i], ra, rb);
    }
    double t = 0;
    for (int i = 0; i < NF; ++i) t += f[i];
    if (t == 123456.789) sink[0] = t;
}
template <int NF> void run(int sms, double* sink, int warps_per_sm)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096; float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<NF><<<sms, warps_per_sm * 32>>>(sink, iters, 0.999, 0.001);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    double flop = 2.0 * NF * 4 * iters * (double)sms * warps_per_sm * 32;
    printf("warps/SM %2d  indep chains %2d : %6.2f TFLOP/s\n", warps_per_sm, NF, flop / best / 1e9);
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink; cudaMalloc(&sink, 64);
    for (int w : {4, 8, 12, 16, 32}) { run<4>(sms, sink, w); run<8>(sms, sink, w); run<16>(sms, sink, w); }
    return 0;
}
