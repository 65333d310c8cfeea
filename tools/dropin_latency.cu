// dropin_latency.cu -- where does a drop-in llz_resample frame spend its time?  (C1: 47,040-byte frame in,
// 51,200 bytes out.)  Times the library call and, beside it, the floor of the same sequence of CUDA calls
// (pinned H2D, an empty kernel, pinned D2H, stream sync) on this box.
//   nvcc -O2 -arch=sm_100a -I include tools/dropin_latency.cu -L llzlab_b200 -lllzfilter_cuda -o tools/dropin_latency
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "llz_fir.h"
#include "llz_resample.h"

__global__ void empty_kernel() {}

// SM clock as the kernels of a latency-bound loop see it: cycles per nanosecond over a fixed spin
__global__ void clock_probe_kernel(double *mhz)
{
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    while (clock64() - c0 < 20000) { }
    const long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    *mhz = 1e3 * (double)(c1 - c0) / (double)(t1 - t0);
}

static double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
    const int L = argc > 1 ? atoi(argv[1]) : 160, M = argc > 2 ? atoi(argv[2]) : 147, frames = 400;
    unsigned long h = llz_resample_filter_init(L, M, 1.0, BLACKMAN);
    if (h == (unsigned long)-1) { fprintf(stderr, "init failed\n"); return 1; }
    const int in_b = llz_get_resample_framelen_bytes(h);
    std::vector<unsigned char> in(in_b), out(1 << 20);
    for (int i = 0; i < in_b; ++i) in[i] = (unsigned char)(i * 37);
    int out_b = 0;
    for (int i = 0; i < 50; ++i) llz_resample(h, in.data(), in_b, out.data(), &out_b);
    double t0 = now_us();
    for (int i = 0; i < frames; ++i) llz_resample(h, in.data(), in_b, out.data(), &out_b);
    const double lib = (now_us() - t0) / frames;
    {
        double *d_mhz, mhz = 0.0;
        cudaMalloc(&d_mhz, sizeof(double));
        clock_probe_kernel<<<1, 1>>>(d_mhz);
        cudaMemcpy(&mhz, d_mhz, sizeof(double), cudaMemcpyDeviceToHost);
        printf("SM clock right after the frame loop: %.0f MHz\n", mhz);
        cudaFree(d_mhz);
    }
    llz_resample_filter_uninit(h);

    {   // the FIR drop-in beside it: 127 taps, frames of 4096 doubles
        unsigned long f = llz_fir_filter_lpf_init(4096, 127, 0.23, HAMMING);
        std::vector<double> xi(4096, 0.25), yo(4096);
        for (int i = 0; i < 50; ++i) llz_fir_filter(f, xi.data(), yo.data(), 4096);
        t0 = now_us();
        for (int i = 0; i < frames; ++i) llz_fir_filter(f, xi.data(), yo.data(), 4096);
        printf("llz_fir_filter 127 taps, frame 4096 doubles: %.1f us/frame\n", (now_us() - t0) / frames);
        llz_fir_filter_uninit(f);
    }

    void *pin_in, *pin_out, *d_in, *d_out;
    cudaStream_t s;
    cudaHostAlloc(&pin_in, in_b, cudaHostAllocDefault);
    cudaHostAlloc(&pin_out, out_b, cudaHostAllocDefault);
    cudaMalloc(&d_in, in_b);
    cudaMalloc(&d_out, out_b);
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    auto seq = [&](int kernels) {
        memcpy(pin_in, in.data(), in_b);
        cudaMemcpyAsync(d_in, pin_in, in_b, cudaMemcpyHostToDevice, s);
        for (int k = 0; k < kernels; ++k) empty_kernel<<<1, 32, 0, s>>>();
        cudaMemcpyAsync(pin_out, d_out, out_b, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        memcpy(out.data(), pin_out, out_b);
    };
    double floor_us[3];
    for (int kernels = 0; kernels < 3; ++kernels) {
        for (int i = 0; i < 50; ++i) seq(kernels);
        t0 = now_us();
        for (int i = 0; i < frames; ++i) seq(kernels);
        floor_us[kernels] = (now_us() - t0) / frames;
    }
    printf("L/M %d/%d: frame %d B in, %d B out: library %.1f us/frame; floor with 0/1/2 empty kernels %.1f / %.1f / %.1f us\n",
           L, M, in_b, out_b, lib, floor_us[0], floor_us[1], floor_us[2]);
    return 0;
}
