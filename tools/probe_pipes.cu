// probe_pipes.cu -- micro-benchmarks behind the kernel design choices in DESIGN.md (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_pipes tools/probe_pipes.cu
// Measures on one B200: register-resident DFMA / FFMA rate, FP64 tensor-core (DMMA) rate, and whether the two
// FP64 paths overlap when interleaved in one instruction stream.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int NF, int NM>
__global__ void __launch_bounds__(256) mix_kernel(double *sink, int iters, double a, double b)
{
    double f[NF > 0 ? NF : 1];
    double c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
    for (int i = 0; i < NM; ++i) { c0[i] = i; c1[i] = threadIdx.x; }
    double ra = a + threadIdx.x, rb = b - threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < NM; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(ra), "d"(rb));
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = fma(f[i], ra, rb);
        }
    }
    double t = 0;
    for (int i = 0; i < NF; ++i) t += f[i];
    for (int i = 0; i < NM; ++i) t += c0[i] + c1[i];
    if (t == 123456.789) sink[0] = t;
}

template <int NF>
__global__ void __launch_bounds__(256) ffma_kernel(float *sink, int iters, float a, float b)
{
    float f[NF];
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
    float ra = a + threadIdx.x, rb = b - threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = fmaf(f[i], ra, rb);
    }
    float t = 0;
    for (int i = 0; i < NF; ++i) t += f[i];
    if (t == 123456.789f) sink[0] = t;
}

// packed FP32 (fma.rn.f32x2 -> SASS FFMA2): NF independent 2-wide accumulators; NL shared-memory loads per NF FFMA2
// show whether the packed form frees issue slots for other pipes
template <int NF, int NL>
__global__ void __launch_bounds__(256) ffma2_kernel(float *sink, int iters, float a, float b)
{
    __shared__ float sm[256 * 4];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = i;
    __syncthreads();
    unsigned long long f[NF];
    for (int i = 0; i < NF; ++i) f[i] = ((unsigned long long)__float_as_uint(threadIdx.x + i) << 32) | __float_as_uint(i + 1.f);
    const unsigned long long ra = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a + 1e-3f);
    const unsigned long long rb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b - 1e-3f);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(f[i]) : "l"(ra), "l"(rb));
#pragma unroll
            for (int i = 0; i < NL; ++i) acc += sm[(threadIdx.x + 32 * i + it) & 1023];
        }
    }
    float t = acc;
    for (int i = 0; i < NF; ++i) t += __uint_as_float((unsigned)f[i]) + __uint_as_float((unsigned)(f[i] >> 32));
    if (t == 123456.789f) sink[0] = t;
}

// scalar FFMA with the same shared-memory load mix, for comparison
template <int NF, int NL>
__global__ void __launch_bounds__(256) ffma_lds_kernel(float *sink, int iters, float a, float b)
{
    __shared__ float sm[256 * 4];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = i;
    __syncthreads();
    float f[NF];
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
    float ra = a + threadIdx.x, rb = b - threadIdx.x, acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = fmaf(f[i], ra, rb);
#pragma unroll
            for (int i = 0; i < NL; ++i) acc += sm[(threadIdx.x + 32 * i + it) & 1023];
        }
    }
    float t = acc;
    for (int i = 0; i < NF; ++i) t += f[i];
    if (t == 123456.789f) sink[0] = t;
}

// legacy tensor-core path: mma.sync m16n8k16 f16 x f16 -> f32, NM independent accumulator tiles per warp
template <int NM>
__global__ void __launch_bounds__(256) hmma_kernel(float *sink, int iters, unsigned a0, unsigned b0)
{
    float c[NM][4];
    for (int i = 0; i < NM; ++i) for (int j = 0; j < 4; ++j) c[i][j] = (float)(threadIdx.x + i + j);
    unsigned a[4] = {a0 + threadIdx.x, a0 ^ 0x1234u, a0 + 7u, a0 * 3u}, b[2] = {b0 + threadIdx.x, b0 ^ 0x4321u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < NM; ++i)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float t = 0;
    for (int i = 0; i < NM; ++i) for (int j = 0; j < 4; ++j) t += c[i][j];
    if (t == 123456.789f) sink[0] = t;
}

// legacy integer tensor-core path: mma.sync m16n8k32 u8 x s8 -> s32 (exact), NM independent accumulator tiles per warp
template <int NM>
__global__ void __launch_bounds__(256) imma_kernel(float *sink, int iters, unsigned a0, unsigned b0)
{
    int c[NM][4];
    for (int i = 0; i < NM; ++i) for (int j = 0; j < 4; ++j) c[i][j] = (int)(threadIdx.x + i + j);
    unsigned a[4] = {a0 + threadIdx.x, a0 ^ 0x1234u, a0 + 7u, a0 * 3u}, b[2] = {b0 + threadIdx.x, b0 ^ 0x4321u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < NM; ++i)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int t = 0;
    for (int i = 0; i < NM; ++i) for (int j = 0; j < 4; ++j) t += c[i][j];
    if (t == 123456789) sink[0] = (float)t;
}

template <typename K, typename... A>
static float time_ms(K kern, int blocks, int threads, A... args)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        kern<<<blocks, threads>>>(args...);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    return best;
}

int main()
{
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double *sink; CK(cudaMalloc(&sink, 64));
    const int blocks = sms * 8, threads = 256, iters = 2048;
    const double nthreads = (double)blocks * threads, nwarps = nthreads / 32;
    auto report = [&](const char *name, float ms, double fma_per_thread, double mma_per_warp) {
        double flop = 2.0 * (fma_per_thread * nthreads + mma_per_warp * nwarps * 256.0) * iters * 4;
        printf("%-28s %8.3f ms  %7.2f TFLOP/s (fma part %.2f, mma part %.2f)\n", name, ms, flop / ms / 1e9,
               2.0 * fma_per_thread * nthreads * iters * 4 / ms / 1e9, 2.0 * mma_per_warp * nwarps * 256 * iters * 4 / ms / 1e9);
    };
    report("dfma x8", time_ms(mix_kernel<8, 0>, blocks, threads, sink, iters, 0.999, 0.001), 8, 0);
    report("dfma x16", time_ms(mix_kernel<16, 0>, blocks, threads, sink, iters, 0.999, 0.001), 16, 0);
    report("dmma x4", time_ms(mix_kernel<0, 4>, blocks, threads, sink, iters, 0.999, 0.001), 0, 4);
    report("dmma x8", time_ms(mix_kernel<0, 8>, blocks, threads, sink, iters, 0.999, 0.001), 0, 8);
    report("dfma x8 + dmma x1", time_ms(mix_kernel<8, 1>, blocks, threads, sink, iters, 0.999, 0.001), 8, 1);
    report("dfma x8 + dmma x2", time_ms(mix_kernel<8, 2>, blocks, threads, sink, iters, 0.999, 0.001), 8, 2);
    report("dfma x8 + dmma x4", time_ms(mix_kernel<8, 4>, blocks, threads, sink, iters, 0.999, 0.001), 8, 4);
    report("dfma x16 + dmma x2", time_ms(mix_kernel<16, 2>, blocks, threads, sink, iters, 0.999, 0.001), 16, 2);
    {
        float ms = time_ms(ffma_kernel<16>, blocks, threads, (float *)sink, iters, 0.999f, 0.001f);
        printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", "ffma x16", ms, 2.0 * 16 * nthreads * iters * 4 / ms / 1e9);
    }
    {
        float ms = time_ms(ffma2_kernel<16, 0>, blocks, threads, (float *)sink, iters, 0.999f, 0.001f);
        printf("%-28s %8.3f ms  %7.2f TFLOP/s (packed fma.rn.f32x2, 2 FMA per lane per instruction)\n", "ffma2 x16", ms,
               2.0 * 2 * 16 * nthreads * iters * 4 / ms / 1e9);
        ms = time_ms(ffma_lds_kernel<16, 8>, blocks, threads, (float *)sink, iters, 0.999f, 0.001f);
        printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", "ffma x16 + 8 lds+fadd", ms, 2.0 * 16 * nthreads * iters * 4 / ms / 1e9);
        ms = time_ms(ffma2_kernel<8, 8>, blocks, threads, (float *)sink, iters, 0.999f, 0.001f);
        printf("%-28s %8.3f ms  %7.2f TFLOP/s (same FMA count as the line above)\n", "ffma2 x8 + 8 lds+fadd", ms,
               2.0 * 2 * 8 * nthreads * iters * 4 / ms / 1e9);
    }
    for (int pass = 0; pass < 2; ++pass) {
        const int nm = pass ? 16 : 8;
        float ms = pass ? time_ms(hmma_kernel<16>, blocks, threads, (float *)sink, iters, 0x3c003c00u, 0x3c003c00u)
                        : time_ms(hmma_kernel<8>, blocks, threads, (float *)sink, iters, 0x3c003c00u, 0x3c003c00u);
        printf("hmma m16n8k16 f16 x%-2d           %8.3f ms  %7.2f TFLOP/s (legacy mma.sync path)\n", nm, ms,
               2.0 * 16 * 8 * 16 * nm * nwarps * iters * 4 / ms / 1e9);
    }
    for (int pass = 0; pass < 2; ++pass) {
        const int nm = pass ? 16 : 8;
        float ms = pass ? time_ms(imma_kernel<16>, blocks, threads, (float *)sink, iters, 0x01020304u, 0x04030201u)
                        : time_ms(imma_kernel<8>, blocks, threads, (float *)sink, iters, 0x01020304u, 0x04030201u);
        printf("imma m16n8k32 u8*s8 x%-2d          %8.3f ms  %7.2f TOP/s (legacy mma.sync path, exact s32 accumulate)\n", nm, ms,
               2.0 * 16 * 8 * 32 * nm * nwarps * iters * 4 / ms / 1e9);
    }
    return 0;
}
