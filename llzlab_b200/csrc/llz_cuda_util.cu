// llz_cuda_util.cu -- synthetic signal generators and machine probes (bench / test helpers of
// libllzfilter_cuda; not on the filtering path).
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "llz_cuda_common.cuh"

namespace llz {

// ---- per-process state ---------------------------------------------------------------------------
namespace {
thread_local const char *t_kernel = "";
thread_local int t_launches = 0;
}
void note_launch(const char *kernel, int launches)
{
    // the helpers either side of the filtering kernel count as launches but do not name the call
    if (strcmp(kernel, "poly_history_kernel") != 0 && strcmp(kernel, "poly_split_planes_kernel") != 0) t_kernel = kernel;
    t_launches += launches;
}
void note_reset() { t_kernel = ""; t_launches = 0; }
const char *noted_kernel() { return t_kernel; }
int noted_launches() { return t_launches; }

int device_sm_count()
{
    constexpr int kMaxDev = 64;
    static std::atomic<int> cache[kMaxDev];                   // zero-initialised
    int dev = 0;
    LLZ_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < kMaxDev) {
        const int c = cache[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int sms = 0;
    LLZ_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < kMaxDev) cache[dev].store(sms, std::memory_order_relaxed);
    return sms;
}

Tunables &tunables()
{
    // C++11 magic static: the lambda (and its getenv calls) runs once, thread-safely
    static Tunables t = [] {
        Tunables v{};
        const char *e = getenv("LLZ_PIPE_SLOT_MB");
        v.pipe_slot_mib = (e && atof(e) >= 1.0) ? atof(e) : 64.0;
        e = getenv("LLZ_SLIDE_RU");
        v.slide_ru = (e && *e) ? atoi(e) : 0;
        e = getenv("LLZ_FFT8K_SKEW");
        v.fft8k_skew = (e && *e) ? atoi(e) : -1;
        e = getenv("LLZ_FFT16K_SKEW");
        v.fft16k_skew = (e && *e) ? atoi(e) : -1;
        v.umma_band_mib = 32;
        v.umma_knife_cycles = 1000;
        e = getenv("LLZ_UMMA_SLAB_MB");
        v.umma_slab_mib = (e && atof(e) >= 1.0) ? atof(e) : 256.0;
        e = getenv("LLZ_FIR_ALGO");
        v.fir_algo = (e && strcmp(e, "direct") == 0) ? LLZ_CUDA_FIR_ALGO_DIRECT
                     : (e && strcmp(e, "fft") == 0)  ? LLZ_CUDA_FIR_ALGO_FFT : LLZ_CUDA_FIR_ALGO_AUTO;
        return v;
    }();
    return t;
}

// ---- per-channel LCG with jump-ahead -----------------------------------------------------------
// s <- s*1664525 + 1013904223 (mod 2^32); element i of a channel is derived from state i+1
// (SURVEY.md section 8d / 9).  An affine map composes as (A2,C2)o(A1,C1) = (A2*A1, A2*C1 + C2), so
// the state after i steps costs O(log i).
__device__ __forceinline__ void lcg_jump(unsigned long long steps, uint32_t &s)
{
    uint32_t accA = 1u, accC = 0u;              // identity
    uint32_t curA = 1664525u, curC = 1013904223u;
    while (steps) {
        if (steps & 1ull) { accC = curA * accC + curC; accA = curA * accA; }
        curC = curA * curC + curC;
        curA = curA * curA;
        steps >>= 1;
    }
    s = accA * s + accC;
}

constexpr int kLcgRun = 16;

template <typename T, int KIND>
__global__ void lcg_kernel(T *out, long long stride, long long first, long long n, uint32_t seed0)
{
    const int ch = blockIdx.y;
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kLcgRun;
    if (i0 >= n) return;
    uint32_t s = seed0 + (uint32_t)ch;
    lcg_jump((unsigned long long)(first + i0), s);
    T *dst = out + (long long)ch * stride + i0;
    const int cnt = (int)min((long long)kLcgRun, n - i0);
    for (int j = 0; j < cnt; ++j) {
        s = s * 1664525u + 1013904223u;
        if constexpr (KIND == 0) dst[j] = (double)((int)(s >> 8) - 8388608) / 8388608.0;
        else if constexpr (KIND == 1) dst[j] = (float)((double)((int)(s >> 8) - 8388608) / 8388608.0);
        else dst[j] = (int16_t)((int)(s >> 17) - 16384);
    }
}

// ---- register-resident FMA throughput ----------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fma_probe_kernel(T *sink, int iters, T a, T b)
{
    T v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if constexpr (sizeof(T) == 8) v[i] = fma(v[i], a, b);
                else v[i] = fmaf(v[i], a, b);
            }
        }
    }
    T total = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) total += v[i];
    if (total == (T)123456789) sink[0] = total;          // never true; keeps the chain alive
}

template <typename T>
static int probe(double *tflops)
{
    int dev = 0, sms = 0;
    LLZ_CUDA_TRY(cudaGetDevice(&dev));
    LLZ_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    T *sink = nullptr;
    LLZ_CUDA_TRY(cudaMalloc(&sink, sizeof(T)));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    LLZ_CUDA_TRY(cudaEventCreate(&e0));
    LLZ_CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        LLZ_CUDA_TRY(cudaEventRecord(e0));
        fma_probe_kernel<T><<<blocks, threads>>>(sink, iters, (T)0.999, (T)0.001);
        LLZ_CUDA_TRY(cudaEventRecord(e1));
        LLZ_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        LLZ_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
        if (rep > 0 && ms > 0.f) best = fmax(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return 0;
}

}  // namespace llz

extern "C" int llz_cuda_tune(const char *key, double value)
{
    using namespace llz;
    if (!key) { llz_set_error("llz_cuda_tune: null key"); return -1; }
    Tunables &t = tunables();
    if (strcmp(key, "pipe_slot_mib") == 0 && value >= 1.0) { t.pipe_slot_mib = value; return 0; }
    if (strcmp(key, "slide_ru") == 0) { t.slide_ru = (int)value; return 0; }
    if (strcmp(key, "fft8k_skew") == 0) { t.fft8k_skew = (int)value; return 0; }
    if (strcmp(key, "fft16k_skew") == 0) { t.fft16k_skew = (int)value; return 0; }
    if (strcmp(key, "umma_knife_cycles") == 0 && value >= 0 && value <= 100000) { t.umma_knife_cycles = (int)value; return 0; }
    if (strcmp(key, "umma_band_mib") == 0 && value >= 1 && value <= 4096) { t.umma_band_mib = (int)value; return 0; }
    if (strcmp(key, "umma_slab_mib") == 0 && value >= 1.0) { t.umma_slab_mib = value; return 0; }
    if (strcmp(key, "fir_algo") == 0 && value >= 0 && value <= 2) { t.fir_algo = (int)value; return 0; }
    llz_set_error("llz_cuda_tune: unknown key or bad value: %s = %g", key, value);
    return -1;
}

extern "C" int llz_cuda_synth_lcg(void *d_out, long long stride, int n_channels, long long n, int kind,
                                  unsigned seed0, llz_cuda_stream_t stream)
{
    return llz_cuda_synth_lcg_at(d_out, stride, n_channels, 0, n, kind, seed0, stream);
}

extern "C" int llz_cuda_synth_lcg_at(void *d_out, long long stride, int n_channels, long long first, long long n,
                                     int kind, unsigned seed0, llz_cuda_stream_t stream)
{
    using namespace llz;
    if (first < 0) { llz_set_error("synth_lcg: negative start index"); return -1; }
    if (n <= 0 || n_channels <= 0) return 0;
    if (n_channels > 65535) { llz_set_error("synth_lcg: too many channels"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    const long long runs = (n + kLcgRun - 1) / kLcgRun;
    dim3 grid((unsigned)((runs + 255) / 256), (unsigned)n_channels);
    switch (kind) {
    case 0: lcg_kernel<double, 0><<<grid, 256, 0, st>>>((double *)d_out, stride, first, n, seed0); break;
    case 1: lcg_kernel<float, 1><<<grid, 256, 0, st>>>((float *)d_out, stride, first, n, seed0); break;
    case 2: lcg_kernel<int16_t, 2><<<grid, 256, 0, st>>>((int16_t *)d_out, stride, first, n, seed0); break;
    default: llz_set_error("synth_lcg: unknown kind %d", kind); return -1;
    }
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int llz_cuda_probe_fma(int dtype, double *tflops)
{
    if (!tflops) { llz_set_error("probe_fma: null output"); return -1; }
    if (dtype == LLZ_CUDA_F32) return llz::probe<float>(tflops);
    if (dtype == LLZ_CUDA_F64 || dtype == LLZ_CUDA_F64_STRICT) return llz::probe<double>(tflops);
    llz_set_error("probe_fma: unknown dtype %d", dtype);
    return -1;
}
