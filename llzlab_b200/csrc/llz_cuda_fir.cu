// llz_cuda_fir.cu -- batched direct-form FIR for sm_100a.
//
// Replaces the inner loops of the reference's llz_fir_filter / llz_conv
// (libllzfilter/llz_fir.c:411-426, 547-584):   y[c][t] = sum_{i<N} h[i] * x[c][t-i].
//
// Grid: x = time tiles of TILE = 256*R outputs, y = channels.  One CTA stages its
// HALO + TILE input samples (HALO = padded tap count) and the taps in shared memory -- interior
// tiles with a single TMA bulk copy each (cp.async.bulk + mbarrier), edge tiles (stream start,
// history splice, ragged end) with a guarded scalar fill -- then every thread runs the
// register-blocked sliding MAC of llz_sliding_mac.cuh over its R outputs and stores them with
// 16-byte vector stores.
//
// Roofline: 2*N flop and 2*sizeof(T) bytes per output.  For N = 127 that
// is 15.9 flop/B (f64) / 31.8 flop/B (f32): above both ridge points, so the FMA pipes bound it.
#include "llz_fir_kernels.h"
#include "llz_sliding_mac.cuh"

namespace llz {

#ifndef LLZ_FIR_THREADS
#define LLZ_FIR_THREADS 128
#endif
constexpr int kFirThreads = LLZ_FIR_THREADS;

template <typename T, int R, bool STRICT, bool BLOCKED>
__global__ void __launch_bounds__(kFirThreads, 512 / kFirThreads)
fir_tile_kernel(FirLaunch<T> a)
{
    using SM = SlidingMac<T, R>;
    constexpr int TILE = kFirThreads * R;
    constexpr int U = SM::U;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *taps_s = reinterpret_cast<T *>(smem_raw + 16);
    T *xs = taps_s + a.ntaps_pad;                       // ntaps_pad*sizeof(T) is a multiple of 16

    const int halo = a.ntaps_pad;
    const long long t0 = (long long)blockIdx.x * TILE;  // first output of this tile
    const int ch = blockIdx.y;
    const T *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
    const int span = halo + TILE;

    // ---- stage taps + input span --------------------------------------------------------------
    const bool interior = a.vec_ok && xc != nullptr && t0 >= halo && t0 + TILE <= a.n;
    if (interior) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            const uint32_t tap_bytes = (uint32_t)(a.ntaps_pad * sizeof(T));
            const uint32_t x_bytes = (uint32_t)(span * sizeof(T));
            mbar_expect_tx(bar, tap_bytes + x_bytes);
            tma_bulk_g2s(taps_s, a.taps, tap_bytes, bar);
            tma_bulk_g2s(xs, xc + (t0 - halo), x_bytes, bar);
        }
        __syncthreads();          // barrier init visible to the waiters
        mbar_wait(bar, 0);
    } else {
        for (int k = threadIdx.x; k < a.ntaps_pad; k += kFirThreads) taps_s[k] = a.taps[k];
        const T *hc = a.hist ? a.hist + (long long)ch * (a.ntaps - 1) : nullptr;
        for (int e = threadIdx.x; e < span; e += kFirThreads) {
            const long long g = t0 - halo + e;          // stream position of this sample
            T v = T(0);
            if (g >= 0) {
                if (g < a.n && xc) v = xc[g];
            } else if (hc && g >= -(long long)(a.ntaps - 1)) {
                v = hc[(a.ntaps - 1) + g];              // the flt_len-1 samples before this call
            }
            xs[e] = v;
        }
        __syncthreads();
    }

    // ---- R outputs per thread -----------------------------------------------------------------
    T acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = T(0);

    const T *win = xs + halo + threadIdx.x * R - U;
    if constexpr (BLOCKED) {
        // partial sums of kBlock taps folded into a second accumulator: keeps the FP32 error of
        // long filters at the short-filter level ("FP32 accuracy at 4095 taps")
        constexpr int kBlock = 128 / SM::GRAN * SM::GRAN;
        T total[R];
#pragma unroll
        for (int r = 0; r < R; ++r) total[r] = T(0);
        for (int k0 = 0; k0 < a.ntaps_pad; k0 += kBlock) {
            const int len = min(kBlock, a.ntaps_pad - k0);
            SM::template run<STRICT>(acc, win - k0, taps_s + k0, len);
#pragma unroll
            for (int r = 0; r < R; ++r) { total[r] += acc[r]; acc[r] = T(0); }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = total[r];
    } else {
        SM::template run<STRICT>(acc, win, taps_s, a.ntaps_pad);
    }

    // ---- store ------------------------------------------------------------------------------------
    const long long t = t0 + (long long)threadIdx.x * R;
    T *yc = a.y + (long long)ch * a.y_stride + t;
    if (a.vec_ok && t + R <= a.n) {
        using V = typename Vec16<T>::type;
#pragma unroll
        for (int r = 0; r < R; r += U) *reinterpret_cast<V *>(yc + r) = pack(&acc[r]);
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (t + r < a.n) yc[r] = acc[r];
    }
}

// new history = the last N-1 samples of (old history ++ this call's input)
template <typename T>
__global__ void fir_history_kernel(const T *x, long long x_stride, long long n, const T *hist_old,
                                   T *hist_new, int hlen)
{
    const int ch = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= hlen) return;
    const long long g = n - hlen + j;                   // position in this call's input
    T v = T(0);
    if (g >= 0) {
        if (x) v = x[(long long)ch * x_stride + g];
    } else if (hist_old) {
        v = hist_old[(long long)ch * hlen + hlen + g];
    }
    hist_new[(long long)ch * hlen + j] = v;
}

template <typename T, int R, bool STRICT, bool BLOCKED>
static int launch_variant(const FirLaunch<T> &a, int n_channels, cudaStream_t stream)
{
    constexpr int TILE = kFirThreads * R;
    const size_t smem = 16 + (size_t)a.ntaps_pad * sizeof(T) + (size_t)(a.ntaps_pad + TILE) * sizeof(T);
    if (smem > 227 * 1024) {
        llz_set_error("FIR with %d taps needs %zu bytes of shared memory (> 227 KiB)", a.ntaps, smem);
        return -1;
    }
    auto kern = fir_tile_kernel<T, R, STRICT, BLOCKED>;
    if (smem > 48 * 1024)
        LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (a.n + TILE - 1) / TILE;
    if (tiles > 0x7fffffffLL || n_channels > 65535) {
        llz_set_error("FIR launch too large: %lld tiles x %d channels", tiles, n_channels);
        return -1;
    }
    dim3 grid((unsigned)tiles, (unsigned)n_channels);
    kern<<<grid, kFirThreads, smem, stream>>>(a);
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T>
int fir_pad_taps(int ntaps, int *variant)
{
    // big tile (R = 7 vectors) unless its coarser tap granularity wastes > 12 % more work
    constexpr int U = Vec16<T>::N;
    const int gran_big = 8 * U, gran_small = 4 * U;
    const int pad_big = (ntaps + gran_big - 1) / gran_big * gran_big;
    const int pad_small = (ntaps + gran_small - 1) / gran_small * gran_small;
    if ((double)pad_big <= 1.12 * pad_small) { *variant = 0; return pad_big; }
    *variant = 1;
    return pad_small;
}

template <typename T>
int fir_launch(FirLaunch<T> a, int n_channels, bool strict, cudaStream_t stream)
{
    if (a.n <= 0 || n_channels <= 0) return 0;
    constexpr int U = Vec16<T>::N;
    int variant = 0;
    a.ntaps_pad = fir_pad_taps<T>(a.ntaps, &variant);
    const bool blocked = sizeof(T) == 4 && a.ntaps_pad > 160;
    if (variant == 0) {
        if (strict)  return launch_variant<T, 7 * U, true, false>(a, n_channels, stream);
        if (blocked) return launch_variant<T, 7 * U, false, true>(a, n_channels, stream);
        return launch_variant<T, 7 * U, false, false>(a, n_channels, stream);
    }
    if (strict)  return launch_variant<T, 3 * U, true, false>(a, n_channels, stream);
    if (blocked) return launch_variant<T, 3 * U, false, true>(a, n_channels, stream);
    return launch_variant<T, 3 * U, false, false>(a, n_channels, stream);
}

template <typename T>
int fir_update_history(const T *x, long long x_stride, long long n, const T *hist_old, T *hist_new,
                       int hlen, int n_channels, cudaStream_t stream)
{
    if (hlen <= 0 || n_channels <= 0) return 0;
    dim3 grid((hlen + 255) / 256, n_channels);
    fir_history_kernel<T><<<grid, 256, 0, stream>>>(x, x_stride, n, hist_old, hist_new, hlen);
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

template int fir_launch<float>(FirLaunch<float>, int, bool, cudaStream_t);
template int fir_launch<double>(FirLaunch<double>, int, bool, cudaStream_t);
template int fir_pad_taps<float>(int, int *);
template int fir_pad_taps<double>(int, int *);
template int fir_update_history<float>(const float *, long long, long long, const float *, float *, int, int, cudaStream_t);
template int fir_update_history<double>(const double *, long long, long long, const double *, double *, int, int, cudaStream_t);

}  // namespace llz
