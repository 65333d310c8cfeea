// llz_cuda_fir.cu -- batched direct-form FIR for sm_100a.
//
// Replaces the inner loops of the reference's llz_fir_filter / llz_conv
// (libllzfilter/llz_fir.c:411-426, 547-584):   y[c][t] = sum_{i<N} h[i] * x[c][t-i].
//
// Grid: x = time tiles of TILE = 256*R outputs, y = channels.  One CTA stages its
// HALO + TILE input samples (HALO = padded tap count) and the taps in shared memory -- a single
// TMA bulk copy each (cp.async.bulk + mbarrier); on edge tiles (stream start, history splice,
// ragged end) the threads fill only the fringes outside the call's input -- then every thread runs the
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
    // Sample e of the span is stream position g0 + e.  The run that lies inside this call's input arrives by one TMA
    // bulk copy (for an interior tile that is the whole span), the taps by another; only the fringes of an edge tile
    // -- the history before x[0], zeros past the end, an odd last sample -- are filled by the threads.  (Edge tiles
    // used to be filled sample by sample; a drop-in frame is mostly edge tiles.)
    {
        constexpr int VU = 16 / (int)sizeof(T);
        const long long g0 = t0 - halo;                     // a multiple of VU: TILE and the padded tap count are
        const long long lo = g0 > 0 ? g0 : 0;
        const long long hi = (g0 + span < a.n) ? g0 + span : a.n;
        long long lo_al = lo, hi_al = hi & ~(long long)(VU - 1);
        const bool tma_x = a.vec_ok && xc != nullptr && hi_al > lo_al;
        if (!tma_x) lo_al = hi_al = lo;
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            const uint32_t tap_bytes = (uint32_t)(a.ntaps_pad * sizeof(T));
            const uint32_t x_bytes = tma_x ? (uint32_t)((hi_al - lo_al) * sizeof(T)) : 0u;
            mbar_expect_tx(bar, tap_bytes + x_bytes);
            tma_bulk_g2s(taps_s, a.taps, tap_bytes, bar);
            if (tma_x) tma_bulk_g2s(xs + (lo_al - g0), xc + lo_al, x_bytes, bar);
        }
        const int e_lo = (int)min(max(lo_al - g0, 0LL), (long long)span);
        const int e_hi = (int)min(max(hi_al - g0, (long long)e_lo), (long long)span);
        if (e_lo > 0 || e_hi < span) {
            const T *hc = a.hist ? a.hist + (long long)ch * (a.ntaps - 1) : nullptr;
            auto sample = [&](int e) -> T {
                const long long g = g0 + e;                 // stream position of this sample
                if (g >= 0) return (g < a.n && xc) ? xc[g] : T(0);
                if (hc && g >= -(long long)(a.ntaps - 1)) return hc[(a.ntaps - 1) + g];   // the flt_len-1 samples before this call
                return T(0);
            };
#pragma unroll 4
            for (int e = threadIdx.x; e < e_lo; e += kFirThreads) xs[e] = sample(e);
#pragma unroll 4
            for (int e = e_hi + threadIdx.x; e < span; e += kFirThreads) xs[e] = sample(e);
        }
        __syncthreads();          // barrier init visible to the waiters, fringes written
        mbar_wait(bar, 0);
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
