// llz_cuda_fir_fft.cu -- overlap-save FIR banks for sm_100a: the tolerance-mode arithmetic of
// llz_fir_filter / llz_conv (libllzfilter/llz_fir.c:411-426, 547-584) when the direct form is
// bound by the FMA pipe rather than by HBM.
//
//   y[c][t] = sum_{i<N} h[i] * x[c][t-i]
//
// A 127-tap filter in direct form is 254 flop per output against 16 bytes (f64): 15.9 flop/B, three
// times the FP64 ridge of a B200, so the direct kernel (llz_cuda_fir.cu) tops out at 36 % of HBM no
// matter how well it feeds the pipe.  Overlap-save with a 1024-point transform does the same
// convolution in ~35 FMA-pipe instructions per output instead of 127:
//
//   * one WARP owns one work item = two consecutive blocks of B = 1024 - halo outputs of one channel
//     (halo = N-1 rounded up to 32 samples, so every 32-lane row of loads and stores is 32-sample aligned), packed as one complex signal  z[n] = xA[n] + i*xB[n]  (h is real, so the real and
//     imaginary parts of  ifft(fft(z) * H)  are the two filtered blocks: no real-FFT split pass);
//   * the 1024-point transform is 32 x 32 (llz_fft32.cuh): each lane holds 32 complex points in
//     registers; lane t gathers  x[s + t + 32 j]  (256 contiguous bytes per warp instruction) -- DFT-32
//     over j -- transpose through shared memory -- DFT-32 over t with the four-step twiddle folded
//     into its butterflies -- multiply by H -- inverse DFT-32 -- transpose -- inverse DFT-32 with the
//     conjugate twiddle folded in, which leaves lane t holding outputs  t + 32 j  again: coalesced
//     streaming stores;
//   * only two transposes per item go through shared memory; the spectrum H (1/1024 folded in) and the
//     16 x 32 folded twiddles sit in shared memory once per CTA (fixed latency, no L1 misses);
//   * no CTA-wide barrier after the table load: warps are independent and loop over items
//     (persistent grid of one CTA per SM), so one warp's loads hide behind the others' butterflies.
//
// Arithmetic stays in the bank's own type (FP64 for double banks): this is a cheaper algorithm, not a
// lower precision; the error against the reference's direct sum is ~1e-15 of full scale (f64) and
// the result is NOT bit-identical -- LLZ_CUDA_F64_STRICT and the drop-in llz_fir_filter keep the
// direct kernel.
//
// Algorithmic cost per item (2B outputs), per lane: 2 x 388 (plain DFT-32) + 2 x 512 (DFT-32 with folded
// twiddles) + 128 (H) = 1928 FMA-pipe instructions.  Shared-memory / LSU wavefronts per item (f64): 512 for the
// two transposes, 128 folded twiddles, 128 spectrum, 128 staged input, 114 global stores: ~1010 against 964
// FP64-pipe cycles -- the two pipes are co-limiters, which is why the twiddles are folded (31 -> 16 table
// reads per pass) rather than multiplied in a separate pass.
#include <stdlib.h>

#include "llz_fft32.cuh"
#include "llz_fir_kernels.h"

namespace llz {

template <typename T> struct Cplx;
template <> struct Cplx<float>  { using type = float2; };
template <> struct Cplx<double> { using type = double2; };

// row pitch of the per-warp transpose buffer: 33 elements keeps both the row-wise stores (lane = row)
// and the column-wise loads (lane = column) bank-conflict-free for 4- and 8-byte elements
constexpr int kFftPitch = kFftR + 1;

// PACK = false: one array at a time through a [32][33] buffer of T (two round trips per exchange);
// PACK = true: (re, im) pairs through a [32][33] buffer of complex elements (one round trip, half the
// shared-memory instructions, twice the buffer).  Both patterns are bank-conflict-free.
template <typename T, bool PACK>
__device__ __forceinline__ void warp_exchange(T (&re)[32], T (&im)[32], void *buf_raw, int lane)
{
    if constexpr (PACK) {
        using C = typename Cplx<T>::type;
        C *buf = reinterpret_cast<C *>(buf_raw);
#pragma unroll
        for (int k = 0; k < 32; ++k) { C v; v.x = re[k]; v.y = im[k]; buf[lane * kFftPitch + k] = v; }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) { const C v = buf[k * kFftPitch + lane]; re[k] = v.x; im[k] = v.y; }
        __syncwarp();
    } else {
        T *buf = reinterpret_cast<T *>(buf_raw);
#pragma unroll
        for (int k = 0; k < 32; ++k) buf[lane * kFftPitch + k] = re[k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) re[k] = buf[k * kFftPitch + lane];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) buf[lane * kFftPitch + k] = im[k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) im[k] = buf[k * kFftPitch + lane];
        __syncwarp();
    }
}

template <typename T>
__device__ __forceinline__ T fir_fft_sample(const FirFftLaunch<T> &a, const T *xc, const T *hc, long long g)
{
    if (g >= 0) return (g < a.n && xc) ? __ldg(xc + g) : T(0);
    if (hc && g >= -(long long)(a.ntaps - 1)) return __ldg(hc + (a.ntaps - 1) + g);
    return T(0);
}

// MODE 0 (kGather): every item is interior (its 1024 + B input samples and 2B outputs lie inside this call's
//   buffers): unguarded loads straight from global memory, optional L2 prefetch of the warp's next item.
// MODE 1 (kStaged): interior items; the warp's NEXT item arrives in a per-warp staging buffer by one TMA bulk copy
//   (cp.async.bulk + mbarrier) issued as soon as the current item has been gathered into registers, so the DRAM
//   latency of the input hides behind a whole item of butterflies.  Needs 16-byte aligned channel rows.
// MODE 2 (kEdge): the first / last items of a channel (history splice, zero fill past the end, flush).
// Separate instantiations rather than branches: the compiler would otherwise clone the forward half of the
// transform behind each load path.
constexpr int kGather = 0, kStaged = 1, kEdge = 2;

template <typename T, int WARPS, bool PACK, int MODE>
struct FftSmem {
    static constexpr size_t tables = (size_t)(kTwistEntries + kFftR) * kFftR * 2 * sizeof(T);     // twist table + H
    static constexpr size_t exch = (size_t)kFftR * kFftPitch * sizeof(T) * (PACK ? 2 : 1);        // per warp
    // per warp: 1024 + B samples, B <= 992 (halo >= 32)
    static constexpr size_t stage = MODE == kStaged ? (size_t)(kFftN + (kFftN - 32)) * sizeof(T) : 0;
    static constexpr size_t bars = 128;
    static constexpr size_t total = tables + bars + WARPS * (exch + stage);
};

template <typename T, int WARPS, bool PACK, int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1)
fir_fft_kernel(FirFftLaunch<T> a)
{
    using C = typename Cplx<T>::type;
    using IO = T;
    using SM = FftSmem<T, WARPS, PACK, MODE>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C *tw_s = reinterpret_cast<C *>(smem_raw);                        // [16][32] folded twiddles (llz_fft32.cuh)
    C *H_s = tw_s + kTwistEntries * kFftR;                            // [32][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + SM::tables) + warp;
    unsigned char *buf = smem_raw + SM::tables + SM::bars + (size_t)warp * SM::exch;
    [[maybe_unused]] IO *stage = reinterpret_cast<IO *>(smem_raw + SM::tables + SM::bars + WARPS * SM::exch + (size_t)warp * SM::stage);

    {
        const C *tw_g = reinterpret_cast<const C *>(a.tw);
        const C *H_g = reinterpret_cast<const C *>(a.H);
        for (int i = threadIdx.x; i < kFftR * kFftR; i += WARPS * 32) {
            if (i < kTwistEntries * kFftR) tw_s[i] = tw_g[i];
            H_s[i] = H_g[i];
        }
    }
    __syncthreads();

    const int hl = a.halo;              // N-1 rounded up to 32 samples: every row of 32 lanes is 32-sample aligned
    const int B = a.B;
    const long long total = a.items_per_channel * a.n_channels;
    const long long item_step = (long long)gridDim.x * WARPS;
    long long item = (long long)blockIdx.x * WARPS + warp;

    // first input sample (block A) of an interior item
    auto item_src = [&](long long it) -> const IO * {
        const int c = (int)(it / a.items_per_channel);
        const long long p = a.first_pair + (it - (long long)c * a.items_per_channel);
        return a.x + (long long)c * a.x_stride + p * (2LL * B) - hl;
    };
    [[maybe_unused]] const uint32_t span_bytes = (uint32_t)((kFftN + B) * sizeof(IO));
    [[maybe_unused]] uint32_t phase = 0;
    if constexpr (MODE == kStaged) {
        if (lane == 0) {
            mbar_init(bar, 1);
            if (item < total) {
                mbar_expect_tx(bar, span_bytes);
                tma_bulk_g2s(stage, item_src(item), span_bytes, bar);
            }
        }
        __syncwarp();
    }

    for (; item < total; item += item_step) {
        const int ch = (int)(item / a.items_per_channel);
        long long pair = a.first_pair + (item - (long long)ch * a.items_per_channel);
        if constexpr (MODE == kEdge) { if (pair >= a.gap_start) pair += a.gap_len; }
        const long long o = pair * (2LL * B);          // first output of block A
        const long long s = o - hl;                    // first input of block A
        const IO *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
        IO *yc = a.y + (long long)ch * a.y_stride;

        T re[32], im[32];
        // ---- gather: lane t holds z[t + 32 j] ----
        if constexpr (MODE == kGather || MODE == kStaged) {
            const IO *p;
            if constexpr (MODE == kStaged) {
                mbar_wait(bar, phase);
                phase ^= 1;
                p = stage + lane;
            } else {
                p = xc + s + lane;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if constexpr (MODE == kStaged) { re[j] = p[32 * j]; im[j] = p[B + 32 * j]; }
                else { re[j] = __ldg(p + 32 * j); im[j] = __ldg(p + B + 32 * j); }
            }
            if constexpr (MODE == kStaged) {
                __syncwarp();                          // every lane has its samples: the buffer may be refilled
                if (lane == 0 && item + item_step < total) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads before the async-proxy refill
                    mbar_expect_tx(bar, span_bytes);
                    tma_bulk_g2s(stage, item_src(item + item_step), span_bytes, bar);
                }
            } else if (a.prefetch && item + item_step < total) {
                // pull this warp's next item into L2 while this one computes
                const char *np = reinterpret_cast<const char *>(item_src(item + item_step));
                for (int off = lane * 128; off < (int)span_bytes; off += 32 * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(np + off));
            }
        } else {
            const T *hc = a.hist ? a.hist + (long long)ch * (a.ntaps - 1) : nullptr;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const long long g = s + lane + 32 * j;
                re[j] = fir_fft_sample(a, xc, hc, g);
                im[j] = fir_fft_sample(a, xc, hc, g + B);
            }
        }

        // the first halo - (N-1) samples of a block reach only discarded outputs; zeroing them makes every kept output
        // a function of its own N-1 predecessors alone, bit for bit: a stream cut at multiples of the block length
        // (time segments with an N-1 halo, pipeline chunks) reproduces the one-shot result exactly
        if (lane < hl - (a.ntaps - 1)) { re[0] = T(0); im[0] = T(0); }

        // ---- forward pass 1: DFT over j -------------------------------------------------------------------
        dft32<T, false>(re, im);
        warp_exchange<T, PACK>(re, im, buf, lane);

        // ---- lane k2: twiddle W_1024^(t*k2) folded into the DFT over t -> Z[k2 + 32 k1]; times H; inverse DFT over k1
        dft32_twisted<T, false>(re, im, tw_s + lane, kFftR);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const C h = H_s[k * kFftR + lane];
            cmul_inplace<T, false>(re[k], im[k], h.x, h.y);
        }
        dft32<T, true>(re, im);
        warp_exchange<T, PACK>(re, im, buf, lane);

        // ---- lane t: conjugate twiddle folded into the inverse DFT over k2 -> y[t + 32 j] ----------------------
        dft32_twisted<T, true>(re, im, tw_s + lane, kFftR);

        // ---- scatter: circular positions m >= halo (whole rows j >= halo/32) are the valid outputs -------------
        IO *q = yc + o - hl + lane;
        const int j0 = hl >> 5;
        if constexpr (MODE != kEdge) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j >= j0) { __stcs(q + 32 * j, re[j]); __stcs(q + B + 32 * j, im[j]); }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const long long tA = o - hl + lane + 32 * j;
                if (j >= j0) {
                    if (tA < a.n) __stcs(q + 32 * j, re[j]);
                    if (tA + B < a.n) __stcs(q + B + 32 * j, im[j]);
                }
            }
        }
    }
}

template <typename T, int WARPS, bool PACK, int MODE>
static int fir_fft_run(FirFftLaunch<T> b, int n_channels, long long first, long long count,
                       long long gap_start, long long gap_len, int sm_count, cudaStream_t stream)
{
    if (count <= 0) return 0;
    constexpr size_t smem = FftSmem<T, WARPS, PACK, MODE>::total;
    static_assert(smem <= 227 * 1024, "overlap-save kernel variant exceeds the shared memory of an SM");
    b.first_pair = first;
    b.items_per_channel = count;
    b.gap_start = gap_start;
    b.gap_len = gap_len;
    auto kern = fir_fft_kernel<T, WARPS, PACK, MODE>;
    LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ctas_needed = (count * n_channels + WARPS - 1) / WARPS;
    const unsigned grid = (unsigned)(ctas_needed < sm_count ? ctas_needed : sm_count);
    kern<<<grid, WARPS * 32, smem, stream>>>(b);
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T, int WARPS, bool PACK, bool STAGE>
static int fir_fft_launch_cfg(const FirFftLaunch<T> &a, int n_channels, long long ppc, long long p_lo, long long p_hi,
                              int sm_count, cudaStream_t stream)
{
    if (fir_fft_run<T, WARPS, PACK, STAGE ? kStaged : kGather>(a, n_channels, p_lo, p_hi - p_lo, ppc, 0, sm_count, stream) != 0)
        return -1;
    return fir_fft_run<T, WARPS, PACK, kEdge>(a, n_channels, 0, ppc - (p_hi - p_lo), p_lo, p_hi - p_lo, sm_count,
                                              a.side ? a.side : stream);
}

template <typename T>
int fir_fft_launch(FirFftLaunch<T> a, int n_channels, cudaStream_t stream)
{
    if (a.n <= 0 || n_channels <= 0) return 0;
    if (a.ntaps < 1 || a.ntaps > kFirFftMaxTaps) {
        llz_set_error("overlap-save FIR kernel takes 1..%d taps, got %d", kFirFftMaxTaps, a.ntaps);
        return -1;
    }
    a.halo = (a.ntaps - 1 + 31) / 32 * 32;
    a.B = kFftN - a.halo;
    a.n_channels = n_channels;
    const long long two_b = 2LL * a.B;
    const long long ppc = (a.n + two_b - 1) / two_b;
    // interior pairs p: p*2B - halo >= 0 and (p+1)*2B <= n, and there is an input buffer at all
    long long p_lo = (a.halo + two_b - 1) / two_b, p_hi = a.n / two_b;
    if (!a.x || p_hi < p_lo) { p_lo = 0; p_hi = 0; }
    const int sm_count = device_sm_count();
    if (sm_count <= 0) return -1;
    // measured on C2 (profiles/r01_sweep_fft.txt): f64 -- 8 warps, plain exchange, the next item staged by TMA (packed
    // exchange and a direct gather when the rows are not 16-byte aligned); f32 -- 20 warps, packed exchange, direct
    // gather + L2 prefetch of the next item.  The other points of that sweep (12 / 16 warps, packed FP32) lost and
    // are no longer built.
    const bool aligned = (reinterpret_cast<uintptr_t>(a.x) & 15u) == 0 && (a.x_stride * sizeof(T)) % 16 == 0;
    a.prefetch = 1;
    if constexpr (sizeof(T) == 8) {
        return aligned ? fir_fft_launch_cfg<T, 8, false, true>(a, n_channels, ppc, p_lo, p_hi, sm_count, stream)
                       : fir_fft_launch_cfg<T, 8, true, false>(a, n_channels, ppc, p_lo, p_hi, sm_count, stream);
    } else {
        return fir_fft_launch_cfg<T, 20, true, false>(a, n_channels, ppc, p_lo, p_hi, sm_count, stream);
    }
}

template int fir_fft_launch<float>(FirFftLaunch<float>, int, cudaStream_t);
template int fir_fft_launch<double>(FirFftLaunch<double>, int, cudaStream_t);

}  // namespace llz
