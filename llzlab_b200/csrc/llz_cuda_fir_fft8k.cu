// llz_cuda_fir_fft8k.cu -- overlap-save FIR banks for long filters (545 .. 6145 taps) on sm_100a:
// the tolerance-mode arithmetic of llz_fir_filter / llz_conv (libllzfilter/llz_fir.c:411-426, 547-584).
//
//   y[c][t] = sum_{i<N} h[i] * x[c][t-i]
//
// A 4095-tap filter in direct form is 8190 flop per output: the direct kernel (llz_cuda_fir.cu) runs the FP64
// pipe at 91 % and still delivers under 1 % of the HBM roof.  Overlap-save with an 8192-point transform needs
// ~81 FMA-pipe instructions per output instead of 4095.
//
//   * one CTA of 256 threads owns one work item = two consecutive blocks of B = 8192 - halo outputs of one
//     channel (halo = N-1 rounded up to 256 samples), packed as one complex signal z = xA + i*xB (h is real);
//   * the transform is 8 x 1024: every thread holds 32 complex points in registers throughout.
//       forward:  thread tid gathers  z[tid + 256 q + 1024 a]  (q < 4, a < 8; 2 KB contiguous per row of 256
//                 threads) -- four DFT-8 over a -- CTA-wide exchange through shared memory, after which WARP b
//                 owns the 1024-point sub-transform of residue b: lane t, register j <-> n_lo = t + 32 j -- DFT-32
//                 over j with the twiddle exp(-2 pi i b n_lo / 8192) folded in (its j part is warp-uniform) --
//                 warp-local exchange -- DFT-32 over t with the rest of that twiddle and the four-step twiddle
//                 folded in (table indexed by (b, lane)) -> bin  k = b + 8 (k2 + 32 k1)  in lane k2, register k1;
//       multiply by the spectrum H (1/8192 folded in, stored in this layout);
//       inverse:  the mirror image -- DFT-32 over k1 -- warp-local exchange -- DFT-32 over k2 with the conjugate
//                 four-step twiddle folded in -- CTA-wide exchange -- four DFT-8 over b with the conjugate
//                 exp(+2 pi i b n_lo / 8192) folded in -> thread tid holds outputs tid + 256 q + 1024 a again:
//                 coalesced streaming stores of the rows at or beyond the halo.
//   * the 128 KB exchange buffer is used as eight 16 KB warp slices for the warp-local exchanges (XOR-swizzled
//     columns, conflict-free) and as [b][j][t] for the CTA-wide ones; three __syncthreads per item;
//   * the buffer's idle phases carry the loads: the NEXT item's 96 KB input span arrives in it by TMA bulk copies
//     (cp.async.bulk + mbarrier) while the current item runs its last DFT-8s and stores, and each warp's 16 KB
//     slice of the spectrum arrives in the warp's own slice while the warp runs its second forward DFT-32; the
//     item after next is prefetched into L2.
//
// Arithmetic stays in the bank's own type; error against the direct sum ~1e-16 of sum|h| (f64), ~1e-8 (f32);
// results are not bit-identical to the reference -- LLZ_CUDA_F64_STRICT keeps the direct kernel.
//
// Per thread and item: 4 x 52 (DFT-8) + 3 x 512 (folded DFT-32) + 388 (plain DFT-32) + 128 (H) + 4 x 80 (folded
// DFT-8) = 2580 FMA-pipe instructions for 32 outputs at B = 4096: 80.6 per output.
#include <stdlib.h>

#include "llz_fft32.cuh"
#include "llz_fir_kernels.h"

namespace llz {

template <typename T> struct Cplx8k;
template <> struct Cplx8k<float>  { using type = float2; };
template <> struct Cplx8k<double> { using type = double2; };

constexpr int kFft8kThreads = 256;

template <typename T>
struct Fft8kSmem {
    static constexpr size_t bars = 128;                       // [0] input staging, [1 + b] spectrum slice of warp b
    static constexpr size_t tabw = (size_t)kTwistEntries * kFftR * 2 * sizeof(T);
    static constexpr size_t tab2 = (size_t)8 * kTwistEntries * kFftR * 2 * sizeof(T);
    static constexpr size_t xbuf = (size_t)kFft8kN * 2 * sizeof(T);
    static constexpr size_t total = bars + tabw + tab2 + xbuf;
};

__device__ __forceinline__ void fence_proxy_async_smem()
{
    // orders this thread's (and, after a barrier, the CTA's) generic-proxy shared-memory accesses before later
    // async-proxy (TMA) accesses to the same locations
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <typename T>
__device__ __forceinline__ T fir_fft8k_sample(const FirFftLaunch<T> &a, const T *xc, const T *hc, long long g)
{
    if (g >= 0) return (g < a.n && xc) ? __ldg(xc + g) : T(0);
    if (hc && g >= -(long long)(a.ntaps - 1)) return __ldg(hc + (a.ntaps - 1) + g);
    return T(0);
}

// EDGE = false: interior items only (unguarded stores; the input span of the NEXT item and the warp's slice of the
// spectrum arrive in the exchange buffer by TMA bulk copies while it is idle); EDGE = true: first / last items of a
// channel (guarded global loads and stores).  STAGE needs 16-byte aligned channel rows.
template <typename T, bool EDGE, bool STAGE>
__global__ void __launch_bounds__(kFft8kThreads, sizeof(T) == 4 ? 2 : 1)
fir_fft8k_kernel(FirFftLaunch<T> a)
{
    using C = typename Cplx8k<T>::type;
    using SM = Fft8kSmem<T>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    C *tabw_s = reinterpret_cast<C *>(smem_raw + SM::bars);                                  // [16][32]
    C *tab2_s = reinterpret_cast<C *>(smem_raw + SM::bars + SM::tabw);                       // [8][16][32]
    C *xbuf = reinterpret_cast<C *>(smem_raw + SM::bars + SM::tabw + SM::tab2);              // 8192 complex
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if constexpr (STAGE) {
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            for (int b = 0; b < 8; ++b) mbar_init(&bars[1 + b], 1);
        }
    }

    for (int i = tid; i < 8 * kTwistEntries * kFftR; i += kFft8kThreads) {
        if (i < kTwistEntries * kFftR) tabw_s[i] = reinterpret_cast<const C *>(a.tw)[i];
        tab2_s[i] = reinterpret_cast<const C *>(a.tw2)[i];
    }
    __syncthreads();

    const C *Hc = reinterpret_cast<const C *>(a.H) + warp * (kFftR * kFftR) + lane;      // [b][k1][k2]
    const C *tab3 = reinterpret_cast<const C *>(a.tw3) + tid;                            // [q][e][tid]
    C *slice = xbuf + warp * (kFftR * kFftR);
    const int hl = a.halo, B = a.B;
    const long long total = a.items_per_channel * a.n_channels;
    const int span_bytes = (kFft8kN + B) * (int)sizeof(T);
    [[maybe_unused]] uint32_t in_phase = 0, h_phase = 0;

    // first input sample (block A) of an interior item
    auto item_src = [&](long long it) -> const T * {
        const int c = (int)(it / a.items_per_channel);
        const long long p = a.first_pair + (it - (long long)c * a.items_per_channel);
        return a.x + (long long)c * a.x_stride + p * (2LL * B) - hl;
    };
    // one thread: bring an item's contiguous input span into the (idle) exchange buffer, four bulk copies
    auto stage_input = [&](long long it) {
        const char *src = reinterpret_cast<const char *>(item_src(it));
        char *dst = reinterpret_cast<char *>(xbuf);
        mbar_expect_tx(&bars[0], (uint32_t)span_bytes);
        const int piece = (span_bytes / 4 + 15) & ~15;
        for (int off = 0; off < span_bytes; off += piece)
            tma_bulk_g2s(dst + off, src + off, (uint32_t)min(piece, span_bytes - off), &bars[0]);
    };
    if constexpr (STAGE) {
        if (tid == 0 && (long long)blockIdx.x < total) stage_input(blockIdx.x);
    }

    for (long long item = blockIdx.x; item < total; item += gridDim.x) {
        const int ch = (int)(item / a.items_per_channel);
        long long pair = a.first_pair + (item - (long long)ch * a.items_per_channel);
        if constexpr (EDGE) { if (pair >= a.gap_start) pair += a.gap_len; }
        const long long o = pair * (2LL * B);
        const long long s = o - hl;
        const T *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
        T *yc = a.y + (long long)ch * a.y_stride;

        T re[32], im[32];
        // ---- gather: register q*8 + a holds z[tid + 256 q + 1024 a] ----------------------------------------
        if constexpr (!EDGE) {
            if constexpr (STAGE) {
                mbar_wait(&bars[0], in_phase);
                in_phase ^= 1;
                const T *p = reinterpret_cast<const T *>(xbuf) + tid;
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int aa = 0; aa < 8; ++aa) {
                        re[q * 8 + aa] = p[256 * q + 1024 * aa];
                        im[q * 8 + aa] = p[B + 256 * q + 1024 * aa];
                    }
                __syncthreads();                           // everyone holds its samples: the buffer turns into the exchange buffer
            } else {
                const T *p = xc + s + tid;
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int aa = 0; aa < 8; ++aa) {
                        re[q * 8 + aa] = __ldg(p + 256 * q + 1024 * aa);
                        im[q * 8 + aa] = __ldg(p + B + 256 * q + 1024 * aa);
                    }
            }
            if (a.prefetch && item + gridDim.x < total) {
                const char *src = reinterpret_cast<const char *>(item_src(item + gridDim.x));
                for (int off = tid * 128; off < span_bytes; off += kFft8kThreads * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(src + off));
            }
        } else {
            const T *hc = a.hist ? a.hist + (long long)ch * (a.ntaps - 1) : nullptr;
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int aa = 0; aa < 8; ++aa) {
                    const long long g = s + tid + 256 * q + 1024 * aa;
                    re[q * 8 + aa] = fir_fft8k_sample(a, xc, hc, g);
                    im[q * 8 + aa] = fir_fft8k_sample(a, xc, hc, g + B);
                }
        }

        // the first halo - (N-1) samples of a block reach only discarded outputs; zeroing them makes every kept output a
        // function of its own N-1 predecessors alone, bit for bit (see llz_cuda_fir_fft.cu)
        if (tid < hl - (a.ntaps - 1)) { re[0] = T(0); im[0] = T(0); }

        // ---- DFT-8 over a; CTA-wide exchange: warp b receives n_lo = t + 32 j of residue b --------------------
        dft8<T, false, 0>(re, im); dft8<T, false, 8>(re, im); dft8<T, false, 16>(re, im); dft8<T, false, 24>(re, im);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                C v; v.x = re[q * 8 + b]; v.y = im[q * 8 + b];
                xbuf[(b * kFftR + warp + 8 * q) * kFftR + lane] = v;
            }
        __syncthreads();
        // Warps w and w + 4 share a scheduler and leave the barrier in the same phase, so they load together and then
        // queue on the FMA pipe together.  Holding warps 4..7 back by about one transform phase lets one warp of every
        // scheduler move data while the other transforms (measured: f64 +7..8 %, f32 +3 %; profiles/r01_sweep_skew.txt).
        if (a.skew > 0 && warp >= 4) {
            const long long t0 = clock64();
            while (clock64() - t0 < a.skew) { }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) { const C v = slice[j * kFftR + lane]; re[j] = v.x; im[j] = v.y; }
        __syncwarp();

        // ---- warp b: 1024-point forward transform of residue b with the outer twiddle folded in --------------
        dft32_twisted<T, false>(re, im, tabw_s + 4 * warp, kFftR);
#pragma unroll
        for (int k = 0; k < 32; ++k) { C v; v.x = re[k]; v.y = im[k]; slice[lane * kFftR + (k ^ lane)] = v; }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) { const C v = slice[k * kFftR + (lane ^ k)]; re[k] = v.x; im[k] = v.y; }
        __syncwarp();
        if constexpr (STAGE) {
            // the slice is idle until the next exchange: fetch this warp's 32 x 32 bins of the spectrum into it
            if (lane == 0) {
                fence_proxy_async_smem();
                mbar_expect_tx(&bars[1 + warp], (uint32_t)(kFftR * kFftR * sizeof(C)));
                tma_bulk_g2s(slice, Hc - lane, (uint32_t)(kFftR * kFftR * sizeof(C)), &bars[1 + warp]);
            }
        }
        dft32_twisted<T, false>(re, im, tab2_s + warp * (kTwistEntries * kFftR) + lane, kFftR);

        // ---- spectrum, inverse 1024-point transform ---------------------------------------------------------------
        if constexpr (STAGE) {
            mbar_wait(&bars[1 + warp], h_phase);
            h_phase ^= 1;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const C h = slice[k * kFftR + lane];
                cmul_inplace<T, false>(re[k], im[k], h.x, h.y);
            }
            __syncwarp();                                  // all lanes are done with the spectrum before the slice is reused
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const C h = __ldg(Hc + k * kFftR);
                cmul_inplace<T, false>(re[k], im[k], h.x, h.y);
            }
        }
        dft32<T, true>(re, im);
#pragma unroll
        for (int k = 0; k < 32; ++k) { C v; v.x = re[k]; v.y = im[k]; slice[lane * kFftR + (k ^ lane)] = v; }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) { const C v = slice[k * kFftR + (lane ^ k)]; re[k] = v.x; im[k] = v.y; }
        __syncwarp();
        dft32_twisted<T, true>(re, im, tabw_s + lane, kFftR);

        // ---- CTA-wide exchange back; DFT-8 over b with the conjugate outer twiddle folded in ------------------
#pragma unroll
        for (int j = 0; j < 32; ++j) { C v; v.x = re[j]; v.y = im[j]; slice[j * kFftR + lane] = v; }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const C v = xbuf[(b * kFftR + warp + 8 * q) * kFftR + lane];
                re[q * 8 + b] = v.x; im[q * 8 + b] = v.y;
            }
        __syncthreads();                                   // the next item's exchange overwrites every slice
        if constexpr (STAGE) {
            if (tid == 0 && item + gridDim.x < total) {
                fence_proxy_async_smem();
                stage_input(item + gridDim.x);             // lands while this item's last DFT-8s and stores run
            }
        }
        dft8_twisted<T, true, 0>(re, im, __ldg(tab3 + 0 * 256), __ldg(tab3 + 1 * 256), __ldg(tab3 + 2 * 256), __ldg(tab3 + 3 * 256));
        dft8_twisted<T, true, 8>(re, im, __ldg(tab3 + 4 * 256), __ldg(tab3 + 5 * 256), __ldg(tab3 + 6 * 256), __ldg(tab3 + 7 * 256));
        dft8_twisted<T, true, 16>(re, im, __ldg(tab3 + 8 * 256), __ldg(tab3 + 9 * 256), __ldg(tab3 + 10 * 256), __ldg(tab3 + 11 * 256));
        dft8_twisted<T, true, 24>(re, im, __ldg(tab3 + 12 * 256), __ldg(tab3 + 13 * 256), __ldg(tab3 + 14 * 256), __ldg(tab3 + 15 * 256));

        // ---- scatter: rows (q + 4 a) at or beyond halo / 256 are the valid outputs -----------------------------------
        T *qy = yc + o - hl + tid;
        const int r0 = hl >> 8;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {
                const int off = 256 * q + 1024 * aa;
                if (q + 4 * aa >= r0) {
                    if constexpr (!EDGE) {
                        __stcs(qy + off, re[q * 8 + aa]);
                        __stcs(qy + B + off, im[q * 8 + aa]);
                    } else {
                        const long long tA = o - hl + tid + off;
                        if (tA < a.n) __stcs(qy + off, re[q * 8 + aa]);
                        if (tA + B < a.n) __stcs(qy + B + off, im[q * 8 + aa]);
                    }
                }
            }
    }
}

template <typename T, bool EDGE, bool STAGE>
static int fir_fft8k_run(FirFftLaunch<T> b, int n_channels, long long first, long long count, long long gap_start,
                         long long gap_len, int sm_count, cudaStream_t stream)
{
    if (count <= 0) return 0;
    constexpr size_t smem = Fft8kSmem<T>::total;
    static_assert(smem <= 227 * 1024, "8192-point overlap-save kernel exceeds the shared memory of an SM");
    b.first_pair = first;
    b.items_per_channel = count;
    b.gap_start = gap_start;
    b.gap_len = gap_len;
    auto kern = fir_fft8k_kernel<T, EDGE, STAGE>;
    LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long items = count * n_channels;
    const long long slots = (long long)sm_count * (sizeof(T) == 4 ? 2 : 1);
    const unsigned grid = (unsigned)(items < slots ? items : slots);
    kern<<<grid, kFft8kThreads, smem, stream>>>(b);
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T>
int fir_fft8k_launch(FirFftLaunch<T> a, int n_channels, cudaStream_t stream)
{
    if (a.n <= 0 || n_channels <= 0) return 0;
    if (a.ntaps < 1 || a.ntaps > kFirFft8kMaxTaps) {
        llz_set_error("8192-point overlap-save FIR kernel takes 1..%d taps, got %d", kFirFft8kMaxTaps, a.ntaps);
        return -1;
    }
    a.halo = (a.ntaps - 1 + 255) / 256 * 256;
    a.B = kFft8kN - a.halo;
    a.n_channels = n_channels;
    const long long two_b = 2LL * a.B;
    const long long ppc = (a.n + two_b - 1) / two_b;
    long long p_lo = (a.halo + two_b - 1) / two_b, p_hi = a.n / two_b;
    if (!a.x || p_hi < p_lo) { p_lo = 0; p_hi = 0; }
    const int sm_count = device_sm_count();
    if (sm_count <= 0) return -1;
    a.prefetch = 1;
    const int sk = tunables().fft8k_skew;                      // llz_cuda_tune("fft8k_skew", cycles); < 0: the measured default
    a.skew = sk >= 0 ? sk : (sizeof(T) == 8 ? 1300 : 1500);
    const bool aligned = (reinterpret_cast<uintptr_t>(a.x) & 15u) == 0 && (a.x_stride * sizeof(T)) % 16 == 0;
    // measured on C5: staging gains 2.5 % in f64 (one CTA per SM) and loses 4 % in f32 (two CTAs per SM hide the loads)
    const bool stage = aligned && sizeof(T) == 8;
    const int rc = stage ? fir_fft8k_run<T, false, true>(a, n_channels, p_lo, p_hi - p_lo, ppc, 0, sm_count, stream)
                         : fir_fft8k_run<T, false, false>(a, n_channels, p_lo, p_hi - p_lo, ppc, 0, sm_count, stream);
    if (rc != 0) return -1;
    return fir_fft8k_run<T, true, false>(a, n_channels, 0, ppc - (p_hi - p_lo), p_lo, p_hi - p_lo, sm_count, a.side ? a.side : stream);
}

template int fir_fft8k_launch<float>(FirFftLaunch<float>, int, cudaStream_t);
template int fir_fft8k_launch<double>(FirFftLaunch<double>, int, cudaStream_t);

}  // namespace llz
