// llz_cuda_resample.cu -- polyphase decimate / interpolate / rational L/M resample for sm_100a.
//
// Replaces the per-frame loops of the reference (libllzfilter/llz_resample.c:425-491 decimate,
// :494-541 interp, :544-609 resample) by whole-signal kernels over the canonical form described in
// llz_poly_kernels.h.  int16 PCM in, int16 PCM out; phase (o % L) and input index
// (floor(o*M/L)) are integer arithmetic, 64-bit, exactly the reference's sequence.
//
// Arithmetic modes
//   ACC_F64         FP64 FMA in any order, then a guard: if gain*sum lies within `thr` of a NON-ZERO
//                   integer (thr = bound on the difference between two FP64 evaluations of the sum for
//                   full-scale input; truncation toward zero is continuous at 0, so 0 needs no guard
//                   and digital silence stays on the fast path), the output is recomputed in the
//                   reference's order with separate multiply and add.  Rows with a single non-zero tap
//                   (the Nyquist phase whose centre tap is 1-2^-53: SURVEY.md "truncation knife-edge")
//                   are exact in any order and skip the guard.  Result: bit-identical int16, proven
//                   rather than observed.
//   ACC_F64_STRICT  every output in the reference's order (verification mode).
//   ACC_F32         FP32 FMA; single-tap rows are evaluated in FP64 so the knife-edge phase stays exact.
//
// Kernels (poly_launch picks one)
//   poly_slide_kernel    L == 1 (decimation): taps split by residue k mod M into M sliding FIRs over
//                        the M de-interleaved input streams, which stay int16 in shared memory;
//                        register-blocked SlidingMac core shared with the FIR kernel.
//   poly_bank_*_kernel   L >= 16 rational banks: llz_cuda_polybank.cu (register-tiled / DMMA / HMMA).
//   poly_general_kernel  everything else (interp, strict mode, few phases) and the correctness
//                        baseline: one output per thread iteration, input span staged in shared
//                        memory already converted, bank read transposed ([tap][phase]) so a warp's
//                        coefficient loads are contiguous.
#include <limits.h>
#include <stdlib.h>

#include "llz_poly_kernels.h"
#include "llz_poly_device.cuh"
#include "llz_sliding_mac.cuh"

namespace llz {

// ---- general kernel ----------------------------------------------------------------------------

constexpr int kPolyThreads = 256;

template <typename TA, int MODE>
__global__ void __launch_bounds__(kPolyThreads)
poly_general_kernel(PolyLaunch a, int tile_out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TA *xs = reinterpret_cast<TA *>(smem_raw);

    const int ch = blockIdx.y;
    const long long ot = (long long)blockIdx.x * tile_out;
    const int cnt = (int)min((long long)tile_out, a.n_out - ot);
    const long long ofirst = a.o0 + ot;
    const long long s_lo = (ofirst * a.M) / a.L + a.shift - (a.ctaps - 1);
    const long long s_hi = ((ofirst + cnt - 1) * a.M) / a.L + a.shift;
    const int span = (int)(s_hi - s_lo + 1);
    long long frame_end = LLONG_MAX;      // interp: tiles never straddle an input frame
    if (a.frame_len > 0) frame_end = ((ofirst / a.L) / a.frame_len + 1) * (long long)a.frame_len;

    const int16_t *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
    const int16_t *hc = a.hist ? a.hist + (long long)ch * a.hist_len : nullptr;

    for (int e = threadIdx.x; e < span; e += kPolyThreads) {
        const long long s = s_lo + e;
        const int v = (s < frame_end) ? poly_sample(a, xc, hc, s) : 0;
        xs[e] = (TA)v;
    }
    const double thr = a.guard_thr;
    __syncthreads();

    int16_t *yc = a.y + (long long)ch * a.y_stride + ot;
    for (int j = threadIdx.x; j < cnt; j += kPolyThreads) {
        const long long o = ofirst + j;
        const int r = (int)(o % a.L);
        const int b = (int)((o * a.M) / a.L + a.shift - s_lo);     // index of the newest sample in xs
        TA acc = TA(0);
        if constexpr (MODE == LLZ_CUDA_ACC_F64_STRICT) {
            for (int t = 0; t < a.order_len; ++t) {
                const int k = a.order[t];
                acc = __dadd_rn(acc, __dmul_rn(xs[b - k], a.cbankT64[(long long)k * a.L + r]));
            }
        } else if constexpr (sizeof(TA) == 8) {
            const double *g = a.cbankT64 + r;
#pragma unroll 4
            for (int k = 0; k < a.ctaps; ++k) acc = fma(xs[b - k], g[(long long)k * a.L], acc);
        } else {
            const float *g = a.cbankT32 + r;
#pragma unroll 4
            for (int k = 0; k < a.ctaps; ++k) acc = fmaf(xs[b - k], g[(long long)k * a.L], acc);
        }
        double single_x = 0.0;
        if constexpr (MODE == LLZ_CUDA_ACC_F32) {
            const int st = a.single_tap[r];
            if (st >= 0) single_x = (double)xs[b - st];
        }
        yc[j] = poly_emit<MODE, TA>(a, xc, hc, o, acc, thr, single_x);
    }
}

// ---- sliding kernel (L == 1) -------------------------------------------------------------------
//
// y[q] = sum_k c[k] X(q*M - k).  With k = i*M + rho:  X(q*M - k) = X_sigma[q - i - d], where
// X_sigma[j] = X(j*M + sigma), sigma = (M - rho) % M and d = (rho > 0).  Stream sigma is stored
// shifted by d so that the newest sample of output q always sits at element HS + (q - qt):
// 16-byte aligned windows for every residue.
// row jj of `src` (M samples) -> element jj-1 of stream 0 and element jj of streams 1..M-1; MC = compile-time M (0: runtime).
// SHARED tells the compiler the address space of src (LDS instead of generic loads).
template <int MC, int NT, bool SHARED>
__device__ __forceinline__ void slide_deinterleave(int16_t *xs, const int16_t *src_generic, int len, int m_rt, int tid)
{
    const int16_t *src = src_generic;
    if constexpr (SHARED) src = reinterpret_cast<const int16_t *>(__cvta_shared_to_generic(__cvta_generic_to_shared(src_generic)));
    const int M = MC ? MC : m_rt;
    const int rows = len + 1;
#pragma unroll 2
    for (int jj = tid; jj < rows; jj += NT) {
        const int16_t *p = src + jj * M;
        if (jj >= 1) xs[jj - 1] = p[0];
        if (jj < len) {
            if constexpr (MC != 0) {
#pragma unroll
                for (int sigma = 1; sigma < MC; ++sigma) xs[sigma * len + jj] = p[sigma];
            } else {
                for (int sigma = 1; sigma < M; ++sigma) xs[sigma * len + jj] = p[sigma];
            }
        }
    }
}

template <typename TA, int R, int MODE, int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
poly_slide_kernel(PolyLaunch a, int ntp /* padded taps per residue */, int tap_stride /* row length of the uploaded taps */)
{
    using SM = SlidingMac<TA, R, int16_t>;                   // streams stay int16 in shared memory
    using V = typename Vec16<TA>::type;
    constexpr int TILE = NT * R;
    constexpr int U = SM::U;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int M = a.M;
    const int HS = ntp;
    const int len = HS + TILE;
    const int rows = len + 1;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    TA *taps_s = reinterpret_cast<TA *>(smem_raw + 16);       // [M][ntp]
    int16_t *xs = reinterpret_cast<int16_t *>(taps_s + (size_t)M * ntp);   // [M][HS + TILE] de-interleaved streams
    int16_t *raw = xs + (((size_t)M * len + 7) & ~(size_t)7);            // [rows*M + 16] input span as it lies in x

    const int tid = threadIdx.x;
    const int ch = blockIdx.y;
    const long long ot = (long long)blockIdx.x * TILE;        // first output of the tile within this call
    const long long qt = a.o0 + ot;                           // its canonical index
    const int16_t *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
    const int16_t *hc = a.hist ? a.hist + (long long)ch * a.hist_len : nullptr;

    // The tile's input span is contiguous in the stream: rows*M samples starting at canonical index s_base.  It is
    // staged as it lies (one TMA bulk copy for the part inside x, the threads fill history / zero fringes) while the
    // taps are staged, then de-interleaved shared memory to shared memory.
    const long long s_base = (qt - HS - 1) * M;
    bool bulk;
    const int raw_off = poly_stage_span<NT>(a, xc, hc, s_base, rows * M, raw, bar, tid, &bulk);
    {
        const TA *src = (sizeof(TA) == 8) ? reinterpret_cast<const TA *>(a.slide64) : reinterpret_cast<const TA *>(a.slide32);
        const int vpr = ntp / U;                              // vectors per row
        for (int k = tid; k < M * vpr; k += NT) {
            const int rho = k / vpr, i = k - rho * vpr;
            reinterpret_cast<V *>(taps_s + (size_t)rho * ntp)[i] = reinterpret_cast<const V *>(src + (size_t)rho * tap_stride)[i];
        }
    }
    __syncthreads();                                          // barrier initialised, fringes written
    if (bulk) mbar_wait(bar, 0);

    // De-interleave into the M streams: row jj of the span holds X((qt - HS - 1 + jj) * M + sigma); thread jj
    // writes element jj-1 (sigma = 0) / jj (sigma > 0) of each stream, so a warp's stores are contiguous
    // within a stream (conflict-free).
    {
        const int16_t *src = raw + raw_off;
        switch (M) {
        case 2: slide_deinterleave<2, NT, true>(xs, src, len, 2, tid); break;
        case 3: slide_deinterleave<3, NT, true>(xs, src, len, 3, tid); break;
        case 4: slide_deinterleave<4, NT, true>(xs, src, len, 4, tid); break;
        default: slide_deinterleave<0, NT, true>(xs, src, len, M, tid); break;
        }
    }
    __syncthreads();

    TA acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = TA(0);
    for (int rho = 0; rho < M; ++rho) {
        const int sigma = rho ? M - rho : 0;
        const int16_t *win = xs + (size_t)sigma * len + HS + tid * R - U;
        SM::template run<false, true>(acc, win, taps_s + (size_t)rho * ntp, ntp);
    }

    // gain / guard / saturate / truncate, staged through shared memory so the global stores are 16-byte vectors
    __syncthreads();                                          // every thread is done reading xs
    uint32_t *ys32 = reinterpret_cast<uint32_t *>(xs);
    const float gain_f = (float)a.gain;
    const bool unit_gain = a.gain == 1.0;                     // x * 1.0 == x exactly: skip the FP64 multiply
    (void)gain_f; (void)unit_gain;
    const long long q0 = ot + (long long)tid * R;             // output index within this call
    // two passes (see poly_bank_dmma_kernel): a straight-line first pass notes near-integer hits, the rare
    // reference-order recompute patches them afterwards
    static_assert(MODE != LLZ_CUDA_ACC_F64 || R <= 32, "guard hits are one bit per output of the thread");
    uint32_t guard_hits = 0;
#pragma unroll
    for (int r = 0; r < R; r += 2) {
        int16_t o2[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if constexpr (MODE == LLZ_CUDA_ACC_F32) {
                // fast mode: gain, saturate and truncate in FP32 (llz_resample.c:594-601 in single precision)
                float vf = __fmul_rn((float)acc[r + i], gain_f);
                vf = fminf(fmaxf(vf, -32768.f), 32767.f);
                o2[i] = (int16_t)__float2int_rz(vf);
                continue;
            }
            const double v = unit_gain ? (double)acc[r + i] : __dmul_rn((double)acc[r + i], a.gain);
            if constexpr (MODE == LLZ_CUDA_ACC_F64) {
                if (poly_near_nonzero_integer(v, a.guard_thr) && q0 + r + i < a.n_out) guard_hits |= 1u << ((r + i) & 31);
            }
            o2[i] = poly_finish(v);
        }
        ys32[(tid * R + r) >> 1] = (uint32_t)(uint16_t)o2[0] | ((uint32_t)(uint16_t)o2[1] << 16);
    }
    while (guard_hits) {                                      // ~1e-8 of the outputs: the reference's own order
        const int r = __ffs(guard_hits) - 1;
        guard_hits &= guard_hits - 1;
        reinterpret_cast<int16_t *>(xs)[tid * R + r] =
            poly_finish(__dmul_rn(poly_reference_order_sum(a, xc, hc, a.o0 + q0 + r), a.gain));
        atomicAdd(a.guard_count, 1ULL);
    }
    __syncthreads();
    const int cnt = (int)min((long long)TILE, a.n_out - ot);
    int16_t *yc = a.y + (long long)ch * a.y_stride + ot;
    const int16_t *ys = reinterpret_cast<const int16_t *>(xs);
    if ((reinterpret_cast<uintptr_t>(yc) & 15u) == 0) {
        const int nv = cnt >> 3;
        for (int v = tid; v < nv; v += NT)
            reinterpret_cast<uint4 *>(yc)[v] = reinterpret_cast<const uint4 *>(ys)[v];
        for (int j = (nv << 3) + tid; j < cnt; j += NT) yc[j] = ys[j];
    } else {
        for (int j = tid; j < cnt; j += NT) yc[j] = ys[j];
    }
}

// ---- history -----------------------------------------------------------------------------------

__global__ void poly_history_kernel(const int16_t *x, long long x_stride, long long n_in,
                                    const int16_t *hist_old, int16_t *hist_new, int hlen)
{
    const int ch = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= hlen) return;
    const long long g = n_in - hlen + j;
    int16_t v = 0;
    if (g >= 0) {
        if (x) v = x[(long long)ch * x_stride + g];
    } else if (hist_old) {
        v = hist_old[(long long)ch * hlen + hlen + g];
    }
    hist_new[(long long)ch * hlen + j] = v;
}

int poly_update_history(const int16_t *x, long long x_stride, long long n_in, const int16_t *hist_old,
                        int16_t *hist_new, int hist_len, int n_channels, cudaStream_t stream)
{
    if (hist_len <= 0 || n_channels <= 0) return 0;
    dim3 grid((hist_len + 255) / 256, n_channels);
    poly_history_kernel<<<grid, 256, 0, stream>>>(x, x_stride, n_in, hist_old, hist_new, hist_len);
    note_launch("poly_history_kernel");
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- launch selection --------------------------------------------------------------------------

namespace {

constexpr size_t kSmemBudget = 226 * 1024;
#ifndef LLZ_SLIDE_THREADS
#define LLZ_SLIDE_THREADS 64
#endif
constexpr int kSlideThreads = LLZ_SLIDE_THREADS;   // threads per CTA of the sliding kernel

// barrier + taps [M][ntp] (accumulator type) + int16 streams [M][ntp + tile] + raw int16 span
inline size_t slide_smem_bytes(int M, int ntp, int tile, size_t elem)
{
    const size_t raw = (((size_t)(ntp + tile + 1) * M + 16) * 2 + 15) & ~(size_t)15;
    const size_t streams = ((((size_t)M * (ntp + tile) + 7) & ~(size_t)7)) * 2;      // int16
    return 16 + (size_t)M * ntp * elem + streams + raw;
}

// taps per residue padded to whole tap vectors (the kernel runs the last, partial unrolled iteration chunk by chunk)
inline int slide_pad(const PolyLaunch &a, int vec)
{
    const int per = (a.ctaps + a.M - 1) / a.M;
    return (per + vec - 1) / vec * vec;
}

// Tile variants of the sliding kernel: R/U in {11, 7, 5, 3} (odd: conflict-free loads), i.e. tap granularity
// {12, 8, 6, 4} vectors.  Larger R amortises the per-chunk loads and conversions over more FMAs.  Pick the least padded one that fits (two CTAs per SM preferred), larger tiles on ties.
// Returns R/U, or 0 for the general kernel.
template <typename TA>
int pick_slide(const PolyLaunch &a, int *ntp_out)
{
    constexpr int U = Vec16<TA>::N;
    const int avail = (sizeof(TA) == 8) ? a.slide_ntp64 : a.slide_ntp32;   // row length of the uploaded taps
    int best = 0, best_ntp = 0;
    double best_cost = 0.0;
    const int ru[4] = {11, 7, 5, 3};
    const int force = tunables().slide_ru;                    // tuning knob: force a tile variant (11, 7, 5 or 3)
    for (int i = 0; i < 4; ++i) {
        if (force && force != ru[i]) continue;
        if (!force && ru[i] == 11 && sizeof(TA) == 4) continue;   // measured: the 44-output float tile is slower than 20 (C3: 4.9 vs 3.9 ms)
        const int R = ru[i] * U, gran = (ru[i] + 1) * U;
        // f64: whole tap vectors (+2.4 % on C3); f32: the variant's own granularity (the tail costs the float tiles 7 %)
        const int ntp = slide_pad(a, sizeof(TA) == 8 ? U : gran);
        if (ntp > avail) continue;
        const size_t smem = slide_smem_bytes(a.M, ntp, kSlideThreads * R, sizeof(TA));
        if (smem > kSmemBudget) continue;
        // work per output ~ padded taps; the fill / halo overhead shrinks with the tile; one CTA per SM hides less
        double cost = (double)ntp * a.M * (1.0 + (double)ntp / (kSlideThreads * R));
        if (smem > 113 * 1024) cost *= 1.25;                   // one CTA per SM: nothing hides the fill
        else if (smem > 75 * 1024) cost *= 1.05;               // two CTAs per SM
        if (best == 0 || cost < best_cost) { best = ru[i]; best_ntp = ntp; best_cost = cost; }
    }
    *ntp_out = best_ntp;
    return best;
}

int pick_kernel(const PolyLaunch &a, int *ntp)
{
    *ntp = 0;
    if (a.L != 1 || a.shift != 0 || a.frame_len != 0 || a.acc == LLZ_CUDA_ACC_F64_STRICT) return 0;
    if (a.acc == LLZ_CUDA_ACC_F32) return a.slide32 ? pick_slide<float>(a, ntp) : 0;
    return a.slide64 ? pick_slide<double>(a, ntp) : 0;
}

template <typename TA, int R, int MODE>
int launch_slide(const PolyLaunch &a, int ntp, int n_channels, cudaStream_t stream)
{
    constexpr int TILE = kSlideThreads * R;
    const int tap_stride = (sizeof(TA) == 8) ? a.slide_ntp64 : a.slide_ntp32;
    const size_t smem = slide_smem_bytes(a.M, ntp, TILE, sizeof(TA));
    auto kern = poly_slide_kernel<TA, R, MODE, kSlideThreads>;
    if (smem > 48 * 1024)
        LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (a.n_out + TILE - 1) / TILE;
    if (tiles > 0x7fffffffLL) { llz_set_error("resample launch too large"); return -1; }
    kern<<<dim3((unsigned)tiles, (unsigned)n_channels), kSlideThreads, smem, stream>>>(a, ntp, tap_stride);
    note_launch("poly_slide_kernel");
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename TA, int MODE>
int launch_general(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    // outputs per tile: as many as keep the staged input span within ~96 KiB, at most 2048;
    // interp tiles must divide the per-frame output count (frame_len * L)
    const size_t budget = 96 * 1024 / sizeof(TA);
    long long tile = 2048;
    while (tile > 32 && (size_t)((tile * a.M) / a.L + a.ctaps + 2) > budget) tile /= 2;
    if ((size_t)((tile * a.M) / a.L + a.ctaps + 2) * sizeof(TA) > kSmemBudget) {
        llz_set_error("polyphase bank with %d taps per phase does not fit shared memory", a.ctaps);
        return -1;
    }
    if (a.frame_len > 0) {
        const long long per_frame = (long long)a.frame_len * a.L;
        while (per_frame % tile) tile /= 2;
    }
    const size_t smem = (size_t)((tile * a.M) / a.L + a.ctaps + 2) * sizeof(TA);
    auto kern = poly_general_kernel<TA, MODE>;
    if (smem > 48 * 1024)
        LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (a.n_out + tile - 1) / tile;
    if (tiles > 0x7fffffffLL) { llz_set_error("resample launch too large"); return -1; }
    kern<<<dim3((unsigned)tiles, (unsigned)n_channels), kPolyThreads, smem, stream>>>(a, (int)tile);
    note_launch("poly_general_kernel");
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace

const char *poly_kernel_name(const PolyLaunch &a)
{
    int ntp = 0;
    return pick_kernel(a, &ntp) ? "sliding" : "general";
}

int poly_launch(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    if (a.n_out <= 0 || n_channels <= 0) return 0;
    if (n_channels > 65535) { llz_set_error("too many channels for one launch (%d)", n_channels); return -1; }
    int ntp = 0;
    const int ru = pick_kernel(a, &ntp);
    if (ru == 0) {
        const int rc = poly_bank_launch(a, n_channels, stream);
        if (rc != 0) return rc < 0 ? -1 : 0;
    }
    switch (a.acc) {
    case LLZ_CUDA_ACC_F64:
        if (ru == 11) return launch_slide<double, 22, LLZ_CUDA_ACC_F64>(a, ntp, n_channels, stream);
        if (ru == 7) return launch_slide<double, 14, LLZ_CUDA_ACC_F64>(a, ntp, n_channels, stream);
        if (ru == 5) return launch_slide<double, 10, LLZ_CUDA_ACC_F64>(a, ntp, n_channels, stream);
        if (ru == 3) return launch_slide<double, 6, LLZ_CUDA_ACC_F64>(a, ntp, n_channels, stream);
        return launch_general<double, LLZ_CUDA_ACC_F64>(a, n_channels, stream);
    case LLZ_CUDA_ACC_F64_STRICT:
        return launch_general<double, LLZ_CUDA_ACC_F64_STRICT>(a, n_channels, stream);
    case LLZ_CUDA_ACC_F32:
        if (ru == 11) return launch_slide<float, 44, LLZ_CUDA_ACC_F32>(a, ntp, n_channels, stream);
        if (ru == 7) return launch_slide<float, 28, LLZ_CUDA_ACC_F32>(a, ntp, n_channels, stream);
        if (ru == 5) return launch_slide<float, 20, LLZ_CUDA_ACC_F32>(a, ntp, n_channels, stream);
        if (ru == 3) return launch_slide<float, 12, LLZ_CUDA_ACC_F32>(a, ntp, n_channels, stream);
        return launch_general<float, LLZ_CUDA_ACC_F32>(a, n_channels, stream);
    default:
        llz_set_error("unknown accumulator mode %d", a.acc);
        return -1;
    }
}

}  // namespace llz
