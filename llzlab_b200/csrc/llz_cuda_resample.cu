// llz_cuda_resample.cu -- polyphase decimate / interpolate / rational L/M resample for sm_100a.
//
// Replaces the per-frame loops of the reference (libllzfilter/llz_resample.c:425-491 decimate,
// :494-541 interp, :544-609 resample) by whole-signal kernels over the canonical form described in
// llz_poly_kernels.h.  int16 PCM in, int16 PCM out; phase (o % L) and input index
// (floor(o*M/L)) are integer arithmetic, 64-bit, exactly the reference's sequence.
//
// Arithmetic modes
//   ACC_F64         FP64 FMA in any order, then a guard: if gain*sum lies within `thr` of an integer
//                   (thr = rounding-error bound of the two summation orders for this tile's peak
//                   sample), the output is recomputed in the reference's order with separate
//                   multiply and add.  Rows with a single non-zero tap (the Nyquist phase whose
//                   centre tap is 1-2^-53: SURVEY.md "truncation knife-edge") are exact in any order
//                   and skip the guard.  Result: bit-identical int16, proven rather than observed.
//   ACC_F64_STRICT  every output in the reference's order (verification mode).
//   ACC_F32         FP32 FMA; single-tap rows are evaluated in FP64 so the knife-edge phase stays exact.
//
// Kernels
//   poly_general_kernel  any L, M: one output per thread iteration, input span staged in shared
//                        memory already converted to the accumulator type, bank read transposed
//                        ([tap][phase]) so a warp's loads are contiguous.
//   poly_slide_kernel    L == 1 (decimation): taps split by residue k mod M into M sliding FIRs over
//                        the M de-interleaved input streams; register-blocked SlidingMac core shared
//                        with the FIR kernel (R outputs per thread, 2 LDS.128 per R*U FMAs).
#include <limits.h>

#include "llz_poly_kernels.h"
#include "llz_sliding_mac.cuh"

namespace llz {

// ---- shared device helpers ---------------------------------------------------------------------

__device__ __forceinline__ int poly_sample(const PolyLaunch &a, const int16_t *xc, const int16_t *hc,
                                           long long s)
{
    const long long sp = s - a.in0;
    if (sp >= 0) return (sp < a.n_in && xc) ? (int)xc[sp] : 0;
    if (hc && sp >= -(long long)a.hist_len) return (int)hc[a.hist_len + sp];
    return 0;
}

// gain, saturate, truncate toward zero: llz_resample.c:594-601
__device__ __forceinline__ int16_t poly_finish(double v)
{
    if (v > 32767) v = 32767;
    if (v < -32768) v = -32768;
    return (int16_t)(int)v;
}

// The reference's own loop for one output: order[] walks the taps as llz_resample.c does, products
// and sums rounded separately.  Reads global memory; only the guard and the strict mode call it.
__device__ __noinline__ double poly_reference_order_sum(const PolyLaunch &a, const int16_t *xc,
                                                        const int16_t *hc, long long o)
{
    const int r = (int)(o % a.L);
    const long long base = (o * a.M) / a.L + a.shift;
    long long frame_end = LLONG_MAX;
    if (a.frame_len > 0) frame_end = ((o / a.L) / a.frame_len + 1) * (long long)a.frame_len;
    const double *row = a.cbank + (long long)r * a.ctaps;
    double acc = 0.0;
    for (int t = 0; t < a.order_len; ++t) {
        const int k = a.order[t];
        const long long s = base - k;
        const double xv = (s < frame_end) ? (double)poly_sample(a, xc, hc, s) : 0.0;
        acc = __dadd_rn(acc, __dmul_rn(xv, row[k]));
    }
    return acc;
}

// turn an accumulated sum into the output sample under the selected mode
template <int MODE, typename TA>
__device__ __forceinline__ int16_t poly_emit(const PolyLaunch &a, const int16_t *xc, const int16_t *hc,
                                             long long o, TA acc, double thr, double single_x)
{
    const int r = (int)(o % a.L);
    if constexpr (MODE == LLZ_CUDA_ACC_F32) {
        double v = (double)acc;
        const int st = a.single_tap[r];
        if (st >= 0) v = __dmul_rn(single_x, a.cbank[(long long)r * a.ctaps + st]);
        return poly_finish(__dmul_rn(v, a.gain));
    } else {
        double v = __dmul_rn((double)acc, a.gain);
        if (MODE == LLZ_CUDA_ACC_F64 && thr > 0.0 && fabs(v - rint(v)) < thr && a.single_tap[r] < 0) {
            v = __dmul_rn(poly_reference_order_sum(a, xc, hc, o), a.gain);
            atomicAdd(a.guard_count, 1ULL);
        }
        return poly_finish(v);
    }
}

// block-wide max of |sample| seen while filling a tile (feeds the guard threshold)
__device__ __forceinline__ int block_peak(int local_peak, int *slot)
{
    for (int d = 16; d > 0; d >>= 1) local_peak = max(local_peak, __shfl_xor_sync(0xffffffffu, local_peak, d));
    if ((threadIdx.x & 31) == 0) atomicMax(slot, local_peak);
    __syncthreads();
    return *slot;
}

// ---- general kernel ----------------------------------------------------------------------------

constexpr int kPolyThreads = 256;

template <typename TA, int MODE>
__global__ void __launch_bounds__(kPolyThreads)
poly_general_kernel(PolyLaunch a, int tile_out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TA *xs = reinterpret_cast<TA *>(smem_raw);
    __shared__ int s_peak;
    if (threadIdx.x == 0) s_peak = 0;
    __syncthreads();

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

    int peak = 0;
    for (int e = threadIdx.x; e < span; e += kPolyThreads) {
        const long long s = s_lo + e;
        const int v = (s < frame_end) ? poly_sample(a, xc, hc, s) : 0;
        peak = max(peak, abs(v));
        xs[e] = (TA)v;
    }
    const double thr = (MODE == LLZ_CUDA_ACC_F64) ? a.guard_thr * (double)block_peak(peak, &s_peak) : 0.0;
    if (MODE != LLZ_CUDA_ACC_F64) __syncthreads();

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
template <typename TA, int R, int MODE>
__global__ void __launch_bounds__(kPolyThreads, 1)
poly_slide_kernel(PolyLaunch a, int ntp /* padded taps per residue */)
{
    using SM = SlidingMac<TA, R>;
    constexpr int TILE = kPolyThreads * R;
    constexpr int U = SM::U;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    TA *taps_s = reinterpret_cast<TA *>(smem_raw);            // [M][ntp]
    TA *xs = taps_s + (size_t)a.M * ntp;                      // [M][HS + TILE]
    __shared__ int s_peak;
    if (threadIdx.x == 0) s_peak = 0;

    const int M = a.M;
    const int HS = ntp;
    const int len = HS + TILE;
    const int ch = blockIdx.y;
    const long long qt = a.o0 + (long long)blockIdx.x * TILE;    // canonical index of the tile's first output
    const int16_t *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
    const int16_t *hc = a.hist ? a.hist + (long long)ch * a.hist_len : nullptr;

    const TA *taps_g = (sizeof(TA) == 8) ? reinterpret_cast<const TA *>(a.slide64)
                                         : reinterpret_cast<const TA *>(a.slide32);
    for (int k = threadIdx.x; k < M * ntp; k += kPolyThreads) taps_s[k] = taps_g[k];

    // de-interleave the contiguous input span into the M streams
    const long long s_base = (qt - HS - 1) * M;
    const int total = (len + 1) * M;
    int peak = 0;
    for (int u = threadIdx.x; u < total; u += kPolyThreads) {
        const int jj = u / M, sigma = u - jj * M;
        const int e = jj - 1 + (sigma ? 1 : 0);
        if (e >= 0 && e < len) {
            const int v = poly_sample(a, xc, hc, s_base + u);
            peak = max(peak, abs(v));
            xs[(size_t)sigma * len + e] = (TA)v;
        }
    }
    __syncthreads();
    const double thr = (MODE == LLZ_CUDA_ACC_F64) ? a.guard_thr * (double)block_peak(peak, &s_peak) : 0.0;

    TA acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = TA(0);
    for (int rho = 0; rho < M; ++rho) {
        const int sigma = rho ? M - rho : 0;
        const TA *win = xs + (size_t)sigma * len + HS + threadIdx.x * R - U;
        SM::template run<false>(acc, win, taps_s + (size_t)rho * ntp, ntp);
    }

    const long long q0 = (long long)blockIdx.x * TILE + threadIdx.x * R;   // output index within this call
    int16_t *yc = a.y + (long long)ch * a.y_stride + q0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (q0 + r < a.n_out) {
            double single_x = 0.0;
            if constexpr (MODE == LLZ_CUDA_ACC_F32) {
                const int st = a.single_tap[0];
                if (st >= 0) single_x = (double)poly_sample(a, xc, hc, (a.o0 + q0 + r) * M - st);
            }
            yc[r] = poly_emit<MODE, TA>(a, xc, hc, a.o0 + q0 + r, acc[r], thr, single_x);
        }
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
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- launch selection --------------------------------------------------------------------------

namespace {

constexpr size_t kSmemBudget = 200 * 1024;

template <typename TA, int R>
size_t slide_smem(const PolyLaunch &a, int ntp)
{
    return ((size_t)a.M * ntp + (size_t)a.M * (ntp + kPolyThreads * R)) * sizeof(TA);
}

// 0 = general, 1 = sliding big tile, 2 = sliding small tile
int pick_kernel(const PolyLaunch &a)
{
    if (a.L != 1 || a.shift != 0 || a.frame_len != 0 || a.acc == LLZ_CUDA_ACC_F64_STRICT) return 0;
    if (a.acc == LLZ_CUDA_ACC_F32) {
        if (!a.slide32) return 0;
        if (slide_smem<float, 28>(a, a.slide_ntp32) <= kSmemBudget / 2) return 1;
        if (slide_smem<float, 12>(a, a.slide_ntp32) <= kSmemBudget) return 2;
        return 0;
    }
    if (!a.slide64) return 0;
    if (slide_smem<double, 14>(a, a.slide_ntp64) <= kSmemBudget / 2) return 1;
    if (slide_smem<double, 6>(a, a.slide_ntp64) <= kSmemBudget) return 2;
    return 0;
}

template <typename TA, int R, int MODE>
int launch_slide(const PolyLaunch &a, int ntp, int n_channels, cudaStream_t stream)
{
    constexpr int TILE = kPolyThreads * R;
    const size_t smem = slide_smem<TA, R>(a, ntp);
    auto kern = poly_slide_kernel<TA, R, MODE>;
    if (smem > 48 * 1024)
        LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (a.n_out + TILE - 1) / TILE;
    if (tiles > 0x7fffffffLL) { llz_set_error("resample launch too large"); return -1; }
    kern<<<dim3((unsigned)tiles, (unsigned)n_channels), kPolyThreads, smem, stream>>>(a, ntp);
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
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace

const char *poly_kernel_name(const PolyLaunch &a)
{
    return pick_kernel(a) ? "sliding" : "general";
}

int poly_launch(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    if (a.n_out <= 0 || n_channels <= 0) return 0;
    if (n_channels > 65535) { llz_set_error("too many channels for one launch (%d)", n_channels); return -1; }
    const int which = pick_kernel(a);
    switch (a.acc) {
    case LLZ_CUDA_ACC_F64:
        if (which == 1) return launch_slide<double, 14, LLZ_CUDA_ACC_F64>(a, a.slide_ntp64, n_channels, stream);
        if (which == 2) return launch_slide<double, 6, LLZ_CUDA_ACC_F64>(a, a.slide_ntp64, n_channels, stream);
        return launch_general<double, LLZ_CUDA_ACC_F64>(a, n_channels, stream);
    case LLZ_CUDA_ACC_F64_STRICT:
        return launch_general<double, LLZ_CUDA_ACC_F64_STRICT>(a, n_channels, stream);
    case LLZ_CUDA_ACC_F32:
        if (which == 1) return launch_slide<float, 28, LLZ_CUDA_ACC_F32>(a, a.slide_ntp32, n_channels, stream);
        if (which == 2) return launch_slide<float, 12, LLZ_CUDA_ACC_F32>(a, a.slide_ntp32, n_channels, stream);
        return launch_general<float, LLZ_CUDA_ACC_F32>(a, n_channels, stream);
    default:
        llz_set_error("unknown accumulator mode %d", a.acc);
        return -1;
    }
}

}  // namespace llz
