// llz_cuda_iir.cu -- direct-form IIR banks: llz_iir_filter (libllzfilter/llz_iir.c:103-145) for n independent channels.
//
//     y[n] = sum_{k<=N} b[k] x[n-k] - sum_{1<=k<=M} a[k] y[n-k]
//
// The feedback makes a channel serial in time and the reference's rounding order (every product and sum rounded on its
// own, b terms first) leaves no freedom to reassociate, so the parallelism is the channels: one thread per channel, one
// warp per 32 channels.  A warp moves tiles of 32 channels x 64 samples between global and shared memory with coalesced
// rows (lanes along time), then every lane walks its own channel's 64 samples out of the tile (pitch 65: conflict-free)
// and writes the outputs in place.  The stream state (the last N inputs and M outputs per channel) lives in device
// memory between calls, so frame-by-frame calls continue the stream exactly like the reference's shift buffers.
#include "llz_cuda_common.cuh"
#include "llz_iir_kernels.h"

namespace llz {

namespace {

constexpr int kIirTile = 64, kIirPitch = kIirTile + 1;

__global__ void __launch_bounds__(32)
iir_bank_kernel(IirLaunch a)
{
    __shared__ double tile[32 * kIirPitch];
    const int lane = threadIdx.x;
    const int c0 = blockIdx.x * 32, ch = c0 + lane;
    const int nch = min(32, a.n_channels - c0);
    const int M = a.M, N = a.N;
    // stream state in registers: xs[i] = x[n - 1 - i], ys[i] = y[n - 1 - i] (newest first)
    double xs[kIirMaxOrder], ys[kIirMaxOrder];
#pragma unroll
    for (int i = 0; i < kIirMaxOrder; ++i) {
        xs[i] = (ch < a.n_channels && i < N) ? a.state[(size_t)ch * a.state_stride + i] : 0.0;
        ys[i] = (ch < a.n_channels && i < M) ? a.state[(size_t)ch * a.state_stride + kIirMaxOrder + i] : 0.0;
    }
    for (long long t0 = 0; t0 < a.n; t0 += kIirTile) {
        const int len = (int)min((long long)kIirTile, a.n - t0);
        // coalesced load: row r = channel c0 + r, lanes along time
        for (int r = 0; r < nch; ++r) {
            const double *src = a.x ? a.x + (size_t)(c0 + r) * a.x_stride + t0 : nullptr;
            for (int t = lane; t < len; t += 32) tile[r * kIirPitch + t] = src ? src[t] : 0.0;
        }
        __syncwarp();
        if (ch < a.n_channels) {
            for (int t = 0; t < len; ++t) {
                const double xin = tile[lane * kIirPitch + t];
                double acc = __dadd_rn(0.0, __dmul_rn(a.b[0], xin));   // y_out = 0.; y_out += b[0] * x  (llz_iir.c:113, :122)
#pragma unroll
                for (int k = 1; k <= kIirMaxOrder; ++k)
                    if (k <= N) acc = __dadd_rn(acc, __dmul_rn(a.b[k], xs[k - 1]));
#pragma unroll
                for (int k = 1; k <= kIirMaxOrder; ++k)
                    if (k <= M) acc = __dsub_rn(acc, __dmul_rn(a.a[k], ys[k - 1]));
#pragma unroll
                for (int i = kIirMaxOrder - 1; i > 0; --i) { xs[i] = xs[i - 1]; ys[i] = ys[i - 1]; }
                xs[0] = xin;
                ys[0] = acc;
                tile[lane * kIirPitch + t] = acc;
            }
        }
        __syncwarp();
        for (int r = 0; r < nch; ++r) {
            double *dst = a.y + (size_t)(c0 + r) * a.y_stride + t0;
            for (int t = lane; t < len; t += 32) dst[t] = tile[r * kIirPitch + t];
        }
        __syncwarp();
    }
    if (ch < a.n_channels) {
#pragma unroll
        for (int i = 0; i < kIirMaxOrder; ++i) {
            a.state[(size_t)ch * a.state_stride + i] = xs[i];
            a.state[(size_t)ch * a.state_stride + kIirMaxOrder + i] = ys[i];
        }
    }
}

}  // namespace

int iir_launch(const IirLaunch &a, cudaStream_t stream)
{
    if (a.n <= 0 || a.n_channels <= 0) return 0;
    iir_bank_kernel<<<(unsigned)((a.n_channels + 31) / 32), 32, 0, stream>>>(a);
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace llz
