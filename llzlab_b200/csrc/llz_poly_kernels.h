// llz_poly_kernels.h -- launch interface of the polyphase kernels (llz_cuda_resample.cu).
//
// Every resampler kind is run in the canonical form of llz_internal.h:
//     y[o] = finish( sum_{k<ctaps} cbank[o % L][k] * X( floor(o*M/L) + shift - k ) )
//     finish(a) = (short) clamp(a * gain, -32768, 32767)           (llz_resample.c:594-601)
// X(s) is the stream sample with canonical index s: this call's input x[s - in0] when
// 0 <= s - in0 < n_in, the stored history just before it, zero elsewhere; for interp
// (frame_len > 0) samples at or beyond the end of the output's own input frame are zero.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "llz_cuda_common.cuh"

namespace llz {

struct PolyLaunch {
    const int16_t *x;          // device planar input of this call (nullptr = zeros)
    long long x_stride;
    long long n_in;
    // Interleaved PCM input (llz_cuda_resample_bank_run_pcm; only the tcgen05 path takes it): pcm_frame_bytes > 0 means x
    // points at frames of that many bytes, channel c's sample at byte c * pcm_sample_bytes of a frame, format pcm_fmt
    // (LLZ_CUDA_PCM_*); a channel's base pointer is then (char *)x + c * pcm_sample_bytes and x_stride is unused.
    int pcm_frame_bytes, pcm_sample_bytes, pcm_fmt;
    const int16_t *hist;       // device [channels][hist_len]: samples just before x[0]; may be nullptr
    int hist_len;
    int16_t *y;
    long long y_stride;
    long long o0;              // canonical index of the first output of this call
    long long n_out;
    long long in0;             // canonical index of x[0]
    int L, M, ctaps, shift, frame_len;
    int acc;                   // LLZ_CUDA_ACC_*
    int tiles;                 // LLZ_CUDA_TILES_*: tile family of the phase-bank kernels (verification / A-B runs)
    double gain;
    double guard_thr;          // |v - nearest integer| below this -> reference-order recompute
    const double *cbank;       // [L][ctaps] row-major
    const double *cbankT64;    // [ctaps][L]
    const float *cbankT32;     // [ctaps][L]
    int bank_pad;              // zero rows before row 0 and after row ctaps-1 of cbankT64 / cbankT32 / cbankT16*
    // the bank as two fp16 planes for the tensor-core fast mode: g * 2^bank16_exp = hi + lo (22 significant bits),
    // same [ctaps][L] layout and padding; nullptr when the bank was not split
    const uint16_t *cbankT16h;
    const uint16_t *cbankT16l;
    int bank16_exp;
    const int *order;          // reference accumulation order, order_len canonical tap indices
    int order_len;
    const int *single_tap;     // [L]
    // sliding (L == 1) kernel: per-residue tap rows, [M][slide_ntp], zero padded
    const double *slide64;
    const float *slide32;
    int slide_ntp64, slide_ntp32;
    unsigned long long *guard_count;
    // integer tensor-core exact mode (llz_cuda_polybank_imma.cu): the bank as int8 digit planes in the kernel's tile
    // layout, g ~ q * imma_scale (imma_scale = 2^-s); nullptr when the bank was not split
    const signed char *imma_tiles;
    int imma_nchunks;          // chunks of 64 taps per phase tile
    int imma_planes;           // 5 digit planes
    double imma_scale;
    double imma_thr;           // guard band: guard_thr + |gain| * (tap rounding bound)
    // tcgen05 exact mode (llz_cuda_polybank_umma.cu): the same digit planes in the K-major SWIZZLE_128B layout of
    // llz_umma_tables.h, and the workspace of the call's expanded input rows; nullptr = not selected for this call
    const signed char *umma_tiles;
    const int *umma_weight;    // [phase tiles of the replicated bank]: relative cost of a tile (shares of the walk)
    long long umma_weight_sum;
    int umma_nchunks;          // chunks of 128 k bytes per phase tile
    int umma_planes;           // digit planes of umma_tiles: 5 (exact mode) or 3 (fast mode)
    double umma_scale;         // g * gain ~ q * umma_scale for those digits (the gain is folded into the tables)
    int umma_ush;              // sum >> umma_ush (<< when negative) = the output value in 32.32 fixed point
    unsigned umma_thr32;       // first-level guard band in units of 2^-32 (exact mode)
    unsigned char *umma_rows;  // workspace (byte planes of a slab): poly_bank_umma_rows_bytes(a, channels, umma_slab_cycles) bytes
    int umma_slab_cycles;      // cycles per slab (multiple of 128)
};

// picks the kernel (sliding for L == 1 when the tile fits, general otherwise) and launches it
int poly_launch(const PolyLaunch &a, int n_channels, cudaStream_t stream);
// register-tiled phase-bank kernel for L > 1 (llz_cuda_polybank.cu): 1 = launched, 0 = not applicable, -1 = error
int poly_bank_launch(const PolyLaunch &a, int n_channels, cudaStream_t stream);
// exact mode on the integer tensor cores (llz_cuda_polybank_imma.cu): same return convention
int poly_bank_imma_launch(const PolyLaunch &a, int n_channels, cudaStream_t stream);
// exact mode on tcgen05 / TMEM (llz_cuda_polybank_umma.cu): same return convention; the workspace must hold
// poly_bank_umma_rows_bytes(a, n_channels, a.umma_slab_cycles) bytes
int poly_bank_umma_launch(const PolyLaunch &a, int n_channels, cudaStream_t stream);
size_t poly_bank_umma_rows_bytes(const PolyLaunch &a, int n_channels, long long cycles);
// name of the kernel poly_launch would pick ("sliding" / "general"), for reporting
const char *poly_kernel_name(const PolyLaunch &a);

// new history = last hist_len samples of (old history ++ x[0..n_in))
int poly_update_history(const int16_t *x, long long x_stride, long long n_in, const int16_t *hist_old,
                        int16_t *hist_new, int hist_len, int n_channels, cudaStream_t stream);

}  // namespace llz
