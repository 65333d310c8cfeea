// llz_fir_kernels.h -- launch interface of the FIR kernels (llz_cuda_fir.cu), used by the shim.
#pragma once

#include <cuda_runtime.h>

#include "llz_cuda_common.cuh"

namespace llz {

template <typename T>
struct FirLaunch {
    const T *x;            // device, planar; nullptr = all-zero input (flush)
    long long x_stride;    // elements between channels
    T *y;
    long long y_stride;
    long long n;           // samples per channel in this call
    const T *hist;         // device [channels][ntaps-1]: samples preceding x[0]; nullptr = zeros
    const T *taps;         // device, zero-padded to a multiple of 32 elements
    int ntaps;             // N
    int ntaps_pad;         // filled in by fir_launch
    int vec_ok;            // x, y, strides all 16-byte aligned
};

template <typename T>
int fir_launch(FirLaunch<T> a, int n_channels, bool strict, cudaStream_t stream);

template <typename T>
int fir_pad_taps(int ntaps, int *variant);

template <typename T>
int fir_update_history(const T *x, long long x_stride, long long n, const T *hist_old, T *hist_new,
                       int hlen, int n_channels, cudaStream_t stream);

// ---- overlap-save (1024-point FFT) FIR, llz_cuda_fir_fft.cu -----------------------------------------
constexpr int kFirFftMinTapsAuto = 48;   // below this the direct kernel is at least as fast
constexpr int kFirFftMaxTaps = 897;      // leaves B = 1024 - halo >= 128 valid outputs per block
constexpr int kFirFft8kMinTapsAuto = 545;  // from here on the 8192-point kernel wins (profiles/r01_crossover_fft.txt)
constexpr int kFirFft8kMaxTaps = 6145;   // leaves B = 8192 - halo >= 2048
// from here on the 16384-point kernel wins (profiles/r02_crossover_fft16k.txt: f64 140 / 137 Gsamples/s at 3073 taps,
// 115 / 134 at 4095; f32 286 / 283 at 2049, 264 / 274 at 2561)
constexpr int kFirFft16kMinTapsAutoF64 = 3329, kFirFft16kMinTapsAutoF32 = 2305;
constexpr int kFirFft16kMaxTaps = 12289; // leaves B = 16384 - halo >= 4096

template <typename T>
struct FirFftLaunch {
    const T *x;            // device, planar; nullptr = all-zero input (flush)
    long long x_stride;
    T *y;
    long long y_stride;
    long long n;
    const T *hist;         // device [channels][ntaps-1] or nullptr
    int ntaps;
    const T *H;            // device [32][32] complex: spectrum of the taps / 1024, H[k1][k2] = bin k2 + 32*k1
    const T *tw;           // device [16][32] pairs: folded twiddles of dft32_twisted (llz_fft32.cuh)
    const T *tw2, *tw3;    // 8192-point kernel only: [8][16][32] and [4][4][256] pairs (llz_fft32.cuh)
    // filled in by fir_fft_launch: item i of a channel is pair first_pair + i, skipping [gap_start, gap_start + gap_len)
    int halo, B;           // halo = N-1 rounded up to 32; B = 1024 - halo valid outputs per block
    int prefetch;          // L2 prefetch of the warp's next item
    int skew;              // 8192- / 16384-point kernels: cycles by which warps 4..7 of a CTA trail warps 0..3 after the first exchange
    int n_channels;
    long long first_pair, items_per_channel, gap_start, gap_len;
    // The edge items (first / last of every channel) are an independent launch: on `side` (already ordered after the
    // caller's stream by the shim, which also joins it back) they run beside the interior items instead of after them.
    cudaStream_t side;     // nullptr: everything on the caller's stream
    T *scratch;            // 16384-point kernel only: fir_fft16k_scratch_bytes<T>() of device memory owned by the bank
};

template <typename T>
int fir_fft_launch(FirFftLaunch<T> a, int n_channels, cudaStream_t stream);

// 8192-point variant for long filters; a.H is the 8192-bin spectrum in the kernel's [8][32][32] layout
template <typename T>
int fir_fft8k_launch(FirFftLaunch<T> a, int n_channels, cudaStream_t stream);

// 16384-point variant (one CTA per item, two rounds of eight residues, half of the item parked in an L2-resident
// scratch); a.H is [16][32][32], a.tw2 [16][16][32], a.tw3 [2][8][512]
template <typename T>
int fir_fft16k_launch(FirFftLaunch<T> a, int n_channels, cudaStream_t stream);
template <typename T>
size_t fir_fft16k_scratch_bytes(int sm_count);

}  // namespace llz
