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

}  // namespace llz
