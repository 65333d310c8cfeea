// llz_iir_kernels.h -- launch interface of the IIR bank kernel (llz_cuda_iir.cu)
#pragma once

#include <cuda_runtime.h>

namespace llz {

constexpr int kIirMaxOrder = 32;

struct IirLaunch {
    int M, N, n_channels;
    double a[kIirMaxOrder + 1], b[kIirMaxOrder + 1];   // a[0] unused (taken as 1, like the reference), b zero-filled beyond N
    const double *x;           // device planar input (nullptr = zeros: the flush)
    long long x_stride;
    double *y;
    long long y_stride;
    long long n;               // samples per channel
    double *state;             // device [channels][state_stride]: the last inputs (newest first), then the last outputs
    int state_stride;          // 2 * kIirMaxOrder
};

int iir_launch(const IirLaunch &a, cudaStream_t stream);

}  // namespace llz
