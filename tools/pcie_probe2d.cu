// pcie_probe2d.cu -- does the host pipeline's row-wise (2-D) copy shape cost PCIe bandwidth?  (not part of the library)
//   nvcc -O3 -o tools/pcie_probe2d tools/pcie_probe2d.cu
// Simultaneous H2D + D2H of the same volume, as contiguous 1-D copies and as 1024 rows of 28 KB out of a
// 3.84 MB pitch (the C2 chunk shape of llz_cuda_fir_bank_run_host).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main()
{
    const size_t rows = 1024, pitch = 480000 * 8, width = 3584 * 8, chunks = 64;
    const size_t total = rows * pitch;
    char *h_in, *h_out, *d_in, *d_out;
    CK(cudaHostAlloc(&h_in, total, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_out, total, cudaHostAllocDefault));
    CK(cudaMalloc(&d_in, rows * width * 3));
    CK(cudaMalloc(&d_out, rows * width * 3));
    cudaStream_t s1, s2;
    CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0, s1);
            cudaStreamWaitEvent(s2, e0, 0);
            for (size_t c = 0; c < chunks; ++c) {
                char *di = d_in + (c % 3) * rows * width, *dO = d_out + (c % 3) * rows * width;
                if (mode == 0) {
                    CK(cudaMemcpyAsync(di, h_in + c * rows * width, rows * width, cudaMemcpyHostToDevice, s1));
                    CK(cudaMemcpyAsync(h_out + c * rows * width, dO, rows * width, cudaMemcpyDeviceToHost, s2));
                } else {
                    CK(cudaMemcpy2DAsync(di, width, h_in + c * width, pitch, width, rows, cudaMemcpyHostToDevice, s1));
                    CK(cudaMemcpy2DAsync(h_out + c * width, pitch, dO, width, width, rows, cudaMemcpyDeviceToHost, s2));
                }
            }
            cudaEventRecord(e1, s1);
            cudaEventRecord(e2, s2);
            CK(cudaDeviceSynchronize());
            float m1, m2;
            cudaEventElapsedTime(&m1, e0, e1);
            cudaEventElapsedTime(&m2, e0, e2);
            const double gb = chunks * rows * width / 1e9;
            if (rep) printf("%s: H2D %.1f GB/s, D2H %.1f GB/s (simultaneous, %.2f GB each way)\n",
                            mode ? "2-D rows of 28 KB, pitch 3.84 MB" : "contiguous 29 MB copies         ", gb / m1 * 1e3, gb / m2 * 1e3, gb);
        }
    }
    return 0;
}
