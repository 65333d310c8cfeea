// probe_l2_dsmem.cu -- what an SM can move per clock on B200 OUTSIDE its own shared-memory crossbar, the question behind
// the layout of a 16384-point overlap-save item (llz_cuda_fir_fft16k*.cu): 16384 complex doubles are 256 KB, twice
// what one SM's exchange buffer holds, so half of every exchange has to travel either through L2 (a private scratch of
// the CTA: st.global + ld.global.cg, never DRAM) or through the partner CTA of a cluster (st.shared::cluster /
// ld.shared::cluster).  One CTA of 256 threads per SM, 16-byte accesses, 512 contiguous bytes per warp instruction,
// 32 accesses per thread and round (= 128 KB per CTA and direction, the size of one exchange).
//   mode 0: write 128 KB to the CTA's scratch, __syncthreads, read it back (another thread's rows)        [L2 spill]
//   mode 1: read 128 KB of a table that every CTA shares (ld.global.nc)                                   [spectrum]
//   mode 2: mode 0 and mode 1 together                                                                    [both]
//   mode 3: cluster of two: push 64 KB into the partner's shared memory, cluster barrier, pull 64 KB back [DSMEM]
//   mode 4: mode 3 with shared-memory traffic of the CTA's own (128 KB written + read) beside it          [DSMEM + local]
//   mode 5: mode 0 with 128 KB ... 1 MB of scratch per CTA (19 ... 148 MB over the chip)                      [L2 capacity]
// Prints bytes per clock and SM (clock64 of CTA 0) and GB/s over the whole chip (CUDA events).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_l2_dsmem probe_l2_dsmem.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int kThreads = 256, kPer = 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 ld_cg(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_sync_ra()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) l2_kernel(double2 *scratch, const double2 *table, int rounds, long long *clk, double *sink)
{
    double2 *mine = scratch + (size_t)blockIdx.x * (kThreads * kPer);
    const int tid = threadIdx.x, other = (tid + 96) & (kThreads - 1);
    double acc = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        double2 v[kPer];
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < kPer; ++i) mine[i * kThreads + tid] = make_double2(acc + i, r);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < kPer; ++i) v[i] = ld_cg(mine + i * kThreads + other);
#pragma unroll
            for (int i = 0; i < kPer; ++i) acc += v[i].x;
            __syncthreads();
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < kPer; ++i) v[i] = __ldg(table + ((i + r) & 31) * kThreads + tid);
#pragma unroll
            for (int i = 0; i < kPer; ++i) acc += v[i].y;
        }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) *clk = t1 - t0;
    if (acc == 12345.678) *sink = acc;
}

// mode 5: the scratch of mode 0 at growing size -- CHUNKS x 128 KB per CTA are written, then read back, per round: where
// does the chip stop holding the CTAs' private scratch in L2?
__global__ void __launch_bounds__(kThreads, 1) footprint_kernel(double2 *scratch, int chunks, int rounds, long long *clk, double *sink)
{
    double2 *mine = scratch + (size_t)blockIdx.x * chunks * (kThreads * kPer);
    const int tid = threadIdx.x, other = (tid + 96) & (kThreads - 1);
    double acc = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        for (int c = 0; c < chunks; ++c) {
            double2 *blk = mine + (size_t)c * (kThreads * kPer);
#pragma unroll
            for (int i = 0; i < kPer; ++i) blk[i * kThreads + tid] = make_double2(acc + i, r);
        }
        __syncthreads();
        for (int c = 0; c < chunks; ++c) {
            const double2 *blk = mine + (size_t)c * (kThreads * kPer);
            double2 v[kPer];
#pragma unroll
            for (int i = 0; i < kPer; ++i) v[i] = ld_cg(blk + i * kThreads + other);
#pragma unroll
            for (int i = 0; i < kPer; ++i) acc += v[i].x;
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) *clk = t1 - t0;
    if (acc == 12345.678) *sink = acc;
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) dsmem_kernel(int rounds, long long *clk, double *sink)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *buf = reinterpret_cast<double2 *>(smem_raw);                 // 8192 x 16 bytes
    uint32_t rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    uint32_t peer;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer) : "r"(smem_u32(buf)), "r"(rank ^ 1u));
    const int tid = threadIdx.x;
    double acc = 0.0;
    cluster_sync_ra();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        // push 16 x 16 bytes per thread (64 KB per CTA) into the partner's upper half
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t a = peer + (uint32_t)((4096 + i * kThreads + tid) * 16);
            asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(acc + i), "d"((double)r) : "memory");
        }
        if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) buf[i * kThreads + tid] = make_double2(acc, i);
        }
        cluster_sync_ra();
        double2 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t a = peer + (uint32_t)((4096 + i * kThreads + ((tid + 96) & 255)) * 16);
            asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v[i].x), "=d"(v[i].y) : "r"(a) : "memory");
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i].x;
        if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc += buf[i * kThreads + ((tid + 32) & 255)].y;
        }
        cluster_sync_ra();
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) *clk = t1 - t0;
    if (acc == 12345.678) *sink = acc;
}

int main()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double2 *scratch, *table;
    long long *clk;
    double *sink;
    CK(cudaMalloc(&scratch, (size_t)sms * kThreads * kPer * sizeof(double2)));
    CK(cudaMalloc(&table, (size_t)kThreads * kPer * sizeof(double2)));
    CK(cudaMemset(scratch, 0, (size_t)sms * kThreads * kPer * sizeof(double2)));
    CK(cudaMemset(table, 0, (size_t)kThreads * kPer * sizeof(double2)));
    CK(cudaMalloc(&clk, sizeof(long long)));
    CK(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int rounds = 2000;
    const double per_dir = (double)kThreads * kPer * 16;       // bytes per CTA, round and direction
    for (int mode = 0; mode < 5; ++mode) {
        float ms = 0.f;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            switch (mode) {
            case 0: l2_kernel<0><<<sms, kThreads>>>(scratch, table, rounds, clk, sink); break;
            case 1: l2_kernel<1><<<sms, kThreads>>>(scratch, table, rounds, clk, sink); break;
            case 2: l2_kernel<2><<<sms, kThreads>>>(scratch, table, rounds, clk, sink); break;
            case 3:
                CK(cudaFuncSetAttribute(dsmem_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
                dsmem_kernel<3><<<sms & ~1, kThreads, 131072>>>(rounds, clk, sink);
                break;
            default:
                CK(cudaFuncSetAttribute(dsmem_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
                dsmem_kernel<4><<<sms & ~1, kThreads, 131072>>>(rounds, clk, sink);
                break;
            }
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        long long c = 0;
        CK(cudaMemcpy(&c, clk, sizeof(c), cudaMemcpyDeviceToHost));
        double bytes;                                           // per CTA and round, outside the CTA's own crossbar
        const char *what;
        switch (mode) {
        case 0: bytes = 2 * per_dir; what = "L2 scratch: 128 KB written + 128 KB read back"; break;
        case 1: bytes = per_dir; what = "L2 shared table: 128 KB read"; break;
        case 2: bytes = 3 * per_dir; what = "scratch write + read + table read"; break;
        case 3: bytes = per_dir; what = "DSMEM: 64 KB pushed + 64 KB pulled"; break;
        default: bytes = per_dir; what = "DSMEM 64 KB + 64 KB beside 64 KB + 64 KB of own shared-memory traffic"; break;
        }
        printf("mode %d  %-72s %7.1f B/clk/SM  %8.1f GB/s chip  (%.0f clk per round, %.3f ms)\n", mode, what,
               bytes * rounds / (double)c, bytes * rounds * sms / (ms * 1e6), (double)c / rounds, ms);
    }
    // mode 5: footprint sweep
    double2 *big;
    const int max_chunks = 8;
    CK(cudaMalloc(&big, (size_t)sms * max_chunks * kThreads * kPer * sizeof(double2)));
    CK(cudaMemset(big, 0, (size_t)sms * max_chunks * kThreads * kPer * sizeof(double2)));
    for (int chunks = 1; chunks <= max_chunks; chunks += (chunks < 4 ? 1 : 2)) {
        float ms = 0.f;
        const int rr = 400;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            footprint_kernel<<<sms, kThreads>>>(big, chunks, rr, clk, sink);
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        long long c = 0;
        CK(cudaMemcpy(&c, clk, sizeof(c), cudaMemcpyDeviceToHost));
        const double bytes = 2.0 * chunks * per_dir;
        printf("mode 5  scratch %4d KB per CTA = %6.1f MB over the chip, written + read back   %7.1f B/clk/SM  %8.1f GB/s chip\n",
               chunks * 128, chunks * 128.0 * sms / 1024, bytes * rr / (double)c, bytes * rr * sms / (ms * 1e6));
    }
    return 0;
}
