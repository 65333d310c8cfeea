// probe_umma_rate.cu -- issue rate of tcgen05.mma on B200 for the shapes the phase-bank kernel uses (operands in shared
// memory, accumulators in TMEM, one CTA per SM, one issuing thread, back-to-back MMAs, tcgen05.commit + mbarrier wait at
// the end).  Prints cycles per MMA and the MAC rate per SM for:
//   kind::i8  M=128, N = 64 / 128 / 256, the kernel's pattern (two A descriptors x five B descriptors -> six accumulators)
//   kind::i8  M=128, one accumulator, one operand pair (pure K loop)
//   kind::f16 M=128, N = 64 / 256 for comparison
// Operand contents are whatever shared memory holds: only timing is measured.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma_rate probe_umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }

template <int KIND>   // 0 = i8, 1 = f16
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc)
{
    if (KIND == 0)
        asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n}\n" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHi) : "memory");
    else
        asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHi) : "memory");
}

template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 db, {%2, %5};\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], db, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void cp_a(uint32_t a_tmem, uint32_t a_lo)
{
    asm volatile("{\n.reg .b64 da;\nmov.b64 da, {%1, %2};\ntcgen05.cp.cta_group::1.128x256b [%0], da;\n}\n" ::"r"(a_tmem), "r"(a_lo), "r"(kDescHi) : "memory");
}

// PATTERN 0: the kernel's (two A, five B, six accumulators of N columns); 1: one accumulator, K loop over 4 steps;
// 2: pattern 0 with the A operand copied to tensor memory first (tcgen05.cp.128x256b, two copies per K step) and read
//    from there by the ten MMAs
template <int KIND, int N, int PATTERN>
__global__ void __launch_bounds__(128, 1) rate_kernel(int rounds, long long *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 3);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    // i8: u8 x s8 -> s32 (bits as in the kernel); f16: f16 x f16 -> f32 (D fmt 1 at bit 4, A/B fmt 0)
    const uint32_t idesc = (KIND == 0 ? ((2u << 4) | (1u << 10)) : (1u << 4)) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = 0, t1 = 0;
    int n_mma = 0;
    if (warp == 0) {
        uint32_t el;
        asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(el));
        const uint32_t sa = smem_u32(smem);
        const uint32_t a0 = desc_lo(sa), a1 = desc_lo(sa + 16384), b0 = desc_lo(sa + 32768);
        constexpr uint32_t kBPlane = (uint32_t)N * 128 >> 4;              // N rows x 128 bytes, >> 4
        constexpr int NACC = (PATTERN == 0) ? (512 / N < 6 ? 512 / N : 6) : 1;
        t0 = clock64();
        if (el) {
            for (int r = 0; r < rounds; ++r) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ko = 2 * ks;
                    if (PATTERN == 2) {
                        const uint32_t ta = tmem + 6 * N + 16 * (ks & 1);          // two K-step slots of 16 columns
                        cp_a(ta, a0 + ko);
                        cp_a(ta + 8, a1 + ko);
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            const uint32_t bi = b0 + (uint32_t)(i % 3) * kBPlane + ko;
                            mma_ts<KIND>(tmem + N * i, ta, bi, idesc, 1u);
                            mma_ts<KIND>(tmem + N * (i + 1), ta + 8, bi, idesc, 1u);
                        }
                    } else if (PATTERN == 0) {
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            const uint32_t bi = b0 + (uint32_t)(i % 3) * kBPlane + ko;
                            mma<KIND>(tmem + N * (i % NACC), a0 + ko, bi, idesc, 1u);
                            mma<KIND>(tmem + N * ((i + 1) % NACC), a1 + ko, bi, idesc, 1u);
                        }
                    } else {
                        mma<KIND>(tmem, a0 + ko, b0 + ko, idesc, 1u);
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        t1 = clock64();
        n_mma = rounds * 4 * (PATTERN == 1 ? 1 : 10);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = n_mma; }
}

template <int KIND, int N, int PATTERN>
int run(const char *name, int grid, long long *d_out)
{
    auto k = rate_kernel<KIND, N, PATTERN>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    const int rounds = 200;
    for (int it = 0; it < 2; ++it) k<<<grid, 128, 160 * 1024>>>(rounds, d_out);
    CK(cudaDeviceSynchronize());
    long long h[2 * 148];
    CK(cudaMemcpy(h, d_out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int i = 0; i < grid; ++i) if ((double)h[2 * i] > worst) worst = (double)h[2 * i];
    const double per = worst / (double)h[1];
    const double kel = KIND == 0 ? 32.0 : 16.0;
    printf("%-44s grid %3d: %7.1f cycles per MMA  (%6.0f MAC/clk/SM; floor 128*N/256 = %d)\n", name, grid, per, 128.0 * N * kel / per, 128 * N / 256);
    return 0;
}

int main()
{
    long long *d_out;
    CK(cudaMalloc(&d_out, sizeof(long long) * 2 * 148));
    for (int grid : {1, 148}) {
        if (run<0, 64, 0>("i8  N=64  kernel pattern (2 A x 5 B -> 6 acc)", grid, d_out)) return 1;
        if (run<0, 64, 2>("i8  N=64  kernel pattern, A copied to TMEM", grid, d_out)) return 1;
        if (run<0, 64, 1>("i8  N=64  one accumulator, K loop", grid, d_out)) return 1;
        if (run<0, 128, 0>("i8  N=128 pattern (4 acc)", grid, d_out)) return 1;
        if (run<0, 128, 1>("i8  N=128 one accumulator", grid, d_out)) return 1;
        if (run<0, 256, 1>("i8  N=256 one accumulator", grid, d_out)) return 1;
        if (run<1, 64, 1>("f16 N=64  one accumulator", grid, d_out)) return 1;
        if (run<1, 256, 1>("f16 N=256 one accumulator", grid, d_out)) return 1;
    }
    printf("PROBE_UMMA_RATE_DONE\n");
    return 0;
}
