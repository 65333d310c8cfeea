// llz_cuda_common.cuh -- small device/host helpers shared by the kernels of libllzfilter_cuda.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "llz_internal.h"

#define LLZ_STR2(x) #x
#define LLZ_STR(x) LLZ_STR2(x)

#define LLZ_CUDA_TRY(expr)                                                                    \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            llz_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                  \
                          cudaGetErrorString(e__));                                           \
            return -1;                                                                        \
        }                                                                                     \
    } while (0)

namespace llz {

// ---- per-process state (llz_cuda_util.cu) --------------------------------------------------------
// SM count of the CURRENT device, cached per device ordinal (a handle runs under a DeviceGuard, so the current device
// is the handle's); -1 + error message on failure.  Thread-safe.
int device_sm_count();
// Every launch site of the resampler kernels reports its kernel here (thread-local); the shim hands the name and the
// launch count of a handle's last call to llz_cuda_resample_bank_last_run (bench.py's gpu_launches and roofline.kernel).
void note_launch(const char *kernel, int launches = 1);
void note_reset();
const char *noted_kernel();
int noted_launches();

// Measurement knobs.  The environment is read ONCE, when the first handle is created; llz_cuda_tune() overrides a value
// afterwards.  Nothing on a launch path calls getenv.
struct Tunables {
    double pipe_slot_mib;   // staging-slot size of the *_run_host pipelines (LLZ_PIPE_SLOT_MB), default 64
    int slide_ru;           // force a tile variant of the sliding kernel: 11, 7, 5, 3; 0 = automatic (LLZ_SLIDE_RU)
    int fft8k_skew, fft16k_skew;   // warp-group skew of the 8192- / 16384-point kernels in cycles; < 0 = measured default
    int umma_knife_cycles;  // extra cost of a phase tile with knife-edge outputs in the walk's shares (tuning)
    int umma_band_mib;      // samples per band of the tcgen05 kernel's tile walk (tuning)
    double umma_slab_mib;   // expanded-row workspace per slab of the tcgen05 phase-bank kernel (LLZ_UMMA_SLAB_MB)
    int fir_algo;           // process default for LLZ_CUDA_FIR_ALGO_AUTO banks: 0 auto, 1 direct, 2 overlap-save (LLZ_FIR_ALGO)
};
Tunables &tunables();

// ---- 16-byte vector views -------------------------------------------------------------------
template <typename T> struct Vec16;
template <> struct Vec16<float>  { using type = float4;  static constexpr int N = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; };

__device__ __forceinline__ void unpack(const float4 &v, float *dst)
{
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
}
__device__ __forceinline__ void unpack(const double2 &v, double *dst)
{
    dst[0] = v.x; dst[1] = v.y;
}
__device__ __forceinline__ float4 pack(const float *s) { return make_float4(s[0], s[1], s[2], s[3]); }
__device__ __forceinline__ double2 pack(const double *s) { return make_double2(s[0], s[1]); }

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    // make the initialised barrier visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// plain arrival (release, CTA scope): publishes this thread's earlier shared-memory writes to the waiters
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// global -> shared bulk copy; dst, src and bytes are multiples of 16
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- arithmetic policies ----------------------------------------------------------------------
// FMA: one rounding, any order is acceptable (tolerance / guard checked elsewhere).
// STRICT: the reference's "y += h*x" compiled without contraction: round(product) then round(sum).
template <typename T, bool STRICT>
__device__ __forceinline__ T mac(T h, T x, T acc)
{
    if constexpr (STRICT) {
        if constexpr (sizeof(T) == 8) return __dadd_rn(acc, __dmul_rn(h, x));
        else                          return __fadd_rn(acc, __fmul_rn(h, x));
    } else {
        if constexpr (sizeof(T) == 8) return fma(h, x, acc);
        else                          return fmaf(h, x, acc);
    }
}

}  // namespace llz
