// llz_f32x2.cuh -- two FP32 values in one 64-bit register, computed with Blackwell's packed FP32 instructions
// (PTX add/sub/mul/fma .f32x2, SASS FADD2 / FMUL2 / FFMA2).
//
// FFMA2 runs at the scalar FFMA rate (74.0 against 72.2 TFLOP/s, profiles/r01_probe_pipes.txt): the packed form does
// not raise the FP32 roof, it halves the ISSUE slots the same arithmetic needs.  The scalar-float overlap-save kernel
// is issue-bound (75 % of issue slots, FP32 pipe 58 %, profiles/r01_c2_f32_fft_ncu_full.txt); with F2 as the
// arithmetic type of the same templates (llz_fft32.cuh) a warp carries two work items in the two halves, every
// butterfly serves both, and table reads (twiddles, spectrum) are shared.  A scalar constant becomes a 32-bit
// immediate that the hardware applies to both halves; each half is an IEEE fma / add / mul, so results are
// bit-identical to the scalar kernel's.
#pragma once

#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define LLZ_F2_HD __host__ __device__ __forceinline__
#else
#define LLZ_F2_HD inline
#endif

namespace llz {

struct F2;
LLZ_F2_HD F2 f2_pack(float lo, float hi);
LLZ_F2_HD void f2_unpack(F2 a, float &lo, float &hi);

struct F2 {
    unsigned long long v;
    F2() = default;
    LLZ_F2_HD explicit F2(double x)
    {
        const float f = (float)x;
        unsigned u;
#if defined(__CUDA_ARCH__)
        u = __float_as_uint(f);
#else
        memcpy(&u, &f, 4);
#endif
        v = ((unsigned long long)u << 32) | u;
    }
    // sign flip of both halves (two LOP3 on the otherwise idle ALU pipe); only used on table values, once per value
    LLZ_F2_HD F2 operator-() const
    {
        F2 r;
        r.v = v ^ 0x8000000080000000ull;
        return r;
    }
};

LLZ_F2_HD F2 f2_pack(float lo, float hi)
{
    F2 r;
#if defined(__CUDA_ARCH__)
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
#else
    unsigned a, b;
    memcpy(&a, &lo, 4);
    memcpy(&b, &hi, 4);
    r.v = ((unsigned long long)b << 32) | a;
#endif
    return r;
}
LLZ_F2_HD void f2_unpack(F2 a, float &lo, float &hi)
{
#if defined(__CUDA_ARCH__)
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
#else
    const unsigned x = (unsigned)a.v, y = (unsigned)(a.v >> 32);
    memcpy(&lo, &x, 4);
    memcpy(&hi, &y, 4);
#endif
}

// device: one packed instruction; host (tests/cpu): the same IEEE operation on each half
#if defined(__CUDA_ARCH__)
#define LLZ_F2_OP2(name, ptx, expr)                                                       \
    LLZ_F2_HD F2 name(F2 a, F2 b)                                                         \
    {                                                                                     \
        F2 d;                                                                             \
        asm(ptx " %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));                          \
        return d;                                                                         \
    }
#else
#define LLZ_F2_OP2(name, ptx, expr)                                                       \
    LLZ_F2_HD F2 name(F2 a, F2 b)                                                         \
    {                                                                                     \
        float x0, x1, y0, y1;                                                             \
        f2_unpack(a, x0, x1);                                                             \
        f2_unpack(b, y0, y1);                                                             \
        return f2_pack(x0 expr y0, x1 expr y1);                                           \
    }
#endif
LLZ_F2_OP2(operator+, "add.rn.f32x2", +)
LLZ_F2_OP2(operator-, "sub.rn.f32x2", -)
LLZ_F2_OP2(operator*, "mul.rn.f32x2", *)
#undef LLZ_F2_OP2

LLZ_F2_HD F2 f2_fma(F2 a, F2 b, F2 c)
{
#if defined(__CUDA_ARCH__)
    F2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return d;
#else
    float a0, a1, b0, b1, c0, c1;
    f2_unpack(a, a0, a1);
    f2_unpack(b, b0, b1);
    f2_unpack(c, c0, c1);
    return f2_pack(fmaf(a0, b0, c0), fmaf(a1, b1, c1));
#endif
}

struct alignas(16) F2x2 { F2 x, y; };        // a (re, im) or (cos, tan) pair: 16 bytes like double2

}  // namespace llz
