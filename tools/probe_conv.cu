// probe_conv.cu -- issue cost (cycles per warp instruction per SM sub-partition) of the FP64 and conversion instructions the
// resampler epilogues use, measured with 2 warps per scheduler (8 warps per CTA, one CTA per SM) and 8 independent chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_conv probe_conv.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int OP>
__global__ void __launch_bounds__(256, 1) k(long long *out, double seed, int iters)
{
    double d[8];
    long long q[8];
    int n[8];
    for (int i = 0; i < 8; ++i) { d[i] = seed + i + threadIdx.x; q[i] = (long long)(seed * 1e6) + i * 977 + threadIdx.x; n[i] = (int)q[i]; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) d[i] = __dadd_rn(d[i], seed);                              // DADD
            if (OP == 1) d[i] = __dmul_rn(d[i], seed);                              // DMUL
            if (OP == 2) d[i] = fma(d[i], seed, seed);                              // DFMA
            if (OP == 3) { q[i] = (long long)__double_as_longlong((double)q[i]) ^ it; }     // I2F.F64.S64
            if (OP == 4) { n[i] = __double2hiint((double)n[i]) ^ it; }              // I2F.F64.S32
            if (OP == 5) { n[i] = __double2int_rn(__hiloint2double(0x40100000 | (n[i] & 0xfffff), n[i])) ^ it; }   // F2I.F64 (rn)
            if (OP == 6) { n[i] = __double2int_rz(__hiloint2double(0x40100000 | (n[i] & 0xfffff), n[i])) ^ it; }   // F2I.F64.TRUNC
            if (OP == 7) { q[i] = q[i] * 0x10001LL + n[i]; }                        // 64-bit multiply-add (IMAD.WIDE pair)
            if (OP == 8) { n[i] = (d[i] > seed) ? n[i] + 1 : n[i] ^ it; d[i] = __longlong_as_double(__double_as_longlong(d[i]) ^ 1); }   // DSETP + select
        }
    }
    const long long t1 = clock64();
    double acc = 0; long long qa = 0;
    for (int i = 0; i < 8; ++i) { acc += d[i]; qa += q[i] + n[i]; }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 1.2345 && qa == 77) out[0] = 0;
}

template <int OP> void run(const char *name, long long *d_out)
{
    const int iters = 2000;
    k<OP><<<148, 256>>>(d_out, 1.000001, iters);
    k<OP><<<148, 256>>>(d_out, 1.000001, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < 148; ++i) if ((double)h[i] > worst) worst = (double)h[i];
    // 2 warps per scheduler, 8 chain steps per iteration
    printf("%-28s %6.2f cycles per warp instruction per scheduler (incl. loop overhead)\n", name, worst / (iters * 8.0 * 2.0));
}

int main()
{
    long long *d_out;
    cudaMalloc(&d_out, sizeof(long long) * 148);
    run<0>("DADD", d_out); run<1>("DMUL", d_out); run<2>("DFMA", d_out); run<3>("I2F.F64.S64 (+xor)", d_out);
    run<4>("I2F.F64.S32 (+xor)", d_out); run<5>("F2I.F64 rn (+3 int)", d_out); run<6>("F2I.F64.TRUNC (+3 int)", d_out);
    run<7>("64-bit mul-add", d_out); run<8>("DSETP + select (+2 int)", d_out);
    printf("PROBE_CONV_DONE\n");
    return 0;
}
