// llz_cuda_fir_fft16k.cu -- overlap-save FIR banks for long filters (3329 / 2305 .. 12289 taps, f64 / f32) on sm_100a:
// a 16384-point transform per CTA, run as TWO ROUNDS of eight 1024-point sub-transforms with the half of the item that
// does not fit the SM parked in L2.  Tolerance-mode arithmetic of llz_fir_filter / llz_conv
// (libllzfilter/llz_fir.c:411-426, 547-584).
//
//   y[c][t] = sum_{i<N} h[i] * x[c][t-i]
//
// At 4095 taps the 8192-point kernel (llz_cuda_fir_fft8k.cu) keeps B = 8192 - 4096 outputs per block: half of every
// transform is overlap.  With 16384 points B = 12288 (75 %): 57-60 instead of 80.6 FMA-pipe instructions and two thirds
// of the shared-memory wavefronts per output -- the two pipes that bound these kernels.  16384 complex doubles are
// 256 KB: twice the exchange buffer an SM can hold and, in registers, twice its register file.  Round 1 of this file
// spread the item over a cluster of two CTAs and exchanged through distributed shared memory; tools/probe_l2_dsmem.cu
// measures that path at 16 B/clk/SM (64 KB pushed + 64 KB pulled cost 8.1 k cycles with their cluster barriers)
// against 44 B/clk/SM for a CTA writing 128 KB to a private L2-resident scratch and reading it back, so the cluster is
// gone: ONE CTA of 256 threads owns the item and processes its 16 residues in two rounds.
//
//   * 16384 = 16 x 1024.  Thread tid gathers z[tid + 256 qq + 1024 a] (qq < 4, a < 16; z = x_A + i x_B, two consecutive
//     blocks of one channel) in four half-passes of 16 points -- the loads of half-pass qq + 1 are in flight while
//     half-pass qq runs its DFT-16 over a and pushes residue b, row j = n_lo / 32 (n_lo = tid + 256 qq): b < 8 into
//     slice b of the 128 KB exchange buffer in shared memory, b >= 8 into the CTA's scratch in global memory (written
//     and read by this CTA only: it lives in L2; the input samples are loaded evict_first so that they do not push it
//     out -- with them at normal priority 40 % more bytes went to DRAM than the call's outputs);
//   * round r (0, 1): warp w transforms residue b = w + 8 r.  It finds the residue in its slice (round 0: pushed by the
//     other warps; round 1: a TMA bulk copy from the scratch, issued when round 0 had finished with the slice), runs
//     the 1024-point transform of llz_cuda_fir_fft.cu with the outer twiddle exp(-2 pi i b n_lo / 16384) folded in
//     (warp-uniform part: the 1024-point table at index 2b; the rest merged with the four-step twiddle into a
//     [b][16][32] table whose 8 KB for residue b arrive in the warp's part of shared memory by a bulk copy issued a
//     whole round earlier), multiplies by the spectrum (its 16 KB slice arrives in the warp's own exchange slice by
//     TMA while the second DFT-32 runs), runs the inverse 1024-point transform and writes the result -- round 0 to the
//     scratch, round 1 to its slice.  There is NO CTA-wide barrier between the push and the pull: 2 x 2052 FMA-pipe
//     instructions per thread during which the eight warps drift apart and one warp's exchanges hide behind another's
//     butterflies (70 % of the FP64 issue rate inside the rounds);
//   * after a barrier every thread pulls its 64 points back in four half-passes (b < 8 from the scratch, b >= 8 from
//     shared memory; again one half-pass ahead), runs a DFT-16 per half-pass with the conjugate outer twiddle -- f32:
//     folded in from the [2][8][512] table; f64: powers of the thread's own root computed on the fly (dft16_powers),
//     because this phase is bound by what the SM can pull out of L2, not by the FP64 pipe -- and holds outputs
//     tid + 256 qq + 1024 a: coalesced streaming stores of the rows at or beyond the halo; f64: half-pass 0 of the
//     NEXT item is loaded under the last of these transforms;
//   * three barriers per item; the next item's input span is prefetched into L2 late (under the pull, not an item
//     ahead: what is prefetched competes with the scratch for L2).
//
// WG (warps per group) is a template parameter: the CTA's eight warps can also form two groups of four that own an item
// each and run four rounds, so that one group's gather / pull phases (L2-bound) run beside the other's butterflies.
// Measured and NOT instantiated: the scratch doubles to 76 MB over the chip, which this L2 no longer holds (DRAM traffic
// 2.4 x the call's bytes, 94-98 against 133 Gsamples/s on C5 f64); one group's 38 MB stay resident.
//
// Per thread and item: 4 x 144 (DFT-16) + 2 x (3 x 512 + 388 + 128) + 4 x 208 (folded DFT-16; f64: 4 x 264) = 5512
// (5736) FMA-pipe instructions for 96 outputs at B = 12288: 57.4 (59.8) per output (8192-point kernel: 80.6).
// L2 traffic per item beyond the samples: 128 KB + 128 KB of scratch written and read, 256 KB of spectrum, 128 KB of
// the [b][16][32] table (f32: half of all that, plus 64 KB of the last pass's table).
// Verified on the host by tests/cpu/fft16k_emulate.cpp (same 16 x 1024 algebra, tables and dft16_powers), on the device
// by tests/test_gpu_fir.py.
#include <stdlib.h>

#include "llz_fft32.cuh"
#include "llz_fir_kernels.h"

namespace llz {

template <typename T> struct Cplx16k;
template <> struct Cplx16k<float>  { using type = float2; };
template <> struct Cplx16k<double> { using type = double2; };

constexpr int kFft16kThreads = 256;
constexpr int kFft16kSlice = kFftR * kFftR;             // points of one residue

template <typename T>
struct Fft16kSmem {
    static constexpr size_t bars = 256;                                               // [w]: spectrum slice, [8 + w]: twiddle table, [16 + w]: residue of warp w
    static constexpr size_t tabw = (size_t)kTwistEntries * kFftR * 2 * sizeof(T);
    static constexpr size_t tab2 = (size_t)8 * kTwistEntries * kFftR * 2 * sizeof(T); // the current round's eight residues
    static constexpr size_t xbuf = (size_t)8 * kFft16kSlice * 2 * sizeof(T);          // eight slices of 32 x 32
    static constexpr size_t total = bars + tabw + tab2 + xbuf;
};

// the CTA's scratch: only this CTA reads it, and only after a __syncthreads that follows the writes -- L2 is the point
// of coherence, so the loads bypass L1 (.cg) and the stores are ordinary write-back stores
__device__ __forceinline__ double2 ld_scratch(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float2 ld_scratch(const float2 *p)
{
    float2 v;
    asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}

// input samples are read once (twice where blocks overlap, by a neighbouring CTA at about the same time): evict_first
// in L2, so that they do not push the CTAs' scratch out to DRAM
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double ld_stream(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_stream(const float *p, uint64_t pol)
{
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}

template <typename T>
__device__ __forceinline__ T fir_fft16k_sample(const FirFftLaunch<T> &a, const T *xc, const T *hc, long long g)
{
    if (g >= 0) return (g < a.n && xc) ? __ldg(xc + g) : T(0);
    if (hc && g >= -(long long)(a.ntaps - 1)) return __ldg(hc + (a.ntaps - 1) + g);
    return T(0);
}

// the WG warps of a group meet on the group's own named barrier (id 1 + g); barrier 0 is used once, before the loop
__device__ __forceinline__ void group_sync(int g, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(threads) : "memory");
}

// EDGE = false: interior items (unguarded loads and stores); EDGE = true: first / last items of a channel.
// WG = warps per group: the CTA's eight warps form 8 / WG groups that each own an item and share nothing but the
// 1024-point twiddle table; a group of WG warps transforms the 16 residues in 16 / WG rounds.  Only WG = 8 (one group,
// two rounds) is instantiated: with two groups one group's gather / pull phases would run beside the other's
// butterflies, but their scratch no longer stays in L2 (see the head of the file).
template <typename T, bool EDGE, int WG>
__global__ void __launch_bounds__(kFft16kThreads, sizeof(T) == 4 ? 2 : 1)
fir_fft16k_kernel(FirFftLaunch<T> a)
{
    using C = typename Cplx16k<T>::type;
    using SM = Fft16kSmem<T>;
    constexpr int GROUPS = 8 / WG, GT = 32 * WG;               // groups per CTA, threads per group
    constexpr int NQ = 1024 / GT;                              // n_lo = gtid + GT qq, qq < NQ: half-pass qq holds 16 points
    constexpr int ROUNDS = 16 / WG;                            // round r: warp wg of the group transforms residue wg + WG r
    constexpr int KEPT = 16 - WG;                              // residues whose results travel through the scratch
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    C *tabw_s = reinterpret_cast<C *>(smem_raw + SM::bars);                                  // [16][32]
    C *tab2_s = reinterpret_cast<C *>(smem_raw + SM::bars + SM::tabw);                       // [8 warps][16][32]
    C *xbuf = reinterpret_cast<C *>(smem_raw + SM::bars + SM::tabw + SM::tab2);              // 8 slices of 32 x 32

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = warp / WG, wg = warp % WG, gtid = tid % GT;
    constexpr uint32_t kTab2Bytes = (uint32_t)(kTwistEntries * kFftR * sizeof(C));
    constexpr uint32_t kSliceBytes = (uint32_t)(kFft16kSlice * sizeof(C));
    C *xg = xbuf + g * (WG * kFft16kSlice);                    // the group's WG slices
    C *slice = xg + wg * kFft16kSlice;
    C *tab2_w = tab2_s + warp * (kTwistEntries * kFftR);
    const C *tab2_g = reinterpret_cast<const C *>(a.tw2) + wg * (kTwistEntries * kFftR);     // + WG r residues
    // the group's scratch, 16 blocks of one residue: block b - WG takes residue b >= WG as pushed and, in place, its
    // result (rounds 1 .. ROUNDS - 2; the last round's results stay in shared memory); blocks KEPT .. 15 take the
    // results of round 0, whose input travelled through shared memory
    C *scr = reinterpret_cast<C *>(a.scratch) + ((size_t)blockIdx.x * GROUPS + g) * (16 * kFft16kSlice);

    if (tid == 0)
        for (int b = 0; b < 24; ++b) mbar_init(&bars[b], 1);
    for (int i = tid; i < kTwistEntries * kFftR; i += kFft16kThreads) tabw_s[i] = reinterpret_cast<const C *>(a.tw)[i];
    __syncthreads();
    // the table of the warp's round-0 residue; from here on every use of the table is followed by the fetch of the next one
    if (lane == 0) {
        mbar_expect_tx(&bars[8 + warp], kTab2Bytes);
        tma_bulk_g2s(tab2_w, tab2_g, kTab2Bytes, &bars[8 + warp]);
    }

    const C *tab3 = reinterpret_cast<const C *>(a.tw3) + gtid;                // [q][e][t], n_lo = t + 512 q
    // FP64: the thread's own outer twiddle exp(+2 pi i gtid / 16384), from the table's (cos, tan) entry of that angle
    [[maybe_unused]] T w0r = T(1), w0i = T(0);
    if constexpr (sizeof(T) == 8) {
        const C cw = __ldg(tab3 + 4 * 512);
        w0r = cw.x; w0i = cw.x * cw.y;
    }
    const int hl = a.halo, B = a.B;
    const int zero_below = hl - (a.ntaps - 1);                                // < 512
    const long long total = a.items_per_channel * a.n_channels;
    const long long stride = (long long)gridDim.x * GROUPS;
    const int span_bytes = (kFft16kN + B) * (int)sizeof(T);
    uint32_t h_phase = 0, t_phase = 0, s_phase = 0;
    [[maybe_unused]] const uint64_t pol_stream = l2_policy_evict_first();
    bool first = true;

    T re[32], im[32];
    // interior items: half-pass 0 of an item is loaded ahead of its iteration -- for the first item here, for every
    // other one under the last transform and the stores of the item before it (registers 0..15 are free there)
    [[maybe_unused]] auto item_base = [&](long long it) -> const T * {
        const int c = (int)(it / a.items_per_channel);
        const long long pr = a.first_pair + (it - (long long)c * a.items_per_channel);
        return a.x + (long long)c * a.x_stride + pr * (2LL * B) - hl + gtid;
    };
    [[maybe_unused]] auto load_half0 = [&](const T *src) {
#pragma unroll
        for (int aa = 0; aa < 16; ++aa) {
            re[aa] = ld_stream(src + 1024 * aa, pol_stream);
            im[aa] = ld_stream(src + B + 1024 * aa, pol_stream);
        }
    };
    constexpr bool AHEAD = !EDGE && sizeof(T) == 8;            // measured: f64 +1 %, f32 -2 %
    if constexpr (AHEAD) {
        if ((long long)blockIdx.x * GROUPS + g < total) load_half0(item_base((long long)blockIdx.x * GROUPS + g));
    }

    for (long long item = (long long)blockIdx.x * GROUPS + g; item < total; item += stride) {
        const int ch = (int)(item / a.items_per_channel);
        long long pair = a.first_pair + (item - (long long)ch * a.items_per_channel);
        if constexpr (EDGE) { if (pair >= a.gap_start) pair += a.gap_len; }
        const long long o = pair * (2LL * B);
        const long long s = o - hl;
        const T *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
        T *yc = a.y + (long long)ch * a.y_stride;

        // ---- gather, DFT-16 over a, push.  Half-pass qq (n_lo = gtid + GT qq) lives in registers 16 (qq & 1) + a; the
        // loads of half-pass qq + 1 are issued before the transform and the push of half-pass qq, so that the global
        // latency of one half hides behind the work of the other (all warps of a group are in this phase together)
        [[maybe_unused]] const T *hc = nullptr;
        if constexpr (EDGE) hc = a.hist ? a.hist + (long long)ch * (a.ntaps - 1) : nullptr;
        auto gather_load = [&](int qq) {
            const int O = 16 * (qq & 1);
            if constexpr (!EDGE) {
                const T *src = xc + s + gtid + GT * qq;
#pragma unroll
                for (int aa = 0; aa < 16; ++aa) {
                    re[O + aa] = ld_stream(src + 1024 * aa, pol_stream);
                    im[O + aa] = ld_stream(src + B + 1024 * aa, pol_stream);
                }
            } else {
#pragma unroll
                for (int aa = 0; aa < 16; ++aa) {
                    const long long gg = s + gtid + GT * qq + 1024 * aa;
                    re[O + aa] = fir_fft16k_sample(a, xc, hc, gg);
                    im[O + aa] = fir_fft16k_sample(a, xc, hc, gg + B);
                }
            }
        };
        if constexpr (!AHEAD) gather_load(0);
#pragma unroll
        for (int qq = 0; qq < NQ; ++qq) {
            if (qq + 1 < NQ) gather_load(qq + 1);
            const int O = 16 * (qq & 1);
            // the first halo - (N-1) samples of a block reach only discarded outputs: zeroed, so that every kept output is
            // a function of its own N-1 predecessors alone, bit for bit (see llz_cuda_fir_fft.cu)
            if (gtid + GT * qq < zero_below) { re[O] = T(0); im[O] = T(0); }
            if (qq & 1) dft16<T, false, 16>(re, im); else dft16<T, false, 0>(re, im);
            // the previous item's pulls have read every slice before this item's pushes overwrite them
            if (qq == 0 && !first) group_sync(g, GT);
            const int row = (wg + WG * qq) * kFftR + lane;                   // j = n_lo / 32, column lane
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                C v; v.x = re[O + b]; v.y = im[O + b];
                if (b < WG) xg[b * kFft16kSlice + row] = v;
                else        scr[(b - WG) * kFft16kSlice + row] = v;
            }
        }
        first = false;
        // the residues pushed to the scratch (generic-proxy stores) are read back by bulk copies (async proxy): one proxy
        // fence per warp on its own stores, before the barrier that orders them before every issuing thread -- there the
        // fence's wait for the stores overlaps the wait for the slowest warp
        __syncwarp();
        if (lane == 0) asm volatile("fence.proxy.async;" ::: "memory");
        group_sync(g, GT);
        // one group: warps 4..7 trail their scheduler partners 0..3 by about one transform phase (see
        // llz_cuda_fir_fft8k.cu); two groups are out of phase by themselves
        if (GROUPS == 1 && a.skew > 0 && warp >= 4) {
            const long long t0 = clock64();
            while (clock64() - t0 < a.skew) { }
        }

        // ---- rounds: warp wg transforms residue b = wg + WG r -----------------------------------------------------
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r) {
            const int b = wg + WG * r;
            C *blk = scr + (size_t)(r == 0 ? KEPT + wg : b - WG) * kFft16kSlice + lane;      // result block (and input, r > 0)
            // rounds after the first find their residue in the slice as well: the bulk copy from the scratch was issued
            // when the previous round had finished with the slice
            if (r > 0) {
                mbar_wait(&bars[16 + warp], s_phase);
                s_phase ^= 1;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) { const C v = slice[j * kFftR + lane]; re[j] = v.x; im[j] = v.y; }
            __syncwarp();

            // 1024-point forward transform of residue b with the outer twiddle folded in
            dft32_twisted<T, false>(re, im, tabw_s + 2 * b, kFftR);
#pragma unroll
            for (int k = 0; k < 32; ++k) { C v; v.x = re[k]; v.y = im[k]; slice[lane * kFftR + (k ^ lane)] = v; }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 32; ++k) { const C v = slice[k * kFftR + (lane ^ k)]; re[k] = v.x; im[k] = v.y; }
            __syncwarp();
            // the slice is idle until the next exchange: fetch the residue's 32 x 32 bins of the spectrum into it
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars[warp], kSliceBytes);
                tma_bulk_g2s(slice, reinterpret_cast<const C *>(a.H) + b * kFft16kSlice, kSliceBytes, &bars[warp]);
            }
            mbar_wait(&bars[8 + warp], t_phase);
            t_phase ^= 1;
            dft32_twisted<T, false>(re, im, tab2_w + lane, kFftR);
            __syncwarp();
            // every lane is done with the table: the next round's residue replaces it (used a whole round from now)
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars[8 + warp], kTab2Bytes);
                tma_bulk_g2s(tab2_w, tab2_g + ((r + 1) % ROUNDS) * (WG * kTwistEntries * kFftR), kTab2Bytes, &bars[8 + warp]);
            }

            // spectrum, inverse 1024-point transform
            mbar_wait(&bars[warp], h_phase);
            h_phase ^= 1;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const C h = slice[k * kFftR + lane];
                cmul_inplace<T, false>(re[k], im[k], h.x, h.y);
            }
            __syncwarp();                                  // all lanes are done with the spectrum before the slice is reused
            dft32<T, true>(re, im);
#pragma unroll
            for (int k = 0; k < 32; ++k) { C v; v.x = re[k]; v.y = im[k]; slice[lane * kFftR + (k ^ lane)] = v; }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 32; ++k) { const C v = slice[k * kFftR + (lane ^ k)]; re[k] = v.x; im[k] = v.y; }
            __syncwarp();
            if (r < ROUNDS - 1 && lane == 0) {
                // the slice is free until the next round: its residue comes in while the last DFT-32 and the stores run
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars[16 + warp], kSliceBytes);
                tma_bulk_g2s(slice, scr + (size_t)b * kFft16kSlice, kSliceBytes, &bars[16 + warp]);
            }
            dft32_twisted<T, true>(re, im, tabw_s + lane, kFftR);

            if (r < ROUNDS - 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) { C v; v.x = re[j]; v.y = im[j]; blk[j * kFftR] = v; }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) { C v; v.x = re[j]; v.y = im[j]; slice[j * kFftR + lane] = v; }
            }
        }
        // the next item's input span into L2 -- now, not a whole item ahead: the CTAs' scratch (256 KB per group) has to
        // stay in L2 beside whatever is prefetched, and an item-long lead (229 KB per group) pushed it out to DRAM
        if constexpr (!EDGE) {
            if (a.prefetch && item + stride < total) {
                const long long nit = item + stride;
                const int nch = (int)(nit / a.items_per_channel);
                const long long np = a.first_pair + (nit - (long long)nch * a.items_per_channel);
                const char *src = reinterpret_cast<const char *>(a.x + (long long)nch * a.x_stride + np * (2LL * B) - hl);
                for (int off = gtid * 128; off < span_bytes; off += GT * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(src + off));
            }
        }
        group_sync(g, GT);

        // ---- pull, DFT-16 over b with the conjugate outer twiddle folded in, scatter ---------------------------------
        const int r0 = hl / GT;                            // rows of GT outputs below the halo are overlap
        // half-pass qq again in registers 16 (qq & 1) + b, the loads of qq + 1 issued before the transform of qq
        [[maybe_unused]] C e[2][8];
        auto pull_load = [&](int qq) {
            const int O = 16 * (qq & 1);
            const int row = (wg + WG * qq) * kFftR + lane;
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                C v;
                if (b < WG)        v = ld_scratch(scr + (KEPT + b) * kFft16kSlice + row);
                else if (b < KEPT) v = ld_scratch(scr + (b - WG) * kFft16kSlice + row);
                else               v = xg[(b - KEPT) * kFft16kSlice + row];
                re[O + b] = v.x; im[O + b] = v.y;
            }
            // n_lo = gtid + GT qq = t + 512 q
            if constexpr (sizeof(T) == 4) {
                const C *t3 = tab3 + (size_t)((GT * qq) >> 9) * (8 * 512) + ((GT * qq) & 511);
#pragma unroll
                for (int i = 0; i < 8; ++i) e[qq & 1][i] = __ldg(t3 + i * 512);
            }
        };
        pull_load(0);
#pragma unroll
        for (int qq = 0; qq < NQ; ++qq) {
            if (qq + 1 < NQ) pull_load(qq + 1);
            if constexpr (AHEAD) {
                static_assert(NQ % 2 == 0, "the last half-pass must leave registers 0..15 free");
                if (qq == NQ - 1 && item + stride < total) load_half0(item_base(item + stride));
            }
            const int O = 16 * (qq & 1);
            if constexpr (sizeof(T) == 4) {
                if (qq & 1) dft16_twisted<T, true, 16>(re, im, e[1]); else dft16_twisted<T, true, 0>(re, im, e[0]);
            } else {
                // FP64: this pass is bound by its L2 reads, not by the FP64 pipe -- the conjugate outer twiddle
                // exp(+2 pi i n_lo / 16384) = w0 * exp(2 pi i GT qq / 16384) and its powers are computed, not read
                constexpr double kCos[8] = {1.0, 0.9987954562051724, 0.9951847266721969, 0.989176509964781, 0.9807852804032304, 0.970031253194544, 0.9569403357322088, 0.9415440651830208};
                constexpr double kSin[8] = {0.0, 0.049067674327418015, 0.0980171403295606, 0.14673047445536175, 0.19509032201612825, 0.24298017990326387, 0.29028467725446233, 0.33688985339222005};       // cos, sin (2 pi k / 128)
                const T cr = (T)kCos[GT * qq / 128], ci = (T)kSin[GT * qq / 128];          // compile-time after unrolling
                const T wr = fma(-w0i, ci, w0r * cr), wi = fma(w0i, cr, w0r * ci);
                if (qq & 1) dft16_powers<T, true, 16>(re, im, wr, wi); else dft16_powers<T, true, 0>(re, im, wr, wi);
            }
            T *qy = yc + o - hl + gtid + GT * qq;
#pragma unroll
            for (int aa = 0; aa < 16; ++aa) {
                const int off = 1024 * aa;
                if (qq + NQ * aa >= r0) {
                    if constexpr (!EDGE) {
                        __stcs(qy + off, re[O + aa]);
                        __stcs(qy + B + off, im[O + aa]);
                    } else {
                        const long long tA = o - hl + gtid + GT * qq + off;
                        if (tA < a.n) __stcs(qy + off, re[O + aa]);
                        if (tA + B < a.n) __stcs(qy + B + off, im[O + aa]);
                    }
                }
            }
        }
    }
    // the table fetch issued by the last round is still in flight: it must land before the CTA's shared memory is released
    mbar_wait(&bars[8 + warp], t_phase);
}

template <typename T>
size_t fir_fft16k_scratch_bytes(int sm_count)
{
    // one region of 16 residues x 1024 points per resident CTA, for the interior and for the edge launch (they run
    // side by side on two streams)
    const size_t slots = (size_t)sm_count * (sizeof(T) == 4 ? 2 : 1);
    return 2 * slots * 16 * kFft16kSlice * 2 * sizeof(T);
}

template <typename T, bool EDGE, int WG>
static int fir_fft16k_run(FirFftLaunch<T> b, int n_channels, long long first, long long count, long long gap_start,
                          long long gap_len, int sm_count, cudaStream_t stream)
{
    if (count <= 0) return 0;
    constexpr size_t smem = Fft16kSmem<T>::total;
    static_assert(smem <= 227 * 1024, "16384-point overlap-save kernel exceeds the shared memory of an SM");
    b.first_pair = first;
    b.items_per_channel = count;
    b.gap_start = gap_start;
    b.gap_len = gap_len;
    auto kern = fir_fft16k_kernel<T, EDGE, WG>;
    constexpr int GROUPS = 8 / WG;
    LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long items = count * n_channels;
    const long long ctas = (long long)sm_count * (sizeof(T) == 4 ? 2 : 1), want = (items + GROUPS - 1) / GROUPS;
    const unsigned grid = (unsigned)(want < ctas ? want : ctas);
    if (EDGE) b.scratch = b.scratch + (size_t)ctas * GROUPS * 16 * kFft16kSlice * 2;   // the second half of the scratch
    kern<<<grid, kFft16kThreads, smem, stream>>>(b);
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T>
int fir_fft16k_launch(FirFftLaunch<T> a, int n_channels, cudaStream_t stream)
{
    if (a.n <= 0 || n_channels <= 0) return 0;
    if (a.ntaps < 1 || a.ntaps > kFirFft16kMaxTaps) {
        llz_set_error("16384-point overlap-save FIR kernel takes 1..%d taps, got %d", kFirFft16kMaxTaps, a.ntaps);
        return -1;
    }
    if (!a.scratch) {
        llz_set_error("internal: the 16384-point overlap-save kernel needs its scratch");
        return -1;
    }
    a.halo = (a.ntaps - 1 + 511) / 512 * 512;
    a.B = kFft16kN - a.halo;
    a.n_channels = n_channels;
    const long long two_b = 2LL * a.B;
    const long long ppc = (a.n + two_b - 1) / two_b;
    long long p_lo = (a.halo + two_b - 1) / two_b, p_hi = a.n / two_b;
    if (!a.x || p_hi < p_lo) { p_lo = 0; p_hi = 0; }
    const int sm_count = device_sm_count();
    if (sm_count <= 0) return -1;
    a.prefetch = 1;
    const int sk = tunables().fft16k_skew;                     // llz_cuda_tune("fft16k_skew", cycles); < 0: the measured default
    a.skew = sk >= 0 ? sk : (sizeof(T) == 8 ? 1300 : 0);
    cudaStream_t es = a.side ? a.side : stream;
    if (fir_fft16k_run<T, false, 8>(a, n_channels, p_lo, p_hi - p_lo, ppc, 0, sm_count, stream) != 0) return -1;
    return fir_fft16k_run<T, true, 8>(a, n_channels, 0, ppc - (p_hi - p_lo), p_lo, p_hi - p_lo, sm_count, es);
}

template int fir_fft16k_launch<float>(FirFftLaunch<float>, int, cudaStream_t);
template int fir_fft16k_launch<double>(FirFftLaunch<double>, int, cudaStream_t);
template size_t fir_fft16k_scratch_bytes<float>(int);
template size_t fir_fft16k_scratch_bytes<double>(int);

}  // namespace llz
