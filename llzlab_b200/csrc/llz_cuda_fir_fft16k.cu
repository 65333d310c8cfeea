// llz_cuda_fir_fft16k.cu -- overlap-save FIR banks for very long filters (4609 .. 12289 taps) on sm_100a: a
// 16384-point transform per CLUSTER OF TWO CTAs, exchanged through distributed shared memory.
// Tolerance-mode arithmetic of llz_fir_filter / llz_conv (libllzfilter/llz_fir.c:411-426, 547-584).
//
//   y[c][t] = sum_{i<N} h[i] * x[c][t-i]
//
// At 4095 taps the 8192-point kernel (llz_cuda_fir_fft8k.cu) keeps B = 8192 - 4096 outputs per block: half of
// every transform is overlap.  A longer transform does not fit one SM in FP64 -- 16384 complex doubles are 256 KB and
// 512 threads x 254 registers are twice the register file -- but it fits TWO: a thread-block cluster of 2 CTAs x 256
// threads holds the 16384 points in the registers of 512 threads, and the two CTAs' 128 KB exchange buffers form one
// distributed buffer.  B = 16384 - 4096 = 12288: 75 % instead of 50 % of every transform is output.
//
//   * 16384 = 16 x 1024.  Cluster-wide thread t (= 256 * rank + tid) gathers z[t + 512 q + 1024 a] (q < 2, a < 16),
//     runs two DFT-16 over a, and PUSHES residue b to the warp that owns it -- warp b mod 8 of CTA b / 8 -- with
//     stores into the cluster's shared-memory window (mapa + st.shared::cluster): half of the 128 KB cross the
//     SM-to-SM network; the hardware cluster barrier (arrive.release / wait.acquire) publishes them;
//   * warp b then runs the 1024-point transform of llz_cuda_fir_fft.cu on its residue with the outer twiddle
//     exp(-2 pi i b n_lo / 16384) folded in (warp-uniform part: the 1024-point table at index 2b; the rest merged with
//     the four-step twiddle into a [b][16][32] table, this CTA's eight residues in shared memory), multiplies by the
//     spectrum (its 16 KB slice arrives in the warp's own slice by TMA while the second DFT-32 runs), runs the
//     inverse 1024-point transform and writes its slice; after a cluster barrier every thread PULLS its 32 points
//     back (half from the peer CTA), and two DFT-16 with the conjugate outer twiddle folded in leave thread t holding
//     outputs t + 512 q + 1024 a: coalesced streaming stores of the rows at or beyond the halo;
//   * three cluster barriers per item; the next item's input span is prefetched into L2.
//
// Per thread and item: 2 x 144 (DFT-16) + 3 x 512 + 388 + 128 (H) + 2 x 208 (folded DFT-16) = 2756 FMA-pipe
// instructions for 48 outputs at B = 12288: 57.4 per output (8192-point kernel: 80.6).
// Verified on the host by tests/cpu/fft16k_emulate.cpp, on the device by tests/test_gpu_fir.py.
#include <stdlib.h>

#include <type_traits>

#include <cooperative_groups.h>

#include "llz_fft32.cuh"
#include "llz_fir_kernels.h"

namespace cg = cooperative_groups;

namespace llz {

template <typename T> struct Cplx16k;
template <> struct Cplx16k<float>  { using type = float2; };
template <> struct Cplx16k<double> { using type = double2; };

constexpr int kFft16kThreads = 256;          // per CTA; the cluster has 512

template <typename T>
struct Fft16kSmem {
    static constexpr size_t bars = 128;                                               // [b]: spectrum slice of warp b
    static constexpr size_t tabw = (size_t)kTwistEntries * kFftR * 2 * sizeof(T);
    static constexpr size_t tab2 = (size_t)8 * kTwistEntries * kFftR * 2 * sizeof(T); // this CTA's eight residues
    static constexpr size_t xbuf = (size_t)8192 * 2 * sizeof(T);                      // half of the cluster's buffer
    static constexpr size_t total = bars + tabw + tab2 + xbuf;
};

// ---- cluster primitives (PTX): the cooperative-groups cluster.sync() adds a full MEMBAR and an error barrier that
// cost ~20 % of this kernel; the release / acquire pair on the hardware cluster barrier is all the exchange needs ------
__device__ __forceinline__ void cluster_sync_ra()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (a shared-memory pointer of this CTA) in CTA `rank` of the cluster, shared::cluster window
__device__ __forceinline__ uint32_t cluster_map(const void *p, int rank)
{
    uint32_t r;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, double2 v)
{
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float2 v)
{
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void ld_cluster(uint32_t addr, double2 &v)
{
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ld_cluster(uint32_t addr, float2 &v)
{
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
}

template <typename T>
__device__ __forceinline__ T fir_fft16k_sample(const FirFftLaunch<T> &a, const T *xc, const T *hc, long long g)
{
    if (g >= 0) return (g < a.n && xc) ? __ldg(xc + g) : T(0);
    if (hc && g >= -(long long)(a.ntaps - 1)) return __ldg(hc + (a.ntaps - 1) + g);
    return T(0);
}

// EDGE = false: interior items (unguarded loads and stores); EDGE = true: first / last items of a channel
template <typename T, bool EDGE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFft16kThreads, 1)
fir_fft16k_kernel(FirFftLaunch<T> a)
{
    using C = typename Cplx16k<T>::type;
    using SM = Fft16kSmem<T>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    C *tabw_s = reinterpret_cast<C *>(smem_raw + SM::bars);                                  // [16][32]
    C *tab2_s = reinterpret_cast<C *>(smem_raw + SM::bars + SM::tabw);                       // [8][16][32]
    C *xbuf = reinterpret_cast<C *>(smem_raw + SM::bars + SM::tabw + SM::tab2);              // 8 slices of 32 x 32

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tc = rank * kFft16kThreads + tid;        // cluster-wide thread
    const int ws = tc >> 5;                            // cluster-wide warp = the residue b this warp transforms
    const uint32_t xb[2] = {cluster_map(xbuf, 0), cluster_map(xbuf, 1)};      // the two halves of the cluster's buffer

    if (tid == 0)
        for (int b = 0; b < 8; ++b) mbar_init(&bars[b], 1);
    for (int i = tid; i < 8 * kTwistEntries * kFftR; i += kFft16kThreads) {
        if (i < kTwistEntries * kFftR) tabw_s[i] = reinterpret_cast<const C *>(a.tw)[i];
        tab2_s[i] = reinterpret_cast<const C *>(a.tw2)[rank * (8 * kTwistEntries * kFftR) + i];
    }
    __syncthreads();
    cluster_sync_ra();                                 // the peer CTA is resident: its shared memory may be written

    const C *Hw = reinterpret_cast<const C *>(a.H) + ws * (kFftR * kFftR);     // [b][k1][k2]
    const C *tab3 = reinterpret_cast<const C *>(a.tw3) + tc;                  // [q][e][t]
    C *slice = xbuf + warp * (kFftR * kFftR);
    const int hl = a.halo, B = a.B;
    const long long total = a.items_per_channel * a.n_channels;
    const long long n_clusters = gridDim.x >> 1;
    const int span_bytes = (kFft16kN + B) * (int)sizeof(T);
    uint32_t h_phase = 0;

    for (long long item = blockIdx.x >> 1; item < total; item += n_clusters) {
        const int ch = (int)(item / a.items_per_channel);
        long long pair = a.first_pair + (item - (long long)ch * a.items_per_channel);
        if constexpr (EDGE) { if (pair >= a.gap_start) pair += a.gap_len; }
        const long long o = pair * (2LL * B);
        const long long s = o - hl;
        const T *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
        T *yc = a.y + (long long)ch * a.y_stride;

        T re[32], im[32];
        // ---- gather: register q*16 + a holds z[tc + 512 q + 1024 a] --------------------------------------------
        if constexpr (!EDGE) {
            const T *p = xc + s + tc;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int aa = 0; aa < 16; ++aa) {
                    re[q * 16 + aa] = __ldg(p + 512 * q + 1024 * aa);
                    im[q * 16 + aa] = __ldg(p + B + 512 * q + 1024 * aa);
                }
            if (a.prefetch && item + n_clusters < total) {
                const long long nit = item + n_clusters;
                const int nch = (int)(nit / a.items_per_channel);
                const long long np = a.first_pair + (nit - (long long)nch * a.items_per_channel);
                const char *src = reinterpret_cast<const char *>(a.x + (long long)nch * a.x_stride + np * (2LL * B) - hl);
                for (int off = tc * 128; off < span_bytes; off += 2 * kFft16kThreads * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(src + off));
            }
        } else {
            const T *hc = a.hist ? a.hist + (long long)ch * (a.ntaps - 1) : nullptr;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int aa = 0; aa < 16; ++aa) {
                    const long long g = s + tc + 512 * q + 1024 * aa;
                    re[q * 16 + aa] = fir_fft16k_sample(a, xc, hc, g);
                    im[q * 16 + aa] = fir_fft16k_sample(a, xc, hc, g + B);
                }
        }
        // the first halo - (N-1) samples of a block reach only discarded outputs: zeroed, so that every kept output is a
        // function of its own N-1 predecessors alone, bit for bit (see llz_cuda_fir_fft.cu)
        if (tc < hl - (a.ntaps - 1)) { re[0] = T(0); im[0] = T(0); }

        // ---- DFT-16 over a; push residue b to its warp: slice b mod 8 of CTA b / 8, row j = ws + 16 q ----------
        dft16<T, false, 0>(re, im);
        dft16<T, false, 16>(re, im);
        // own residues through ordinary shared-memory stores, the peer's through the cluster window
        auto push = [&](auto rk) {
            constexpr int RK = decltype(rk)::value;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int off = (b * kFftR + ws + 16 * q) * kFftR + lane;
                    C v; v.x = re[q * 16 + 8 * RK + b]; v.y = im[q * 16 + 8 * RK + b];
                    xbuf[off] = v;
                    C w; w.x = re[q * 16 + 8 * (RK ^ 1) + b]; w.y = im[q * 16 + 8 * (RK ^ 1) + b];
                    st_cluster(xb[RK ^ 1] + (uint32_t)(off * sizeof(C)), w);
                }
        };
        if (rank == 0) push(std::integral_constant<int, 0>{}); else push(std::integral_constant<int, 1>{});
        cluster_sync_ra();
        // warps 4..7 trail their scheduler partners 0..3 by about one transform phase (see llz_cuda_fir_fft8k.cu)
        if (a.skew > 0 && warp >= 4) {
            const long long t0 = clock64();
            while (clock64() - t0 < a.skew) { }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) { const C v = slice[j * kFftR + lane]; re[j] = v.x; im[j] = v.y; }
        __syncwarp();

        // ---- warp ws: 1024-point forward transform of residue ws with the outer twiddle folded in -----------
        dft32_twisted<T, false>(re, im, tabw_s + 2 * ws, kFftR);
#pragma unroll
        for (int k = 0; k < 32; ++k) { C v; v.x = re[k]; v.y = im[k]; slice[lane * kFftR + (k ^ lane)] = v; }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) { const C v = slice[k * kFftR + (lane ^ k)]; re[k] = v.x; im[k] = v.y; }
        __syncwarp();
        // the slice is idle until the next exchange: fetch this warp's 32 x 32 bins of the spectrum into it
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bars[warp], (uint32_t)(kFftR * kFftR * sizeof(C)));
            tma_bulk_g2s(slice, Hw, (uint32_t)(kFftR * kFftR * sizeof(C)), &bars[warp]);
        }
        dft32_twisted<T, false>(re, im, tab2_s + warp * (kTwistEntries * kFftR) + lane, kFftR);

        // ---- spectrum, inverse 1024-point transform ---------------------------------------------------------------
        mbar_wait(&bars[warp], h_phase);
        h_phase ^= 1;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const C h = slice[k * kFftR + lane];
            cmul_inplace<T, false>(re[k], im[k], h.x, h.y);
        }
        __syncwarp();                                      // all lanes are done with the spectrum before the slice is reused
        dft32<T, true>(re, im);
#pragma unroll
        for (int k = 0; k < 32; ++k) { C v; v.x = re[k]; v.y = im[k]; slice[lane * kFftR + (k ^ lane)] = v; }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) { const C v = slice[k * kFftR + (lane ^ k)]; re[k] = v.x; im[k] = v.y; }
        __syncwarp();
        dft32_twisted<T, true>(re, im, tabw_s + lane, kFftR);

        // ---- write the slice; every thread pulls its 32 points back (half from the peer CTA) ----------------------
#pragma unroll
        for (int j = 0; j < 32; ++j) { C v; v.x = re[j]; v.y = im[j]; slice[j * kFftR + lane] = v; }
        cluster_sync_ra();
        auto pull = [&](auto rk) {
            constexpr int RK = decltype(rk)::value;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int off = (b * kFftR + ws + 16 * q) * kFftR + lane;
                    C w;
                    ld_cluster(xb[RK ^ 1] + (uint32_t)(off * sizeof(C)), w);
                    re[q * 16 + 8 * (RK ^ 1) + b] = w.x; im[q * 16 + 8 * (RK ^ 1) + b] = w.y;
                    const C v = xbuf[off];
                    re[q * 16 + 8 * RK + b] = v.x; im[q * 16 + 8 * RK + b] = v.y;
                }
        };
        if (rank == 0) pull(std::integral_constant<int, 0>{}); else pull(std::integral_constant<int, 1>{});
        cluster_sync_ra();                                 // the next item's pushes overwrite every slice of both CTAs

        // ---- DFT-16 over b with the conjugate outer twiddle folded in ------------------------------------------------
        {
            C e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = __ldg(tab3 + i * 512);
            dft16_twisted<T, true, 0>(re, im, e);
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = __ldg(tab3 + (8 + i) * 512);
            dft16_twisted<T, true, 16>(re, im, e);
        }

        // ---- scatter: rows (q + 2 a) at or beyond halo / 512 are the valid outputs ------------------------------------
        T *qy = yc + o - hl + tc;
        const int r0 = hl >> 9;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int aa = 0; aa < 16; ++aa) {
                const int off = 512 * q + 1024 * aa;
                if (q + 2 * aa >= r0) {
                    if constexpr (!EDGE) {
                        __stcs(qy + off, re[q * 16 + aa]);
                        __stcs(qy + B + off, im[q * 16 + aa]);
                    } else {
                        const long long tA = o - hl + tc + off;
                        if (tA < a.n) __stcs(qy + off, re[q * 16 + aa]);
                        if (tA + B < a.n) __stcs(qy + B + off, im[q * 16 + aa]);
                    }
                }
            }
    }
}

template <typename T, bool EDGE>
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
    auto kern = fir_fft16k_kernel<T, EDGE>;
    LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long items = count * n_channels;
    const long long slots = sm_count / 2;              // one cluster of two CTAs per pair of SMs
    const unsigned clusters = (unsigned)(items < slots ? items : slots);
    kern<<<2 * clusters, kFft16kThreads, smem, stream>>>(b);
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
    a.skew = sk >= 0 ? sk : (sizeof(T) == 8 ? 1300 : 0);       // f64 +5 %, f32 none (profiles/r01_sweep_skew.txt)
    if (fir_fft16k_run<T, false>(a, n_channels, p_lo, p_hi - p_lo, ppc, 0, sm_count, stream) != 0) return -1;
    return fir_fft16k_run<T, true>(a, n_channels, 0, ppc - (p_hi - p_lo), p_lo, p_hi - p_lo, sm_count, stream);
}

template int fir_fft16k_launch<float>(FirFftLaunch<float>, int, cudaStream_t);
template int fir_fft16k_launch<double>(FirFftLaunch<double>, int, cudaStream_t);

}  // namespace llz
