// llz_cuda_polybank_imma.cu -- the exact mode of the phase-bank resampler on the INTEGER tensor cores of sm_100a.
//
// Same tile algebra as llz_cuda_polybank.cu (libllzfilter/llz_resample.c:583-603 written as Y = G'^T X' per tile of
// 64 phases x 64 cycles), different arithmetic.  The reference accumulates sum_k g[l][k] * x in FP64 and truncates;
// the exact mode only has to land on the same side of every integer, and the near-integer guard
// recomputes the rare outputs that come close.  So the bulk evaluation may be ANY evaluation with a small, known error
// bound -- and for int16 samples and fixed taps that can be an exact integer one:
//
//   * taps:    g * 2^s rounded to an integer q, |q| < 2^38, written in five signed base-256 digits (int8 planes p0..p4);
//   * samples: x = 256 * xh + xl with xh = x >> 8 (int8) and xl = x & 255 (uint8);
//   * the ten digit products p_i * {xl, xh} run as IMMA.16832 (mma.sync.m16n8k32, s8 x u8 / s8 x s8 -> s32): every
//     product and every 32-bit sum is exact (|sum| <= 2 * 404 * 128 * 255 < 2^25); products of equal weight 2^(8(i+j))
//     share an accumulator, so a thread carries six s32 accumulator sets;
//   * epilogue: sum_d acc_d * 256^d as a 64-bit integer, one conversion, scale 2^-s, gain, guard, saturate, truncate.
//
// The only error is the tap rounding: |sum - exact| <= Q * 32768 * 2^-(s+1) (1.5e-5 for config C4's Q = 257, s = 38),
// which widens the guard band from ~1e-9 to ~3e-5 of the outputs; those get a second look as a warp-cooperative FP64 dot
// product, and only what is still near an integer goes to the reference's serial order.  Single-tap (knife-edge) rows
// are evaluated as one exact FP64 product.
// B200 rates (tools/probe_pipes.cu): IMMA 1144 TOP/s = 57 T exact MACs/s after the ten-way split, against 17-18.5 T
// MACs/s for DFMA / DMMA.
//
// Layout.  With the tap index reversed, k'' = K'-1-k', the operand rows are k-contiguous as mma's row.col fragments
// want them: X''[j][k''] = span[j*M + k''] (an ascending run of the staged input span per cycle) and
// G''[l][k''] = g[l][Q-1 + (c_l - c_lo) - k''].  G'' depends on the phase tile only, so the host lays every tile out
// once, chunk by chunk (64 k'' bytes per row, rows padded to 80 bytes: conflict-free ldmatrix), and a chunk of all
// five planes arrives by ONE TMA bulk copy; X'' chunks are cut from two byte planes of the span (split once per tile)
// with funnel shifts.  Four-stage mbarrier pipeline between a producer warpgroup and two consumer warpgroups.
//
// Status (round 1): bit-identical on every resampler test and the default for the exact mode (LLZ_BANK_NO_IMMA=1 keeps
// the FP64 tiles): C4 83.3 against 55.6 Gsamples/s.  History: 35.0 (every guard hit recomputed serially) -> 49.0
// (warp-cooperative second look) -> 51.3 (integer epilogue, 64-tap chunks, taps fetched before the span) -> 56.8
// (persistent CTAs) -> 58.3 (four stages) -> 74.8 (producer warpgroup, see the kernel) -> 75.0 (guard band from the taps'
// actual rounding errors) -> 83.3 (interior-tile epilogue without validity tests and 64-bit index arithmetic).  Tensor
// pipe ~48 % active; what is left exposed is the consumers' epilogue.  Next: two consumer
// groups alternating tiles, and the taps of a phase tile multicast across a cluster (the G'' stream is 23 B/clk/SM at the
// IMMA rate: L2-bound if every CTA fetches its own).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "llz_imma_tables.h"
#include "llz_poly_device.cuh"

namespace llz {

namespace {

constexpr int kIHeader = 128;

struct ImmaGeom {
    long long jc0;
    int n_cycle_tiles, n_phase_tiles, n_channels;
    int raw_cap;                                      // int16 elements reserved for the staged span (multiple of 16)
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void *p)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}

// A = tap digits (s8), B = sample bytes: low plane u8, high plane s8
__device__ __forceinline__ void imma_s8u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void imma_s8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// geometry of one tile (64 phases x 64 cycles of one channel)
struct ImmaTile {
    int ch, l0, pbv, cspan, KP, nchunks, need, tile_p;
    long long j0, S0;
};

__device__ __forceinline__ ImmaTile imma_tile(const PolyLaunch &a, const ImmaGeom &geo, long long t)
{
    ImmaTile T;
    T.tile_p = (int)(t % geo.n_phase_tiles);                   // phase tiles fastest: neighbours share the input span in L2
    const long long r = t / geo.n_phase_tiles;
    const int tile_j = (int)(r % geo.n_cycle_tiles);
    T.ch = (int)(r / geo.n_cycle_tiles);
    const int L = a.L, M = a.M, Q = a.ctaps;
    T.l0 = T.tile_p * kIPB;
    T.pbv = min(kIPB, L - T.l0);
    const int c_lo = (int)(((long long)T.l0 * M) / L);
    const int c_hi = (int)(((long long)(T.l0 + T.pbv - 1) * M) / L);
    T.cspan = c_hi - c_lo;
    T.KP = Q + T.cspan;
    T.nchunks = (T.KP + kIKC - 1) / kIKC;
    T.j0 = geo.jc0 + (long long)tile_j * kIJB;
    T.S0 = T.j0 * M + c_lo - (Q - 1);
    const long long jc_last = (a.o0 + a.n_out - 1) / L;
    const int jv = (int)min((long long)kIJB, jc_last - T.j0 + 1);   // cycles of this tile that hold outputs of the call
    T.need = (jv - 1) * M + T.cspan + Q;
    return T;
}

// Persistent and warp-specialised: one CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The six
// accumulator sets cost ~230 registers, so only one CTA fits an SM and nothing else would hide a tile's prologue and
// chunk production.  Warps 0-7 (two warpgroups, setmaxnreg 232) only wait for full stages, multiply and run the epilogue;
// warps 8-11 (one warpgroup, setmaxnreg 40) stage and split the spans, cut the X'' chunks and issue the G'' copies, up to
// kIStages chunks ahead of the consumers -- across tile boundaries, so the next tile's first chunks are ready while the
// consumers still finish the current tile's outputs.  The two sides meet only on the stages' full / empty mbarriers.
constexpr int kIConsumers = 256, kIProducers = 128;   // consumers: 8 warps = 4 (phases) x 2 (cycles), warp tile 16 x 32

template <int MODE, int P>
__global__ void __launch_bounds__(kIConsumers + kIProducers, 1)
poly_bank_imma_kernel(PolyLaunch a, ImmaGeom geo)
{
    constexpr int kIPlanes = P, kIGStage = imma_gstage(P), kIStage = imma_stage(P);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *span_bar = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *s_full = span_bar + 1, *s_empty = s_full + kIStages;
    unsigned char *stages = smem_raw + kIHeader;
    int16_t *raw = reinterpret_cast<int16_t *>(stages + kIStages * kIStage);      // [raw_cap]
    unsigned char *rawl = reinterpret_cast<unsigned char *>(raw + geo.raw_cap);   // [raw_cap + 32] low bytes, by span index
    unsigned char *rawh = rawl + geo.raw_cap + 32;                                // [raw_cap + 32] high bytes

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = a.L, M = a.M, Q = a.ctaps;
    const long long total = (long long)geo.n_phase_tiles * geo.n_cycle_tiles * geo.n_channels;

    if (tid == 0) {
        mbar_init(span_bar, 1);
        for (int i = 0; i < kIStages; ++i) {
            mbar_init(&s_full[i], kIProducers + 1);            // every producer thread's share of X'' + the expect_tx of the G'' copy
            mbar_init(&s_empty[i], kIConsumers);
        }
    }
    __syncthreads();

    if (warp >= kIConsumers / 32) {
        // ================================ producers ================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int ptid = tid - kIConsumers;
        auto chan_x = [&](int ch) { return a.x ? a.x + (long long)ch * a.x_stride : nullptr; };
        long long g = 0;                                       // global chunk counter
        uint32_t span_phase = 0;
        ImmaTile T = imma_tile(a, geo, blockIdx.x);
        if (ptid == 0) {
            const int16_t *xc = chan_x(T.ch);
            const PolySpanPlan sp = poly_span_plan(a, xc, T.S0, T.need);
            if (sp.tma) poly_span_issue(sp, xc, raw, span_bar);
        }
        for (long long t = blockIdx.x; t < total; t += gridDim.x) {
            const int16_t *xc = chan_x(T.ch);
            const int16_t *hc = a.hist ? a.hist + (long long)T.ch * a.hist_len : nullptr;
            // the span: bulk part in flight, fringes by the threads, then split into byte planes
            const PolySpanPlan sp = poly_span_plan(a, xc, T.S0, T.need);
            poly_span_fringes<kIProducers>(a, xc, hc, T.S0, T.need, sp, raw, ptid);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (sp.tma) { mbar_wait(span_bar, span_phase); span_phase ^= 1; }
            const int raw_off = sp.off;
            const int nvec = (raw_off + T.need + 7) >> 3;      // eight samples per step, both planes indexed like raw
            for (int v = ptid; v < nvec; v += kIProducers) {
                const uint4 w = reinterpret_cast<const uint4 *>(raw)[v];
                uint2 lo, hi;
                lo.x = __byte_perm(w.x, w.y, 0x6420); lo.y = __byte_perm(w.z, w.w, 0x6420);
                hi.x = __byte_perm(w.x, w.y, 0x7531); hi.y = __byte_perm(w.z, w.w, 0x7531);
                reinterpret_cast<uint2 *>(rawl)[v] = lo;
                reinterpret_cast<uint2 *>(rawh)[v] = hi;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // the next tile's span may now overwrite raw
            const long long tn = t + gridDim.x;
            const int tile_p = T.tile_p, nchunks = T.nchunks;
            if (tn < total) {
                T = imma_tile(a, geo, tn);
                if (ptid == 0) {
                    const int16_t *xn = chan_x(T.ch);
                    const PolySpanPlan spn = poly_span_plan(a, xn, T.S0, T.need);
                    if (spn.tma) poly_span_issue(spn, xn, raw, span_bar);
                }
            }
            for (int c = 0; c < nchunks; ++c, ++g) {
                const int buf = (int)(g % kIStages);
                if (g >= kIStages) mbar_wait(&s_empty[buf], (uint32_t)((g / kIStages - 1) & 1));
                unsigned char *st = stages + buf * kIStage;
                if (ptid == 0) {                               // the chunk's G'' planes: one bulk copy
                    mbar_expect_tx(&s_full[buf], (uint32_t)kIGStage);
                    tma_bulk_g2s(st, a.imma_tiles + ((size_t)tile_p * a.imma_nchunks + c) * kIGStage, (uint32_t)kIGStage, &s_full[buf]);
                }
#pragma unroll
                for (int q2 = 0; q2 < 4; ++q2) {               // this thread's 4 x 16 bytes of the chunk's X'' planes
                    const int task = ptid + q2 * kIProducers;  // (cycle, plane, 16-byte quarter)
                    const int xj = task & 63, xplane = (task >> 6) & 1, xq = task >> 7;
                    const unsigned char *src = (xplane ? rawh : rawl) + raw_off + xj * M + c * kIKC + 16 * xq;   // any alignment
                    const uint32_t addr = smem_u32(src);
                    const uint32_t sh = (addr & 3u) * 8u;
                    const uint32_t *w = reinterpret_cast<const uint32_t *>(src - (addr & 3u));
                    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
                    uint4 o;
                    o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh);
                    o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
                    *reinterpret_cast<uint4 *>(st + kIGStage + xplane * (kIJB * kIPitch) + xj * kIPitch + 16 * xq) = o;
                }
                mbar_arrive(&s_full[buf]);
            }
            // the planes are rewritten at the top of the next iteration: every producer is past its last read of them
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        return;
    }

    // ================================ consumers ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int wm = warp >> 1, wn = warp & 1;                   // warp tile: phases [16*wm, +16) x cycles [32*wn, +32)
    // ldmatrix row addresses: lane -> matrix lane/8, row lane%8.  A (16 phases x 32 k-bytes): matrices (rows 0-7 | 8-15)
    // x (bytes 0-15 | 16-31); B (8 cycles x 32 k-bytes per n tile, two n tiles per x4): (n tile) x (bytes 0-15 | 16-31)
    const int mat = lane >> 3, mr = lane & 7;
    const int a_off = (16 * wm + mr + 8 * (mat & 1)) * kIPitch + 16 * (mat >> 1);
    const int b_off = (32 * wn + mr + 8 * (mat >> 1)) * kIPitch + 16 * (mat & 1);
    long long g0 = 0;                                          // global chunk index of the current tile's chunk 0
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const ImmaTile T = imma_tile(a, geo, t);
        const int16_t *xc = a.x ? a.x + (long long)T.ch * a.x_stride : nullptr;
        const int16_t *hc = a.hist ? a.hist + (long long)T.ch * a.hist_len : nullptr;
        const int l0 = T.l0, pbv = T.pbv, nchunks = T.nchunks;
        const long long j0 = T.j0;

        // the two phase rows of this thread: is either a single-tap row?  (fetched now, needed in the epilogue)
        int st_row[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int l = 16 * wm + 8 * h + (lane >> 2);
            st_row[h] = (l < pbv) ? __ldg(a.single_tap + l0 + l) : -1;
        }

        int acc[kIPlanes + 1][4][4];                           // [weight 2^(8d)][n tile][c fragment]
#pragma unroll
        for (int d = 0; d <= kIPlanes; ++d)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[d][ni][e] = 0;

        for (int c = 0; c < nchunks; ++c) {
            const long long g = g0 + c;
            const int buf = (int)(g % kIStages);
            mbar_wait(&s_full[buf], (uint32_t)((g / kIStages) & 1));
            const unsigned char *gs = stages + buf * kIStage;
            const unsigned char *xs = gs + kIGStage;
            // the asm statements keep their order: the tap fragments are double-buffered by hand, so that plane i+1 is
            // on its way from shared memory while plane i's eight IMMAs issue
#pragma unroll
            for (int ks = 0; ks < kIKC / 32; ++ks) {
                uint32_t bl[2][4], bh[2][4], af[2][4];
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    ldsm_x4(bl[np], xs + b_off + np * (16 * kIPitch) + 32 * ks);
                    ldsm_x4(bh[np], xs + kIJB * kIPitch + b_off + np * (16 * kIPitch) + 32 * ks);
                }
                ldsm_x4(af[0], gs + a_off + 32 * ks);
#pragma unroll
                for (int i = 0; i < kIPlanes; ++i) {
                    if (i + 1 < kIPlanes) ldsm_x4(af[(i + 1) & 1], gs + (i + 1) * (kIPB * kIPitch) + a_off + 32 * ks);
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
                        const int np = ni >> 1, q = (ni & 1) * 2;
                        imma_s8u8(acc[i][ni], af[i & 1], bl[np][q], bl[np][q + 1]);
                        imma_s8s8(acc[i + 1][ni], af[i & 1], bh[np][q], bh[np][q + 1]);
                    }
                }
            }
            mbar_arrive(&s_empty[buf]);
        }
        g0 += nchunks;

        // ---- epilogue: C[row = lane/4 (+8)][col = 2*(lane%4) + {0,1}] of n tile ni ---------------------------------------
        const long long o_end = a.o0 + a.n_out;
        int16_t *ych = a.y + (long long)T.ch * a.y_stride;
        // First pass, straight-line: sum_d acc_d 256^d fits a 64-bit integer (|.| < 2^62), so the six accumulators are
        // combined with integer multiply-adds and converted once; every output is finished and stored, near-integer hits
        // are only noted.  (Converting and combining in FP64 behind a branch per output cost as much as the MMA loop.)
        uint32_t guard_hits = 0;                               // bit ni*4 + e
        const double out_scale = a.imma_scale * 16777216.0;    // the high half carries 256^3
        auto combine = [&](int ni, int e) -> double {          // the tile's sum for accumulator element (ni, e), times gain
            const long long lo = (long long)acc[0][ni][e] + (long long)acc[1][ni][e] * 256 + (long long)acc[2][ni][e] * 65536;
            double s;
            if constexpr (P == kIPlanesExact) {
                const long long hi = (long long)acc[3][ni][e] + (long long)acc[4][ni][e] * 256 + (long long)acc[5][ni][e] * 65536;
                s = fma((double)hi, out_scale, (double)lo * a.imma_scale);
            } else {
                static_assert(P == kIPlanesFast, "three or five tap digits");
                s = (double)(lo + (long long)acc[3][ni][e] * 16777216) * a.imma_scale;   // < 2^51: exact
            }
            return __dmul_rn(s, a.gain);
        };
        const long long tile_first = j0 * (long long)L + l0;   // output index of (cycle 0, phase 0) of the tile
        const bool interior = pbv == kIPB && tile_first >= a.o0 && tile_first + 63LL * L + (kIPB - 1) < o_end;
        if (interior) {
            // every output of the tile belongs to the call (all but the first and last tiles of a channel): no validity
            // tests, 32-bit offsets from one 64-bit row pointer, straight-line stores
            const int jl = (32 * wn + 2 * (lane & 3)) * L;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int st = st_row[h];
                int16_t *yrow = ych + (tile_first - a.o0) + (16 * wm + 8 * h + (lane >> 2)) + jl;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int e2 = 0; e2 < 2; ++e2) {
                        const int e = 2 * h + e2;
                        const double v = combine(ni, e);
                        if (MODE == LLZ_CUDA_ACC_F64 && st < 0 && poly_near_nonzero_integer(v, a.imma_thr))
                            guard_hits |= 1u << (ni * 4 + e);
                        yrow[(8 * ni + e2) * L] = poly_finish(v);
                    }
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int l = 16 * wm + 8 * h + (lane >> 2);
                const bool l_ok = l < pbv;
                const int st = st_row[h];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int e2 = 0; e2 < 2; ++e2) {
                        const int e = 2 * h + e2;
                        const int j = 32 * wn + 8 * ni + 2 * (lane & 3) + e2;
                        const long long o = (j0 + j) * (long long)L + l0 + l;
                        const bool valid = l_ok && o >= a.o0 && o < o_end;
                        const double v = combine(ni, e);
                        if (MODE == LLZ_CUDA_ACC_F64 && valid && st < 0 && poly_near_nonzero_integer(v, a.imma_thr))
                            guard_hits |= 1u << (ni * 4 + e);
                        if (valid) ych[o - a.o0] = poly_finish(v);
                    }
            }
        }
        // knife-edge phases (one tap, 1 - 2^-53 for the L-th band prototype): one exact FP64 product per output
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int l = 16 * wm + 8 * h + (lane >> 2);
            const int st = h ? st_row[1] : st_row[0];
            if (st < 0) continue;
            for (int idx = 0; idx < 8; ++idx) {
                const int j = 32 * wn + 8 * (idx >> 1) + 2 * (lane & 3) + (idx & 1);
                const long long o = (j0 + j) * (long long)L + l0 + l;
                if (o < a.o0 || o >= o_end) continue;
                const long long base = (o * M) / L;
                ych[o - a.o0] = poly_finish(__dmul_rn(__dmul_rn((double)poly_sample(a, xc, hc, base - st), a.cbank[(long long)(l0 + l) * Q + st]), a.gain));
            }
        }
        __syncwarp();

        // Second look at the outputs that came within the (wide) integer-evaluation band of an integer, ~3e-5 of them: the
        // whole warp evaluates such an output again as an FP64 dot product (lane-strided FMAs, shuffle reduction), which is
        // good to the narrow band of the FP64 kernels; only what is STILL near an integer (~1e-8) goes to the reference's
        // own serial order.  (Sending every first-level hit to the serial recompute -- 257 dependent global loads on one
        // lane -- cost more than the tile's tensor work.)
        unsigned pending = __ballot_sync(0xffffffffu, guard_hits != 0);
        while (pending) {
            const int src = __ffs(pending) - 1;
            int idx = (lane == src) ? __ffs(guard_hits) - 1 : 0;
            idx = __shfl_sync(0xffffffffu, idx, src);
            const int ni = idx >> 2, e = idx & 3;
            const int l = 16 * wm + 8 * (e >> 1) + (src >> 2);
            const int j = 32 * wn + 8 * ni + 2 * (src & 3) + (e & 1);
            const long long o = (j0 + j) * (long long)L + l0 + l;         // warp-uniform
            const long long base = (o * M) / L;
            const double *row = a.cbank + (long long)(l0 + l) * Q;
            double part = 0.0;
            for (int k = lane; k < Q; k += 128) {              // four independent sample / tap loads in flight per lane
                int xv[4];
                double gv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int kk = k + 32 * u;
                    const bool in = kk < Q;
                    xv[u] = in ? poly_sample(a, xc, hc, base - kk) : 0;
                    gv[u] = in ? row[kk] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) part = fma((double)xv[u], gv[u], part);
            }
#pragma unroll
            for (int m = 16; m; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
            if (lane == src) {
                double v = __dmul_rn(part, a.gain);
                if (poly_near_nonzero_integer(v, a.guard_thr)) {
                    v = __dmul_rn(poly_reference_order_sum(a, xc, hc, o), a.gain);
                    atomicAdd(a.guard_count, 1ULL);
                }
                ych[o - a.o0] = poly_finish(v);
                guard_hits &= guard_hits - 1;
            }
            pending = __ballot_sync(0xffffffffu, guard_hits != 0);
        }
    }
}

}  // namespace

// 1 = launched, 0 = not applicable, -1 = error
int poly_bank_imma_launch(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    if (!a.imma_tiles || a.imma_nchunks <= 0) return 0;
    ImmaGeom geo{};
    geo.jc0 = a.o0 / a.L;
    const long long jc_last = (a.o0 + a.n_out - 1) / a.L;
    geo.n_cycle_tiles = (int)((jc_last - geo.jc0 + 1 + kIJB - 1) / kIJB);
    geo.n_phase_tiles = (a.L + kIPB - 1) / kIPB;
    const int cspan_max = (int)(((long long)kIPB * a.M) / a.L) + 2;
    // the byte planes are read up to 31 bytes past the last needed sample (chunk padding meets zero taps)
    geo.raw_cap = ((kIJB - 1) * a.M + cspan_max + a.ctaps + kIKC + 16 + 15) & ~15;
    geo.n_channels = n_channels;
    if (a.acc != LLZ_CUDA_ACC_F64 || a.imma_planes != kIPlanesExact) return 0;
    const size_t smem = kIHeader + (size_t)kIStages * imma_stage(a.imma_planes) + (size_t)geo.raw_cap * 2 + 2 * ((size_t)geo.raw_cap + 32);
    if (smem > 226 * 1024) return 0;
    auto kern = poly_bank_imma_kernel<LLZ_CUDA_ACC_F64, kIPlanesExact>;
    LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = device_sm_count();
    if (sms <= 0) return -1;
    const long long tiles = (long long)geo.n_cycle_tiles * geo.n_phase_tiles * n_channels;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);           // persistent: one CTA per SM
    kern<<<grid, kIConsumers + kIProducers, smem, stream>>>(a, geo);
    note_launch("poly_bank_imma_kernel");
    LLZ_CUDA_TRY(cudaGetLastError());
    return 1;
}

}  // namespace llz
