// llz_cuda_polybank.cu -- register-tiled phase-bank kernel for rational L/M resampling (L > 1) on sm_100a.
//
// Replaces the output loop of the reference's llz_resample (libllzfilter/llz_resample.c:583-603) for banks
// with many phases (BASELINE configs C1: L=160, Q=45 and C4: L=320, Q=257).
//
// Write output m = j*L + l (cycle j, phase l).  The reference's index stepping gives
//     y[j][l] = sum_{k<Q} g[l][k] * x[j*M + c_l - k],        c_l = floor(l*M/L)          (llz_resample.c:586-592)
// i.e. for a tile of phases every cycle j is one column of a small matrix product.  With the tap index shifted
// per phase, k' = k + (c_hi - c_l) (c_hi = the largest c_l of the tile), all phases of the tile read the SAME
// input column X'[k'][j] = x[j*M + c_hi - k'], and the tile becomes
//     Y[l][j] = sum_{k'<K'} G'[k'][l] * X'[k'][j],   K' = Q + (c_hi - c_lo),   G'[k'][l] = g[l][k' - (c_hi - c_l)] or 0.
// A CTA owns PB phases x JB cycles.  Its contiguous int16 input span arrives once in shared memory (TMA bulk
// copy for interior tiles), then K' is walked in chunks of KC rows: each chunk of G' (gathered from the
// transposed bank in L2) and X' (expanded from the staged span, converted to the accumulator type once) is
// staged in a three-deep mbarrier pipeline in shared memory while every thread accumulates a TP x TJ register
// tile.  The binding resource is the shared-memory -> register return path (128 B/clk/SM: an LDS.128 costs four
// cycles of it however much of it is broadcast), so the tile is large (f32 16x8, f64 8x8) and X' stays int16 in
// shared memory (one LDS.128 = 8 cycles' worth of samples), converted in registers by the otherwise idle
// conversion unit: 80 bytes loaded per thread per TP*TJ FMAs.
//
// Phase (m mod L) and input index (floor(m*M/L)) are pure integer arithmetic, identical to the reference's
// sequence; the FP64 mode carries the same near-integer guard as the other kernels (bit-identical int16).
// Padding cost: K'/Q (e.g. 286/257 for C4 with 64-phase tiles).
#include <stdlib.h>

#include <type_traits>

#include <cuda_fp16.h>

#include "llz_poly_device.cuh"

namespace llz {

namespace {

constexpr int kStages = 3;        // chunk pipeline depth: a warp may run two chunks ahead of the slowest one
constexpr int kRawSlack = 64;      // X' rows beyond K' index up to KC samples before the span: finite garbage times zero taps

struct BankGeom {
    long long jc0;                 // first cycle touched by this call = floor(o0 / L)
    int n_cycle_tiles;
    int n_phase_tiles;
    int raw_cap;                   // int16 elements reserved for the staged span
};

// CTA = PG phase groups x CG cycle groups of threads; thread tile TP phases x TJ consecutive cycles.
template <typename TA, int TP, int TJ, int PG, int CG, int KC, bool XI16, int MODE>
__global__ void __launch_bounds__(PG * CG, (PG * CG == 64) ? 4 : (sizeof(TA) == 4 ? 3 : 2))
poly_bank_kernel(PolyLaunch a, BankGeom geo)
{
    constexpr int kPG = PG, kCG = CG, kBankThreads = PG * CG;
    // X' element type in shared memory: int16 (converted in registers, one LDS.128 = 8 cycles) or the accumulator type
    using XT = typename std::conditional<XI16, int16_t, TA>::type;
    static_assert(!XI16 || TJ == 8, "int16 X': a thread's cycles are one 16-byte vector of samples");
    constexpr int PB = kPG * TP;
    constexpr int JB = kCG * TJ;
    constexpr int VU = 16 / (int)sizeof(TA);          // elements per 16-byte vector
    static_assert(TP % VU == 0, "a thread's phases must be whole vectors");
    using V = typename Vec16<TA>::type;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    TA *Gs = reinterpret_cast<TA *>(smem_raw + 16);            // [kStages][KC][PB]
    XT *Xs = reinterpret_cast<XT *>(Gs + kStages * KC * PB);   // [kStages][KC][JB]
    int16_t *raw = reinterpret_cast<int16_t *>(Xs + kStages * KC * JB) + kRawSlack;   // kRawSlack elements of slack in front
    __shared__ int s_shift[PB];                                // c_hi - c_l per phase of the tile
    __shared__ uint64_t s_full[kStages], s_empty[kStages];     // chunk pipeline: stage filled / stage drained

    const int tid = threadIdx.x;
    const int pg = tid / kCG, cg = tid % kCG;
    const int tile_p = blockIdx.x % geo.n_phase_tiles;         // phase tiles fastest: neighbours share the input span in L2
    const int tile_j = blockIdx.x / geo.n_phase_tiles;
    const int ch = blockIdx.y;
    const int L = a.L, M = a.M, Q = a.ctaps;

    const int l0 = tile_p * PB;
    const int pbv = min(PB, L - l0);                           // phases that exist in this tile
    const int c_lo = (int)(((long long)l0 * M) / L);
    const int c_hi = (int)(((long long)(l0 + pbv - 1) * M) / L);
    const int cspan = c_hi - c_lo;
    const int KP = Q + cspan;                                  // K'
    const long long j0 = geo.jc0 + (long long)tile_j * JB;     // first cycle of the tile
    const int rawn = (JB - 1) * M + cspan + Q;                 // staged samples
    const long long S0 = j0 * M + c_lo - (Q - 1);              // canonical index of raw[0]

    const int16_t *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
    const int16_t *hc = a.hist ? a.hist + (long long)ch * a.hist_len : nullptr;

    // ---- stage the input span ---------------------------------------------------------------------------
    // cycles of this tile that hold outputs of the call, and the samples they read
    const long long jc_last = (a.o0 + a.n_out - 1) / L;
    const int jv = (int)min((long long)JB, jc_last - j0 + 1);
    const int need = min(rawn, (jv - 1) * M + cspan + Q);
    bool bulk;
    const int raw_off = poly_stage_span<kBankThreads>(a, xc, hc, S0, need, raw, bar, tid, &bulk);
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&s_full[i], kBankThreads);
            mbar_init(&s_empty[i], kBankThreads);
        }
    }
    for (int l = tid; l < PB; l += kBankThreads)
        s_shift[l] = (l < pbv) ? c_hi - (int)(((long long)(l0 + l) * M) / L) : 0;   // phases past L: any finite taps
    __syncthreads();
    if (bulk) mbar_wait(bar, 0);
    const int16_t *rawp = raw + raw_off;

    // ---- chunk builders ---------------------------------------------------------------------------------
    // The transposed bank [k][L] is uploaded with zero rows before row 0 and after row Q-1 (PolyLaunch::bank_pad),
    // so G'[k'][l] = bankT[(k' - shift_l)*L + l0 + l] needs no bounds checks; rows k' >= K' of X' meet only zeros.
    const TA *bankT = (sizeof(TA) == 8) ? reinterpret_cast<const TA *>(a.cbankT64) : reinterpret_cast<const TA *>(a.cbankT32);
    static_assert(kBankThreads % PB == 0, "G' builder: a pass of the CTA covers whole rows");
    static_assert(kBankThreads % JB == 0 || JB % kBankThreads == 0, "X' builder: whole rows per pass or whole passes per row");
    constexpr int GR = kBankThreads / PB;                      // G' rows covered by one pass of the CTA
    constexpr int GE = KC / GR;                                // G' elements per thread per chunk
    constexpr int XC = (JB > kBankThreads) ? JB / kBankThreads : 1;      // X' columns per thread
    constexpr int XR = (JB > kBankThreads) ? 1 : kBankThreads / JB;      // X' rows covered by one pass
    constexpr int XE = KC / XR;                                // X' rows per thread per chunk
    static_assert(KC % GR == 0 && KC % XR == 0, "tile shape");
    const int gl = tid % PB, gk = tid / PB;                    // this thread's G' column and first row
    const TA *gsrc = bankT + ((long long)gk - s_shift[gl]) * L + l0 + gl;
    const int xj = tid % JB, xk = (JB > kBankThreads) ? 0 : tid / JB;    // this thread's first X' column and first row
    const int16_t *xsrc = rawp + xj * M + (KP - 1) - xk;
    TA gpre[GE];

    auto load_g = [&](int chunk) {                             // global -> registers (latency hidden behind the MAC loop)
        const TA *src = gsrc + (long long)chunk * KC * L;
#pragma unroll
        for (int i = 0; i < GE; ++i) gpre[i] = src[(long long)i * GR * L];
    };
    auto store_g = [&](int buf) {
        TA *dst = Gs + buf * KC * PB + tid;
#pragma unroll
        for (int i = 0; i < GE; ++i) dst[i * kBankThreads] = gpre[i];
    };
    auto build_x = [&](int chunk, int buf) {                   // X'[k'][j] = raw[j*M + (K'-1) - k'] (still int16)
        XT *dst = Xs + buf * KC * JB + xk * JB + xj;
        const int16_t *src = xsrc - chunk * KC;
#pragma unroll
        for (int i = 0; i < XE; ++i)
#pragma unroll
            for (int c = 0; c < XC; ++c) dst[i * XR * JB + c * kBankThreads] = (XT)src[c * kBankThreads * M - i * XR];
    };

    // ---- main loop ----------------------------------------------------------------------------------------
    TA acc[TP][TJ];
#pragma unroll
    for (int p = 0; p < TP; ++p)
#pragma unroll
        for (int t = 0; t < TJ; ++t) acc[p][t] = TA(0);

    // Chunk c lives in stage c % kStages.  Every thread builds its share of chunk c+2 after its own MACs of chunk c
    // and signals s_full; it starts chunk c when all shares have arrived.  No CTA-wide barrier in the loop: warps
    // drift up to two chunks apart, so one warp's build / shared-memory latency is covered by the others' FMAs.
    const int nchunks = (KP + KC - 1) / KC;
#pragma unroll
    for (int c0 = 0; c0 < 2; ++c0) {
        if (c0 < nchunks) {
            load_g(c0);
            store_g(c0);
            build_x(c0, c0);
            mbar_arrive(&s_full[c0]);
        }
    }
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c % kStages;
        const int nxt = (c + 2) % kStages;
        const bool produce = c + 2 < nchunks;
        if (produce) load_g(c + 2);                            // global loads in flight during the MACs
        mbar_wait(&s_full[buf], (c / kStages) & 1);
        const TA *gb = Gs + buf * KC * PB + pg * TP;
        const XT *xb = Xs + buf * KC * JB + cg * TJ;
        // fragments double-buffered in registers: the LDS of row kk+1 are in flight while row kk's FMAs issue
        constexpr int NB = (TP * TJ <= 64) ? 2 : 1;            // big tiles: no room for a second fragment set
        constexpr int XW = TJ * (int)sizeof(XT) / 16;          // 16-byte vectors of X' per thread per row
        TA g[NB][TP];
        uint4 xr[NB][XW];
        auto load_row = [&](int kk, int slot) {
#pragma unroll
            for (int p = 0; p < TP; p += VU) unpack(*reinterpret_cast<const V *>(gb + kk * PB + p), &g[slot][p]);
#pragma unroll
            for (int q = 0; q < XW; ++q) xr[slot][q] = reinterpret_cast<const uint4 *>(xb + kk * JB)[q];
        };
        if (NB == 2) load_row(0, 0);
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            if (NB == 2) {
                if (kk + 1 < KC) load_row(kk + 1, (kk + 1) & 1);
            } else {
                load_row(kk, 0);
            }
            TA xv[TJ];
            if constexpr (XI16) {
                const uint4 w = xr[kk & (NB - 1)][0];
                xv[0] = (TA)(short)(w.x & 0xffffu); xv[1] = (TA)(short)(w.x >> 16);
                xv[2] = (TA)(short)(w.y & 0xffffu); xv[3] = (TA)(short)(w.y >> 16);
                xv[4] = (TA)(short)(w.z & 0xffffu); xv[5] = (TA)(short)(w.z >> 16);
                xv[6] = (TA)(short)(w.w & 0xffffu); xv[7] = (TA)(short)(w.w >> 16);
            } else {
#pragma unroll
                for (int q = 0; q < XW; ++q) unpack(*reinterpret_cast<const V *>(&xr[kk & (NB - 1)][q]), &xv[q * VU]);
            }
#pragma unroll
            for (int p = 0; p < TP; ++p)
#pragma unroll
                for (int t = 0; t < TJ; ++t) acc[p][t] = mac<TA, false>(g[kk & (NB - 1)][p], xv[t], acc[p][t]);
        }
        mbar_arrive(&s_empty[buf]);                            // this thread is done reading stage buf
        if (produce) {
            // stage nxt last held chunk c-1: wait until every thread has drained it
            if (c >= 1) mbar_wait(&s_empty[nxt], ((c - 1) / kStages) & 1);
            store_g(nxt);
            build_x(c + 2, nxt);
            mbar_arrive(&s_full[nxt]);
        }
    }

    // ---- gain / guard / saturate / truncate / store --------------------------------------------------------------
    const long long o_end = a.o0 + a.n_out;
    const bool unit_gain = a.gain == 1.0;                      // x * 1.0 == x exactly: skip the FP64 multiply
    (void)unit_gain;
#pragma unroll
    for (int t = 0; t < TJ; ++t) {
        const int j = cg * TJ + t;                             // cycle within the tile
        const long long obase = (j0 + j) * (long long)L + l0 + pg * TP;
        int16_t outv[TP];
        bool all_valid = true;
#pragma unroll
        for (int p = 0; p < TP; ++p) {
            const int l = pg * TP + p;
            const long long o = obase + p;
            const bool valid = l < pbv && o >= a.o0 && o < o_end;
            all_valid = all_valid && valid;
            int16_t r16 = 0;
            if (valid) {
                const int st = a.single_tap[l0 + l];
                if constexpr (MODE == LLZ_CUDA_ACC_F32) {
                    double v = (double)acc[p][t];
                    if (st >= 0) {                             // knife-edge phase: one exact FP64 product
                        const long long base = (o * M) / L;
                        v = __dmul_rn((double)poly_sample(a, xc, hc, base - st), a.cbank[(long long)(l0 + l) * Q + st]);
                    }
                    r16 = poly_finish(__dmul_rn(v, a.gain));
                } else {
                    double v = unit_gain ? (double)acc[p][t] : __dmul_rn((double)acc[p][t], a.gain);
                    if (st < 0 && poly_near_nonzero_integer(v, a.guard_thr)) {
                        v = __dmul_rn(poly_reference_order_sum(a, xc, hc, o), a.gain);
                        atomicAdd(a.guard_count, 1ULL);
                    }
                    r16 = poly_finish(v);
                }
            }
            outv[p] = r16;
        }
        int16_t *yp = a.y + (long long)ch * a.y_stride + (obase - a.o0);
        if (all_valid && (reinterpret_cast<uintptr_t>(yp) & 15u) == 0 && TP % 8 == 0) {
#pragma unroll
            for (int p = 0; p + 8 <= TP; p += 8) {
                uint4 w;
                w.x = (uint16_t)outv[p + 0] | ((uint32_t)(uint16_t)outv[p + 1] << 16);
                w.y = (uint16_t)outv[p + 2] | ((uint32_t)(uint16_t)outv[p + 3] << 16);
                w.z = (uint16_t)outv[p + 4] | ((uint32_t)(uint16_t)outv[p + 5] << 16);
                w.w = (uint16_t)outv[p + 6] | ((uint32_t)(uint16_t)outv[p + 7] << 16);
                *reinterpret_cast<uint4 *>(yp + p) = w;
            }
        } else if (all_valid && (reinterpret_cast<uintptr_t>(yp) & 3u) == 0) {
            {
#pragma unroll
                for (int p = 0; p < TP; p += 2)
                    *reinterpret_cast<uint32_t *>(yp + p) = (uint16_t)outv[p] | ((uint32_t)(uint16_t)outv[p + 1] << 16);
            }
        } else {
#pragma unroll
            for (int p = 0; p < TP; ++p) {
                const long long o = obase + p;
                if (pg * TP + p < pbv && o >= a.o0 && o < o_end) yp[p] = outv[p];
            }
        }
    }
}

// ---- FP64 tensor-core variant (exact mode) -----------------------------------------------------------------------
//
// Same tile, staging and chunk pipeline as poly_bank_kernel, but the 64-phase x 128-cycle product Y = G'^T X' is
// issued as DMMA.8x8x4 (mma.sync.m8n8k4.f64).  Why, on B200: (1) tools/probe_pipes.cu measures DMMA at 37.1 TFLOP/s
// against 34.1 for DFMA, and the two do not overlap, so the tensor path is simply the faster way into the same FP64
// units; (2) a warp tile of 32 phases x 32 cycles needs only 32 accumulator doubles (64 registers) and loads 64
// bytes per thread per 128 FMAs (0.5 B/FMA against 1.25 for the 8x8 register tile), which lifts both limiters ncu
// found for the scalar version: the shared-memory -> register return path and, through the register count, the
// number of resident warps.  The result is still IEEE FP64 multiply-add in some order, so the near-integer guard
// makes the int16 output bit-identical exactly as before.
//
// 8 warps = 2 (phases) x 4 (cycles).  Per k4 step a thread loads 4 A values (G'[k0 + lane%4][m0 + 8*mi + lane/4])
// and one 8-byte vector of 4 int16 B values (X'[k0 + lane%4][n0 + 4*(lane/4) .. +4]); n-tile `ni` of a warp owns the
// cycles n0 + 4*col + ni, so those four samples are the thread's B operand for ni = 0..3.  Row pitches of G' (PB + 4
// doubles) and X' (JB + 8 int16) make both loads bank-conflict-free.
//
// WN = warps along the cycle axis (4, 2 or 1: CTA tile 64 phases x 128 / 64 / 32 cycles, 256 / 128 / 64 threads).  The
// narrow tiles exist for small calls -- one drop-in frame of config C1 is 160 phases x 160 cycles, six 64 x 128 tiles
// on 148 SMs -- where the launch is latency-bound and more, smaller CTAs finish sooner (launch_bank_dmma picks).
constexpr int kDmmaThreads = 256;
constexpr int kDPB = 64, kDJB = 128;                             // CTA tile: phases x cycles (WN = 4; also the fp16 variant's)
constexpr int kGP = kDPB + 4;                                    // G' row pitch in doubles  (== 4 mod 16)
template <int WN>
struct DmmaXP { static constexpr int value = WN == 4 ? 136 : 72; };   // X' row pitch in int16 (== 4 words mod 32)

template <int KC, int WN>
constexpr size_t dmma_stage_bytes() { return (size_t)KC * (kGP * 8 + DmmaXP<WN>::value * 2); }

template <int KC, int MODE, int WN>
__global__ void __launch_bounds__(64 * WN, 8 / WN)
poly_bank_dmma_kernel(PolyLaunch a, BankGeom geo)
{
    constexpr int PB = kDPB, JB = 32 * WN, NT = 64 * WN, kXP = DmmaXP<WN>::value;
    static_assert(WN == 4 || WN == 2 || WN == 1, "cycle-axis warps");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *Gs = reinterpret_cast<double *>(smem_raw + 16);               // [kStages][KC][kGP]
    int16_t *Xs = reinterpret_cast<int16_t *>(Gs + kStages * KC * kGP);   // [kStages][KC][kXP]
    int16_t *raw = Xs + kStages * KC * kXP + kRawSlack;
    __shared__ int s_shift[PB];
    __shared__ uint64_t s_full[kStages], s_empty[kStages];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;                              // warp tile: phases [32*wm, +32) x cycles [32*wn, +32)
    const int tile_p = blockIdx.x % geo.n_phase_tiles;
    const int tile_j = blockIdx.x / geo.n_phase_tiles;
    const int ch = blockIdx.y;
    const int L = a.L, M = a.M, Q = a.ctaps;

    const int l0 = tile_p * PB;
    const int pbv = min(PB, L - l0);
    const int c_lo = (int)(((long long)l0 * M) / L);
    const int c_hi = (int)(((long long)(l0 + pbv - 1) * M) / L);
    const int cspan = c_hi - c_lo;
    const int KP = Q + cspan;
    const long long j0 = geo.jc0 + (long long)tile_j * JB;
    const int rawn = (JB - 1) * M + cspan + Q;
    const long long S0 = j0 * M + c_lo - (Q - 1);

    const int16_t *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
    const int16_t *hc = a.hist ? a.hist + (long long)ch * a.hist_len : nullptr;

    // ---- stage the input span (as in poly_bank_kernel) ----  [dmma]
    // cycles of this tile that hold outputs of the call, and the samples they read
    const long long jc_last = (a.o0 + a.n_out - 1) / L;
    const int jv = (int)min((long long)JB, jc_last - j0 + 1);
    const int need = min(rawn, (jv - 1) * M + cspan + Q);
    bool bulk;
    const int raw_off = poly_stage_span<NT>(a, xc, hc, S0, need, raw, bar, tid, &bulk);
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&s_full[i], NT);
            mbar_init(&s_empty[i], NT);
        }
    }
    for (int l = tid; l < PB; l += NT)
        s_shift[l] = (l < pbv) ? c_hi - (int)(((long long)(l0 + l) * M) / L) : 0;
    __syncthreads();
    if (bulk) mbar_wait(bar, 0);
    const int16_t *rawp = raw + raw_off;

    // ---- chunk builders: thread -> one G' column (4 rows apart) and one X' column (2 rows apart) ----
    constexpr int GR = NT / PB, GE = KC / GR;                              // WN rows per pass
    constexpr int XR = NT / JB, XE = KC / XR;                              // 2 rows per pass
    static_assert(KC % GR == 0 && KC % XR == 0 && KC % 4 == 0, "chunk shape");
    const int gl = tid % PB, gk = tid / PB;
    const double *gsrc = a.cbankT64 + ((long long)gk - s_shift[gl]) * L + l0 + gl;
    const int xj = tid % JB, xk = tid / JB;
    const int16_t *xsrc = rawp + xj * M + (KP - 1) - xk;
    double gpre[GE];
    auto load_g = [&](int chunk) {
        const double *src = gsrc + (long long)chunk * KC * L;
#pragma unroll
        for (int i = 0; i < GE; ++i) gpre[i] = src[(long long)i * GR * L];
    };
    auto store_g = [&](int buf) {
        double *dst = Gs + buf * KC * kGP + gk * kGP + gl;
#pragma unroll
        for (int i = 0; i < GE; ++i) dst[i * GR * kGP] = gpre[i];
    };
    auto build_x = [&](int chunk, int buf) {
        int16_t *dst = Xs + buf * KC * kXP + xk * kXP + xj;
        const int16_t *src = xsrc - chunk * KC;
#pragma unroll
        for (int i = 0; i < XE; ++i) dst[i * XR * kXP] = src[-i * XR];
    };

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    const int nchunks = (KP + KC - 1) / KC;
#pragma unroll
    for (int c0 = 0; c0 < 2; ++c0) {
        if (c0 < nchunks) {
            load_g(c0);
            store_g(c0);
            build_x(c0, c0);
            mbar_arrive(&s_full[c0]);
        }
    }
    const int a_off = (lane & 3) * kGP + wm * 32 + (lane >> 2);          // A: row k0 + lane%4, phase 32*wm + 8*mi + lane/4
    const int b_off = (lane & 3) * kXP + wn * 32 + 4 * (lane >> 2);      // B: row k0 + lane%4, cycles 32*wn + 4*(lane/4) + ni
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c % kStages;
        const int nxt = (c + 2) % kStages;
        const bool produce = c + 2 < nchunks;
        if (produce) load_g(c + 2);
        mbar_wait(&s_full[buf], (c / kStages) & 1);
        const double *gb = Gs + buf * KC * kGP + a_off;
        const int16_t *xb = Xs + buf * KC * kXP + b_off;
#pragma unroll
        for (int k0 = 0; k0 < KC; k0 += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) af[mi] = gb[k0 * kGP + 8 * mi];
            const uint2 w = *reinterpret_cast<const uint2 *>(xb + k0 * kXP);
            bf[0] = (double)(short)(w.x & 0xffffu); bf[1] = (double)(short)(w.x >> 16);
            bf[2] = (double)(short)(w.y & 0xffffu); bf[3] = (double)(short)(w.y >> 16);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1]) : "d"(af[mi]), "d"(bf[ni]));
        }
        mbar_arrive(&s_empty[buf]);
        if (produce) {
            if (c >= 1) mbar_wait(&s_empty[nxt], ((c - 1) / kStages) & 1);
            store_g(nxt);
            build_x(c + 2, nxt);
            mbar_arrive(&s_full[nxt]);
        }
    }

    // ---- gain / guard / saturate / truncate / store: C[row = lane/4][col = 2*(lane%4) + e] of tile (mi, ni) ----
    // Two passes: the first is straight-line (every output finished and stored, near-integer hits only noted), so its
    // 64 outputs overlap in the pipes; the guard's reference-order recompute, which fires on ~1e-8 of the outputs, runs
    // in a second pass that almost no thread enters.  (One pass with the recompute inside made every output a
    // serial chain behind a branch: 11 k cycles of a 30 k-cycle drop-in frame.)
    const long long o_end = a.o0 + a.n_out;
    const bool unit_gain = a.gain == 1.0;
    int16_t *ych = a.y + (long long)ch * a.y_stride;
    uint32_t guard_hits = 0;                                               // bit (mi*4 + ni)*2 + e
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int l = wm * 32 + 8 * mi + (lane >> 2);                      // phase within the tile
        const bool l_ok = l < pbv;
        const int st = l_ok ? a.single_tap[l0 + l] : 0;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = wn * 32 + 4 * (2 * (lane & 3) + e) + ni;    // cycle within the tile
                const long long o = (j0 + j) * (long long)L + l0 + l;
                const bool valid = l_ok && o >= a.o0 && o < o_end;
                const double v = unit_gain ? acc[mi][ni][e] : __dmul_rn(acc[mi][ni][e], a.gain);
                if (MODE == LLZ_CUDA_ACC_F64 && valid && st < 0 && poly_near_nonzero_integer(v, a.guard_thr))
                    guard_hits |= 1u << ((mi * 4 + ni) * 2 + e);
                if (valid) ych[o - a.o0] = poly_finish(v);
            }
        }
    }
    while (guard_hits) {
        const int idx = __ffs(guard_hits) - 1;
        guard_hits &= guard_hits - 1;
        const int mi = idx >> 3, ni = (idx >> 1) & 3, e = idx & 1;
        const int l = wm * 32 + 8 * mi + (lane >> 2);
        const int j = wn * 32 + 4 * (2 * (lane & 3) + e) + ni;
        const long long o = (j0 + j) * (long long)L + l0 + l;
        ych[o - a.o0] = poly_finish(__dmul_rn(poly_reference_order_sum(a, xc, hc, o), a.gain));
        atomicAdd(a.guard_count, 1ULL);
    }
}

template <int KC, int MODE, int WN>
int launch_bank_dmma_tile(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    constexpr int JB = 32 * WN;
    BankGeom geo{};
    geo.jc0 = a.o0 / a.L;
    const long long jc_last = (a.o0 + a.n_out - 1) / a.L;
    geo.n_cycle_tiles = (int)((jc_last - geo.jc0 + 1 + JB - 1) / JB);
    geo.n_phase_tiles = (a.L + kDPB - 1) / kDPB;
    const int cspan_max = (int)(((long long)kDPB * a.M) / a.L) + 2;
    geo.raw_cap = (JB - 1) * a.M + cspan_max + a.ctaps + 16;
    const size_t smem = 16 + kStages * dmma_stage_bytes<KC, WN>() + ((((size_t)geo.raw_cap + kRawSlack) * 2 + 15) & ~(size_t)15);
    if (smem > 226 * 1024) return 0;
    auto kern = poly_bank_dmma_kernel<KC, MODE, WN>;
    if (smem > 48 * 1024)
        LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = (long long)geo.n_cycle_tiles * geo.n_phase_tiles;
    if (blocks > 0x7fffffffLL) { llz_set_error("resample launch too large"); return -1; }
    kern<<<dim3((unsigned)blocks, (unsigned)n_channels), 64 * WN, smem, stream>>>(a, geo);
    note_launch("poly_bank_dmma_kernel");
    LLZ_CUDA_TRY(cudaGetLastError());
    return 1;
}

// Cycle tile by size of the call: 128 cycles when that already gives every SM a few CTAs (the throughput shape), else 64
// or 32 cycles so that a small call -- a drop-in frame -- spreads over more SMs.
template <int KC, int MODE>
int launch_bank_dmma(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    const long long cycles = (a.o0 + a.n_out - 1) / a.L - a.o0 / a.L + 1;
    const long long per_cycle_tile = (long long)((a.L + kDPB - 1) / kDPB) * n_channels;
    const int sms = device_sm_count();
    if (sms <= 0) return -1;
    int wn = 4;
    while (wn > 1 && (cycles + 32 * wn - 1) / (32 * wn) * per_cycle_tile < 2LL * sms) wn >>= 1;
    if (wn == 4) return launch_bank_dmma_tile<KC, MODE, 4>(a, n_channels, stream);
    if (wn == 2) return launch_bank_dmma_tile<KC, MODE, 2>(a, n_channels, stream);
    return launch_bank_dmma_tile<KC, MODE, 1>(a, n_channels, stream);
}

// ---- FP16 tensor-core variant (fast mode, ACC_F32) --------------------------------------------------------------
//
// The fast mode of the same 64 x 128 tile on the tensor cores (mma.sync.m16n8k16, f16 operands, f32 accumulate), used
// because ncu shows the CUDA-core f32 version bound by the shared-memory return path and issue slots at 40 % of the
// FFMA pipe (profiles/r01_c4_f32_ncu_full.txt), i.e. the long-tap bank is FMA-side bound, not HBM bound.
// Operands are split so that every product is EXACT: a sample x = xh + xl with xh = x & ~255 and xl = x & 255, both
// exactly representable in fp16; a tap g*2^e = gh + gl (two fp16 planes prepared by the host, 22 significant bits).
// Four MMAs (xh*gh, xh*gl, xl*gh, xl*gl) accumulate into one f32 tile, so the only inexact steps are the f32
// accumulation -- the same as the CUDA-core fast mode -- and the 2^-22 tap truncation.  Knife-edge (single-tap)
// phases are recomputed exactly in the epilogue as in the other fast kernels.  Tolerance: |diff| <= 1 LSB.
// B200 rate of this legacy tensor path: 558 TFLOP/s (tools/probe_pipes.cu), i.e. 139 TFLOP/s of useful FMAs after
// the four-way split, against 72 TFLOP/s for the FFMA pipe.
//
// 8 warps = 2 (phases) x 4 (cycles), warp tile 32 x 32 = 2 m16 x 4 n8.  G' planes are staged [k][phase] and X' planes
// [k][cycle] (k-major rows), both read with ldmatrix.trans; row pitches (72 / 136 halves) keep ldmatrix conflict-free.
constexpr int kHGP = kDPB + 8;                                   // G' fp16 row pitch (144 B: 4 banks per row)
constexpr int kHXP = kDJB + 8;                                   // X' fp16 row pitch (272 B: 4 banks per row)

template <int KC>
constexpr size_t hmma_stage_bytes() { return (size_t)KC * (2 * kHGP + 2 * kHXP) * 2; }

__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void *p)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}

__device__ __forceinline__ void hmma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KC>
__global__ void __launch_bounds__(kDmmaThreads, 2)
poly_bank_hmma_kernel(PolyLaunch a, BankGeom geo)
{
    constexpr int PB = kDPB, JB = kDJB, NT = kDmmaThreads;
    static_assert(KC == 16, "one m16n8k16 step per chunk row block");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    uint16_t *Gh = reinterpret_cast<uint16_t *>(smem_raw + 16);           // per stage: Gh[KC][kHGP], Gl, Xh[KC][kHXP], Xl
    constexpr int kStageHalves = KC * (2 * kHGP + 2 * kHXP);
    int16_t *raw = reinterpret_cast<int16_t *>(Gh + kStages * kStageHalves) + kRawSlack;
    __shared__ int s_shift[PB];
    __shared__ uint64_t s_full[kStages], s_empty[kStages];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int tile_p = blockIdx.x % geo.n_phase_tiles;
    const int tile_j = blockIdx.x / geo.n_phase_tiles;
    const int ch = blockIdx.y;
    const int L = a.L, M = a.M, Q = a.ctaps;

    const int l0 = tile_p * PB;
    const int pbv = min(PB, L - l0);
    const int c_lo = (int)(((long long)l0 * M) / L);
    const int c_hi = (int)(((long long)(l0 + pbv - 1) * M) / L);
    const int cspan = c_hi - c_lo;
    const int KP = Q + cspan;
    const long long j0 = geo.jc0 + (long long)tile_j * JB;
    const int rawn = (JB - 1) * M + cspan + Q;
    const long long S0 = j0 * M + c_lo - (Q - 1);

    const int16_t *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
    const int16_t *hc = a.hist ? a.hist + (long long)ch * a.hist_len : nullptr;

    // ---- stage the input span (as in poly_bank_kernel) ----
    // cycles of this tile that hold outputs of the call, and the samples they read
    const long long jc_last = (a.o0 + a.n_out - 1) / L;
    const int jv = (int)min((long long)JB, jc_last - j0 + 1);
    const int need = min(rawn, (jv - 1) * M + cspan + Q);
    bool bulk;
    const int raw_off = poly_stage_span<NT>(a, xc, hc, S0, need, raw, bar, tid, &bulk);
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&s_full[i], NT);
            mbar_init(&s_empty[i], NT);
        }
    }
    for (int l = tid; l < PB; l += NT)
        s_shift[l] = (l < pbv) ? c_hi - (int)(((long long)(l0 + l) * M) / L) : 0;
    __syncthreads();
    if (bulk) mbar_wait(bar, 0);
    const int16_t *rawp = raw + raw_off;

    // ---- chunk builders ----
    constexpr int GR = NT / PB, GE = KC / GR;                              // 4 rows per pass, 4 passes
    constexpr int XR = NT / JB, XE = KC / XR;                              // 2 rows per pass, 8 passes
    const int gl = tid % PB, gk = tid / PB;
    const long long goff = ((long long)gk - s_shift[gl]) * L + l0 + gl;
    const int xj = tid % JB, xk = tid / JB;
    const int16_t *xsrc = rawp + xj * M + (KP - 1) - xk;
    uint16_t gph[GE], gpl[GE];
    auto load_g = [&](int chunk) {
        const long long o = goff + (long long)chunk * KC * L;
#pragma unroll
        for (int i = 0; i < GE; ++i) {
            gph[i] = a.cbankT16h[o + (long long)i * GR * L];
            gpl[i] = a.cbankT16l[o + (long long)i * GR * L];
        }
    };
    auto store_g = [&](int buf) {
        uint16_t *dh = Gh + buf * kStageHalves + gk * kHGP + gl;
        uint16_t *dl = dh + KC * kHGP;
#pragma unroll
        for (int i = 0; i < GE; ++i) {
            dh[i * GR * kHGP] = gph[i];
            dl[i * GR * kHGP] = gpl[i];
        }
    };
    auto build_x = [&](int chunk, int buf) {
        __half *dh = reinterpret_cast<__half *>(Gh + buf * kStageHalves + 2 * KC * kHGP) + xk * kHXP + xj;
        __half *dl = dh + KC * kHXP;
        const int16_t *src = xsrc - chunk * KC;
#pragma unroll
        for (int i = 0; i < XE; ++i) {
            const int x = src[-i * XR];
            dh[i * XR * kHXP] = __short2half_rn((short)(x & ~255));       // multiple of 256, |.| <= 32768: exact in fp16
            dl[i * XR * kHXP] = __ushort2half_rn((unsigned short)(x & 255));
        }
    };

    float acc[2][4][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mi][ni][e] = 0.f;

    const int nchunks = (KP + KC - 1) / KC;
#pragma unroll
    for (int c0 = 0; c0 < 2; ++c0) {
        if (c0 < nchunks) {
            load_g(c0);
            store_g(c0);
            build_x(c0, c0);
            mbar_arrive(&s_full[c0]);
        }
    }
    // ldmatrix row addresses.  A (m16 x k16, from [k][phase]): matrix id = lane/8 -> (m block = id&1, k block = id>>1).
    const int a_row = ((lane >> 4) << 3) + (lane & 7), a_col = wm * 32 + (((lane >> 3) & 1) << 3);
    // B (k16 x n8 pairs, from [k][cycle]): one x4 covers two n-tiles: id -> (k block = id&1, n tile = id>>1)
    const int b_row = (((lane >> 3) & 1) << 3) + (lane & 7), b_col = wn * 32 + ((lane >> 4) << 3);
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c % kStages;
        const int nxt = (c + 2) % kStages;
        const bool produce = c + 2 < nchunks;
        if (produce) load_g(c + 2);
        mbar_wait(&s_full[buf], (c / kStages) & 1);
        const uint16_t *sGh = Gh + buf * kStageHalves, *sGl = sGh + KC * kHGP;
        const uint16_t *sXh = sGl + KC * kHGP, *sXl = sXh + KC * kHXP;
        uint32_t ah[2][4], al[2][4], bh[2][4], bl[2][4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
            ldsm_x4_trans(ah[mi], sGh + a_row * kHGP + a_col + 16 * mi);
            ldsm_x4_trans(al[mi], sGl + a_row * kHGP + a_col + 16 * mi);
        }
#pragma unroll
        for (int np = 0; np < 2; ++np) {                                   // n-tile pair (2*np, 2*np+1)
            ldsm_x4_trans(bh[np], sXh + b_row * kHXP + b_col + 16 * np);
            ldsm_x4_trans(bl[np], sXl + b_row * kHXP + b_col + 16 * np);
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int np = ni >> 1, q = (ni & 1) * 2;
                hmma_16816(acc[mi][ni], al[mi], bl[np][q], bl[np][q + 1]);   // smallest products first
                hmma_16816(acc[mi][ni], al[mi], bh[np][q], bh[np][q + 1]);
                hmma_16816(acc[mi][ni], ah[mi], bl[np][q], bl[np][q + 1]);
                hmma_16816(acc[mi][ni], ah[mi], bh[np][q], bh[np][q + 1]);
            }
        mbar_arrive(&s_empty[buf]);
        if (produce) {
            if (c >= 1) mbar_wait(&s_empty[nxt], ((c - 1) / kStages) & 1);
            store_g(nxt);
            build_x(c + 2, nxt);
            mbar_arrive(&s_full[nxt]);
        }
    }

    // ---- epilogue: C[row = lane/4 (+8)][col = 2*(lane%4) + {0,1}] of tile (mi, ni) ----
    const long long o_end = a.o0 + a.n_out;
    const double unscale = ldexp(a.gain, -a.bank16_exp);
    int16_t *ych = a.y + (long long)ch * a.y_stride;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int l = wm * 32 + 16 * mi + 8 * h + (lane >> 2);
            const bool l_ok = l < pbv;
            const int st = l_ok ? a.single_tap[l0 + l] : -1;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = wn * 32 + 8 * ni + 2 * (lane & 3) + e;
                    const long long o = (j0 + j) * (long long)L + l0 + l;
                    if (!l_ok || o < a.o0 || o >= o_end) continue;
                    double v = (double)acc[mi][ni][2 * h + e] * unscale;
                    if (st >= 0) {                                         // knife-edge phase: one exact FP64 product
                        const long long base = (o * M) / L;
                        v = __dmul_rn(__dmul_rn((double)poly_sample(a, xc, hc, base - st), a.cbank[(long long)(l0 + l) * Q + st]), a.gain);
                    }
                    ych[o - a.o0] = poly_finish(v);
                }
        }
}

template <int KC>
int launch_bank_hmma(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    BankGeom geo{};
    geo.jc0 = a.o0 / a.L;
    const long long jc_last = (a.o0 + a.n_out - 1) / a.L;
    geo.n_cycle_tiles = (int)((jc_last - geo.jc0 + 1 + kDJB - 1) / kDJB);
    geo.n_phase_tiles = (a.L + kDPB - 1) / kDPB;
    const int cspan_max = (int)(((long long)kDPB * a.M) / a.L) + 2;
    geo.raw_cap = (kDJB - 1) * a.M + cspan_max + a.ctaps + 16;
    const size_t smem = 16 + kStages * hmma_stage_bytes<KC>() + ((((size_t)geo.raw_cap + kRawSlack) * 2 + 15) & ~(size_t)15);
    if (smem > 226 * 1024) return 0;
    auto kern = poly_bank_hmma_kernel<KC>;
    if (smem > 48 * 1024)
        LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = (long long)geo.n_cycle_tiles * geo.n_phase_tiles;
    if (blocks > 0x7fffffffLL) { llz_set_error("resample launch too large"); return -1; }
    kern<<<dim3((unsigned)blocks, (unsigned)n_channels), kDmmaThreads, smem, stream>>>(a, geo);
    note_launch("poly_bank_hmma_kernel");
    LLZ_CUDA_TRY(cudaGetLastError());
    return 1;
}

template <typename TA, int TP, int TJ, int PG, int CG, int KC, bool XI16>
size_t bank_smem(const PolyLaunch &a)
{
    constexpr int PB = PG * TP, JB = CG * TJ;
    const int cspan_max = (int)(((long long)PB * a.M) / a.L) + 2;
    const size_t raw_cap = (size_t)(JB - 1) * a.M + cspan_max + a.ctaps + 16;
    return 16 + (size_t)kStages * KC * (PB * sizeof(TA) + JB * (XI16 ? 2 : sizeof(TA))) + (((raw_cap + kRawSlack) * 2 + 15) & ~(size_t)15);
}

template <typename TA, int TP, int TJ, int PG, int CG, int KC, bool XI16, int MODE>
int launch_bank(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    constexpr int PB = PG * TP, JB = CG * TJ, kBankThreads = PG * CG;
    BankGeom geo{};
    geo.jc0 = a.o0 / a.L;
    const long long jc_last = (a.o0 + a.n_out - 1) / a.L;
    const long long cycles = jc_last - geo.jc0 + 1;
    geo.n_cycle_tiles = (int)((cycles + JB - 1) / JB);
    geo.n_phase_tiles = (a.L + PB - 1) / PB;
    const int cspan_max = (int)(((long long)PB * a.M) / a.L) + 2;
    geo.raw_cap = (JB - 1) * a.M + cspan_max + a.ctaps + 16;
    const size_t smem = bank_smem<TA, TP, TJ, PG, CG, KC, XI16>(a);
    auto kern = poly_bank_kernel<TA, TP, TJ, PG, CG, KC, XI16, MODE>;
    if (smem > 48 * 1024)
        LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = (long long)geo.n_cycle_tiles * geo.n_phase_tiles;
    if (blocks > 0x7fffffffLL) { llz_set_error("resample launch too large"); return -1; }
    kern<<<dim3((unsigned)blocks, (unsigned)n_channels), kBankThreads, smem, stream>>>(a, geo);
    note_launch("poly_bank_kernel");
    LLZ_CUDA_TRY(cudaGetLastError());
    return 1;
}

}  // namespace

// 1 = launched, 0 = not applicable (caller falls back to the general kernel), -1 = error
int poly_bank_launch(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    if (a.shift != 0 || a.frame_len != 0 || a.acc == LLZ_CUDA_ACC_F64_STRICT) return 0;
    if (a.L < 16) return 0;                                    // few phases: the phase tile would be mostly padding
    if (a.bank_pad < (int)(64.0 * a.M / a.L) + 2 + 32) return 0;   // transposed bank not padded for 64-phase tiles
    // padded work K'/Q must stay reasonable: c-span of a 64-phase tile against the taps per phase
    const double cspan = 64.0 * a.M / a.L;
    if ((a.ctaps + cspan) / a.ctaps > 2.5) return 0;
    constexpr size_t kLimit = 226 * 1024;
    if (a.acc == LLZ_CUDA_ACC_F32) {
        // fast mode: exact-product fp16 split on the tensor cores unless the handle asks for the FFMA tile
        // (llz_cuda_resample_bank_set_tiles).
        if (a.cbankT16h && a.cbankT16l && a.tiles != LLZ_CUDA_TILES_CUDA_CORE) {
            const int rc = launch_bank_hmma<16>(a, n_channels, stream);
            if (rc != 0) return rc;
        }
        // f32: X' as float (the conversion unit cannot feed 8 I2F per 64 FFMA), 16 phases x 4 cycles per thread
        if (bank_smem<float, 16, 4, 4, 32, 16, false>(a) > kLimit) return 0;
        return launch_bank<float, 16, 4, 4, 32, 16, false, LLZ_CUDA_ACC_F32>(a, n_channels, stream);
    }
    // f64: exact integer evaluation on the INT8 tensor cores (llz_cuda_polybank_imma.cu) unless the handle asks for
    // another tile family (llz_cuda_resample_bank_set_tiles) or the span of a tile does not fit beside its stages
    // (extreme M / L)
    // (large calls never get here: the shim runs them on the tcgen05 kernel, llz_cuda_polybank_umma.cu)
    if (a.imma_tiles && (a.tiles == LLZ_CUDA_TILES_AUTO || a.tiles == LLZ_CUDA_TILES_INT8 || a.tiles == LLZ_CUDA_TILES_INT8_TCGEN05)) {
        const int rc = poly_bank_imma_launch(a, n_channels, stream);
        if (rc != 0) return rc;
    }
    // f64: FP64 tensor-core tiles (DMMA.8x8x4) unless the handle asks for the scalar DFMA tile
    if (a.tiles != LLZ_CUDA_TILES_CUDA_CORE) {
        const int rc = launch_bank_dmma<16, LLZ_CUDA_ACC_F64>(a, n_channels, stream);
        if (rc != 0) return rc;
    }
    // scalar tile: X' as int16 (I2F.F64 in registers), 8 phases x 8 cycles per thread: 80 bytes loaded per 64 DFMA
    if (bank_smem<double, 8, 8, 8, 16, 8, true>(a) > kLimit) return 0;
    return launch_bank<double, 8, 8, 8, 16, 8, true, LLZ_CUDA_ACC_F64>(a, n_channels, stream);
}

}  // namespace llz
