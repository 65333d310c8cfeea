// llz_cuda_polybank_umma.cu -- the exact mode of the phase-bank resampler on the 5th-generation tensor cores of sm_100a:
// tcgen05.mma.kind::i8 issued by one thread, operands staged by TMA, accumulators in tensor memory.
//
// Same arithmetic as llz_cuda_polybank_imma.cu (libllzfilter/llz_resample.c:583-603 as an exact integer evaluation: taps
// as five signed base-256 digit planes, samples as a low (u8) and a high (s8) byte plane, the ten digit products
// accumulated exactly in s32, products of equal weight sharing an accumulator; then one conversion, the two-level
// near-integer guard and the reference's truncation) -- but the products are UTCIMMA instructions: a 128-cycle x
// 64-phase tile per CTA, M = 128, N = 64, K = 32 bytes per instruction, six accumulators of 64 TMEM columns.
//
//   * A operand (samples).  The MMA wants K-major rows; row j of a tile is the run of input bytes starting at sample
//     j*M + c_lo - (Q-1), and TMA traps on an unaligned row start or innermost box coordinate (tools/probe_umma_i8.cu).
//     The kernel therefore runs the bank replicated r times (llz_umma_tables.h: L' = r L, M' = r M, the same outputs) with
//     M' a multiple of 16, which puts every row start on a 16-byte boundary of the PLAIN byte planes of the input: the
//     operand of a (phase tile, chunk) is one box {128 bytes x 128 rows} of a 4-D tensor map whose rows overlap (row
//     stride M' bytes), landing in shared memory as sixteen SWIZZLE_128B atoms.  A pre-pass (poly_split_planes_kernel)
//     splits the int16 input of a slab into its low and high byte planes once -- 2 bytes written per sample -- and
//     resolves history, zeros beyond the input and the call's ragged ends, so the tile kernel has no edge cases on its
//     input side.  (The first versions wrote one expanded row per cycle instead: 5.7 bytes per sample for config C4.)
//   * B operand (taps): host-built digit planes in the same swizzled layout (llz_umma_tables.h), one bulk copy per chunk.
//   * Warp roles: warp 0 = TMA producer (one thread), warp 1 = MMA issuer (one thread) and TMEM owner, warps 4-11 =
//     epilogue (tcgen05.ld 32x32b: one accumulator row = one cycle per thread, 32 consecutive phases = 64 contiguous
//     output bytes; the accumulators are released as soon as they are combined in registers, so the rest of the
//     epilogue overlaps the next tile's MMAs).  Three-stage full/empty mbarrier ring between producer and issuer
//     (tcgen05.commit frees a stage), a full/empty pair on the accumulators between issuer and epilogue.  Persistent
//     grid, phase tiles fastest.
//   * The fast mode (LLZ_CUDA_ACC_F32) is the same kernel with three digit planes (22-bit taps, four accumulators) and
//     no guard: every product and sum is still exact, the only error is the tap rounding (<= 1 LSB of the output).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "llz_poly_device.cuh"
#include "llz_umma_tables.h"

namespace llz {

namespace {

struct UmmaGeom {
    long long jc0;                 // first cycle of the slab (absolute)
    int n_cycles;                  // cycles of the slab = rows of the expanded operand
    int n_cycle_tiles, n_phase_tiles, n_channels;
    int nchunk_max;                // chunks per phase tile in the tap tables
    long long plane_len;           // bytes of one channel's byte plane of the slab (a multiple of 16)
    // Frames.  llz_interp restarts its window at every input frame (samples at or beyond the end of an output's own frame
    // read as zero, llz_resample.c:515-523), so its planes hold the frames one after the other, each followed by zeros:
    // frame f at byte f * frame_pitch, rows_per_frame rows each.  Everything else is one frame.
    long long frame_pitch;         // bytes between frames in a plane (a multiple of 16); plane_len when there is one frame
    long long frame_samples;       // samples of a frame that come from the stream (the rest of the pitch is zeros)
    int n_frames, rows_per_frame;
    int band;                      // row blocks (cycle tile x channel) per band of the tile walk
    long long weight_sum;          // sum of a.umma_weight over the phase tiles
    int n_stages;                  // A-operand stages that fit beside the resident taps (2..4)
};

constexpr int kUThreads = 384, kUEpiThreads = 256;   // warpgroup 0: producer + issuer (96 registers), warpgroups 1-2: epilogue (200)
constexpr int kUTmemCols = 512;
constexpr int kUParkCol = 384;              // first of the 128 columns that hold the drained sums (two per output)
static_assert((kUPlanesExact + 1) * kUPB <= kUParkCol, "accumulators overlap the parking columns");
static_assert(2 * (kUPlanesFast + 1) * kUPB <= kUTmemCols, "two accumulator sets of the fast mode exceed tensor memory");
constexpr int kUMaxStages = 4;

// ---- pre-pass: byte planes --------------------------------------------------------------------------------------------
// planes[plane][channel][f * frame_pitch + e] = byte `plane` of X(S_base + f * frame_samples + e) for e < frame_samples,
// zero for the rest of the pitch;  S_base = jc0*M - (Q-1) + shift,  X = the stream sample of llz_poly_kernels.h.  One
// CTA splits kSplitSpan consecutive bytes of one frame of one channel: the span arrives by a bulk copy (poly_stage_span:
// history, zeros and ragged ends are resolved there) and leaves as two coalesced byte streams.  With interleaved PCM input
// (llz_cuda_resample_bank_run_pcm: s16 / s24 / f32 frames of several channels) the same stage de-interleaves and converts:
// the threads gather the channel's samples out of the frames, so no planar copy of the input is ever written.
constexpr int kSplitSpan = 8192, kSplitThreads = 256;

__global__ void __launch_bounds__(kSplitThreads)
poly_split_planes_kernel(PolyLaunch a, UmmaGeom geo, unsigned char *planes, int spans_per_frame)
{
    __shared__ __align__(16) int16_t raw[kSplitSpan + 16];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int ch = blockIdx.y;
    const long long f = blockIdx.x / spans_per_frame;
    const long long e0 = (long long)(blockIdx.x % spans_per_frame) * kSplitSpan;    // first byte of the span within its frame
    const int span = (int)min((long long)kSplitSpan, geo.frame_pitch - e0);          // a multiple of 16
    const int need = (int)max(0LL, min((long long)span, geo.frame_samples - e0));    // samples that come from the stream
    const int16_t *xc = poly_channel_base(a, ch);              // planar row, or the channel's samples inside interleaved PCM frames
    const int16_t *hc = a.hist ? a.hist + (long long)ch * a.hist_len : nullptr;
    const long long S0 = geo.jc0 * (long long)a.M - (a.ctaps - 1) + a.shift + f * geo.frame_samples + e0;
    bool bulk = false;
    int off = 0;
    if (need > 0) off = poly_stage_span<kSplitThreads>(a, xc, hc, S0, need, raw, &bar, tid, &bulk);
    __syncthreads();
    if (bulk) mbar_wait(&bar, 0);
    unsigned char *lo = planes + (size_t)ch * geo.plane_len + f * geo.frame_pitch + e0;
    unsigned char *hi = lo + (size_t)geo.n_channels * geo.plane_len;
    for (int e = 4 * tid; e < span; e += 4 * kSplitThreads) {
        uint32_t lw = 0, hw = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t v = (e + b < need) ? (uint32_t)(uint16_t)raw[off + e + b] : 0u;
            lw |= (v & 255u) << (8 * b);
            hw |= (v >> 8) << (8 * b);
        }
        *reinterpret_cast<uint32_t *>(lo + e) = lw;
        *reinterpret_cast<uint32_t *>(hi + e) = hw;
    }
}

// ---- tcgen05 helpers ---------------------------------------------------------------------------------------------------
// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row atoms of 128 bytes, 1024 bytes between atoms): bits 0-13
// start address >> 4, 16-29 leading byte offset (unused for swizzled K-major operands: 1), 32-45 stride byte offset >> 4
// (1024: the next 8-row group), 46 descriptor version of sm_100, 61-63 swizzle mode (2 = 128 B).  See umma_desc_lo /
// kUDescHi below.
// instruction descriptor: s32 accumulators; A = u8 (0) or s8 (1), B = s8; both K-major; N = 64, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int a_signed)
{
    return (2u << 4) | ((uint32_t)a_signed << 7) | (1u << 10) | ((uint32_t)(kUPB >> 3) << 17) | ((uint32_t)(kUJB >> 4) << 24);
}

// The descriptors differ in their low word only (start address >> 4, LBO); the high word (SBO, version, swizzle) is the
// same for every operand.  Passing 32-bit words keeps the per-MMA address arithmetic in 32-bit uniform registers.
constexpr uint32_t kUDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint32_t desc_a_lo, uint32_t desc_b_lo, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n}\n"
                 ::"r"(tmem_d), "r"(desc_a_lo), "r"(desc_b_lo), "r"(idesc), "r"(accumulate), "r"(kUDescHi) : "memory");
}

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}

// arrives on the barrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, int c4, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                   "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

// Profiling build (make EXTRA_NVFLAGS=-DLLZ_UMMA_TRACE, tools/umma_trace.py): clock stamps of CTA 0's first tiles per
// role, read back through llz_debug_umma_trace().  Compiled out of the product library.
#ifdef LLZ_UMMA_TRACE
__device__ long long g_umma_trace[16 * 16 * 16];
__device__ long long g_umma_cta[2 * 148];      // per CTA: clock at entry and at exit of the last launch
__device__ int g_umma_trace_cta = 74;
#define UTRACE(r, n, k) do { if (blockIdx.x == g_umma_trace_cta && (n) < 16 && lane == 0) g_umma_trace[((r) * 16 + (n)) * 16 + (k)] = clock64(); } while (0)
#else
#define UTRACE(r, n, k) do { } while (0)
#endif

// The reference's finish step (llz_resample.c:594-601: scale, saturate, truncate toward zero) and the near-integer test of
// the guard, on the exact integer sum.  The gain is folded into the digit planes (q = round(g * gain * 2^s)), so the
// output value is T * 2^-s, a dyadic rational: shifted to 32.32 fixed point it splits into floor and fraction with integer
// instructions only.  (The FP64 version -- I2F.F64, a multiply, F2I.F64 twice per output -- cost more than the tile's
// MMAs: those conversions issue at a quarter of the DADD rate on B200, tools/probe_conv.cu.)  The bits dropped by the
// shift matter only within 2^-32 of an integer, which is inside the guard band (exact mode) and below the tap rounding
// (fast mode).
__device__ __forceinline__ int umma_finish(uint32_t lo, uint32_t hi, int ush, uint32_t thr32, bool *near_nonzero_integer)
{
    const long long T = (long long)(((unsigned long long)hi << 32) | lo);
    const long long U = ush >= 0 ? (T >> ush) : (T << -ush);   // warp-uniform choice
    const int nf = (int)(U >> 32);                             // floor(v)
    const uint32_t f = (uint32_t)U;                            // fraction * 2^32
    const bool up = (int)f < 0;                                // fraction >= 1/2
    const int n = nf + (int)up;                                // nearest integer
    const uint32_t dist = up ? 0u - f : f;
    *near_nonzero_integer = n != 0 && dist < thr32;
    const int t = nf + (int)(nf < 0 && f != 0u);               // toward zero
    return min(max(t, -32768), 32767);
}

// first sample index an output must read as zero: the end of its own input frame (llz_interp), nothing otherwise
__device__ __forceinline__ long long umma_frame_end(const PolyLaunch &a, long long o)
{
    return a.frame_len > 0 ? ((o * a.M) / a.L / a.frame_len + 1) * (long long)a.frame_len : LLONG_MAX;
}

// Single-tap output: the integer sum is q * x exactly (q = round(g gain 2^s)); recover x = sum / q, redo the reference's two
// products (x * g, then * gain), saturate and truncate (llz_resample.c:594-601).  No division and no conversion instruction
// (inv = 2^-s / (g gain) is computed once per phase tile; int64 -> double through the 2^52 trick, rounding and truncation
// through magic-number additions): an interpolator by L has one such output in L, a call per output to a division routine
// made llz_interp's finish pass three times as long as its MMAs.
__device__ __forceinline__ int umma_single_tap(uint32_t lo, uint32_t hi, double inv, double g, double gain)
{
    const double kMagic = 6755399441055744.0;                  // 1.5 * 2^52
    const double dhi = __hiloint2double(0x43300000, (int)(hi ^ 0x80000000u)) - 4503601774854144.0;   // 2^52 + 2^31
    const double dlo = __hiloint2double(0x43300000, (int)lo) - 4503599627370496.0;                   // 2^52
    const double dT = fma(dhi, 4294967296.0, dlo);             // exact: |q x| < 2^53
    const double xr = (dT * inv + kMagic) - kMagic;            // rint: |x| <= 2^15
    const double v = __dmul_rn(__dmul_rn(xr, g), gain);
    if (!(fabs(v) < 32769.0)) return v > 0.0 ? 32767 : -32768;
    const double w = v + kMagic;
    const int n = __double2loint(w);                           // rint(v)
    const double d = v - (w - kMagic);
    const int t = n - (int)(d < 0.0 && n > 0) + (int)(d > 0.0 && n < 0);   // toward zero
    return min(max(t, -32768), 32767);
}

struct UmmaTile {
    int ch, tile_p, tile_j;
    UmmaPhaseTile pt;
};

// The tile walk.  A row block (cycle tile x channel, index r) is read by every phase tile; a phase tile's taps should stay
// resident for many tiles.  So the slab is cut into BANDS of geo.band row blocks; within a band the tiles are numbered
// phase-major, t = p * band_width + (r - band_start), and every CTA takes the same contiguous share of every band.  All CTAs
// are then inside one band at any time -- its rows (tens of MB) are read from DRAM once and from L2 by all phase tiles --
// and a CTA meets the same one or two phase tiles band after band (taps reloaded at most twice per band).  (Walking
// the whole slab phase-major instead spread the 148 CTAs over as many different row blocks: with the 80 phase tiles of
// config C4 every tile's samples then came from DRAM, 4.5 GB per 300 MB of planes by ncu.)
// The walk costs a few additions per tile: the divisions (five 64-bit ones for a tile's coordinates and its phase-tile
// geometry) are done once per band share and once per phase tile -- per tile they took longer than llz_interp's MMAs.
// Shares are cut by COST, not by tile count: a.umma_weight[p] (host: K steps of the phase tile, + 25 % when it contains a
// knife-edge phase, whose epilogue is slower) -- a CTA keeps its phase tile, so with equal counts the CTAs that own the
// expensive tiles finished 23 % after the mean (per-CTA clocks of the profiling build) and set the kernel's time.
template <typename Body>
__device__ __forceinline__ void umma_walk(const PolyLaunch &a, const UmmaGeom &geo, const int *weight, Body &&body)
{
    const long long R = (long long)geo.n_cycle_tiles * geo.n_channels;
    UmmaTile T;
    T.tile_p = -1;
    for (long long b0 = 0; b0 < R; b0 += geo.band) {
        const long long bw = min((long long)geo.band, R - b0);
        const long long wtot = geo.weight_sum * bw;
        const UmmaCut lo = umma_locate(weight, geo.n_phase_tiles, bw, wtot * blockIdx.x / gridDim.x);
        const UmmaCut hi = blockIdx.x + 1 == gridDim.x ? UmmaCut{geo.n_phase_tiles, 0}
                                                       : umma_locate(weight, geo.n_phase_tiles, bw, wtot * (blockIdx.x + 1) / gridDim.x);
        int p = lo.p;
        long long rb = lo.rb;                                  // row block within the band
        T.tile_j = (int)((b0 + rb) % geo.n_cycle_tiles);
        T.ch = (int)((b0 + rb) / geo.n_cycle_tiles);
        const int tj0 = (int)(b0 % geo.n_cycle_tiles), ch0 = (int)(b0 / geo.n_cycle_tiles);   // first row block of the band
        while (p < hi.p || (p == hi.p && rb < hi.rb)) {
            if (p != T.tile_p) { T.tile_p = p; T.pt = umma_phase_tile(a.L, a.M, a.ctaps, p); }
            body(T);
            if (++rb == bw) { rb = 0; ++p; T.tile_j = tj0; T.ch = ch0; }
            else if (++T.tile_j == geo.n_cycle_tiles) { T.tile_j = 0; ++T.ch; }
        }
    }
}

// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer and TMEM owner (both walk their loops as whole
// warps -- warp-uniform control flow keeps descriptors and coordinates in uniform registers -- and one elected lane
// issues), warps 4-11 = epilogue.
//
// Shared memory: the taps of the CTA's current phase tile stay RESIDENT (nchunks x PLANES x 8 KB: 120 KB for C4's exact
// mode), only the sample operand streams through a ring of 32 KB stages.  (The first versions streamed the taps with the
// samples: 221 KB from L2 per 8192 outputs, and clock stamps showed the issuer waiting ~2400 cycles for every 74 KB
// stage -- 30 B/clk/SM, the whole chip at the L2's throughput limit -- while the MMAs of a stage take 1900.)
//
// Epilogue warp e reads TMEM lane quarter e & 3 (= warp index % 4, the quarter the hardware lets it read) and the column
// half e >> 2: one accumulator row = one cycle per thread, 32 consecutive phases = 64 contiguous output bytes.  A pass
// has two parts: DRAIN -- tcgen05.ld the PLANES + 1 accumulators, combine them into one 64-bit integer per output and
// convert once (32 doubles per thread stay in registers) -- after which the thread releases the accumulators, and
// FINISH -- near-integer test, truncation, stores, the guard's second look -- which runs while the issuer is already
// multiplying the next tile.  (The first version ran the whole epilogue on four warps between two tiles.)
template <int PLANES, bool GUARD>
__global__ void __launch_bounds__(kUThreads, 1)
poly_bank_umma_kernel(const __grid_constant__ CUtensorMap rows_map, PolyLaunch a, UmmaGeom geo)
{
    constexpr int kBStage = umma_b_stage(PLANES);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *taps = smem_raw;                                            // nchunk_max x [PLANES x 8 KB]
    unsigned char *stages = smem_raw + (size_t)geo.nchunk_max * kBStage;        // n_stages x [A lo | A hi]
    uint64_t *s_full = reinterpret_cast<uint64_t *>(stages + (size_t)geo.n_stages * kUAStage);
    uint64_t *s_empty = s_full + kUMaxStages;
    // accumulator hand-over.  Exact mode (six accumulators, one set): t_full[0] = "the tile's sums are complete", and one
    // acc_free barrier PER ACCUMULATOR, released as soon as the epilogue has read that accumulator -- the issuer walks the
    // first chunk plane-major, so it needs accumulator i + 1 only when it reaches digit plane i and starts the next tile
    // after a third of the drain instead of all of it.  Fast mode (four accumulators): TWO sets of 256 columns,
    // t_full[s] / t_empty[s]: the epilogue of a tile runs entirely under the MMAs of the next one.
    uint64_t *t_full = s_empty + kUMaxStages, *t_empty = t_full + 2, *acc_free = t_full + 4;
    uint64_t *b_full = acc_free + (kUPlanesExact + 1), *b_free = b_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(b_free + 1);
    int *s_weight = reinterpret_cast<int *>(tmem_slot + 2);                     // [n_phase_tiles]: tile costs for the walk (read per band)
    constexpr bool kTwoSets = PLANES == kUPlanesFast;          // 2 x 4 x 64 columns fit, 2 x 6 x 64 do not

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);            // warp-uniform for the compiler, too
#ifdef LLZ_UMMA_TRACE
    if (tid == 0 && blockIdx.x < 148) g_umma_cta[2 * blockIdx.x] = clock64();
#endif
    const int S = geo.n_stages;

    for (int i = tid; i < geo.n_phase_tiles; i += kUThreads) s_weight[i] = __ldg(a.umma_weight + i);
    if (tid == 0) {
        for (int i = 0; i < kUMaxStages; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], kUEpiThreads); }
        for (int i = 0; i <= kUPlanesExact; ++i) mbar_init(&acc_free[i], kUEpiThreads);
        mbar_init(b_full, 1);
        mbar_init(b_free, 1);
    }
    __syncwarp();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kUTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;

    // the epilogue's accumulator-major drain keeps 32 64-bit sums live: its two warpgroups take the registers that the
    // warpgroup of the two single-thread roles does not need
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    if (warp == 0) {
        // ================================ TMA producer ================================
        const bool leader = elect_one();
        int buf = 0, run = 0, prev_p = -1;
        uint32_t ph = 0, tile_n = 0;
        umma_walk(a, geo, s_weight, [&](const UmmaTile &T) {
            if (T.tile_p != prev_p) {
                // new phase tile: its taps replace the resident ones once every MMA of the previous run has read them
                if (run > 0) mbar_wait(b_free, (uint32_t)((run - 1) & 1));
                if (leader) {
                    mbar_expect_tx(b_full, (uint32_t)(T.pt.nchunks * kBStage));
                    for (int c = 0; c < T.pt.nchunks; ++c)
                        tma_bulk_g2s(taps + (size_t)c * kBStage, a.umma_tiles + ((size_t)T.tile_p * geo.nchunk_max + c) * kBStage,
                                     (uint32_t)kBStage, b_full);
                }
                __syncwarp();
                prev_p = T.tile_p;
                ++run;
            }
            const int j0 = T.tile_j * kUJB;
            const int f0 = j0 / geo.rows_per_frame, i0 = j0 - f0 * geo.rows_per_frame;   // first frame and row of the tile's box
            for (int c = 0; c < T.pt.nchunks; ++c) {
                mbar_wait(&s_empty[buf], ph ^ 1u);                             // passes at once on the first lap
                UTRACE(0, tile_n, c);
                if (leader) {
                    unsigned char *st = stages + (size_t)buf * kUAStage;
                    mbar_expect_tx(&s_full[buf], (uint32_t)kUAStage);
                    const int b0 = T.pt.w0 + kUKC * c;
                    tma_load_5d(st, &rows_map, b0, i0, f0, T.ch, 0, &s_full[buf]);
                    tma_load_5d(st + kUAPlane, &rows_map, b0, i0, f0, T.ch, 1, &s_full[buf]);
                }
                __syncwarp();
                if (++buf == S) { buf = 0; ph ^= 1u; }
            }
            ++tile_n;
        });
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        // With the loops inside `if (lane == 0)` every MMA sat in a divergence "waterfall" (ELECT / R2UR.BROADCAST /
        // BRA.U.ANY: ~20 instructions per MMA); now ten UTCIMMA follow one another.
        const bool leader = elect_one();
        int buf = 0, run = 0, prev_p = -1;
        uint32_t ph = 0, tile_n = 0;
        const uint32_t taps_desc = umma_desc_lo(smem_u32(taps));
        umma_walk(a, geo, s_weight, [&](const UmmaTile &T) {
            if (T.tile_p != prev_p) {
                if (run > 0 && leader) umma_commit(b_free);                    // the old taps are free once the MMAs so far are done
                __syncwarp();
                mbar_wait(b_full, (uint32_t)(run & 1));
                prev_p = T.tile_p;
                ++run;
            }
            const uint32_t set = kTwoSets ? (tile_n & 1u) : 0u;
            const uint32_t tacc = tmem + set * 256u;
            if constexpr (kTwoSets) {
                // the epilogue has finished with this set (two tiles ago; passes at once for the first two tiles)
                mbar_wait(&t_empty[set], ((tile_n >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
            }
            UTRACE(1, tile_n, 0);
            for (int c = 0; c < T.pt.nchunks; ++c) {
                mbar_wait(&s_full[buf], ph);
                UTRACE(1, tile_n, 1 + c);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t sa = smem_u32(stages) + (uint32_t)buf * kUAStage;
                const uint32_t d_lo = umma_desc_lo(sa), d_hi = umma_desc_lo(sa + kUAPlane);
                const uint32_t d_b = taps_desc + (uint32_t)c * (kBStage >> 4);
                const int ks_n = min(4, T.pt.ksteps - 4 * c);
                if (leader) {
                    // plane-major: digit plane i of every K step of the chunk, then plane i + 1 (sums are exact: any order)
#pragma unroll
                    for (int i = 0; i < PLANES; ++i) {
                        if (!kTwoSets && c == 0) {
                            // accumulators i and i + 1 of the previous tile have been read (passes at once for the first tile)
                            if (i == 0) mbar_wait(&acc_free[0], (tile_n & 1u) ^ 1u);
                            mbar_wait(&acc_free[i + 1], (tile_n & 1u) ^ 1u);
                            asm volatile("tcgen05.fence::after_thread_sync;");
                        }
                        const uint32_t b_i = d_b + (uint32_t)(i * (kUBPlane >> 4));
                        for (int ks = 0; ks < ks_n; ++ks) {
                            const uint32_t koff = (uint32_t)(2 * ks);          // 32 bytes of K per step, >> 4
                            const bool first = c == 0 && ks == 0;
                            // digit i x low byte -> weight 256^i, digit i x high byte -> weight 256^(i+1)
                            umma_i8(tacc + kUPB * i, d_lo + koff, b_i + koff, umma_idesc(0), (first && i == 0) ? 0u : 1u);
                            umma_i8(tacc + kUPB * (i + 1), d_hi + koff, b_i + koff, umma_idesc(1), first ? 0u : 1u);
                        }
                    }
                    umma_commit(&s_empty[buf]);                                // the stage is free once these MMAs have read it
                }
                __syncwarp();
                if (++buf == S) { buf = 0; ph ^= 1u; }
            }
            if (leader) umma_commit(&t_full[set]);                             // the tile's accumulators are complete
            __syncwarp();
            UTRACE(1, tile_n, 8);
            ++tile_n;
        });
    }
    } else {
        // ================================ epilogue ================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int q = warp & 3;                                // TMEM lane quarter this warp may read
        const int h = (warp - 4) >> 2;                         // column half: phases 32h .. 32h + 31 of the tile
        const int m = 32 * q + lane;                           // accumulator row = cycle within the tile
        const int L = a.L, M = a.M, Q = a.ctaps;
        const long long o_end = a.o0 + a.n_out;
        uint32_t tile_n = 0;
        int st_p = -1, my_st = -1;
        double my_g = 0.0, my_inv = 0.0, st_g0 = 0.0, st_inv0 = 0.0;
        unsigned st_mask = 0, st_slow = 0;
        bool st_same = true;
        umma_walk(a, geo, s_weight, [&](const UmmaTile &T) {
            const int l0 = T.pt.l0 + 32 * h, pbv = max(0, min(32, T.pt.pbv - 32 * h));
            const long long j = geo.jc0 + (long long)T.tile_j * kUJB + m;          // this thread's cycle
            const int16_t *xc = poly_channel_base(a, T.ch);
            const int16_t *hc = a.hist ? a.hist + (long long)T.ch * a.hist_len : nullptr;
            int16_t *ych = a.y + (long long)T.ch * a.y_stride;
            // Single-tap (knife-edge) phases of this half tile, one bit per phase; lane e keeps the tap of phase l0 + e.  Such
            // an output is the reference's x * g * gain, two roundings (g = 1 - 2^-53 for an L-th band prototype: x - 1 after
            // truncation for positive x).  The integer sum of such a row is q * x exactly, q = round(g gain 2^s), so x is
            // recovered from it and the two products are redone in the reference's order -- no memory access.  (Fetching x
            // and g per output from global memory made the warps that own phase 0 the slowest part of the whole kernel:
            // 8700 cycles of finish per tile against 1800.)  A tap too small for the division to recover x exactly keeps
            // the slow path.
            if (T.tile_p != st_p) {                            // once per phase tile, not per tile (two dependent loads)
                st_p = T.tile_p;
                my_st = lane < pbv ? __ldg(a.single_tap + l0 + lane) : -1;
                my_g = my_st >= 0 ? __ldg(a.cbank + (long long)(l0 + lane) * Q + my_st) : 0.0;
                my_inv = my_st >= 0 ? a.umma_scale / (my_g * a.gain) : 0.0;
                st_mask = __ballot_sync(0xffffffffu, my_st >= 0);
                st_slow = __ballot_sync(0xffffffffu, my_st >= 0 && !(fabs(my_g * a.gain) >= a.umma_scale * 1048576.0));
                // the usual case (rows of one knife-edge phase, replicated): one tap value for all of them, no shuffles
                const int first = st_mask ? __ffs((int)st_mask) - 1 : 0;
                st_g0 = __shfl_sync(0xffffffffu, my_g, first);
                st_inv0 = __shfl_sync(0xffffffffu, my_inv, first);
                st_same = __all_sync(0xffffffffu, my_st < 0 || my_g == st_g0);
            }

            const uint32_t set = kTwoSets ? (tile_n & 1u) : 0u;
            UTRACE(warp - 2, tile_n, 0);
            mbar_wait(&t_full[set], kTwoSets ? ((tile_n >> 1) & 1u) : (tile_n & 1u));
            UTRACE(warp - 2, tile_n, 1);
            __syncwarp();                                      // the lanes leave the wait loop one by one; tcgen05.ld is .aligned
            asm volatile("tcgen05.fence::after_thread_sync;");

            const uint32_t trow = tmem + set * 256u + ((uint32_t)(32 * q) << 16) + 32 * h;
            const uint32_t tpark = tmem + ((uint32_t)(32 * q) << 16) + kUParkCol + 64 * h;
            if constexpr (!kTwoSets) {
                // ---- drain (exact mode): accumulator by accumulator -> one 64-bit integer per output, each accumulator
                // released as soon as it is read; the sums are then parked in the 128 TMEM columns the accumulators leave free
                // (two columns per output).  Keeping them in registers through the finish pass instead needs that pass unrolled
                // 32 times; that code (and the other roles' loops) no longer fit the instruction cache, and clock stamps
                // showed 60 cycles per output.
                long long T64[32];
#pragma unroll
                for (int d = 0; d <= PLANES; ++d) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {     // sixteen columns at a time: 64 + 16 live registers
                        uint32_t acc[16];
                        tmem_ld16(trow + kUPB * d + 16 * half, acc);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            T64[16 * half + e] = d == 0 ? (long long)(int)acc[e] : T64[16 * half + e] + ((long long)(int)acc[e] << (8 * d));
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    mbar_arrive(&acc_free[d]);
                }
#pragma unroll
                for (int cg = 0; cg < 4; ++cg) {
                    uint32_t park[16];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { park[2 * e] = (uint32_t)T64[8 * cg + e]; park[2 * e + 1] = (uint32_t)(T64[8 * cg + e] >> 32); }
                    tmem_st16(tpark + 16 * cg, park);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            UTRACE(warp - 2, tile_n, 2);

            // ---- finish ----
            const long long o_row = j * (long long)L + l0;     // output index of this thread's first phase
            const bool row_in = (m < geo.n_cycles - T.tile_j * kUJB);
            const bool interior = row_in && pbv == 32 && o_row >= a.o0 && o_row + 31 < o_end;
            int16_t *yrow = ych + (o_row - a.o0);
            const bool vec_ok = interior && ((reinterpret_cast<uintptr_t>(yrow) & 15u) == 0);
            unsigned hits = 0;                                 // near-integer outputs of this row, one bit per phase
            const unsigned st_fast = st_mask & ~st_slow;
#pragma unroll 1
            for (int cg = 0; cg < 4; ++cg) {
                uint32_t park[16];
                if constexpr (kTwoSets) {
                    // fast mode: straight from this set's accumulators (the other set is being multiplied into)
                    int acc[PLANES + 1][8];
#pragma unroll
                    for (int d = 0; d <= PLANES; ++d) tmem_ld8(trow + kUPB * d + 8 * cg, acc[d]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        long long t64 = (long long)acc[0][e];
#pragma unroll
                        for (int d = 1; d <= PLANES; ++d) t64 += (long long)acc[d][e] << (8 * d);
                        park[2 * e] = (uint32_t)t64;
                        park[2 * e + 1] = (uint32_t)(t64 >> 32);
                    }
                } else {
                    tmem_ld16(tpark + 16 * cg, park);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                }
                int y[8];
                unsigned hit8 = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    bool hit;
                    y[e] = umma_finish(park[2 * e], park[2 * e + 1], a.umma_ush, a.umma_thr32, &hit);
                    if (GUARD) hit8 |= (unsigned)hit << e;
                }
                if ((st_fast >> (8 * cg)) & 255u) {            // warp-uniform, one tile in L / 64
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if ((st_fast >> (8 * cg + e)) & 1u) {
                            const double g = st_same ? st_g0 : __shfl_sync(0xffffffffu, my_g, 8 * cg + e);
                            const double inv = st_same ? st_inv0 : __shfl_sync(0xffffffffu, my_inv, 8 * cg + e);
                            y[e] = umma_single_tap(park[2 * e], park[2 * e + 1], inv, g, a.gain);
                        }
                }
                uint32_t packed[4];
#pragma unroll
                for (int e = 0; e < 8; e += 2) packed[e >> 1] = (uint32_t)(uint16_t)y[e] | ((uint32_t)y[e + 1] << 16);
                if (vec_ok) {
                    // streaming store: the outputs must not push the band's samples (read by every phase tile) out of L2
                    __stcs(reinterpret_cast<uint4 *>(yrow) + cg, make_uint4(packed[0], packed[1], packed[2], packed[3]));
                } else if (row_in) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int l = 8 * cg + e;
                        const long long o = o_row + l;
                        const bool valid = l < pbv && o >= a.o0 && o < o_end;
                        if (valid) ych[o - a.o0] = (int16_t)(packed[e >> 1] >> (16 * (e & 1)));
                        else hit8 &= ~(1u << e);
                    }
                }
                hits |= hit8 << (8 * cg);
            }
            if constexpr (kTwoSets) {                          // this set's accumulators have been read: free for the tile after next
                asm volatile("tcgen05.fence::before_thread_sync;");
                mbar_arrive(&t_empty[set]);
            }
            if (!row_in) hits = 0;

            // knife-edge phases (one tap, 1 - 2^-53 for the L-th band prototype): one exact FP64 product per output
            hits &= ~st_mask;
            if (row_in && st_slow) {
                unsigned sm = st_slow;
                while (sm) {
                    const int l = __ffs((int)sm) - 1;
                    sm &= sm - 1;
                    const long long o = o_row + l;
                    if (o < a.o0 || o >= o_end) continue;
                    const int st = __ldg(a.single_tap + l0 + l);
                    const long long sidx = (o * M) / L + a.shift - st;
                    const double xv = sidx < umma_frame_end(a, o) ? (double)poly_sample(a, xc, hc, sidx) : 0.0;
                    ych[o - a.o0] = poly_finish(__dmul_rn(__dmul_rn(xv, a.cbank[(long long)(l0 + l) * Q + st]), a.gain));
                }
            }
            UTRACE(warp - 2, tile_n, 3);
            ++tile_n;
            if (!GUARD) return;
            __syncwarp();

            // Second look at the outputs that came within the (wide) band of the integer evaluation: the whole warp evaluates
            // such an output again as an FP64 dot product; only what is STILL near an integer goes to the reference's own
            // serial order (llz_cuda_polybank_imma.cu has the history of this two-level scheme).
            unsigned pending = __ballot_sync(0xffffffffu, hits != 0);
            while (pending) {
                const int src = __ffs(pending) - 1;
                int l = (lane == src) ? __ffs((int)hits) - 1 : 0;
                l = __shfl_sync(0xffffffffu, l, src);
                const long long o = (geo.jc0 + (long long)T.tile_j * kUJB + 32 * q + src) * (long long)L + l0 + l;   // warp-uniform
                const long long base = (o * M) / L + a.shift;
                const long long frame_end = umma_frame_end(a, o);
                const double *row = a.cbank + (long long)(l0 + l) * Q;
                double part = 0.0;
                for (int k = lane; k < Q; k += 128) {          // four independent sample / tap loads in flight per lane
                    int xv[4];
                    double gv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int kk = k + 32 * u;
                        const bool in = kk < Q && base - kk < frame_end;
                        xv[u] = in ? poly_sample(a, xc, hc, base - kk) : 0;
                        gv[u] = in ? row[kk] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) part = fma((double)xv[u], gv[u], part);
                }
#pragma unroll
                for (int sh = 16; sh; sh >>= 1) part += __shfl_xor_sync(0xffffffffu, part, sh);
                if (lane == src) {
                    double vv = __dmul_rn(part, a.gain);
                    if (poly_near_nonzero_integer(vv, a.guard_thr)) {
                        vv = __dmul_rn(poly_reference_order_sum(a, xc, hc, o), a.gain);
                        atomicAdd(a.guard_count, 1ULL);
                    }
                    ych[o - a.o0] = poly_finish(vv);
                    hits &= hits - 1;
                }
                pending = __ballot_sync(0xffffffffu, hits != 0);
            }
        });
    }

#ifdef LLZ_UMMA_TRACE
    if (tid == 128 && blockIdx.x < 148) g_umma_cta[2 * blockIdx.x + 1] = clock64();
#endif
    // ---- teardown: every MMA has completed (the epilogue waited for the last tile), nobody touches TMEM any more ----
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kUTmemCols));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

}  // namespace

namespace {

// plane geometry of a slab of `cycles` cycles (whole tiles of 128): one frame, or llz_interp's frames with their zero tails
struct UmmaPlanes {
    long long plane_len, frame_pitch, frame_samples;
    int n_frames, rows_per_frame, ext;
    bool ok;
};

UmmaPlanes umma_planes(const PolyLaunch &a, long long cycles)
{
    UmmaPlanes g{};
    g.ext = umma_row_extent(a.L, a.M, a.ctaps);
    const long long rows = (cycles + kUJB - 1) / kUJB * kUJB;
    if (a.frame_len > 0) {
        // rows must not straddle frames, and a tile's 128 rows must be whole frames or part of one
        if (a.frame_len % a.M != 0) return g;
        g.rows_per_frame = a.frame_len / a.M;
        if (kUJB % g.rows_per_frame != 0 && g.rows_per_frame % kUJB != 0) return g;
        g.n_frames = (int)((rows + g.rows_per_frame - 1) / g.rows_per_frame);
        g.frame_samples = a.frame_len;
        g.frame_pitch = ((long long)a.frame_len + g.ext + 15) & ~15LL;
        g.plane_len = g.frame_pitch * g.n_frames;
    } else {
        if (rows > 0x7fffffffLL) return g;
        g.rows_per_frame = (int)rows;
        g.n_frames = 1;
        g.plane_len = g.frame_pitch = g.frame_samples = (rows * a.M + g.ext + 15) & ~15LL;
    }
    g.ok = true;
    return g;
}

}  // namespace

// bytes of the byte-plane workspace for a slab of `cycles` cycles (of the replicated bank a.L / a.M); 0 = not applicable
size_t poly_bank_umma_rows_bytes(const PolyLaunch &a, int n_channels, long long cycles)
{
    const UmmaPlanes g = umma_planes(a, cycles);
    return g.ok ? (size_t)2 * n_channels * (size_t)g.plane_len : 0;
}

namespace {

template <int PLANES, bool GUARD>
int umma_launch_slabs(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { llz_set_error("cuTensorMapEncodeTiled is not available from this driver"); return -1; }
    const int sms = device_sm_count();
    if (sms <= 0) return -1;
    if (a.M % 16 != 0) return 0;                               // rows must start on 16-byte boundaries (umma_replication)
    const long long jc_first = a.o0 / a.L, jc_last = (a.o0 + a.n_out - 1) / a.L;
    // shared memory: the resident taps, then as many 32 KB sample stages as fit (at least two)
    const int n_ptiles = (a.L + kUPB - 1) / kUPB;
    if (n_ptiles > 1024) return 0;
    constexpr size_t kSmemMax = 227 * 1024;
    const size_t kBarBytes = 256 + (((size_t)n_ptiles * sizeof(int) + 127) & ~(size_t)127);   // barriers, TMEM slot, tile weights
    const size_t taps_bytes = (size_t)a.umma_nchunks * umma_b_stage(PLANES);
    if (taps_bytes + 2 * kUAStage + kBarBytes + 1024 > kSmemMax) return 0;
    int n_stages = (int)((kSmemMax - kBarBytes - 1024 - taps_bytes) / kUAStage);
    if (n_stages > kUMaxStages) n_stages = kUMaxStages;
    const size_t smem = taps_bytes + (size_t)n_stages * kUAStage + kBarBytes;
    auto kern = poly_bank_umma_kernel<PLANES, GUARD>;
    LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (long long jc0 = jc_first; jc0 <= jc_last; jc0 += a.umma_slab_cycles) {
        UmmaGeom geo{};
        geo.jc0 = jc0;
        geo.n_cycles = (int)min((long long)a.umma_slab_cycles, jc_last - jc0 + 1);
        geo.n_cycle_tiles = (geo.n_cycles + kUJB - 1) / kUJB;
        geo.n_phase_tiles = (a.L + kUPB - 1) / kUPB;
        geo.n_channels = n_channels;
        geo.nchunk_max = a.umma_nchunks;
        geo.n_stages = n_stages;
        // rows of a partial last cycle tile are read too (and discarded): the planes cover whole tiles
        const UmmaPlanes pl = umma_planes(a, geo.n_cycles);
        if (!pl.ok) return 0;
        geo.plane_len = pl.plane_len;
        geo.frame_pitch = pl.frame_pitch;
        geo.frame_samples = pl.frame_samples;
        geo.n_frames = pl.n_frames;
        geo.rows_per_frame = pl.rows_per_frame;
        // band of the tile walk: row blocks whose samples (both planes) make up ~32 MB, at least 8
        const long long block_bytes = 2LL * kUJB * a.M;
        geo.weight_sum = a.umma_weight_sum;
        geo.band = (int)max(8LL, min((long long)geo.n_cycle_tiles * n_channels, ((long long)tunables().umma_band_mib << 20) / block_bytes));
        // 1. byte planes of the slab
        const int spans_per_frame = (int)((geo.frame_pitch + kSplitSpan - 1) / kSplitSpan);
        dim3 sgrid((unsigned)((long long)spans_per_frame * geo.n_frames), (unsigned)n_channels);
        poly_split_planes_kernel<<<sgrid, kSplitThreads, 0, stream>>>(a, geo, a.umma_rows, spans_per_frame);
        LLZ_CUDA_TRY(cudaGetLastError());
        note_launch("poly_split_planes_kernel");
        // 2. the tensor map over them: [plane][channel][frame][row][byte], rows overlapping (stride M bytes); a tile's box is
        // 128 rows of one frame, or whole frames of fewer rows each
        CUtensorMap map;
        const int box_rows = geo.rows_per_frame < kUJB ? geo.rows_per_frame : kUJB;
        const cuuint64_t dims[5] = {(cuuint64_t)pl.ext, (cuuint64_t)geo.rows_per_frame, (cuuint64_t)geo.n_frames, (cuuint64_t)n_channels, 2};
        const cuuint64_t strides[4] = {(cuuint64_t)a.M, (cuuint64_t)geo.frame_pitch, (cuuint64_t)geo.plane_len,
                                       (cuuint64_t)geo.plane_len * n_channels};
        const cuuint32_t box[5] = {(cuuint32_t)kUKC, (cuuint32_t)box_rows, (cuuint32_t)(kUJB / box_rows), 1, 1}, estr[5] = {1, 1, 1, 1, 1};
        const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, a.umma_rows, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { llz_set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)cr); return -1; }
        // 3. the tiles
        const long long tiles = (long long)geo.n_cycle_tiles * geo.n_phase_tiles * n_channels;
        const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);           // persistent: one CTA per SM
        kern<<<grid, kUThreads, smem, stream>>>(map, a, geo);
        LLZ_CUDA_TRY(cudaGetLastError());
        note_launch(PLANES == kUPlanesExact ? "poly_bank_umma_kernel<5>" : "poly_bank_umma_kernel<3>");
    }
    return 1;
}

}  // namespace

// 1 = launched, 0 = not applicable, -1 = error.  a.L / a.M / a.cbank / a.single_tap describe the REPLICATED bank.
int poly_bank_umma_launch(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    if (!a.umma_tiles || a.umma_nchunks <= 0 || !a.umma_rows || a.umma_slab_cycles <= 0 || !a.umma_weight || a.umma_weight_sum <= 0) return 0;
    if (a.n_out <= 0) return 0;
    if (a.acc == LLZ_CUDA_ACC_F64 && a.umma_planes == kUPlanesExact) return umma_launch_slabs<kUPlanesExact, true>(a, n_channels, stream);
    if (a.acc == LLZ_CUDA_ACC_F32 && a.umma_planes == kUPlanesFast) return umma_launch_slabs<kUPlanesFast, false>(a, n_channels, stream);
    return 0;
}

}  // namespace llz

#ifdef LLZ_UMMA_TRACE
extern "C" int llz_debug_umma_trace(long long *host_out)
{
    return cudaMemcpyFromSymbol(host_out, llz::g_umma_trace, sizeof(llz::g_umma_trace)) == cudaSuccess ? 0 : -1;
}
extern "C" int llz_debug_umma_select(int cta)
{
    return cudaMemcpyToSymbol(llz::g_umma_trace_cta, &cta, sizeof(int)) == cudaSuccess ? 0 : -1;
}
extern "C" int llz_debug_umma_cta(long long *host_out)
{
    return cudaMemcpyFromSymbol(host_out, llz::g_umma_cta, sizeof(llz::g_umma_cta)) == cudaSuccess ? 0 : -1;
}
#endif
