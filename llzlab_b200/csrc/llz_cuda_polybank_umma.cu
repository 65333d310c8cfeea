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
//     j*M + c_lo - (Q-1), i.e. at an arbitrary byte offset, and TMA traps on an unaligned innermost box coordinate
//     (tools/probe_umma_i8.cu).  A pre-pass (poly_expand_rows_kernel) therefore writes the input once as "expanded rows":
//     one row of RL = c_hi_max + Q bytes per output cycle and byte plane, so that every row starts aligned and the operand
//     of a (phase tile, chunk) is ONE box {128 bytes x 128 rows} of a plain 4-D tensor, landing in shared memory as
//     sixteen SWIZZLE_128B atoms.  The pre-pass also resolves history, zeros beyond the input and the call's ragged
//     ends, so the tile kernel has no edge cases on its input side.  The call is cut into slabs of cycles so that the
//     expanded rows of a slab stay small (tunable; they are re-read by the five phase tiles of a cycle tile).
//   * B operand (taps): host-built digit planes in the same swizzled layout (llz_umma_tables.h), one bulk copy per chunk.
//   * Warp roles: warp 0 = TMA producer (one thread), warp 1 = MMA issuer (one thread) and TMEM owner, warps 2-5 =
//     epilogue (tcgen05.ld 32x32b: one accumulator row = one cycle per thread, 64 consecutive phases = 128 contiguous
//     output bytes).  Three-stage full/empty mbarrier ring between producer and issuer (tcgen05.commit frees a stage),
//     a full/empty pair on the accumulators between issuer and epilogue.  Persistent grid, phase tiles fastest.
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
    int row_len;                   // RL
};

constexpr int kUThreads = 192;
constexpr int kUTmemCols = 512;

// ---- pre-pass: expanded rows -----------------------------------------------------------------------------------------
// rows[plane][channel][j][b] = byte `plane` of X((jc0 + j)*M - (Q-1) + b),  X = the stream sample of llz_poly_kernels.h
constexpr int kERows = 8, kEThreads = 256;

__global__ void __launch_bounds__(kEThreads)
poly_expand_rows_kernel(PolyLaunch a, UmmaGeom geo, unsigned char *rows)
{
    extern __shared__ __align__(16) int16_t span[];
    const int ch = blockIdx.y;
    const int j_first = blockIdx.x * kERows;
    const int nrows = min(kERows, geo.n_cycles - j_first);
    const int RL = geo.row_len;
    const int16_t *xc = a.x ? a.x + (long long)ch * a.x_stride : nullptr;
    const int16_t *hc = a.hist ? a.hist + (long long)ch * a.hist_len : nullptr;
    const long long S0 = (geo.jc0 + j_first) * (long long)a.M - (a.ctaps - 1);
    const int need = (nrows - 1) * a.M + RL;
    for (int e = threadIdx.x; e < need; e += kEThreads) span[e] = (int16_t)poly_sample(a, xc, hc, S0 + e);
    __syncthreads();
    const size_t plane_stride = (size_t)geo.n_channels * geo.n_cycles * RL;
    unsigned char *lo = rows + ((size_t)ch * geo.n_cycles + j_first) * RL;
    unsigned char *hi = lo + plane_stride;
    const int vec_per_row = RL >> 4;
    for (int v = threadIdx.x; v < nrows * vec_per_row; v += kEThreads) {
        const int r = v / vec_per_row, c = v - r * vec_per_row;
        const int16_t *src = span + r * a.M + 16 * c;
        uint32_t l[4], h[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t lw = 0, hw = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t s = (uint16_t)src[4 * w + b];
                lw |= (s & 255u) << (8 * b);
                hw |= (s >> 8) << (8 * b);
            }
            l[w] = lw; h[w] = hw;
        }
        *reinterpret_cast<uint4 *>(lo + (size_t)r * RL + 16 * c) = make_uint4(l[0], l[1], l[2], l[3]);
        *reinterpret_cast<uint4 *>(hi + (size_t)r * RL + 16 * c) = make_uint4(h[0], h[1], h[2], h[3]);
    }
}

// ---- tcgen05 helpers ---------------------------------------------------------------------------------------------------
// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row atoms of 128 bytes, 1024 bytes between atoms
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);           // start address >> 4
    d |= (uint64_t)1 << 16;                                // leading byte offset: unused for swizzled K-major operands
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset: the next 8-row group
    d |= (uint64_t)1 << 46;                                // descriptor version of sm_100
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}

// instruction descriptor: s32 accumulators; A = u8 (0) or s8 (1), B = s8; both K-major; N = 64, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int a_signed)
{
    return (2u << 4) | ((uint32_t)a_signed << 7) | (1u << 10) | ((uint32_t)(kUPB >> 3) << 17) | ((uint32_t)(kUJB >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// arrives on the barrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

struct UmmaTile {
    int ch, tile_p, tile_j;
    UmmaPhaseTile pt;
};

__device__ __forceinline__ UmmaTile umma_tile(const PolyLaunch &a, const UmmaGeom &geo, long long t)
{
    UmmaTile T;
    T.tile_p = (int)(t % geo.n_phase_tiles);                   // phase tiles fastest: neighbours share the expanded rows in L2
    const long long r = t / geo.n_phase_tiles;
    T.tile_j = (int)(r % geo.n_cycle_tiles);
    T.ch = (int)(r / geo.n_cycle_tiles);
    T.pt = umma_phase_tile(a.L, a.M, a.ctaps, T.tile_p);
    return T;
}

__global__ void __launch_bounds__(kUThreads, 1)
poly_bank_umma_kernel(const __grid_constant__ CUtensorMap rows_map, PolyLaunch a, UmmaGeom geo)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *stages = smem_raw;                                          // kUStages x [A lo | A hi | B planes]
    uint64_t *s_full = reinterpret_cast<uint64_t *>(smem_raw + kUStages * kUStage);
    uint64_t *s_empty = s_full + kUStages;
    uint64_t *t_full = s_empty + kUStages, *t_empty = t_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long total = (long long)geo.n_phase_tiles * geo.n_cycle_tiles * geo.n_channels;

    if (tid == 0) {
        for (int i = 0; i < kUStages; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        mbar_init(t_full, 1);
        mbar_init(t_empty, 128);
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

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            long long g = 0;                                                   // global chunk counter
            for (long long t = blockIdx.x; t < total; t += gridDim.x) {
                const UmmaTile T = umma_tile(a, geo, t);
                for (int c = 0; c < T.pt.nchunks; ++c, ++g) {
                    const int buf = (int)(g % kUStages);
                    if (g >= kUStages) mbar_wait(&s_empty[buf], (uint32_t)((g / kUStages - 1) & 1));
                    unsigned char *st = stages + (size_t)buf * kUStage;
                    mbar_expect_tx(&s_full[buf], (uint32_t)kUStage);
                    const int b0 = T.pt.w0 + kUKC * c, j0 = T.tile_j * kUJB;
                    tma_load_4d(st, &rows_map, b0, j0, T.ch, 0, &s_full[buf]);
                    tma_load_4d(st + kUAPlane, &rows_map, b0, j0, T.ch, 1, &s_full[buf]);
                    tma_bulk_g2s(st + kUAStage, a.umma_tiles + ((size_t)T.tile_p * geo.nchunk_max + c) * kUBStage, (uint32_t)kUBStage,
                                 &s_full[buf]);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            long long g = 0;
            uint32_t tile_n = 0;
            for (long long t = blockIdx.x; t < total; t += gridDim.x, ++tile_n) {
                const UmmaTile T = umma_tile(a, geo, t);
                // the epilogue has drained the accumulators of the previous tile (passes at once for the first tile)
                mbar_wait(t_empty, (tile_n & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                for (int c = 0; c < T.pt.nchunks; ++c, ++g) {
                    const int buf = (int)(g % kUStages);
                    mbar_wait(&s_full[buf], (uint32_t)((g / kUStages) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t sa = smem_u32(stages + (size_t)buf * kUStage);
                    const int ks_n = min(4, T.pt.ksteps - 4 * c);
                    for (int ks = 0; ks < ks_n; ++ks) {
                        const uint64_t koff = (uint64_t)((32 * ks) >> 4);
                        const uint64_t a_lo = umma_smem_desc(sa) + koff, a_hi = umma_smem_desc(sa + kUAPlane) + koff;
                        const bool first = c == 0 && ks == 0;
#pragma unroll
                        for (int i = 0; i < kUPlanes; ++i) {
                            const uint64_t b_i = umma_smem_desc(sa + kUAStage + i * kUBPlane) + koff;
                            // digit i x low byte -> weight 256^i, digit i x high byte -> weight 256^(i+1)
                            umma_i8(tmem + kUPB * i, a_lo, b_i, umma_idesc(0), (first && i == 0) ? 0u : 1u);
                            umma_i8(tmem + kUPB * (i + 1), a_hi, b_i, umma_idesc(1), first ? 0u : 1u);
                        }
                    }
                    umma_commit(&s_empty[buf]);                                // the stage is free once these MMAs have read it
                }
                umma_commit(t_full);                                           // the tile's accumulators are complete
            }
        }
    } else {
        // ================================ epilogue ================================
        const int q = warp & 3;                                // TMEM lane quarter this warp may read
        const int m = 32 * q + lane;                           // accumulator row = cycle within the tile
        const int L = a.L, M = a.M, Q = a.ctaps;
        const long long o_end = a.o0 + a.n_out;
        const double out_scale = a.imma_scale * 16777216.0;    // the high half carries 256^3
        uint32_t tile_n = 0;
        for (long long t = blockIdx.x; t < total; t += gridDim.x, ++tile_n) {
            const UmmaTile T = umma_tile(a, geo, t);
            const int l0 = T.pt.l0, pbv = T.pt.pbv;
            const long long j = geo.jc0 + (long long)T.tile_j * kUJB + m;          // this thread's cycle
            const int16_t *xc = a.x ? a.x + (long long)T.ch * a.x_stride : nullptr;
            const int16_t *hc = a.hist ? a.hist + (long long)T.ch * a.hist_len : nullptr;
            int16_t *ych = a.y + (long long)T.ch * a.y_stride;
            // single-tap (knife-edge) phases of the tile, one bit per phase
            const unsigned st_lo = __ballot_sync(0xffffffffu, lane < pbv && __ldg(a.single_tap + l0 + lane) >= 0);
            const unsigned st_hi = __ballot_sync(0xffffffffu, lane + 32 < pbv && __ldg(a.single_tap + l0 + lane + 32) >= 0);
            const unsigned long long st_mask = ((unsigned long long)st_hi << 32) | st_lo;

            mbar_wait(t_full, tile_n & 1u);
            __syncwarp();                                      // the lanes leave the wait loop one by one; tcgen05.ld is .aligned
            asm volatile("tcgen05.fence::after_thread_sync;");

            const long long o_row = j * (long long)L + l0;     // output index of this thread's phase l0
            const bool row_in = (m < geo.n_cycles - T.tile_j * kUJB);
            const bool interior = row_in && pbv == kUPB && o_row >= a.o0 && o_row + (kUPB - 1) < o_end;
            int16_t *yrow = ych + (o_row - a.o0);
            const bool vec_ok = interior && ((reinterpret_cast<uintptr_t>(yrow) & 15u) == 0);
            unsigned long long hits = 0;                       // near-integer outputs of this row, one bit per phase
            const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
#pragma unroll 1
            for (int cg = 0; cg < 4; ++cg) {
                int acc[kUPlanes + 1][16];
#pragma unroll
                for (int d = 0; d <= kUPlanes; ++d) tmem_ld16(trow + kUPB * d + 16 * cg, acc[d]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                int16_t outv[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const long long lo = (long long)acc[0][e] + (long long)acc[1][e] * 256 + (long long)acc[2][e] * 65536;
                    const long long hi = (long long)acc[3][e] + (long long)acc[4][e] * 256 + (long long)acc[5][e] * 65536;
                    const double s = fma((double)hi, out_scale, (double)lo * a.imma_scale);
                    const double v = __dmul_rn(s, a.gain);
                    if (poly_near_nonzero_integer(v, a.imma_thr)) hits |= 1ull << (16 * cg + e);
                    outv[e] = poly_finish(v);
                }
                if (vec_ok) {
                    uint4 w0, w1;
                    w0.x = (uint16_t)outv[0] | ((uint32_t)(uint16_t)outv[1] << 16);   w0.y = (uint16_t)outv[2] | ((uint32_t)(uint16_t)outv[3] << 16);
                    w0.z = (uint16_t)outv[4] | ((uint32_t)(uint16_t)outv[5] << 16);   w0.w = (uint16_t)outv[6] | ((uint32_t)(uint16_t)outv[7] << 16);
                    w1.x = (uint16_t)outv[8] | ((uint32_t)(uint16_t)outv[9] << 16);   w1.y = (uint16_t)outv[10] | ((uint32_t)(uint16_t)outv[11] << 16);
                    w1.z = (uint16_t)outv[12] | ((uint32_t)(uint16_t)outv[13] << 16); w1.w = (uint16_t)outv[14] | ((uint32_t)(uint16_t)outv[15] << 16);
                    reinterpret_cast<uint4 *>(yrow + 16 * cg)[0] = w0;
                    reinterpret_cast<uint4 *>(yrow + 16 * cg)[1] = w1;
                } else if (interior) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) yrow[16 * cg + e] = outv[e];
                } else if (row_in) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int l = 16 * cg + e;
                        const long long o = o_row + l;
                        const bool valid = l < pbv && o >= a.o0 && o < o_end;
                        if (valid) ych[o - a.o0] = outv[e];
                        else hits &= ~(1ull << l);
                    }
                }
            }
            if (!row_in) hits = 0;
            // the accumulators are in registers / stored: the issuer may start the next tile
            asm volatile("tcgen05.fence::before_thread_sync;");
            mbar_arrive(t_empty);

            // knife-edge phases (one tap, 1 - 2^-53 for the L-th band prototype): one exact FP64 product per output
            hits &= ~st_mask;
            if (row_in) {
                unsigned long long sm = st_mask;
                while (sm) {
                    const int l = __ffsll((long long)sm) - 1;
                    sm &= sm - 1;
                    const long long o = o_row + l;
                    if (o < a.o0 || o >= o_end) continue;
                    const int st = __ldg(a.single_tap + l0 + l);
                    const long long base = (o * M) / L;
                    ych[o - a.o0] = poly_finish(__dmul_rn(__dmul_rn((double)poly_sample(a, xc, hc, base - st), a.cbank[(long long)(l0 + l) * Q + st]), a.gain));
                }
            }
            __syncwarp();

            // Second look at the outputs that came within the (wide) band of the integer evaluation: the whole warp evaluates
            // such an output again as an FP64 dot product; only what is STILL near an integer goes to the reference's own
            // serial order (llz_cuda_polybank_imma.cu has the history of this two-level scheme).
            unsigned pending = __ballot_sync(0xffffffffu, hits != 0);
            while (pending) {
                const int src = __ffs(pending) - 1;
                int l = (lane == src) ? __ffsll((long long)hits) - 1 : 0;
                l = __shfl_sync(0xffffffffu, l, src);
                const long long o = (geo.jc0 + (long long)T.tile_j * kUJB + 32 * q + src) * (long long)L + l0 + l;   // warp-uniform
                const long long base = (o * M) / L;
                const double *row = a.cbank + (long long)(l0 + l) * Q;
                double part = 0.0;
                for (int k = lane; k < Q; k += 128) {          // four independent sample / tap loads in flight per lane
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
                for (int sh = 16; sh; sh >>= 1) part += __shfl_xor_sync(0xffffffffu, part, sh);
                if (lane == src) {
                    double v = __dmul_rn(part, a.gain);
                    if (poly_near_nonzero_integer(v, a.guard_thr)) {
                        v = __dmul_rn(poly_reference_order_sum(a, xc, hc, o), a.gain);
                        atomicAdd(a.guard_count, 1ULL);
                    }
                    ych[o - a.o0] = poly_finish(v);
                    hits &= hits - 1;
                }
                pending = __ballot_sync(0xffffffffu, hits != 0);
            }
        }
    }

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

size_t poly_bank_umma_rows_bytes(const PolyLaunch &a, int n_channels, long long cycles)
{
    return (size_t)2 * n_channels * (size_t)cycles * umma_row_len(a.L, a.M, a.ctaps);
}

// 1 = launched, 0 = not applicable, -1 = error
int poly_bank_umma_launch(const PolyLaunch &a, int n_channels, cudaStream_t stream)
{
    if (!a.umma_tiles || a.umma_nchunks <= 0 || !a.umma_rows || a.umma_slab_cycles <= 0) return 0;
    if (a.acc != LLZ_CUDA_ACC_F64 || a.shift != 0 || a.frame_len != 0 || a.n_out <= 0) return 0;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { llz_set_error("cuTensorMapEncodeTiled is not available from this driver"); return -1; }
    const int sms = device_sm_count();
    if (sms <= 0) return -1;
    const int RL = umma_row_len(a.L, a.M, a.ctaps);
    const long long jc_first = a.o0 / a.L, jc_last = (a.o0 + a.n_out - 1) / a.L;
    constexpr size_t smem = (size_t)kUStages * kUStage + 128;
    static_assert(smem <= 227 * 1024, "pipeline stages exceed the shared memory of an SM");
    LLZ_CUDA_TRY(cudaFuncSetAttribute(poly_bank_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (long long jc0 = jc_first; jc0 <= jc_last; jc0 += a.umma_slab_cycles) {
        UmmaGeom geo{};
        geo.jc0 = jc0;
        geo.n_cycles = (int)min((long long)a.umma_slab_cycles, jc_last - jc0 + 1);
        geo.n_cycle_tiles = (geo.n_cycles + kUJB - 1) / kUJB;
        geo.n_phase_tiles = (a.L + kUPB - 1) / kUPB;
        geo.n_channels = n_channels;
        geo.nchunk_max = a.umma_nchunks;
        geo.row_len = RL;
        // 1. expanded rows of the slab
        const size_t esm = ((size_t)(kERows - 1) * a.M + RL) * sizeof(int16_t);
        if (esm > 48 * 1024) return 0;
        dim3 egrid((unsigned)((geo.n_cycles + kERows - 1) / kERows), (unsigned)n_channels);
        poly_expand_rows_kernel<<<egrid, kEThreads, esm, stream>>>(a, geo, a.umma_rows);
        LLZ_CUDA_TRY(cudaGetLastError());
        // 2. the tensor map over them: [plane][channel][cycle][byte]
        CUtensorMap map;
        const cuuint64_t dims[4] = {(cuuint64_t)RL, (cuuint64_t)geo.n_cycles, (cuuint64_t)n_channels, 2};
        const cuuint64_t strides[3] = {(cuuint64_t)RL, (cuuint64_t)RL * geo.n_cycles, (cuuint64_t)RL * geo.n_cycles * n_channels};
        const cuuint32_t box[4] = {(cuuint32_t)kUKC, (cuuint32_t)kUJB, 1, 1}, estr[4] = {1, 1, 1, 1};
        const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, a.umma_rows, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { llz_set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)cr); return -1; }
        // 3. the tiles
        const long long tiles = (long long)geo.n_cycle_tiles * geo.n_phase_tiles * n_channels;
        const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);           // persistent: one CTA per SM
        poly_bank_umma_kernel<<<grid, kUThreads, smem, stream>>>(map, a, geo);
        LLZ_CUDA_TRY(cudaGetLastError());
    }
    return 1;
}

}  // namespace llz
