// llz_umma_tables.h -- tile geometry and host-side tap tables of the tcgen05 phase-bank kernel
// (llz_cuda_polybank_umma.cu).  Plain C++ (no CUDA): the kernel, the shim and the host emulation test
// (tests/cpu/umma_emulate.cpp) all read the layout from here.
//
// Tile = 128 cycles (rows of the A operand = TMEM lanes) x 64 phases (rows of the B operand = TMEM columns).
// A operand: row j (one output cycle) is the run of input bytes of the samples  s = j*M - (Q-1) + b,  two byte planes.  TMA
// traps on an unaligned innermost box coordinate or row start (tools/probe_umma_i8.cu), so the kernel runs the bank
// REPLICATED r times (L' = r L phases, M' = r M samples per cycle: the same outputs, see umma_replication) with M' a
// multiple of 16: every row then starts at a 16-byte boundary of the plain byte planes of the input and the operand is a
// tensor map with OVERLAPPING rows (row stride M' bytes < row extent) -- no expanded copy of the input exists.  A phase
// tile reads the window  b in [w0, w0 + 128*chunks)  of the rows, w0 = 16*floor(c_lo/16); the c_lo % 16 bytes in front of
// its first useful sample meet zero taps.
// B operand: per phase tile and 128-byte chunk, five (exact mode) or three (fast mode) signed base-256 digit planes of the taps, each a [64 x 128 byte]
// K-major SWIZZLE_128B block (8-row atoms of 1024 bytes, 16-byte chunk c of row r at position c ^ (r & 7)) -- the
// layout a TMA box with CU_TENSOR_MAP_SWIZZLE_128B produces and the one the shared-memory matrix descriptor names.
#pragma once

#include <math.h>
#include <stddef.h>

#include <vector>

#if defined(__CUDACC__)
#define LLZ_UMMA_HD __host__ __device__
#else
#define LLZ_UMMA_HD
#endif

namespace llz {

constexpr int kUPB = 64;                  // phases per tile (MMA N)
constexpr int kUJB = 128;                 // cycles per tile (MMA M)
constexpr int kUKC = 128;                 // k bytes per pipeline chunk = four K steps of 32
// signed base-256 digits of a tap: 5 planes (38 bits) for the exact mode, 3 planes (22 bits) for the fast mode
constexpr int kUPlanesExact = 5, kUPlanesFast = 3;
constexpr int kUBPlane = kUPB * kUKC;     // 8192 bytes: one digit plane of one chunk
constexpr int kUAPlane = kUJB * kUKC;     // 16,384 bytes: one byte plane of the samples of one chunk
constexpr int kUAStage = 2 * kUAPlane;
LLZ_UMMA_HD constexpr int umma_tap_bits(int planes) { return 8 * planes - 2; }
LLZ_UMMA_HD constexpr int umma_b_stage(int planes) { return planes * kUBPlane; }            // 40,960 / 24,576 bytes of taps per chunk

// geometry of phase tile p: everything the producer, the MMA issuer and the host table builder must agree on
struct UmmaPhaseTile {
    int l0, pbv;           // first phase, valid phases
    int c_lo, cspan;       // floor(l0*M/L) and the spread of floor(l*M/L) over the tile
    int w0, off;           // window start within an expanded row (multiple of 16) and c_lo - w0
    int ksteps, nchunks;   // K steps of 32 bytes that hold taps, chunks of 128 bytes
};

LLZ_UMMA_HD inline UmmaPhaseTile umma_phase_tile(int L, int M, int Q, int p)
{
    UmmaPhaseTile t;
    t.l0 = p * kUPB;
    t.pbv = (L - t.l0 < kUPB) ? L - t.l0 : kUPB;
    t.c_lo = (int)(((long long)t.l0 * M) / L);
    const int c_hi = (int)(((long long)(t.l0 + t.pbv - 1) * M) / L);
    t.cspan = c_hi - t.c_lo;
    t.w0 = t.c_lo & ~15;
    t.off = t.c_lo - t.w0;
    const int kk = Q + t.cspan + t.off;                        // window bytes that can meet a non-zero tap
    t.ksteps = (kk + 31) / 32;
    t.nchunks = (t.ksteps + 3) / 4;
    return t;
}

// Shares of the kernel's tile walk.  Within a band of bw row blocks the tiles are numbered phase-major; tile (p, rb) starts
// at weight  bw * sum_{p' < p} w[p'] + rb * w[p].  umma_locate returns the first tile whose starting weight is >= target, so
// consecutive targets  W c / G  (W = bw * sum w) cut the band into G shares that cover every tile exactly once
// (tests/cpu/umma_emulate.cpp checks that).
struct UmmaCut { int p; long long rb; };

LLZ_UMMA_HD inline UmmaCut umma_locate(const int *weight, int n_phase_tiles, long long bw, long long target)
{
    long long cum = 0;
    for (int p = 0; p < n_phase_tiles; ++p) {
        const long long w = weight[p];
        if (target < cum + w * bw) {
            UmmaCut c;
            c.p = p;
            c.rb = (target - cum + w - 1) / w;
            if (c.rb == bw) { c.p = p + 1; c.rb = 0; }
            return c;
        }
        cum += w * bw;
    }
    return UmmaCut{n_phase_tiles, 0};
}

// bytes of a row that the widest-reaching phase tile reads (its boxes end at w0 + 128 * chunks)
inline int umma_row_extent(int L, int M, int Q)
{
    int ext = 0;
    for (int p = 0; p * kUPB < L; ++p) {
        const UmmaPhaseTile t = umma_phase_tile(L, M, Q, p);
        if (t.w0 + kUKC * t.nchunks > ext) ext = t.w0 + kUKC * t.nchunks;
    }
    return ext;
}

// Replication factor r of the bank for this kernel: r M a multiple of 16 (aligned rows), r L >= 64 (one full phase tile),
// then doubled while that removes padding of the last phase tile (r L a multiple of 64) and the tables stay small.
inline int umma_replication(int L, int M)
{
    int r = 1;
    while (((long long)r * M) % 16 != 0) r *= 2;
    while ((long long)r * L < kUPB) r *= 2;
    while (((long long)r * L) % kUPB != 0 && (long long)r * L < 16384) r *= 2;
    return r;
}

// byte offset of element (row n, k byte kk) inside a [64 x 128] K-major SWIZZLE_128B block
LLZ_UMMA_HD inline int umma_b_offset(int n, int kk)
{
    return (n >> 3) * 1024 + (n & 7) * 128 + ((((kk >> 4) ^ (n & 7)) << 4) | (kk & 15));
}

// Host: the bank [L][Q] as int8 digit planes, [phase tile][chunk][plane][64 x 128 swizzled].  Returns the chunks per
// tile (0 when the bank cannot be split); *shift = s with g ~ q * 2^-s, *eps = bound on |sum_k (g - q 2^-s) x| for |x| <= 32768,
// *qsum_max = the largest sum_k |q| of a row (the integer sum of a row is below 32768 times that).  The caller passes the
// taps already multiplied by the gain.
inline int poly_umma_build_tables(const double *cb, int L, int M, int Q, int planes, std::vector<signed char> *out, int *shift, double *eps,
                                  double *qsum_max)
{
    double gmax = 0.0;
    for (size_t i = 0; i < (size_t)L * Q; ++i) gmax = fmax(gmax, fabs(cb[i]));
    if (!(gmax > 0.0) || !isfinite(gmax)) return 0;
    int e2 = 0;
    frexp(gmax, &e2);                                          // gmax < 2^e2
    const int s = umma_tap_bits(planes) - e2;
    const int b_stage = umma_b_stage(planes);
    const int n_tiles = (L + kUPB - 1) / kUPB;
    int nchunks = 0;
    for (int p = 0; p < n_tiles; ++p) {
        const UmmaPhaseTile t = umma_phase_tile(L, M, Q, p);
        if (t.nchunks > nchunks) nchunks = t.nchunks;
    }
    out->assign((size_t)n_tiles * nchunks * b_stage, 0);
    for (int p = 0; p < n_tiles; ++p) {
        const UmmaPhaseTile t = umma_phase_tile(L, M, Q, p);
        for (int l = 0; l < t.pbv; ++l) {
            const int d = (int)(((long long)(t.l0 + l) * M) / L) - t.c_lo;
            for (int k = 0; k < Q; ++k) {
                long long q = llrint(ldexp(cb[(size_t)(t.l0 + l) * Q + k], s));
                // window byte that meets tap k of phase l: sample index  j*M + c_l - k  sits at row byte  c_l + (Q-1) - k
                const int kappa = t.off + (Q - 1) + d - k;
                signed char *dst = out->data() + ((size_t)p * nchunks + kappa / kUKC) * b_stage + umma_b_offset(l, kappa % kUKC);
                for (int pl = 0; pl < planes; ++pl) {
                    const int dg = (int)((((q % 256) + 256 + 128) % 256) - 128);   // signed digit in [-128, 127]
                    dst[(size_t)pl * kUBPlane] = (signed char)dg;
                    q = (q - dg) / 256;
                }
                if (q != 0) return 0;                          // cannot happen for |g 2^s| < 2^(8 planes - 2)
            }
        }
    }
    *shift = s;
    double worst = 0.0, qsum = 0.0;                            // the taps' actual rounding errors and sum |q|, worst row
    for (int l = 0; l < L; ++l) {
        double row = 0.0, qs = 0.0;
        for (int k = 0; k < Q; ++k) {
            const double g = cb[(size_t)l * Q + k];
            const double qd = (double)llrint(ldexp(g, s));
            row += fabs(g - ldexp(qd, -s));
            qs += fabs(qd);
        }
        worst = fmax(worst, row);
        qsum = fmax(qsum, qs);
    }
    *qsum_max = qsum;
    *eps = 32768.0 * worst * (1.0 + 1e-9) + (double)Q * 32768.0 * ldexp(1.0, -(s + 40));
    return nchunks;
}

}  // namespace llz
