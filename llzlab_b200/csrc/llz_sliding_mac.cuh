// llz_sliding_mac.cuh -- the register-blocked sliding-window multiply-accumulate that every
// "sliding" kernel of libllzfilter_cuda is built on (direct-form FIR, and each polyphase
// branch of the decimating resampler).
//
// One thread owns R consecutive outputs  y[r] += sum_k h[k] * x[base + r - k],  r < R.
// Taps are consumed in chunks of U = 16 bytes / sizeof(T) (one LDS.128, broadcast to the warp);
// the R+U input samples a chunk touches live in registers as a ring of NS = R/U + 1 slots of U
// samples, and each chunk replaces exactly one slot with one LDS.128 -- so a chunk costs
// R*U FMAs + 2 shared-memory loads (f32, R=28: 112 FFMA : 2 LDS).
//
// R/U is odd on purpose: neighbouring lanes start R/U 16-byte vectors apart, and an odd vector
// stride makes every quarter-warp hit 8 distinct 16-byte bank groups -- conflict-free LDS.128
// with a dense (unpadded, unswizzled) tile, which is what lets the tile arrive by one TMA bulk
// copy.
//
// The chunk loop is unrolled NS deep so every ring index is a compile-time constant.
#pragma once

#include "llz_cuda_common.cuh"

namespace llz {

// S = storage type of the samples in shared memory: T itself (FIR tiles) or int16_t (the decimating resampler keeps
// its de-interleaved PCM as int16 and converts in the ring refill: a U-sample refill is then a 4- or 8-byte load,
// one shared-memory return cycle instead of four, and the tile is 4x / 2x smaller, which buys resident warps).
template <typename T, int R, typename S = T>
struct SlidingMac {
    using V = typename Vec16<T>::type;
    static constexpr int U = Vec16<T>::N;       // taps per chunk
    static constexpr int NS = R / U + 1;        // ring slots
    static constexpr int GRAN = NS * U;         // taps per unrolled iteration
    static_assert(R % U == 0, "R must be a whole number of 16-byte vectors");
    static_assert((R / U) % 2 == 1, "R/U must be odd (bank-conflict-free lane stride)");

    // U consecutive samples at p (aligned to U*sizeof(S)) -> dst[0..U)
    static __device__ __forceinline__ void load_samples(const S *p, T *dst)
    {
        if constexpr (sizeof(S) == sizeof(T)) {
            unpack(*reinterpret_cast<const V *>(p), dst);
        } else if constexpr (U == 2) {
            const uint32_t w = *reinterpret_cast<const uint32_t *>(p);
            dst[0] = (T)(short)(w & 0xffffu);
            dst[1] = (T)(short)(w >> 16);
        } else {
            static_assert(U == 4, "int16 storage: 2 or 4 samples per refill");
            const uint2 w = *reinterpret_cast<const uint2 *>(p);
            dst[0] = (T)(short)(w.x & 0xffffu);
            dst[1] = (T)(short)(w.x >> 16);
            dst[2] = (T)(short)(w.y & 0xffffu);
            dst[3] = (T)(short)(w.y >> 16);
        }
    }

    // One chunk: ring slot for logical window index j (element j of the R+U samples, lowest
    // address first) at chunk CC (mod NS) is ((j/U - CC) mod NS).
    template <int CC, bool STRICT>
    static __device__ __forceinline__ void chunk(T (&acc)[R], const T (&ring)[NS * U], const T (&h)[U])
    {
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = r - u + U;                      // 1 .. R+U-1
                const int slot = ((j / U) - CC + NS) % NS;
                acc[r] = mac<T, STRICT>(h[u], ring[slot * U + (j % U)], acc[r]);
            }
        }
    }

    template <int CC, bool STRICT>
    static __device__ __forceinline__ void step(T (&acc)[R], T (&ring)[NS * U], const S *xp, const T *hp)
    {
        // refill the slot that becomes the lowest-address slot of this chunk
        load_samples(xp - CC * U, &ring[((NS - CC) % NS) * U]);
        T h[U];
        unpack(*reinterpret_cast<const V *>(hp + CC * U), h);
        chunk<CC, STRICT>(acc, ring, h);
        if constexpr (CC + 1 < NS) step<CC + 1, STRICT>(acc, ring, xp, hp);
    }

    // the last, partial iteration: the first `chunks` (< NS) chunks of step<0>; the ring is not used afterwards
    template <int CC, bool STRICT>
    static __device__ __forceinline__ void tail(T (&acc)[R], T (&ring)[NS * U], const S *xp, const T *hp, int chunks)
    {
        load_samples(xp - CC * U, &ring[((NS - CC) % NS) * U]);
        T h[U];
        unpack(*reinterpret_cast<const V *>(hp + CC * U), h);
        chunk<CC, STRICT>(acc, ring, h);
        if constexpr (CC + 2 < NS) {
            if (CC + 1 < chunks) tail<CC + 1, STRICT>(acc, ring, xp, hp, chunks);
        }
    }

    // Accumulate `ntaps` taps; the tail of `taps` is zero-padded to a multiple of GRAN, or -- TAIL = true, which
    // costs a second copy of the unrolled code -- only to a multiple of U (45 taps per residue of config C3 run as
    // 46 instead of 48).
    //   win  : shared-memory address of sample (base - U), i.e. one vector below the first output's
    //          newest sample; aligned to U samples.  Reads reach down to win - ntaps + U ... up to win+R+U-1.
    //   taps : shared-memory tap array, taps[k] multiplies x[base + r - k]; 16-byte aligned.
    template <bool STRICT, bool TAIL = false>
    static __device__ __forceinline__ void run(T (&acc)[R], const S *win, const T *taps, int ntaps)
    {
        T ring[NS * U];
#pragma unroll
        for (int s = 1; s < NS; ++s) load_samples(win + s * U, &ring[s * U]);
        const S *xp = win;
        const T *hp = taps;
        int k = 0;
        for (; k + GRAN <= ntaps; k += GRAN) {
            step<0, STRICT>(acc, ring, xp, hp);
            xp -= GRAN;
            hp += GRAN;
        }
        if constexpr (TAIL) {
            if (k < ntaps) tail<0, STRICT>(acc, ring, xp, hp, (ntaps - k) / U);
        }
    }
};

}  // namespace llz
