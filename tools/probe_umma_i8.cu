// probe_umma_i8.cu -- the hardware assumptions of the tcgen05 phase-bank kernel (llz_cuda_polybank_umma.cu), checked one by
// one on a B200 before the kernel relies on them:
//   1. a TMA tensor map over a byte plane whose dim-1 stride (16*M bytes) is SMALLER than the dim-0 extent (overlapping
//      rows) is accepted by the driver;
//   2. the innermost box coordinate may be any byte offset (rows of the operand start at j*M, M = 147);
//   3. an {128 bytes x 8 rows} box with CU_TENSOR_MAP_SWIZZLE_128B lands as one canonical K-major SWIZZLE_128B atom
//      (16-byte chunk c of row r at chunk position c ^ r), which is also how the host lays out the tap operand;
//   4. tcgen05.mma.cta_group::1.kind::i8 with M = 128, N = 64, K = 32, A = u8 or s8 (instruction descriptor bits 7-9),
//      B = s8, s32 accumulators in TMEM; K steps by advancing the descriptor start address by 32 bytes;
//   5. tcgen05.ld.32x32b: lane = accumulator row, column = accumulator column; row m of the A tile = atom m / 8, row m % 8.
//   6. (mode 4) the A operand from tensor memory: tcgen05.cp.128x256b copies one K step (128 rows x 32 bytes) of the
//      swizzled shared-memory tile into 8 TMEM columns, tcgen05.mma [d], [a_tmem], b_desc reads it from there; the
//      products must equal mode 3's, and the copied columns are dumped (row m in lane m, bytes 4c..4c+3 of the K step in
//      column c).
// Prints PROBE_UMMA_I8_OK when the device result equals the host integer product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma_i8 probe_umma_i8.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int kM = 147;            // input step per cycle (config C4)
constexpr int kRows = 128, kN = 64, kKB = 128;   // tile: 128 cycles x 64 phases, 128 k-bytes per chunk
constexpr int kPlaneLen = 16 * kM * 12;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row atoms of 128 bytes, 1024 bytes between atoms
__device__ __forceinline__ uint64_t umma_desc(const void *p)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(p) >> 4) & 0x3FFF);         // start address
    d |= (uint64_t)1 << 16;                                // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}

// instruction descriptor: s32 accumulate, A format (0 = u8, 1 = s8), B = s8, both K-major, N = 64, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int a_signed)
{
    return (2u << 4) | ((uint32_t)a_signed << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap, const signed char *b_tiles, int first_byte, unsigned char *a_dump, int *d_out, int mode)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                         // [2 planes][16 atoms][1024]
    unsigned char *sB = smem + 2 * 16 * 1024;         // [64 rows x 128 bytes], host-swizzled
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + kN * kKB);
    uint64_t *mma_bar = bar + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) { mbar_init(bar, 1); mbar_init(mma_bar, 1); }
    __syncwarp();
    if (warp == 0 && mode >= 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = mode >= 1 ? *tmem_slot : 0u;

    if (mode == -1) return;
    if (mode == -2) {
        if (tid == 0) {
            mbar_expect_tx(bar, kN * kKB);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sB)), "l"(b_tiles), "r"(kN * kKB), "r"(smem_u32(bar)) : "memory");
        }
        mbar_wait(bar, 0);
        return;
    }
    if (tid == 0) {
        mbar_expect_tx(bar, 2 * 16 * 1024 + kN * kKB);
        for (int plane = 0; plane < 2; ++plane)
            for (int r = 0; r < 16; ++r) {
                // rows j = r + 16 i (i = 0..7) of the operand: bytes plane[first_byte + j*M + k], k < 128
                const int c0 = first_byte >= 0 ? first_byte + r * kM : 16 * r, c1 = 0, c2 = 0, c3 = plane;
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                             ::"r"(smem_u32(sA + (plane * 16 + r) * 1024)), "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                             : "memory");
            }
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sB)), "l"(b_tiles), "r"(kN * kKB), "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    for (int i = tid; i < 2 * 16 * 1024; i += 128) a_dump[i] = sA[i];
    __syncthreads();
    if (mode < 2) {
        if (warp == 0 && mode == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
        return;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) {
        for (int plane = 0; plane < 2; ++plane)
            for (int ks = 0; ks < kKB / 32; ++ks) {
                const uint64_t da = umma_desc(sA + plane * 16 * 1024) + (uint64_t)((32 * ks) >> 4);
                const uint64_t db = umma_desc(sB) + (uint64_t)((32 * ks) >> 4);
                const uint32_t idesc = umma_idesc(plane);                    // plane 0: unsigned bytes, plane 1: signed
                const uint32_t acc = ks > 0 ? 1u : 0u;
                if (mode >= 4) {
                    const uint32_t ta = tmem + 128 + 8 * (plane * 4 + ks);
                    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(ta), "l"(da) : "memory");
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n"
                                 ::"r"(tmem + 64 * plane), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
                } else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem + 64 * plane), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mma_bar)) : "memory");
    }
    mbar_wait(mma_bar, 0);
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (mode < 3) {
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
        return;
    }
    // thread (warp w, lane l) owns accumulator row 32 w + l
    for (int plane = 0; plane < 2; ++plane)
        for (int cg = 0; cg < 4; ++cg) {
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + 64 * plane + 16 * cg;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int e = 0; e < 16; ++e) d_out[(plane * kRows + tid) * kN + 16 * cg + e] = (int)v[e];
        }
    if (mode >= 4) {                                  // dump the 64 columns the copies wrote: [lane][col]
        for (int cg = 0; cg < 4; ++cg) {
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + 128 + 16 * cg;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int e = 0; e < 16; ++e) d_out[2 * kRows * kN + tid * 64 + 16 * cg + e] = (int)v[e];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main(int argc, char **argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 3;     // 0: TMA only, 1: + TMEM alloc, 2: + MMA, 3: + TMEM load and checks
    // byte planes: [2 planes][1 channel][pitch]
    const size_t pitch = ((size_t)kPlaneLen + 8 * 16 * kM + 256 + 15) & ~(size_t)15;
    std::vector<unsigned char> planes(2 * pitch);
    for (size_t i = 0; i < planes.size(); ++i) planes[i] = (unsigned char)((i * 2654435761u) >> 13);
    unsigned char *d_planes = nullptr;
    CK(cudaMalloc(&d_planes, planes.size()));
    CK(cudaMemcpy(d_planes, planes.data(), planes.size(), cudaMemcpyHostToDevice));

    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
    CUtensorMap tmap;
    const int variant = argc > 2 ? atoi(argv[2]) : 0;   // 1: dim-0 extent = dim-1 stride (no overlap; boxes that cross it are zero filled)
    const cuuint64_t dims[4] = {(cuuint64_t)(variant == 1 ? 16 * kM : kPlaneLen), 8, 1, 2};
    const cuuint64_t strides[3] = {(cuuint64_t)16 * kM, (cuuint64_t)pitch, (cuuint64_t)pitch};   // dim 1 overlaps dim 0
    const cuuint32_t box[4] = {128, 8, 1, 1}, estr[4] = {1, 1, 1, 1};
    CUresult cr = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d_planes, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 (argc > 3 && (atoi(argv[3]) & 2)) ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled (dim-1 stride %d < dim-0 extent %d): %s\n", 16 * kM, kPlaneLen, cr == CUDA_SUCCESS ? "accepted" : "REJECTED");
    if (cr != CUDA_SUCCESS) { printf("CUresult %d\n", (int)cr); return 1; }

    // taps: [64 rows][128 bytes] signed, laid out as canonical SWIZZLE_128B atoms by the host
    std::vector<signed char> b(kN * kKB), b_sw(kN * kKB);
    for (int n = 0; n < kN; ++n)
        for (int k = 0; k < kKB; ++k) {
            b[n * kKB + k] = (signed char)(((n * 131 + k * 7) * 2654435761u) >> 24);
            b_sw[(n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 4) ^ (n & 7)) << 4) | (k & 15))] = b[n * kKB + k];
        }
    signed char *d_b = nullptr;
    CK(cudaMalloc(&d_b, b_sw.size()));
    CK(cudaMemcpy(d_b, b_sw.data(), b_sw.size(), cudaMemcpyHostToDevice));
    unsigned char *d_dump = nullptr;
    int *d_out = nullptr;
    CK(cudaMalloc(&d_dump, 2 * 16 * 1024));
    CK(cudaMalloc(&d_out, (2 * kRows * kN + kRows * 64) * sizeof(int)));
    CK(cudaMemset(d_out, 0xff, (2 * kRows * kN + kRows * 64) * sizeof(int)));
    const int flags = argc > 3 ? atoi(argv[3]) : 0;    // 1: 16-byte aligned box coordinates; 2: no swizzle
    const int first_byte = (flags & 1) ? -1 : 3 * kM + 5;   // default: an unaligned start
    const size_t smem = 2 * 16 * 1024 + kN * kKB + 64;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(tmap, d_b, first_byte, d_dump, d_out, mode);
    printf("mode %d\n", mode);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());

    auto row_start = [&](int r, int i) { return first_byte >= 0 ? first_byte + (r + 16 * i) * kM : 16 * r + 16 * kM * i; };
    std::vector<unsigned char> dump(2 * 16 * 1024);
    std::vector<int> out(2 * kRows * kN + kRows * 64);
    CK(cudaMemcpy(dump.data(), d_dump, dump.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(int), cudaMemcpyDeviceToHost));
    // 2 + 3: what the TMA boxes left in shared memory
    long bad_layout = 0;
    for (int plane = 0; plane < 2; ++plane)
        for (int r = 0; r < 16; ++r)
            for (int i = 0; i < 8; ++i)
                for (int k = 0; k < 128; ++k) {
                    const unsigned char want = planes[plane * pitch + row_start(r, i) + k];
                    const unsigned char got = dump[(plane * 16 + r) * 1024 + i * 128 + ((((k >> 4) ^ i) << 4) | (k & 15))];
                    bad_layout += want != got;
                }
    printf("TMA boxes at unaligned byte coordinates, SWIZZLE_128B atoms: %ld mismatching bytes\n", bad_layout);
    // 4 + 5: the products
    long bad_mma = 0;
    for (int plane = 0; plane < 2; ++plane)
        for (int m = 0; m < kRows; ++m) {
            for (int n = 0; n < kN; ++n) {
                long long s = 0;
                for (int k = 0; k < kKB; ++k) {
                    const unsigned char raw = planes[plane * pitch + row_start(m >> 3, m & 7) + k];
                    const int av = plane ? (int)(signed char)raw : (int)raw;
                    s += (long long)av * b[n * kKB + k];
                }
                bad_mma += out[(plane * kRows + m) * kN + n] != (int)s;
            }
        }
    printf("tcgen05.mma kind::i8 (u8 x s8 and s8 x s8, M 128, N 64, 4 K steps) against the host product: %ld mismatches\n", bad_mma);
    if (bad_mma) printf("  sample: got %d %d %d, row 0\n", out[0], out[1], out[2]);
    if (mode >= 4) {
        // the copied operand: lane m, column 8*(plane*4+ks)+c should hold bytes k = 32 ks + 4 c .. + 3 of row m
        long bad_cp = 0;
        for (int plane = 0; plane < 2; ++plane)
            for (int m = 0; m < kRows; ++m)
                for (int ks = 0; ks < 4; ++ks)
                    for (int c = 0; c < 8; ++c) {
                        uint32_t want = 0;
                        for (int b4 = 0; b4 < 4; ++b4) want |= (uint32_t)planes[plane * pitch + row_start(m >> 3, m & 7) + 32 * ks + 4 * c + b4] << (8 * b4);
                        bad_cp += (uint32_t)out[2 * kRows * kN + m * 64 + 8 * (plane * 4 + ks) + c] != want;
                    }
        printf("tcgen05.cp.128x256b into TMEM (row m -> lane m, 4 bytes per column): %ld mismatching words\n", bad_cp);
        if (bad_cp) {
            printf("  lane 0, plane 0 columns:"); for (int c = 0; c < 16; ++c) printf(" %08x", (unsigned)out[2 * kRows * kN + c]); printf("\n  want bytes row 0:");
            for (int k = 0; k < 64; ++k) printf("%s%02x", k % 4 ? "" : " ", planes[row_start(0, 0) + k]); printf("\n");
            printf("  lane 1, plane 0 columns:"); for (int c = 0; c < 8; ++c) printf(" %08x", (unsigned)out[2 * kRows * kN + 64 + c]); printf("\n  want bytes row 1:");
            for (int k = 0; k < 32; ++k) printf("%s%02x", k % 4 ? "" : " ", planes[row_start(0, 1) + k]); printf("\n");
        }
    }
    printf(bad_layout == 0 && bad_mma == 0 ? "PROBE_UMMA_I8_OK\n" : "PROBE_UMMA_I8_FAILED\n");
    return (bad_layout == 0 && bad_mma == 0) ? 0 : 1;
}
