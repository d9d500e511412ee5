// K6: dense layers on the 5th-generation tensor cores (tcgen05), TF32 inputs / fp32 accumulate.
//
//   C[M, N] = act(A[M, K] . W[N, K]^T + bias[N])
//
// the tower Linear layers of torchctr/models/dnn.py:35-46 and the x.W^T contraction of the DCN-v2
// cross layer -- the only GEMM-shaped work on the path.  Both operands are row-major
// with the reduction dimension contiguous (K-major), read as fp32 and fed to the tensor cores as TF32.
//
// One CTA per 128 x 128 output tile, six warps:
//   warp 0   TMA producer: cp.async.bulk.tensor 128 x 32-float boxes of A and W into a 3-stage ring of
//            128-byte-swizzled shared memory tiles, completion on the stage's `full` mbarrier;
//   warp 1   allocates 128 TMEM columns, then one elected lane issues tcgen05.mma.kind::tf32
//            (M = 128, N = 128, K = 8 per instruction, 4 per stage) with the accumulator in TMEM;
//            tcgen05.commit releases the stage (`empty`) and finally signals `accum_full`;
//   warps 2-5 epilogue: tcgen05.ld the accumulator (32 lanes x 16 columns per load), add bias,
//            optional ReLU, 128-bit stores to C.
// Out-of-range rows / columns / K are zero-filled by TMA and masked in the epilogue, so M, N, K
// need no padding beyond 16-byte row pitches.  Two CTAs fit per SM (96 KB smem, 128 of 512 TMEM
// columns each), so one tile's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace ctr {

constexpr int kBM = 128, kBK = 32;               // tile; kBK floats = 128 bytes = one swizzle row
constexpr int kUmmaK = 8;                       // tf32: 32 bytes per MMA along K
constexpr int kGemmThreads = 192;
constexpr int kPairThreads = 320;              // the CTA-pair kernel: producer + MMA + eight epilogue warps
constexpr int kTileABytes = kBM * kBK * 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);        // start address, 16-byte units
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct GemmArgs {
    float *C;
    const float *bias;
    int64_t ldc;
    int M, N, K, act;
    int tiles_m, tiles_n;
    const float *A;      // for the L2 prefetch of whole A tiles
    int64_t lda;
    float *stats;        // optional (pair kernel): per 32-row block column sums of C and of C^2, [blocks][2][N] -- the batch
                         // statistics of the BatchNorm that follows the Linear come out of the GEMM epilogue
    int block_cols;      // > 0 (pair kernel): C is stored column-BLOCKED -- columns [j * block_cols, (j + 1) * block_cols) are the
    int64_t block_stride;//     contiguous matrix C + j * block_stride, [M, block_cols] row-major (ctr_linear_fwd_blocked)
};

// pull a contiguous global range towards L2 (no data comes back to the SM): 16-byte aligned, multiple of 16 bytes
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct GemmArgs;
__device__ __forceinline__ void prefetch_a_rows(const GemmArgs &g, int tile_m);

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// rows [tile_m * 128, +128) of A are one contiguous range: stream them into L2
__device__ __forceinline__ void prefetch_a_rows(const GemmArgs &g, int tile_m) {
    const int r0 = tile_m * kBM;
    if (r0 >= g.M) return;
    const int nr = g.M - r0 < kBM ? g.M - r0 : kBM;
    const char *p = reinterpret_cast<const char *>(g.A + (int64_t)r0 * g.lda);
    const int64_t bytes = ((int64_t)(nr - 1) * g.lda + g.K) * 4 / 16 * 16;
    for (int64_t off = 0; off < bytes; off += 32768)
        l2_prefetch_bulk(p + off, (uint32_t)(bytes - off < 32768 ? bytes - off : 32768));
}

// Persistent, warp-specialised: one CTA per SM walks the output tiles (n fastest, so the CTAs that run side by side
// share their A tile in L2).  The accumulator is double-buffered in TMEM, so the epilogue of tile i overlaps the TMA /
// MMA main loop of tile i + 1; BN up to 256 columns keeps the re-reads of A (once per n tile) low.  Measured on
// the 65536 x 256 x 432 layer (profiles/): 41 us with stores and MMAs switched off, i.e. the kernel is bound by the
// L2 -> SM operand stream (344 MB: every 128-row tile re-reads its W tile), not by the tensor pipe (22 % busy) nor
// by HBM; a 2-CTA (cta_group::2) variant that halves the W stream is the next step.
template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
    linear_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const GemmArgs g) {
    constexpr int kSt = 4;
    constexpr int kTileB = BN * kBK * 4;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B wants 1024-byte alignment
    const uint32_t a_smem = base, b_smem = base + kSt * kTileABytes;
    const uint32_t bars = b_smem + kSt * kTileB;
    const uint32_t full0 = bars, empty0 = bars + 8 * kSt, acc_full0 = bars + 16 * kSt, acc_empty0 = acc_full0 + 16;
    const uint32_t tmem_slot = acc_empty0 + 16;
    const uint32_t epi_smem = bars + 256;                                  // 4 epilogue warps x 4 KB of staging
    uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(gen_base + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (g.K + kBK - 1) / kBK;
    const int num_tiles = g.tiles_m * g.tiles_n;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kSt; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full0 + 8 * b, 1);
            mbar_init(acc_empty0 + 8 * b, 4);      // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / g.tiles_n) * kBM, n0 = (tile % g.tiles_n) * BN;
                // The TMA boxes below take 128 bytes from each of 128 rows -- 128 DRAM pages per box.  The rows of
                // a tile are one contiguous range of A, so ask L2 for the NEXT tile's rows as one stream now; its
                // boxes then hit L2.  (Only when the next tile is another row block: n tiles share their A rows.)
                const int next = tile + gridDim.x;
                if (tile == (int)blockIdx.x) prefetch_a_rows(g, tile / g.tiles_n);
                if (next < num_tiles && next / g.tiles_n != tile / g.tiles_n) prefetch_a_rows(g, next / g.tiles_n);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kSt;
                    if (it >= kSt) mbar_wait(empty0 + 8 * s, ((it / kSt) - 1) & 1);
                    mbar_expect_tx(full0 + 8 * s, kTileABytes + kTileB);
                    tma_load_2d(a_smem + s * kTileABytes, &map_a, full0 + 8 * s, kb * kBK, m0);
                    tma_load_2d(b_smem + s * kTileB, &map_b, full0 + 8 * s, kb * kBK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            const uint32_t idesc = umma_idesc_tf32(kBM, BN);
            int it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
                const int buf = tl & 1;
                if (tl >= 2) {                                  // the epilogue must have drained this accumulator
                    mbar_wait(acc_empty0 + 8 * buf, ((tl >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kSt;
                    mbar_wait(full0 + 8 * s, (it / kSt) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = umma_desc(a_smem + s * kTileABytes), db = umma_desc(b_smem + s * kTileB);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k)  // advance 32 bytes along K inside the swizzle atom
                        umma_tf32(tmem_acc, da + (uint64_t)(k * kUmmaK * 4 >> 4), db + (uint64_t)(k * kUmmaK * 4 >> 4), idesc,
                                  (kb | k) ? 1u : 0u);
                    umma_commit(empty0 + 8 * s);       // stage free once these MMAs have read it
                }
                umma_commit(acc_full0 + 8 * buf);      // accumulator complete
            }
        }
    } else {  // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int q = warp & 3;
        const bool vec_ok = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15u) == 0);
        const bool bias_vec = g.bias != nullptr && (reinterpret_cast<uintptr_t>(g.bias) & 15u) == 0;
        int tl = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
            const int buf = tl & 1;
            const int m0 = (tile / g.tiles_n) * kBM, n0 = (tile % g.tiles_n) * BN;
            mbar_wait(acc_full0 + 8 * buf, (tl >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * BN);
            // 32 columns at a time: lane = row out of TMEM, bias / activation in registers, then through this warp's
            // 4 KB of shared memory (float4 slots xor-swizzled by row) so that the global stores are whole 128-byte
            // row pieces, four rows per instruction, instead of 16 bytes to each of 32 rows
            float4 *stage = reinterpret_cast<float4 *>(gen_base + (epi_smem - base)) + q * 256;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                if (n0 + c >= g.N) break;              // warp-uniform: nothing left in this tile
                float v[32];
                tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);   // all lanes take part (.sync.aligned)
                const int col = n0 + c;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias_vec && col + j + 4 <= g.N) {
                        b4 = __ldg(reinterpret_cast<const float4 *>(g.bias + col + j));
                    } else if (g.bias != nullptr) {
                        if (col + j + 0 < g.N) b4.x = __ldg(g.bias + col + j + 0);
                        if (col + j + 1 < g.N) b4.y = __ldg(g.bias + col + j + 1);
                        if (col + j + 2 < g.N) b4.z = __ldg(g.bias + col + j + 2);
                        if (col + j + 3 < g.N) b4.w = __ldg(g.bias + col + j + 3);
                    }
                    float4 o = make_float4(v[j] + b4.x, v[j + 1] + b4.y, v[j + 2] + b4.z, v[j + 3] + b4.w);
                    if (g.act == 1) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    stage[lane * 8 + ((j >> 2) ^ (lane & 7))] = o;
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 32; i += 4) {       // rows i .. i + 3 of this warp's 32, 8 lanes (128 bytes) per row
                    const int r = i + (lane >> 3), slot = lane & 7;
                    const float4 o = stage[r * 8 + (slot ^ (r & 7))];
                    const int grow = m0 + q * 32 + r, gcol = col + 4 * slot;
                    if (grow < g.M && gcol < g.N) {
                        float *dst = g.C + (int64_t)grow * g.ldc + gcol;
                        if (vec_ok && gcol + 4 <= g.N) {
                            *reinterpret_cast<float4 *>(dst) = o;
                        } else {
                            dst[0] = o.x;
                            if (gcol + 1 < g.N) dst[1] = o.y;
                            if (gcol + 2 < g.N) dst[2] = o.z;
                            if (gcol + 3 < g.N) dst[3] = o.w;
                        }
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty0 + 8 * buf);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
    }
}


// ---- the same GEMM on CTA PAIRS (tcgen05.mma.cta_group::2) --------------------------------------------------------------
// Measured on the 1-CTA kernel above: the tensor pipe is ~25 % busy and the kernel runs at the speed of the L2 -> SM operand
// stream, because every 128-row CTA streams its own copy of the W tile.  A pair of CTAs (a 2-CTA cluster = the two SMs of a
// TPC) computes a 256 x BN tile instead: CTA r loads ITS 128 rows of A and only HALF of the W tile (rows r * BN/2 .. of it);
// one thread of the leader CTA issues M = 256 MMAs that read A and W halves out of both CTAs' shared memory and write each
// CTA's 128 accumulator rows into that CTA's TMEM.  Per 256 x BN x 32 step the pair moves 32 KB + 32 KB x BN / 256 instead of
// 32 KB + 64 KB x BN / 256: the W stream halves, and the same shared memory holds 6 stages instead of 4.
//   producer (warp 0, both CTAs)   cp.async.bulk.tensor ... .cta_group::2: the bytes of BOTH CTAs complete on the LEADER's
//                                  `full` barrier; a CTA re-fills a stage when ITS `empty` barrier fires;
//   MMA (warp 1, leader only)      waits `full`, issues, tcgen05.commit.multicast -> `empty` of both CTAs, finally `acc_full`
//                                  of both;
//   epilogue (warps 2-5, both)     drain their own TMEM half, then arrive on the LEADER's `acc_empty` (8 arrivals).
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {     // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
    linear_tf32_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmArgs g) {
    constexpr int kSt = BN == 256 ? 5 : (BN == 128 ? 7 : 8);
    constexpr int kHalfB = (BN / 2) * kBK * 4;                             // this CTA's half of the W tile
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base, b_smem = base + kSt * kTileABytes;
    const uint32_t bars = b_smem + kSt * kHalfB;
    const uint32_t full0 = bars, empty0 = bars + 8 * kSt, acc_full0 = bars + 16 * kSt, acc_empty0 = acc_full0 + 16;
    const uint32_t tmem_slot = acc_empty0 + 16;
    const uint32_t epi_smem = bars + 256;
    uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(gen_base + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_kb = (g.K + kBK - 1) / kBK;
    const int tiles_m2 = (g.M + 2 * kBM - 1) / (2 * kBM);
    const int num_tiles = tiles_m2 * g.tiles_n;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kSt; ++s) {
            mbar_init(full0 + 8 * s, 1);           // the leader's arrive.expect_tx; bytes of both CTAs land here
            mbar_init(empty0 + 8 * s, 1);          // one multicast commit per use
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full0 + 8 * b, 1);
            mbar_init(acc_empty0 + 8 * b, 16);     // eight epilogue warps of each CTA
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    cluster_sync_all();                            // barriers of both CTAs exist before anybody signals across
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer (both CTAs)
            int it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int m0 = (tile / g.tiles_n) * 2 * kBM + (int)rank * kBM;
                const int n0 = (tile % g.tiles_n) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kSt;
                    if (it >= kSt) mbar_wait(empty0 + 8 * s, ((it / kSt) - 1) & 1);
                    const uint32_t lead_full = mapa_rank(full0 + 8 * s, 0);
                    if (leader) mbar_expect_tx(full0 + 8 * s, 2 * (kTileABytes + kHalfB));
                    tma_load_2d_pair(a_smem + s * kTileABytes, &map_a, lead_full, kb * kBK, m0);
                    tma_load_2d_pair(b_smem + s * kHalfB, &map_b, lead_full, kb * kBK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {  // ---- MMA issuer (leader CTA only)
            const uint32_t idesc = umma_idesc_tf32(2 * kBM, BN);
            int it = 0, tl = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++tl) {
                const int buf = tl & 1;
                if (tl >= 2) {
                    mbar_wait(acc_empty0 + 8 * buf, ((tl >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kSt;
                    mbar_wait(full0 + 8 * s, (it / kSt) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = umma_desc(a_smem + s * kTileABytes), db = umma_desc(b_smem + s * kHalfB);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k)
                        umma_tf32_pair(tmem_acc, da + (uint64_t)(k * kUmmaK * 4 >> 4), db + (uint64_t)(k * kUmmaK * 4 >> 4), idesc,
                                       (kb | k) ? 1u : 0u);
                    umma_commit_pair(empty0 + 8 * s);       // both CTAs may re-fill the stage
                }
                umma_commit_pair(acc_full0 + 8 * buf);      // both CTAs' epilogues may read their accumulator half
            }
        }
    } else {  // ---- epilogue warps 2..9 of both CTAs: TMEM lane quarter = warp % 4, two warps per quarter taking alternate
              //      32-column chunks (a lone warp per scheduler cannot hide its own TMEM / shared-memory latencies)
        const int q = warp & 3, half = (warp - 2) >> 2, ew = warp - 2;
        const bool vec_ok = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15u) == 0);
        const bool bias_vec = g.bias != nullptr && (reinterpret_cast<uintptr_t>(g.bias) & 15u) == 0;
        int tl = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++tl) {
            const int buf = tl & 1;
            const int m0 = (tile / g.tiles_n) * 2 * kBM + (int)rank * kBM, n0 = (tile % g.tiles_n) * BN;
            // this tile's bias slice goes to shared memory ONCE (while the main loop still runs): the chunk loop below then
            // reads it with broadcast LDS instead of eight dependent global loads per 32 columns (measured: those loads,
            // serialised behind the TMEM load, made the epilogue -- not the tensor pipe -- the pace of the whole kernel)
            float *bias_s = reinterpret_cast<float *>(gen_base + (epi_smem - base) + 8 * 4096) + ew * BN;
            for (int j = lane; j < BN; j += 32) bias_s[j] = (g.bias != nullptr && n0 + j < g.N) ? __ldg(g.bias + n0 + j) : 0.f;
            __syncwarp();
            mbar_wait(acc_full0 + 8 * buf, (tl >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * BN);
            float4 *stage = reinterpret_cast<float4 *>(gen_base + (epi_smem - base)) + ew * 256;
#pragma unroll 1
            for (int c = half * 32; c < BN; c += 64) {
                if (n0 + c >= g.N) break;
                float v[32];
                tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
                const int col = n0 + c;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(bias_s + c + j);
                    float4 o = make_float4(v[j] + b4.x, v[j + 1] + b4.y, v[j + 2] + b4.z, v[j + 3] + b4.w);
                    if (g.act == 1) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    stage[lane * 8 + ((j >> 2) ^ (lane & 7))] = o;
                }
                __syncwarp();
                float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), cq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int r = i + (lane >> 3), slot = lane & 7;
                    const float4 o = stage[r * 8 + (slot ^ (r & 7))];
                    const int grow = m0 + q * 32 + r, gcol = col + 4 * slot;
                    if (grow < g.M && gcol < g.N) {
                        // (blocked: block_cols is a multiple of 4 that divides N, so a float4 never straddles two blocks)
                        float *dst = g.block_cols > 0
                                         ? g.C + (int64_t)(gcol / g.block_cols) * g.block_stride + (int64_t)grow * g.block_cols + gcol % g.block_cols
                                         : g.C + (int64_t)grow * g.ldc + gcol;
                        if (vec_ok && gcol + 4 <= g.N) {
                            *reinterpret_cast<float4 *>(dst) = o;
                        } else {
                            dst[0] = o.x;
                            if (gcol + 1 < g.N) dst[1] = o.y;
                            if (gcol + 2 < g.N) dst[2] = o.z;
                            if (gcol + 3 < g.N) dst[3] = o.w;
                        }
                        cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
                        cq.x = fmaf(o.x, o.x, cq.x); cq.y = fmaf(o.y, o.y, cq.y); cq.z = fmaf(o.z, o.z, cq.z); cq.w = fmaf(o.w, o.w, cq.w);
                    }
                }
                if (g.stats != nullptr) {      // column sums over this warp's 32 rows: lanes with the same slot hold rows 8 apart
#pragma unroll
                    for (int off = 8; off < 32; off <<= 1) {
                        cs.x += __shfl_xor_sync(kFull, cs.x, off); cs.y += __shfl_xor_sync(kFull, cs.y, off);
                        cs.z += __shfl_xor_sync(kFull, cs.z, off); cs.w += __shfl_xor_sync(kFull, cs.w, off);
                        cq.x += __shfl_xor_sync(kFull, cq.x, off); cq.y += __shfl_xor_sync(kFull, cq.y, off);
                        cq.z += __shfl_xor_sync(kFull, cq.z, off); cq.w += __shfl_xor_sync(kFull, cq.w, off);
                    }
                    const int gcol = col + 4 * lane;
                    if (lane < 8 && gcol + 4 <= g.N) {
                        float *dst = g.stats + ((size_t)((m0 >> 5) + q) * 2) * g.N + gcol;
                        *reinterpret_cast<float4 *>(dst) = cs;
                        *reinterpret_cast<float4 *>(dst + g.N) = cq;
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_rank(acc_empty0 + 8 * buf, 0));     // the leader's MMA warp waits for all 8
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();                                  // the single-lane roles rejoin their warps (barrier.cluster is .aligned)
    cluster_sync_all();                            // nobody frees TMEM / exits while the peer may still touch this CTA
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
    }
}

template <int BN>
static int launch_linear_pair(const CUtensorMap &ma, const CUtensorMap &mb, GemmArgs g, cudaStream_t stream) {
    constexpr int kSt = BN == 256 ? 5 : (BN == 128 ? 7 : 8);
    constexpr int smem = kSt * (kTileABytes + (BN / 2) * kBK * 4) + 1024 + 256 + 8 * 4096 + 8 * BN * 4;
    static bool configured[64] = {};
    int dev = 0;
    CTR_CUDA_OK(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CTR_CUDA_OK(cudaFuncSetAttribute(linear_tf32_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[dev & 63] = true;
    }
    g.tiles_m = (g.M + 2 * kBM - 1) / (2 * kBM);
    g.tiles_n = (g.N + BN - 1) / BN;
    const int tiles = g.tiles_m * g.tiles_n;
    const int pairs = tiles < kNumSMs / 2 ? tiles : kNumSMs / 2;
    note_launch(), linear_tf32_pair_kernel<BN><<<2 * pairs, kPairThreads, smem, stream>>>(ma, mb, g);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

template <int BN>
static int launch_linear(const CUtensorMap &ma, const CUtensorMap &mb, GemmArgs g, cudaStream_t stream) {
    constexpr int smem = 4 * (kTileABytes + BN * kBK * 4) + 1024 + 256 + 4 * 4096;
    static bool configured[64] = {};            // the attribute is per device: a process may drive several
    int dev = 0;
    CTR_CUDA_OK(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CTR_CUDA_OK(cudaFuncSetAttribute(linear_tf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[dev & 63] = true;
    }
    g.tiles_m = (g.M + kBM - 1) / kBM;
    g.tiles_n = (g.N + BN - 1) / BN;
    const int tiles = g.tiles_m * g.tiles_n;
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    note_launch(), linear_tf32_kernel<BN><<<grid, kGemmThreads, smem, stream>>>(ma, mb, g);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// rows x cols fp32, row pitch ld floats; box = 32 floats x box_rows, 128-byte swizzle, OOB reads give zeros
static int make_map(CUtensorMap *map, const float *ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return CTR_E_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    // L2 promotion: a box takes 128 bytes from each of its rows; asking L2 to fetch 256 bytes per request halves the DRAM
    // transactions of the A stream (the other half is the next k-block's data) -- tuning knob CTR_TMA_L2PROMO = 128 | 256
    static const int promo_env = getenv("CTR_TMA_L2PROMO") ? atoi(getenv("CTR_TMA_L2PROMO")) : 256;
    const CUtensorMapL2promotion promo = promo_env == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld)", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return CTR_E_CUDA;
    }
    return CTR_OK;
}


// ---- weight gradient: dW[N, K] = sum_b G[b, N] X[b, K] -------------------------------------------------------
// The reduction runs over the batch, so both operands are MN-major for the tensor core (the row index b is the MMA's
// K): TMA drops [32 batch rows x 32 floats] boxes (128-byte swizzle with 32-byte atoms -- the only shared-memory
// layout tcgen05 accepts for MN-major 32-bit operands) and the descriptors say "MN-major, SWIZZLE_128B_BASE32B":
// 32-float groups along M / N are kWgBoxBytes apart (leading byte offset), 4-row groups along the batch 512 bytes.  The output is tiny (N x K) and the batch huge, so the batch is split over the
// CTAs (one output tile x one batch slice each, ~one CTA per SM); every CTA leaves its partial tile in a workspace
// and a second kernel adds the slices in order (deterministic).  cuBLAS needs 130 us for the 256 x 432 x 65536 case,
// HBM time is 28 us.
constexpr int kWgBM = 128;             // rows of dW per tile (columns of G)
constexpr int kWgBNMax = 256;          // columns of dW per tile (columns of X): one TMEM accumulator of 256 columns
constexpr int kWgRows = 32;            // batch rows per stage
constexpr int kWgStages = 4;
constexpr int kWgBoxBytes = kWgRows * 128;                      // one [32 rows x 32 floats] box
constexpr int kWgABytes = (kWgBM / 32) * kWgBoxBytes;           // 16 KB
constexpr int kWgBBytes = (kWgBNMax / 32) * kWgBoxBytes;        // 32 KB
constexpr int kWgSmem = kWgStages * (kWgABytes + kWgBBytes) + 1024 + 256;

__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);        // start address, 16-byte units
    d |= (uint64_t)(kWgBoxBytes >> 4) << 16;            // leading byte offset: next 32-float group along M / N
    d |= (uint64_t)(512 >> 4) << 32;                    // stride byte offset: next 4 rows of the batch
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;                             // SWIZZLE_128B_BASE32B: the one layout 32-bit MN-major operands may use
    return d;
}
// D = F32, A = B = TF32, both MN-major (bits 15, 16)
__device__ __forceinline__ uint32_t umma_idesc_tf32_mn(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct WgradArgs {
    float *partial;      // [splits, N, K]
    int B, N, K;
    int tiles_n;         // tiles along K (columns of dW)
    int rows_per_split;  // multiple of kWgRows
};

__global__ void __launch_bounds__(kGemmThreads, 1)
    wgrad_tf32_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x, const WgradArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base, b_smem = base + kWgStages * kWgABytes;
    const uint32_t bars = b_smem + kWgStages * kWgBBytes;
    const uint32_t full0 = bars, empty0 = bars + 8 * kWgStages, accum_full = bars + 16 * kWgStages;
    const uint32_t tmem_slot = accum_full + 8;
    uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(gen_base + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x / g.tiles_n, tile_n = blockIdx.x % g.tiles_n;
    const int n0 = tile_m * kWgBM;                         // first row of dW (column of G)
    const int k0 = tile_n * kWgBNMax;                      // first column of dW (column of X)
    int n_mma = g.K - k0;
    n_mma = n_mma > kWgBNMax ? kWgBNMax : (n_mma + 15) / 16 * 16;
    const int b_boxes = (n_mma + 31) / 32;
    const int split = blockIdx.y;
    const int row0 = split * g.rows_per_split;
    int rows = g.B - row0;
    rows = rows > g.rows_per_split ? g.rows_per_split : rows;
    const int num_st = rows > 0 ? (rows + kWgRows - 1) / kWgRows : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kWgStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kWgBNMax));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer: 4 boxes of G and b_boxes boxes of X per stage
            for (int st = 0; st < num_st; ++st) {
                const int s = st % kWgStages;
                if (st >= kWgStages) mbar_wait(empty0 + 8 * s, ((st / kWgStages) - 1) & 1);
                mbar_expect_tx(full0 + 8 * s, (kWgBM / 32 + b_boxes) * kWgBoxBytes);
                const int r = row0 + st * kWgRows;
                for (int j = 0; j < kWgBM / 32; ++j)
                    tma_load_2d(a_smem + s * kWgABytes + j * kWgBoxBytes, &map_g, full0 + 8 * s, n0 + 32 * j, r);
                for (int j = 0; j < b_boxes; ++j)
                    tma_load_2d(b_smem + s * kWgBBytes + j * kWgBoxBytes, &map_x, full0 + 8 * s, k0 + 32 * j, r);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer: 4 instructions of 8 batch rows per stage
            const uint32_t idesc = umma_idesc_tf32_mn(kWgBM, n_mma);
            for (int st = 0; st < num_st; ++st) {
                const int s = st % kWgStages;
                mbar_wait(full0 + 8 * s, (st / kWgStages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = umma_desc_mn(a_smem + s * kWgABytes), db = umma_desc_mn(b_smem + s * kWgBBytes);
#pragma unroll
                for (int k = 0; k < kWgRows / kUmmaK; ++k)   // next 8 batch rows: 1024 bytes further inside every box
                    umma_tf32(tmem_acc, da + (uint64_t)(k * 1024 >> 4), db + (uint64_t)(k * 1024 >> 4), idesc, (st | k) ? 1u : 0u);
                umma_commit(empty0 + 8 * s);
            }
            umma_commit(accum_full);
        }
    } else {  // ---- epilogue warps 2..5: this split's partial tile -> workspace
        const int q = warp & 3;
        const int row = n0 + q * 32 + lane;
        float *dst_row = g.partial + ((size_t)split * g.N + row) * g.K + k0;
        if (num_st > 0) {
            mbar_wait(accum_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
#pragma unroll 1
        for (int c = 0; c < n_mma; c += 16) {
            float v[16];
            if (num_st > 0) {
                tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            if (row < g.N) {
                if (k0 + c + 16 <= g.K && (g.K % 4 == 0)) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4 *>(dst_row + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                    for (int j = 0; j < 16 && k0 + c + j < g.K; ++j) dst_row[c + j] = v[j];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(kWgBNMax));
    }
}

// dW[n, k] = sum over the batch slices, in slice order
__global__ void __launch_bounds__(256)
    wgrad_reduce_kernel(const float *__restrict__ partial, int splits, int64_t elems, int K, float *__restrict__ dW, int64_t lddw) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < splits; ++p) s += partial[(size_t)p * elems + i];
        dW[(i / K) * lddw + (i % K)] = s;
    }
}

static void wgrad_plan(int B, int N, int K, int *tiles_m, int *tiles_n, int *splits, int *rows_per_split) {
    *tiles_m = (N + kWgBM - 1) / kWgBM;
    *tiles_n = (K + kWgBNMax - 1) / kWgBNMax;
    const int tiles = *tiles_m * *tiles_n;
    int sp = (kNumSMs + tiles - 1) / tiles;                       // ~ one CTA per SM
    int rps = ((B + sp - 1) / sp + kWgRows - 1) / kWgRows * kWgRows;
    if (rps < kWgRows) rps = kWgRows;
    *rows_per_split = rps;
    *splits = (B + rps - 1) / rps;
    if (*splits < 1) *splits = 1;
}

}  // namespace ctr

using namespace ctr;

static int linear_fwd_impl(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C, int64_t ldc,
                           int32_t M, int32_t N, int32_t K, int32_t act, float *stats, void *stream, int32_t block_cols = 0,
                           int64_t block_stride = 0);

// C = act(A W^T + bias) stored column-blocked: columns [j * block_cols, (j + 1) * block_cols) form the contiguous matrix
// C + j * block_stride, [M, block_cols] row-major.  What the embedding update wants dL/d(pooled output) in: feature j's
// gradient slices are then a dense [B, D] block (L2-sized) instead of 64-byte pieces strided over a [B, F * D] matrix.
extern "C" int ctr_linear_fwd_blocked(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C,
                                      int32_t block_cols, int64_t block_stride, int32_t M, int32_t N, int32_t K, int32_t act,
                                      void *stream) {
    CTR_REQUIRE(block_cols >= 4 && block_cols % 4 == 0 && N % block_cols == 0, "block_cols=%d must be a multiple of 4 that divides N=%d",
                block_cols, N);
    CTR_REQUIRE(block_stride >= (int64_t)M * block_cols && block_stride % 4 == 0, "block_stride=%lld must be >= M * block_cols and a multiple of 4",
                (long long)block_stride);
    CTR_REQUIRE(M > kBM, "the column-blocked output needs M > %d (the CTA-pair kernel)", kBM);
    CTR_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15u) == 0, "C must be 16-byte aligned");
    return linear_fwd_impl(A, lda, W, ldw, bias, C, N, M, N, K, act, nullptr, stream, block_cols, block_stride);
}

extern "C" int ctr_linear_fwd(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C,
                              int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t act, void *stream) {
    return linear_fwd_impl(A, lda, W, ldw, bias, C, ldc, M, N, K, act, nullptr, stream);
}

// number of 32-row blocks the statistics variant writes partial sums for (0: this M takes a path without them)
extern "C" int32_t ctr_linear_stats_blocks(int32_t M) { return M > kBM ? ((M + 2 * kBM - 1) / (2 * kBM)) * 8 : 0; }

extern "C" int ctr_linear_fwd_stats(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C,
                                    int64_t ldc, int32_t M, int32_t N, int32_t K, float *stats, int64_t stats_floats, void *stream) {
    const int64_t blocks = ctr_linear_stats_blocks(M);
    CTR_REQUIRE(blocks > 0, "M=%d takes the single-CTA path, which does not produce statistics", M);
    CTR_REQUIRE(stats != nullptr && stats_floats >= blocks * 2 * N, "stats buffer too small: %lld < %lld floats", (long long)stats_floats,
                (long long)(blocks * 2 * N));
    CTR_REQUIRE(N % 4 == 0 && (reinterpret_cast<uintptr_t>(stats) & 15u) == 0, "statistics need N %% 4 == 0 and a 16-byte aligned buffer");
    return linear_fwd_impl(A, lda, W, ldw, bias, C, ldc, M, N, K, 0, stats, stream);
}

static int linear_fwd_impl(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C, int64_t ldc,
                           int32_t M, int32_t N, int32_t K, int32_t act, float *stats, void *stream, int32_t block_cols,
                           int64_t block_stride) {
    CTR_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad GEMM shape M=%d N=%d K=%d", M, N, K);
    if (M == 0) return CTR_OK;
    CTR_REQUIRE(A != nullptr && W != nullptr && C != nullptr, "null pointer");
    CTR_REQUIRE(lda >= K && ldw >= K && ldc >= N, "leading dimension smaller than the row length");
    CTR_REQUIRE(lda % 4 == 0 && ldw % 4 == 0, "lda and ldw must be multiples of 4 floats (16-byte TMA row pitch)");
    CTR_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(W) & 15u) == 0,
                "A and W must be 16-byte aligned");
    CTR_REQUIRE(act == 0 || act == 1, "act must be 0 (none) or 1 (relu)");
    // widest tile the output needs: fewer passes over A (one per n tile)
    const int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
    static const int pair_env = getenv("CTR_GEMM_2CTA") ? atoi(getenv("CTR_GEMM_2CTA")) : 1;
    const bool pair = (pair_env != 0 || stats != nullptr || block_cols > 0) && M > kBM;   // CTA pairs (256-row tiles) unless there is a single 128-row tile
    CUtensorMap ma, mb;
    int rc = make_map(&ma, A, M, K, lda, kBM);
    if (rc != CTR_OK) return rc;
    rc = make_map(&mb, W, N, K, ldw, pair ? bn / 2 : bn);
    if (rc != CTR_OK) return rc;
    GemmArgs g{C, bias, ldc, M, N, K, act, 0, 0, A, lda, stats, block_cols, block_stride};
    if (pair) {
        if (bn == 256) return launch_linear_pair<256>(ma, mb, g, (cudaStream_t)stream);
        if (bn == 128) return launch_linear_pair<128>(ma, mb, g, (cudaStream_t)stream);
        return launch_linear_pair<64>(ma, mb, g, (cudaStream_t)stream);
    }
    if (bn == 256) return launch_linear<256>(ma, mb, g, (cudaStream_t)stream);
    if (bn == 128) return launch_linear<128>(ma, mb, g, (cudaStream_t)stream);
    return launch_linear<64>(ma, mb, g, (cudaStream_t)stream);
}

extern "C" int64_t ctr_linear_wgrad_workspace_bytes(int32_t B, int32_t N, int32_t K) {
    int tm, tn, sp, rps;
    wgrad_plan(B < 1 ? 1 : B, N, K, &tm, &tn, &sp, &rps);
    return (int64_t)sp * N * K * (int64_t)sizeof(float) + 256;
}

extern "C" int ctr_linear_wgrad(const float *G, int64_t ldg, const float *X, int64_t ldx, int32_t B, int32_t N, int32_t K,
                                float *dW, int64_t lddw, void *workspace, int64_t workspace_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CTR_REQUIRE(B >= 1 && N >= 1 && K >= 1, "bad wgrad shape B=%d N=%d K=%d", B, N, K);
    CTR_REQUIRE(G != nullptr && X != nullptr && dW != nullptr && workspace != nullptr, "null pointer");
    CTR_REQUIRE(ldg >= N && ldx >= K && lddw >= K, "leading dimension smaller than the row length");
    CTR_REQUIRE(ldg % 4 == 0 && ldx % 4 == 0, "ldg and ldx must be multiples of 4 floats (16-byte TMA row pitch)");
    CTR_REQUIRE((reinterpret_cast<uintptr_t>(G) & 15u) == 0 && (reinterpret_cast<uintptr_t>(X) & 15u) == 0 &&
                    (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0,
                "G, X and the workspace must be 16-byte aligned");
    int tm, tn, sp, rps;
    wgrad_plan(B, N, K, &tm, &tn, &sp, &rps);
    CTR_REQUIRE(workspace_bytes >= (int64_t)sp * N * K * (int64_t)sizeof(float), "workspace too small");
    CUtensorMap mg, mx;
    int rc = make_map(&mg, G, B, N, ldg, kWgRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc != CTR_OK) return rc;
    rc = make_map(&mx, X, B, K, ldx, kWgRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc != CTR_OK) return rc;
    static bool configured[64] = {};
    int dev = 0;
    CTR_CUDA_OK(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CTR_CUDA_OK(cudaFuncSetAttribute(wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
        configured[dev & 63] = true;
    }
    WgradArgs a{static_cast<float *>(workspace), B, N, K, tn, rps};
    note_launch(), wgrad_tf32_kernel<<<dim3(tm * tn, sp), kGemmThreads, kWgSmem, stream>>>(mg, mx, a);
    const int64_t elems = (int64_t)N * K;
    int64_t rb = (elems + 255) / 256;
    if (rb > kNumSMs * 8) rb = kNumSMs * 8;
    note_launch(), wgrad_reduce_kernel<<<(unsigned)rb, 256, 0, stream>>>(a.partial, sp, elems, K, dW, lddw);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}
