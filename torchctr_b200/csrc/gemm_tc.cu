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

#include "common.cuh"

namespace ctr {

constexpr int kBM = 128, kBN = 128, kBK = 32;   // tile; kBK floats = 128 bytes = one swizzle row
constexpr int kStages = 3;
constexpr int kUmmaK = 8;                       // tf32: 32 bytes per MMA along K
constexpr int kGemmThreads = 192;
constexpr int kTileABytes = kBM * kBK * 4, kTileBBytes = kBN * kBK * 4;
constexpr int kGemmSmem = kStages * (kTileABytes + kTileBBytes) + 1024 /*align*/ + 256 /*barriers*/;

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
};

__global__ void __launch_bounds__(kGemmThreads, 2)
    linear_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B wants 1024-byte alignment
    const uint32_t a_smem = base, b_smem = base + kStages * kTileABytes;
    const uint32_t bars = b_smem + kStages * kTileBBytes;
    const uint32_t full0 = bars, empty0 = bars + 8 * kStages, accum_full = bars + 16 * kStages;
    const uint32_t tmem_slot = accum_full + 8;
    uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(gen_base + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
    const int num_kb = (g.K + kBK - 1) / kBK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kBN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                if (kb >= kStages) mbar_wait(empty0 + 8 * s, ((kb / kStages) - 1) & 1);
                mbar_expect_tx(full0 + 8 * s, kTileABytes + kTileBBytes);
                tma_load_2d(a_smem + s * kTileABytes, &map_a, full0 + 8 * s, kb * kBK, m0);
                tma_load_2d(b_smem + s * kTileBBytes, &map_b, full0 + 8 * s, kb * kBK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            const uint32_t idesc = umma_idesc_tf32(kBM, kBN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                mbar_wait(full0 + 8 * s, (kb / kStages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = umma_desc(a_smem + s * kTileABytes), db = umma_desc(b_smem + s * kTileBBytes);
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k)  // advance 32 bytes along K inside the swizzle atom
                    umma_tf32(tmem_acc, da + (uint64_t)(k * kUmmaK * 4 >> 4), db + (uint64_t)(k * kUmmaK * 4 >> 4), idesc,
                              (kb | k) ? 1u : 0u);
                umma_commit(empty0 + 8 * s);       // stage free once these MMAs have read it
            }
            umma_commit(accum_full);               // accumulator complete
        }
    } else {  // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4
        mbar_wait(accum_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const bool vec_ok = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15u) == 0);
#pragma unroll 1
        for (int c = 0; c < kBN; c += 16) {
            float v[16];
            tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);   // all lanes take part (.sync.aligned)
            const int col = n0 + c;
            if (row < g.M && col < g.N) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float b = (g.bias != nullptr && col + j < g.N) ? __ldg(g.bias + col + j) : 0.f;
                    float x = v[j] + b;
                    v[j] = (g.act == 1 && x < 0.f) ? 0.f : x;
                }
                float *dst = g.C + (int64_t)row * g.ldc + col;
                if (vec_ok && col + 16 <= g.N) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                    for (int j = 0; j < 16 && col + j < g.N; ++j) dst[j] = v[j];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(kBN));
    }
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
static int make_map(CUtensorMap *map, const float *ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return CTR_E_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld)", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return CTR_E_CUDA;
    }
    return CTR_OK;
}

}  // namespace ctr

using namespace ctr;

extern "C" int ctr_linear_fwd(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C,
                              int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t act, void *stream) {
    CTR_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad GEMM shape M=%d N=%d K=%d", M, N, K);
    if (M == 0) return CTR_OK;
    CTR_REQUIRE(A != nullptr && W != nullptr && C != nullptr, "null pointer");
    CTR_REQUIRE(lda >= K && ldw >= K && ldc >= N, "leading dimension smaller than the row length");
    CTR_REQUIRE(lda % 4 == 0 && ldw % 4 == 0, "lda and ldw must be multiples of 4 floats (16-byte TMA row pitch)");
    CTR_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(W) & 15u) == 0,
                "A and W must be 16-byte aligned");
    CTR_REQUIRE(act == 0 || act == 1, "act must be 0 (none) or 1 (relu)");
    CUtensorMap ma, mb;
    int rc = make_map(&ma, A, M, K, lda, kBM);
    if (rc != CTR_OK) return rc;
    rc = make_map(&mb, W, N, K, ldw, kBN);
    if (rc != CTR_OK) return rc;
    static bool configured = false;
    if (!configured) {
        CTR_CUDA_OK(cudaFuncSetAttribute(linear_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        configured = true;
    }
    GemmArgs g{C, bias, ldc, M, N, K, act};
    dim3 grid((N + kBN - 1) / kBN, (M + kBM - 1) / kBM);
    note_launch(), linear_tf32_kernel<<<grid, kGemmThreads, kGemmSmem, (cudaStream_t)stream>>>(ma, mb, g);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}
