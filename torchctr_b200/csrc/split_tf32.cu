// Error-compensated TF32 ("3xTF32") operand preparation for the tcgen05 GEMMs of gemm_tc.cu.
//
// The tensor cores read fp32 operands as TF32 (10-bit mantissa), which cannot meet the 1e-5 parity bound of the
// tower of torchctr/models/dnn.py:35-46.  Every fp32 value is therefore split into two TF32-representable parts
//     hi = rna_tf32(x),   lo = rna_tf32(x - hi)          (x - hi is exact in fp32)
// and the product A.W^T is formed as  lo_a.hi_w + hi_a.lo_w + hi_a.hi_w  (the dropped lo.lo term is 2^-22 relative)
// with fp32 accumulation in TMEM.  The two small correction terms come FIRST: the tensor core truncates when it adds
// into the accumulator (measured: the error grows linearly with the number of accumulation steps), so corrections added
// to an already large accumulator would each lose up to one ulp of it; added from zero they lose an ulp of themselves.
// Rather than a second GEMM kernel, the three terms are laid out as ONE GEMM with a three times longer reduction
// dimension: this kernel writes the segments
//     role 0 (left operand):  [lo | hi | hi]        role 1 (right operand): [hi | lo | hi]
// side by side along the reduction axis -- along the columns for the K-major operands of ctr_linear_fwd (axis 1) or
// stacked along the rows for the batch-reduced operands of ctr_linear_wgrad (axis 0) -- and the unchanged tcgen05
// kernels contract over them.  Both parts have their low 13 mantissa bits clear, so the result does not depend on how
// the hardware narrows fp32 to TF32.
#include "common.cuh"

namespace ctr {

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// one thread per float4 of the source (cols4 = ceil(cols / 4) vectors per row; the tail past `cols` is written as 0)
__global__ void __launch_bounds__(256)
    split_tf32_kernel(const float *__restrict__ x, int64_t ldx, int rows, int cols, int cols4, float *__restrict__ out,
                      int64_t ldo, int64_t seg_stride, int role, bool vec_in) {
    const int64_t total = (int64_t)rows * cols4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols4), c = (int)(i % cols4) * 4;
        const float *src = x + (int64_t)r * ldx + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec_in && c + 4 <= cols) {
            v = __ldg(reinterpret_cast<const float4 *>(src));
        } else {
            if (c + 0 < cols) v.x = __ldg(src + 0);
            if (c + 1 < cols) v.y = __ldg(src + 1);
            if (c + 2 < cols) v.z = __ldg(src + 2);
            if (c + 3 < cols) v.w = __ldg(src + 3);
        }
        float4 hi, lo;
        hi.x = rna_tf32(v.x); hi.y = rna_tf32(v.y); hi.z = rna_tf32(v.z); hi.w = rna_tf32(v.w);
        lo.x = rna_tf32(v.x - hi.x); lo.y = rna_tf32(v.y - hi.y); lo.z = rna_tf32(v.z - hi.z); lo.w = rna_tf32(v.w - hi.w);
        float *dst = out + (int64_t)r * ldo + c;
        *reinterpret_cast<float4 *>(dst) = role == 0 ? lo : hi;
        *reinterpret_cast<float4 *>(dst + seg_stride) = role == 0 ? hi : lo;
        *reinterpret_cast<float4 *>(dst + 2 * seg_stride) = hi;
    }
}

}  // namespace ctr

using namespace ctr;

extern "C" int ctr_split_tf32(const float *x, int64_t ldx, int32_t rows, int32_t cols, float *out, int64_t ldo, int32_t axis,
                              int32_t role, void *stream) {
    CTR_REQUIRE(rows >= 0 && cols >= 1, "bad shape rows=%d cols=%d", rows, cols);
    if (rows == 0) return CTR_OK;
    CTR_REQUIRE(x != nullptr && out != nullptr, "null pointer");
    CTR_REQUIRE(axis == 0 || axis == 1, "axis must be 0 (segments stacked along rows) or 1 (along columns)");
    CTR_REQUIRE(role == 0 || role == 1, "role must be 0 (lo, hi, hi) or 1 (hi, lo, hi)");
    const int cols4 = (cols + 3) / 4;
    const int64_t seg = (int64_t)cols4 * 4;
    CTR_REQUIRE(ldx >= cols, "ldx smaller than the row length");
    CTR_REQUIRE(ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, "out must be 16-byte aligned with a pitch multiple of 4");
    CTR_REQUIRE(ldo >= (axis == 1 ? 3 * seg : seg), "ldo too small for the segments");
    const bool vec_in = ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
    const int64_t seg_stride = axis == 1 ? seg : (int64_t)rows * ldo;
    const int64_t total = (int64_t)rows * cols4;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    note_launch(), split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, cols, cols4, out, ldo,
                                                                                          seg_stride, role, vec_in);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}
