// K4: FM second-order interaction and the DCN-v2 cross-layer epilogue.
//
// Neither exists in the reference (torchctr/models/__init__.py exports only DNN); the
// definitions are the ones of SURVEY.md section 8c / oracle/models.py:
//   FM      out[b] = 0.5 * sum_d[(sum_f v[b,f,d])^2 - sum_f v[b,f,d]^2]   (+ first-order terms)
//   cross   y = x0 * (x W^T + bias) + x
// Both are HBM-bound element/row-wise work: a team of G lanes owns one sample, keeps the
// per-column sums in registers and walks the F fields with all loads of a field group in
// flight; the team reduction is a handful of xor-shuffles.
#include "common.cuh"

namespace ctr {

struct FmArgs {
    const float *x;
    int64_t x_stride;
    int B, F, D;
    int vec, G;          // lanes per sample and floats per lane
    const float *first;  // [B, nfirst] first-order terms (may be null)
    int nfirst;
    int64_t first_stride;
    float *out;
    int64_t out_stride;
    int accumulate;
    // backward
    const float *gout;
    int64_t gout_stride;
    float *gx;
    int64_t gx_stride;
    float *gfirst;
    int64_t gfirst_stride;
};

__device__ __forceinline__ float4 fm_load(const float *p, int vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec == 4) v = __ldg(reinterpret_cast<const float4 *>(p));
    else v.x = __ldg(p);
    return v;
}

__global__ void __launch_bounds__(256) fm_fwd_kernel(const FmArgs a) {
    const int G = a.G;
    const int lane = threadIdx.x & 31;
    const int g_lane = lane & (G - 1);
    const bool col_ok = g_lane * a.vec < a.D;
    const int64_t teams_total = (int64_t)gridDim.x * (256 / G);
    const int64_t nb = ((int64_t)a.B + (kWarp / G) - 1) / (kWarp / G) * (kWarp / G);  // whole warps iterate together
    for (int64_t b = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G; b < nb; b += teams_total) {
        const bool ok = b < a.B;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
        if (ok && col_ok) {
            const float *row = a.x + b * a.x_stride + g_lane * a.vec;
#pragma unroll 8
            for (int f = 0; f < a.F; ++f) {
                const float4 v = fm_load(row + f * a.D, a.vec);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y);
                q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
            }
        }
        float r = 0.5f * ((s.x * s.x - q.x) + (s.y * s.y - q.y) + (s.z * s.z - q.z) + (s.w * s.w - q.w));
        if (ok && a.first != nullptr)
            for (int j = g_lane; j < a.nfirst; j += G) r += __ldg(a.first + b * a.first_stride + j);
        for (int off = 1; off < G; off <<= 1) r += __shfl_xor_sync(kFull, r, off);
        if (ok && g_lane == 0) {
            float *o = a.out + b * a.out_stride;
            *o = a.accumulate ? *o + r : r;
        }
    }
}

__global__ void __launch_bounds__(256) fm_bwd_kernel(const FmArgs a) {
    const int G = a.G;
    const int lane = threadIdx.x & 31;
    const int g_lane = lane & (G - 1);
    const bool col_ok = g_lane * a.vec < a.D;
    const int64_t teams_total = (int64_t)gridDim.x * (256 / G);
    for (int64_t b = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G; b < a.B; b += teams_total) {
        const float go = __ldg(a.gout + b * a.gout_stride);
        if (a.gfirst != nullptr)
            for (int j = g_lane; j < a.nfirst; j += G) a.gfirst[b * a.gfirst_stride + j] = go;
        if (!col_ok) continue;
        const float *row = a.x + b * a.x_stride + g_lane * a.vec;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int f = 0; f < a.F; ++f) {
            const float4 v = fm_load(row + f * a.D, a.vec);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        float *grow = a.gx + b * a.gx_stride + g_lane * a.vec;
#pragma unroll 4
        for (int f = 0; f < a.F; ++f) {
            const float4 v = fm_load(row + f * a.D, a.vec);  // second touch: L1 / L2 hit
            float4 gr = make_float4(go * (s.x - v.x), go * (s.y - v.y), go * (s.z - v.z), go * (s.w - v.w));
            float *dst = grow + f * a.D;
            if (a.vec == 4) {
                float4 *d4 = reinterpret_cast<float4 *>(dst);
                if (a.accumulate) {
                    const float4 o = *d4;
                    gr.x += o.x; gr.y += o.y; gr.z += o.z; gr.w += o.w;
                }
                *d4 = gr;
            } else {
                *dst = a.accumulate ? *dst + gr.x : gr.x;
            }
        }
    }
}

// ---- warp-per-sample variants (the main path) ----------------------------------------------------------------
// A team of G lanes per sample reads 64-byte pieces of 8 different rows per warp instruction; HBM likes wider
// requests.  Here a whole warp walks ONE sample: lane l loads the float4 columns l, l + 32, ... of the row (512
// contiguous bytes per instruction), so column c belongs to field c / G and to the d-part c % G = l % G.  The
// per-d sums over the fields are therefore a register accumulation over k followed by an xor-butterfly over the
// lane bits above G.  Needs D % 4 == 0, G = D / 4 a power of two and F * G <= 32 * kFmK float4 columns.
constexpr int kFmK = 8;

__device__ __forceinline__ float4 fm_field_sums(float4 s, int G) {
    for (int off = G; off < kWarp; off <<= 1) {
        s.x += __shfl_xor_sync(kFull, s.x, off);
        s.y += __shfl_xor_sync(kFull, s.y, off);
        s.z += __shfl_xor_sync(kFull, s.z, off);
        s.w += __shfl_xor_sync(kFull, s.w, off);
    }
    return s;
}

__global__ void __launch_bounds__(256) fm_fwd_row_kernel(const FmArgs a) {
    const int lane = threadIdx.x & 31;
    const int cols = a.F * a.G;                       // float4 columns of the FM part of a row
    const int64_t warps_total = (int64_t)gridDim.x * (256 / kWarp);
    for (int64_t b = (int64_t)blockIdx.x * (256 / kWarp) + (threadIdx.x >> 5); b < a.B; b += warps_total) {
        const float4 *row = reinterpret_cast<const float4 *>(a.x + b * a.x_stride);
        float4 v[kFmK];
#pragma unroll
        for (int k = 0; k < kFmK; ++k) {
            const int c = lane + 32 * k;
            v[k] = c < cols ? __ldg(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < kFmK; ++k) {
            s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w;
            q = fmaf(v[k].x, v[k].x, q); q = fmaf(v[k].y, v[k].y, q); q = fmaf(v[k].z, v[k].z, q); q = fmaf(v[k].w, v[k].w, q);
        }
        s = fm_field_sums(s, a.G);
        // every d-part once (lanes < G), minus all the squares; the first-order terms ride on the same reduction
        float t = 0.5f * ((lane < a.G ? s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w : 0.f) - q);
        if (a.first != nullptr)
            for (int j = lane; j < a.nfirst; j += kWarp) t += __ldg(a.first + b * a.first_stride + j);
        for (int off = 16; off; off >>= 1) t += __shfl_xor_sync(kFull, t, off);
        if (lane == 0) {
            float *o = a.out + b * a.out_stride;
            *o = a.accumulate ? *o + t : t;
        }
    }
}

__global__ void __launch_bounds__(256) fm_bwd_row_kernel(const FmArgs a) {
    const int lane = threadIdx.x & 31;
    const int cols = a.F * a.G;
    const int64_t warps_total = (int64_t)gridDim.x * (256 / kWarp);
    for (int64_t b = (int64_t)blockIdx.x * (256 / kWarp) + (threadIdx.x >> 5); b < a.B; b += warps_total) {
        const float go = __ldg(a.gout + b * a.gout_stride);
        if (a.gfirst != nullptr)
            for (int j = lane; j < a.nfirst; j += kWarp) a.gfirst[b * a.gfirst_stride + j] = go;
        const float4 *row = reinterpret_cast<const float4 *>(a.x + b * a.x_stride);
        float4 *grow = reinterpret_cast<float4 *>(a.gx + b * a.gx_stride);
        float4 v[kFmK], o[kFmK];
#pragma unroll
        for (int k = 0; k < kFmK; ++k) {
            const int c = lane + 32 * k;
            v[k] = c < cols ? __ldg(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            o[k] = (a.accumulate && c < cols) ? grow[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < kFmK; ++k) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
        s = fm_field_sums(s, a.G);
#pragma unroll
        for (int k = 0; k < kFmK; ++k) {
            const int c = lane + 32 * k;
            if (c < cols)
                grow[c] = make_float4(fmaf(go, s.x - v[k].x, o[k].x), fmaf(go, s.y - v[k].y, o[k].y),
                                      fmaf(go, s.z - v[k].z, o[k].z), fmaf(go, s.w - v[k].w, o[k].w));
        }
    }
}

static bool fm_row_path(const FmArgs &a) { return a.vec == 4 && a.G * 4 == a.D && a.F * a.G <= kWarp * kFmK; }

static int fm_lower(FmArgs &a) {
    CTR_REQUIRE(a.B >= 0 && a.F >= 1 && a.D >= 1, "bad FM shape B=%d F=%d D=%d", a.B, a.F, a.D);
    CTR_REQUIRE(a.x != nullptr || a.B == 0, "x is null");
    CTR_REQUIRE((int64_t)a.F * a.D <= a.x_stride, "F*D exceeds the row stride");
    const bool v4 = a.D % 4 == 0 && a.x_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15u) == 0 &&
                    (a.gx == nullptr || (a.gx_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(a.gx) & 15u) == 0));
    a.vec = v4 ? 4 : 1;
    a.G = pow2_ceil(a.D / a.vec);
    CTR_REQUIRE(a.G <= 32, "FM: D=%d too wide for one warp", a.D);
    return CTR_OK;
}

// ---- DCN-v2 cross epilogue --------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    cross_combine_fwd_kernel(const float *__restrict__ x0, const float *__restrict__ x, const float *__restrict__ u,
                             const float *__restrict__ bias, int B, int d, int64_t stride, float *__restrict__ y) {
    const int64_t total = (int64_t)B * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / d;
        const int c = (int)(i - b * d);
        const int64_t o = b * stride + c;
        y[o] = fmaf(x0[o], u[o] + __ldg(bias + c), x[o]);
    }
}

__global__ void __launch_bounds__(256)
    cross_combine_bwd_kernel(const float *__restrict__ x0, const float *__restrict__ u, const float *__restrict__ bias,
                             const float *__restrict__ gy, int B, int d, int64_t stride, float *__restrict__ gu,
                             float *__restrict__ gx0, int accumulate) {
    const int64_t total = (int64_t)B * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / d;
        const int c = (int)(i - b * d);
        const int64_t o = b * stride + c;
        const float g = gy[o];
        gu[o] = g * x0[o];
        const float t = g * (u[o] + __ldg(bias + c));
        gx0[o] = accumulate ? gx0[o] + t : t;
    }
}

static int ew_grid(int64_t total) {
    int64_t blocks = (total + 1023) / 1024;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    return blocks < 1 ? 1 : (int)blocks;
}


// ---- target attention pooling (torchctr/nn/functional.py:46-74) --------------------------------------------------------
// out[b, :] = sum_n softmax_n(cand[b, n, :] . target[b, :]) cand[b, n, :].  The reference materialises scores, weights and a
// second [B, N, E] tensor; here a warp owns a row b, G = E / 4 lanes hold one candidate (float4 each, R = 32 / G candidates per
// trip), the softmax is computed online (running max / sum per candidate slot, slots merged with xor shuffles at the end),
// so cand is read once forward and once backward.  The scores and the row's (max, sum) are kept for the backward pass.
// honor_mask = 0 reproduces the reference exactly: its masked_fill is not in place (functional.py:63), so the mask has no
// effect; honor_mask = 1 is the intended semantics (masked candidates get weight 0; a row with no candidate gives 0).
struct AttnArgs {
    const float *target;   // [B, E]
    const float *cand;     // [B, N, E]
    const float *mask;     // [B, N] or null (0 = masked)
    float *scores;         // [B, N]
    float *row_max;        // [B]
    float *row_sum;        // [B]
    int B, N, E, G, honor_mask;
};

__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__global__ void __launch_bounds__(256) attn_pool_fwd_kernel(const AttnArgs a, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int G = a.G, R = 32 / G;
    const int t = lane % G, r = lane / G;
    for (int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < a.B; b += gridDim.x * (blockDim.x >> 5)) {
        const float4 tg = __ldg(reinterpret_cast<const float4 *>(a.target + (size_t)b * a.E) + t);
        float m = -INFINITY, l = 0.f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int n0 = 0; n0 < a.N; n0 += R) {
            const int n = n0 + r;
            const bool in = n < a.N;
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in) c = __ldg(reinterpret_cast<const float4 *>(a.cand + ((size_t)b * a.N + n) * a.E) + t);
            float s = dot4(c, tg);
            for (int off = 1; off < G; off <<= 1) s += __shfl_xor_sync(kFull, s, off);
            bool valid = in;
            if (in && a.honor_mask && a.mask != nullptr) valid = __ldg(a.mask + (size_t)b * a.N + n) != 0.f;
            if (in && t == 0) a.scores[(size_t)b * a.N + n] = s;
            if (valid) {
                const float mn = fmaxf(m, s);
                const float sc = m == -INFINITY ? 0.f : __expf(m - mn);
                const float p = __expf(s - mn);
                l = l * sc + p;
                acc.x = acc.x * sc + p * c.x; acc.y = acc.y * sc + p * c.y; acc.z = acc.z * sc + p * c.z; acc.w = acc.w * sc + p * c.w;
                m = mn;
            }
        }
        for (int off = G; off < 32; off <<= 1) {       // merge the candidate slots
            const float m2 = __shfl_xor_sync(kFull, m, off), l2 = __shfl_xor_sync(kFull, l, off);
            float4 a2;
            a2.x = __shfl_xor_sync(kFull, acc.x, off); a2.y = __shfl_xor_sync(kFull, acc.y, off);
            a2.z = __shfl_xor_sync(kFull, acc.z, off); a2.w = __shfl_xor_sync(kFull, acc.w, off);
            const float mn = fmaxf(m, m2);
            const float s1 = m == -INFINITY ? 0.f : __expf(m - mn), s2 = m2 == -INFINITY ? 0.f : __expf(m2 - mn);
            l = l * s1 + l2 * s2;
            acc.x = acc.x * s1 + a2.x * s2; acc.y = acc.y * s1 + a2.y * s2; acc.z = acc.z * s1 + a2.z * s2; acc.w = acc.w * s1 + a2.w * s2;
            m = mn;
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        if (r == 0) reinterpret_cast<float4 *>(out + (size_t)b * a.E)[t] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        if (lane == 0) {
            a.row_max[b] = m;
            a.row_sum[b] = l;
        }
    }
}

// d cand_n = w_n gout + ds_n target, d target = sum_n ds_n cand_n, ds_n = w_n (gout . cand_n - gout . out), w_n = exp(s_n - max) / sum
__global__ void __launch_bounds__(256)
    attn_pool_bwd_kernel(const AttnArgs a, const float *__restrict__ out, const float *__restrict__ gout, float *__restrict__ gtarget,
                         float *__restrict__ gcand) {
    const int lane = threadIdx.x & 31;
    const int G = a.G, R = 32 / G;
    const int t = lane % G, r = lane / G;
    for (int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < a.B; b += gridDim.x * (blockDim.x >> 5)) {
        const float4 tg = __ldg(reinterpret_cast<const float4 *>(a.target + (size_t)b * a.E) + t);
        const float4 go = __ldg(reinterpret_cast<const float4 *>(gout + (size_t)b * a.E) + t);
        const float4 o = __ldg(reinterpret_cast<const float4 *>(out + (size_t)b * a.E) + t);
        float gdo = dot4(go, o);
        for (int off = 1; off < G; off <<= 1) gdo += __shfl_xor_sync(kFull, gdo, off);
        const float m = a.row_max[b], l = a.row_sum[b];
        const float inv = l > 0.f ? 1.f / l : 0.f;
        float4 gt = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int n0 = 0; n0 < a.N; n0 += R) {
            const int n = n0 + r;
            const bool in = n < a.N;
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in) c = __ldg(reinterpret_cast<const float4 *>(a.cand + ((size_t)b * a.N + n) * a.E) + t);
            float gc = dot4(go, c);
            for (int off = 1; off < G; off <<= 1) gc += __shfl_xor_sync(kFull, gc, off);
            bool valid = in;
            if (in && a.honor_mask && a.mask != nullptr) valid = __ldg(a.mask + (size_t)b * a.N + n) != 0.f;
            const float w = valid ? __expf(__ldg(a.scores + (size_t)b * a.N + n) - m) * inv : 0.f;
            const float ds = w * (gc - gdo);
            if (in)
                reinterpret_cast<float4 *>(gcand + ((size_t)b * a.N + n) * a.E)[t] =
                    make_float4(w * go.x + ds * tg.x, w * go.y + ds * tg.y, w * go.z + ds * tg.z, w * go.w + ds * tg.w);
            gt.x += ds * c.x; gt.y += ds * c.y; gt.z += ds * c.z; gt.w += ds * c.w;
        }
        for (int off = G; off < 32; off <<= 1) {
            gt.x += __shfl_xor_sync(kFull, gt.x, off); gt.y += __shfl_xor_sync(kFull, gt.y, off);
            gt.z += __shfl_xor_sync(kFull, gt.z, off); gt.w += __shfl_xor_sync(kFull, gt.w, off);
        }
        if (r == 0) reinterpret_cast<float4 *>(gtarget + (size_t)b * a.E)[t] = gt;
    }
}

static int attn_args(AttnArgs *a, const float *target, const float *cand, const float *mask, float *scores, float *row_max,
                     float *row_sum, int B, int N, int E, int honor_mask) {
    CTR_REQUIRE(B >= 0 && N >= 0, "bad shape B=%d N=%d", B, N);
    CTR_REQUIRE(E >= 4 && E <= 128 && (E & (E - 1)) == 0, "E=%d unsupported (a power of two in [4, 128])", E);
    CTR_REQUIRE(target && (cand || N == 0 || B == 0) && scores && row_max && row_sum, "null pointer");
    CTR_REQUIRE(((reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(cand)) & 15u) == 0, "target / cand must be 16-byte aligned");
    *a = AttnArgs{target, cand, mask, scores, row_max, row_sum, B, N, E, E / 4, honor_mask ? 1 : 0};
    return CTR_OK;
}

}  // namespace ctr

using namespace ctr;

extern "C" int ctr_fm_fwd(const float *x, int64_t x_stride, int32_t B, int32_t F, int32_t D, const float *first,
                          int32_t nfirst, int64_t first_stride, float *out, int64_t out_stride, int32_t accumulate,
                          void *stream) {
    FmArgs a{};
    a.x = x; a.x_stride = x_stride; a.B = B; a.F = F; a.D = D;
    a.first = nfirst > 0 ? first : nullptr; a.nfirst = nfirst; a.first_stride = first_stride;
    a.out = out; a.out_stride = out_stride; a.accumulate = accumulate;
    int rc = fm_lower(a);
    if (rc != CTR_OK) return rc;
    if (B == 0) return CTR_OK;
    CTR_REQUIRE(out != nullptr, "out is null");
    CTR_REQUIRE(nfirst == 0 || first != nullptr, "first is null");
    if (fm_row_path(a)) {
        int64_t blocks = ((int64_t)B + 7) / 8;
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        note_launch(), fm_fwd_row_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
        CTR_CUDA_OK(cudaGetLastError());
        return CTR_OK;
    }
    const int teams_per_block = 256 / a.G;
    int64_t blocks = ((int64_t)B + teams_per_block - 1) / teams_per_block;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    note_launch(), fm_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_fm_bwd(const float *x, int64_t x_stride, int32_t B, int32_t F, int32_t D, const float *gout,
                          int64_t gout_stride, float *gx, int64_t gx_stride, int32_t accumulate, float *gfirst,
                          int32_t nfirst, int64_t gfirst_stride, void *stream) {
    FmArgs a{};
    a.x = x; a.x_stride = x_stride; a.B = B; a.F = F; a.D = D;
    a.gout = gout; a.gout_stride = gout_stride; a.gx = gx; a.gx_stride = gx_stride; a.accumulate = accumulate;
    a.gfirst = nfirst > 0 ? gfirst : nullptr; a.nfirst = nfirst; a.gfirst_stride = gfirst_stride;
    int rc = fm_lower(a);
    if (rc != CTR_OK) return rc;
    if (B == 0) return CTR_OK;
    CTR_REQUIRE(gout != nullptr && gx != nullptr, "null gradient pointer");
    CTR_REQUIRE((int64_t)F * D <= gx_stride, "F*D exceeds the gradient row stride");
    if (fm_row_path(a)) {
        int64_t blocks = ((int64_t)B + 7) / 8;
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        note_launch(), fm_bwd_row_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
        CTR_CUDA_OK(cudaGetLastError());
        return CTR_OK;
    }
    const int teams_per_block = 256 / a.G;
    int64_t blocks = ((int64_t)B + teams_per_block - 1) / teams_per_block;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    note_launch(), fm_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_cross_combine_fwd(const float *x0, const float *x, const float *u, const float *bias, int32_t B,
                                     int32_t d, int64_t stride, float *y, void *stream) {
    CTR_REQUIRE(B >= 0 && d >= 1 && stride >= d, "bad cross shape");
    if (B == 0) return CTR_OK;
    CTR_REQUIRE(x0 && x && u && bias && y, "null pointer");
    note_launch(), cross_combine_fwd_kernel<<<ew_grid((int64_t)B * d), 256, 0, (cudaStream_t)stream>>>(x0, x, u, bias, B, d, stride, y);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_cross_combine_bwd(const float *x0, const float *u, const float *bias, const float *gy, int32_t B,
                                     int32_t d, int64_t stride, float *gu, float *gx0, int32_t accumulate_gx0,
                                     void *stream) {
    CTR_REQUIRE(B >= 0 && d >= 1 && stride >= d, "bad cross shape");
    if (B == 0) return CTR_OK;
    CTR_REQUIRE(x0 && u && bias && gy && gu && gx0, "null pointer");
    note_launch(), cross_combine_bwd_kernel<<<ew_grid((int64_t)B * d), 256, 0, (cudaStream_t)stream>>>(x0, u, bias, gy, B, d, stride, gu,
                                                                                     gx0, accumulate_gx0);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}


extern "C" int ctr_target_attention_fwd(const float *target, const float *cand, const float *mask, int32_t B, int32_t N, int32_t E,
                                        int32_t honor_mask, float *out, float *scores, float *row_max, float *row_sum, void *stream) {
    AttnArgs a;
    int rc = attn_args(&a, target, cand, mask, scores, row_max, row_sum, B, N, E, honor_mask);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(out != nullptr && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, "out is null or not 16-byte aligned");
    if (B == 0) return CTR_OK;
    int blocks = (B + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    note_launch(), attn_pool_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, out);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_target_attention_bwd(const float *target, const float *cand, const float *mask, int32_t B, int32_t N, int32_t E,
                                        int32_t honor_mask, const float *out, const float *scores, const float *row_max,
                                        const float *row_sum, const float *gout, float *gtarget, float *gcand, void *stream) {
    AttnArgs a;
    int rc = attn_args(&a, target, cand, mask, const_cast<float *>(scores), const_cast<float *>(row_max), const_cast<float *>(row_sum), B,
                       N, E, honor_mask);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(out && gout && gtarget && (gcand || N == 0), "null pointer");
    CTR_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(gtarget) |
                  reinterpret_cast<uintptr_t>(gcand)) & 15u) == 0, "operands must be 16-byte aligned");
    if (B == 0) return CTR_OK;
    int blocks = (B + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    note_launch(), attn_pool_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, out, gout, gtarget, gcand);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}
