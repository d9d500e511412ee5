// K6b: the element-wise / column-reduction half of a tower block, fused.
//
// The reference tower (torchctr/models/dnn.py:35-46) is [Linear, BatchNorm1d, ReLU, Dropout(0.5)] x k + Linear.  In
// training mode torch runs, per block, ~5 kernels forward (batch statistics, normalise, clamp, dropout mask,
// scale) and ~6 backward (mask, threshold, two BatchNorm passes, bias-gradient reduction ...), each a full pass over
// the [B, N] activation.  Here:
//   forward   col_stats (one read of z = x W^T + b)  ->  finalize (mean, rstd, running statistics)
//             -> bn_act_fwd: y = dropout(relu((z - mean) * rstd * gamma + beta))        (one read, one write)
//   backward  bn_act_bwd_reduce: sum g, sum g * zhat per column, g = gy * dropout * relu'   (reads gy, z)
//             -> finalize (dgamma, dbeta, the two means)
//             -> bn_act_bwd_apply: gz = gamma * rstd * (g - mean(g) - zhat * mean(g zhat)), and the column sums of
//                gz (the Linear's bias gradient) in the same pass
// The dropout mask is never stored: both directions recompute it from a counter-based generator keyed by
// (device-resident step seed, layer, element), so a captured CUDA graph draws a fresh mask every replay.
// All of it is HBM-bound fp32 streaming work; the GEMMs on either side are in gemm_tc.cu.
#include <stdlib.h>

#include "common.cuh"

namespace ctr {

constexpr int kTowerThreads = 256;
constexpr int kTowerMaxBlocks = kNumSMs * 4;

// geometry shared by every kernel: a thread owns 4 consecutive columns (CG = N / 4 column groups) and every
// RL-th row of the block's row range
struct TowerGeom {
    int B, N, CG, RL, blocks, rows_per_block;
};

static TowerGeom tower_geom(int B, int N) {
    TowerGeom g{};
    g.B = B;
    g.N = N;
    g.CG = N / 4;
    g.RL = kTowerThreads / g.CG;
    if (g.RL < 1) g.RL = 1;
    int blocks = (B + g.RL * 8 - 1) / (g.RL * 8);   // at least 8 rows per row lane
    if (blocks > kTowerMaxBlocks) blocks = kTowerMaxBlocks;
    if (blocks < 1) blocks = 1;
    g.rows_per_block = (B + blocks - 1) / blocks;
    g.blocks = (B + g.rows_per_block - 1) / g.rows_per_block;
    return g;
}

__device__ __forceinline__ float4 ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// keep-mask of 4 consecutive elements: one 64-bit hash, 16 bits per element
__device__ __forceinline__ void drop_keep4(uint64_t seed, uint64_t elem4, uint32_t thresh16, bool keep[4]) {
    const uint64_t h = mix64(seed ^ (elem4 * 0x9e3779b97f4a7c15ull + 0x632be59bd9b4e019ull));
#pragma unroll
    for (int j = 0; j < 4; ++j) keep[j] = (uint32_t)((h >> (16 * j)) & 0xffffu) >= thresh16;
}

// block-level fixed-order reduction of NQ float4 accumulators over the RL row lanes; the block's partial goes to
// partial[(block * NQ + q) * N + col]
template <int NQ>
__device__ __forceinline__ void reduce_rows_and_store(float4 (&acc)[NQ], const TowerGeom &g, int cg, int rl, float4 *red,
                                                      float *__restrict__ partial) {
    const bool active = rl < g.RL;      // threads past RL * CG only keep the barriers company
    for (int q = 0; q < NQ; ++q) {
        __syncthreads();
        if (active) red[rl * g.CG + cg] = acc[q];
        __syncthreads();
        if (rl == 0) {
            float4 s = red[cg];
            for (int r = 1; r < g.RL; ++r) {
                const float4 o = red[r * g.CG + cg];
                s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
            }
            reinterpret_cast<float4 *>(partial + ((size_t)blockIdx.x * NQ + q) * g.N)[cg] = s;
        }
    }
}

// ---- forward -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTowerThreads)
    col_stats_kernel(const float *__restrict__ z, int64_t ldz, const TowerGeom g, float *__restrict__ partial) {
    __shared__ float4 red[kTowerThreads];
    const int cg = threadIdx.x % g.CG, rl = threadIdx.x / g.CG;
    float4 acc[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
    if (rl < g.RL) {
        const int r0 = blockIdx.x * g.rows_per_block;
        const int r1 = min(r0 + g.rows_per_block, g.B);
        for (int r = r0 + rl; r < r1; r += 4 * g.RL) {       // four rows in flight
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = r + j * g.RL < r1 ? ld4(z + (size_t)(r + j * g.RL) * ldz + 4 * cg) : make_float4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[0].x += v[j].x; acc[0].y += v[j].y; acc[0].z += v[j].z; acc[0].w += v[j].w;
                acc[1].x = fmaf(v[j].x, v[j].x, acc[1].x); acc[1].y = fmaf(v[j].y, v[j].y, acc[1].y);
                acc[1].z = fmaf(v[j].z, v[j].z, acc[1].z); acc[1].w = fmaf(v[j].w, v[j].w, acc[1].w);
            }
        }
    }
    reduce_rows_and_store<2>(acc, g, cg, rl, red, partial);
}

// Finalize kernels: a block of kFinCols x kFinLanes threads owns kFinCols columns; lane j adds the partials of
// blocks j, j + kFinLanes, ... in double, the lanes are then added in lane order (fixed order: deterministic).
constexpr int kFinCols = 16, kFinLanes = 64;
// The finalize kernels ON the critical path (BatchNorm statistics forward, dgamma / dbeta / means backward) own kFinColsFast
// columns per block instead: 256-thread blocks find a free SM at once while the second stream's weight-gradient kernel or the
// sort occupies the machine (the 1024-thread ones waited ~9 us for one, CUPTI timeline of round 2), and there are four times
// as many of them.  Same lanes per column, same order of every sum: bit-identical results.
constexpr int kFinColsFast = 4;

template <int NQ, int LANES = kFinLanes, int COLS = kFinCols>
__device__ __forceinline__ bool column_totals(const float *__restrict__ partial, int blocks, int N, double (&tot)[NQ], int *col) {
    static_assert(LANES % 8 == 0, "lanes are folded eight at a time");
    __shared__ double sh[NQ][LANES][COLS];
    __shared__ double sh2[NQ][LANES / 8][COLS];
    const int c = threadIdx.x % COLS, j = threadIdx.x / COLS;
    const int n = blockIdx.x * COLS + c;
    double acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
    if (n < N) {
        // (an explicit "all loads of a batch first, then the adds" form of this loop measured 2x SLOWER in the step -- 95.7 vs 41.2 us
        // for the three statistics finalisations of a cfg2 step: the rolling window the compiler builds here keeps loads in flight
        // while earlier ones are consumed)
#pragma unroll 8
        for (int b = j; b < blocks; b += LANES) {       // independent loads: eight blocks in flight per thread
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[q] += (double)partial[((size_t)b * NQ + q) * N + n];
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) sh[q][j][c] = acc[q];
    __syncthreads();
    if (j < LANES / 8) {                                // lanes 8 j .. 8 j + 7, in lane order
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            double t = 0.0;
#pragma unroll
            for (int l = 0; l < 8; ++l) t += sh[q][j * 8 + l][c];
            sh2[q][j][c] = t;
        }
    }
    __syncthreads();
    *col = n;
    if (j != 0 || n >= N) return false;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        double t = 0.0;
#pragma unroll
        for (int l = 0; l < LANES / 8; ++l) t += sh2[q][l][c];
        tot[q] = t;
    }
    return true;
}

// partials -> mean, rstd (+ running statistics, torch.nn.BatchNorm1d semantics).  LANES threads share a column's partials.
template <int LANES>
__global__ void __launch_bounds__(kFinColsFast * LANES)
    bn_finalize_fwd_kernel(const float *__restrict__ partial, int blocks, int B, int N, float eps, float momentum,
                           float *__restrict__ mean, float *__restrict__ rstd, float *running_mean, float *running_var,
                           long long *num_batches_tracked) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
    double tot[2];
    int n;
    if (!column_totals<2, LANES, kFinColsFast>(partial, blocks, N, tot, &n)) return;
    const double s = tot[0], ss = tot[1];
    const double m = s / B;
    double var = ss / B - m * m;
    if (var < 0.0) var = 0.0;
    mean[n] = (float)m;
    rstd[n] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) {
        const double unbiased = B > 1 ? var * ((double)B / (double)(B - 1)) : var;
        running_mean[n] = (float)((1.0 - momentum) * running_mean[n] + momentum * m);
        running_var[n] = (float)((1.0 - momentum) * running_var[n] + momentum * unbiased);
    }
}

struct ActArgs {
    const float *mean, *rstd, *gamma, *beta;
    const unsigned long long *seed_dev;   // step seed on the device (may be null: seed 0)
    unsigned long long seed_offset;       // layer id
    float p_drop;
};

__global__ void __launch_bounds__(kTowerThreads)
    bn_act_fwd_kernel(const float *__restrict__ z, int64_t ldz, const TowerGeom g, const ActArgs a, float *__restrict__ y,
                      int64_t ldy) {
    const int cg = threadIdx.x % g.CG, rl = threadIdx.x / g.CG;
    if (rl >= g.RL) return;
    const float4 mu = ld4(a.mean + 4 * cg), rs = ld4(a.rstd + 4 * cg), ga = ld4(a.gamma + 4 * cg), be = ld4(a.beta + 4 * cg);
    const float4 sc = make_float4(rs.x * ga.x, rs.y * ga.y, rs.z * ga.z, rs.w * ga.w);
    const bool drop = a.p_drop > 0.f;
    const float keep_scale = drop ? 1.f / (1.f - a.p_drop) : 1.f;
    const uint32_t thresh = (uint32_t)(a.p_drop * 65536.f);
    const uint64_t seed = (a.seed_dev != nullptr ? *a.seed_dev : 0ull) * 0x100000001b3ull + a.seed_offset;
    const int r0 = blockIdx.x * g.rows_per_block;
    const int r1 = min(r0 + g.rows_per_block, g.B);
#pragma unroll 4
    for (int r = r0 + rl; r < r1; r += g.RL) {
        const float4 v = ld4(z + (size_t)r * ldz + 4 * cg);
        float o[4] = {fmaf(v.x - mu.x, sc.x, be.x), fmaf(v.y - mu.y, sc.y, be.y), fmaf(v.z - mu.z, sc.z, be.z),
                      fmaf(v.w - mu.w, sc.w, be.w)};
        bool keep[4] = {true, true, true, true};
        if (drop) drop_keep4(seed, (uint64_t)r * g.CG + cg, thresh, keep);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (o[j] > 0.f && keep[j]) ? o[j] * keep_scale : 0.f;
        *reinterpret_cast<float4 *>(y + (size_t)r * ldy + 4 * cg) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---- backward ----------------------------------------------------------------------------------------------
// g = gy * keep_scale where the unit was kept and active, else 0; zhat = (z - mean) * rstd
__device__ __forceinline__ void act_grad4(const float4 gy, const float4 v, const float4 mu, const float4 rs, const float4 ga,
                                          const float4 be, bool drop, float keep_scale, uint64_t seed, uint64_t elem4,
                                          uint32_t thresh, float (&gq)[4], float (&zh)[4]) {
    zh[0] = (v.x - mu.x) * rs.x; zh[1] = (v.y - mu.y) * rs.y; zh[2] = (v.z - mu.z) * rs.z; zh[3] = (v.w - mu.w) * rs.w;
    const float pre[4] = {fmaf(v.x - mu.x, rs.x * ga.x, be.x), fmaf(v.y - mu.y, rs.y * ga.y, be.y),
                          fmaf(v.z - mu.z, rs.z * ga.z, be.z), fmaf(v.w - mu.w, rs.w * ga.w, be.w)};
    bool keep[4] = {true, true, true, true};
    if (drop) drop_keep4(seed, elem4, thresh, keep);
    const float gin[4] = {gy.x, gy.y, gy.z, gy.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) gq[j] = (pre[j] > 0.f && keep[j]) ? gin[j] * keep_scale : 0.f;
}

__global__ void __launch_bounds__(kTowerThreads)
    bn_act_bwd_reduce_kernel(const float *__restrict__ gy, int64_t ldgy, const float *__restrict__ z, int64_t ldz,
                             const TowerGeom g, const ActArgs a, float *__restrict__ partial) {
    __shared__ float4 red[kTowerThreads];
    const int cg = threadIdx.x % g.CG, rl = threadIdx.x / g.CG;
    float4 acc[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
    if (rl < g.RL) {
        const float4 mu = ld4(a.mean + 4 * cg), rs = ld4(a.rstd + 4 * cg), ga = ld4(a.gamma + 4 * cg), be = ld4(a.beta + 4 * cg);
        const bool drop = a.p_drop > 0.f;
        const float keep_scale = drop ? 1.f / (1.f - a.p_drop) : 1.f;
        const uint32_t thresh = (uint32_t)(a.p_drop * 65536.f);
        const uint64_t seed = (a.seed_dev != nullptr ? *a.seed_dev : 0ull) * 0x100000001b3ull + a.seed_offset;
        const int r0 = blockIdx.x * g.rows_per_block;
        const int r1 = min(r0 + g.rows_per_block, g.B);
#pragma unroll 4
        for (int r = r0 + rl; r < r1; r += g.RL) {
            float gq[4], zh[4];
            act_grad4(ld4(gy + (size_t)r * ldgy + 4 * cg), ld4(z + (size_t)r * ldz + 4 * cg), mu, rs, ga, be, drop, keep_scale,
                      seed, (uint64_t)r * g.CG + cg, thresh, gq, zh);
            acc[0].x += gq[0]; acc[0].y += gq[1]; acc[0].z += gq[2]; acc[0].w += gq[3];
            acc[1].x = fmaf(gq[0], zh[0], acc[1].x); acc[1].y = fmaf(gq[1], zh[1], acc[1].y);
            acc[1].z = fmaf(gq[2], zh[2], acc[1].z); acc[1].w = fmaf(gq[3], zh[3], acc[1].w);
        }
    }
    reduce_rows_and_store<2>(acc, g, cg, rl, red, partial);
}

// dbeta = sum g, dgamma = sum g zhat; c1 = mean(g), c2 = mean(g zhat) for the apply pass
__global__ void __launch_bounds__(kFinColsFast * kFinLanes)
    bn_finalize_bwd_kernel(const float *__restrict__ partial, int blocks, int B, int N, float *__restrict__ dgamma,
                           float *__restrict__ dbeta, float *__restrict__ c1, float *__restrict__ c2) {
    double tot[2];
    int n;
    if (!column_totals<2, kFinLanes, kFinColsFast>(partial, blocks, N, tot, &n)) return;
    const double s = tot[0], sz = tot[1];
    dbeta[n] = (float)s;
    dgamma[n] = (float)sz;
    c1[n] = (float)(s / B);
    c2[n] = (float)(sz / B);
}

__global__ void __launch_bounds__(kTowerThreads, 4)      // <= 64 registers: all kTowerMaxBlocks blocks resident in one wave
    bn_act_bwd_apply_kernel(const float *__restrict__ gy, int64_t ldgy, const float *__restrict__ z, int64_t ldz,
                            const TowerGeom g, const ActArgs a, const float *__restrict__ c1, const float *__restrict__ c2,
                            float *__restrict__ gz, int64_t ldgz, float *__restrict__ partial) {
    __shared__ float4 red[kTowerThreads];
    const int cg = threadIdx.x % g.CG, rl = threadIdx.x / g.CG;
    float4 acc[1] = {make_float4(0, 0, 0, 0)};
    if (rl < g.RL) {
        const float4 mu = ld4(a.mean + 4 * cg), rs = ld4(a.rstd + 4 * cg), ga = ld4(a.gamma + 4 * cg), be = ld4(a.beta + 4 * cg);
        const float4 m1 = ld4(c1 + 4 * cg), m2 = ld4(c2 + 4 * cg);
        const float k[4] = {ga.x * rs.x, ga.y * rs.y, ga.z * rs.z, ga.w * rs.w};
        const float mm1[4] = {m1.x, m1.y, m1.z, m1.w}, mm2[4] = {m2.x, m2.y, m2.z, m2.w};
        const bool drop = a.p_drop > 0.f;
        const float keep_scale = drop ? 1.f / (1.f - a.p_drop) : 1.f;
        const uint32_t thresh = (uint32_t)(a.p_drop * 65536.f);
        const uint64_t seed = (a.seed_dev != nullptr ? *a.seed_dev : 0ull) * 0x100000001b3ull + a.seed_offset;
        const int r0 = blockIdx.x * g.rows_per_block;
        const int r1 = min(r0 + g.rows_per_block, g.B);
#pragma unroll 4
        for (int r = r0 + rl; r < r1; r += g.RL) {
            float gq[4], zh[4], o[4];
            act_grad4(ld4(gy + (size_t)r * ldgy + 4 * cg), ld4(z + (size_t)r * ldz + 4 * cg), mu, rs, ga, be, drop, keep_scale,
                      seed, (uint64_t)r * g.CG + cg, thresh, gq, zh);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = k[j] * (gq[j] - mm1[j] - zh[j] * mm2[j]);
            *reinterpret_cast<float4 *>(gz + (size_t)r * ldgz + 4 * cg) = make_float4(o[0], o[1], o[2], o[3]);
            acc[0].x += o[0]; acc[0].y += o[1]; acc[0].z += o[2]; acc[0].w += o[3];
        }
    }
    reduce_rows_and_store<1>(acc, g, cg, rl, red, partial);
}

// column sums from one-quantity partials (bias gradient of the Linear in front of the BatchNorm)
__global__ void __launch_bounds__(kFinCols * kFinLanes)
    col_sum_finalize_kernel(const float *__restrict__ partial, int blocks, int N, float *__restrict__ out) {
    double tot[1];
    int n;
    if (!column_totals<1>(partial, blocks, N, tot, &n)) return;
    out[n] = (float)tot[0];
}

// ---- logit head: final Linear(H -> 1) + extra logit terms + BCE-with-logits (mean), forward and backward ---------
// torchctr/models/dnn.py:46,68 (last Linear of the tower) and :75 (binary_cross_entropy_with_logits).  torch runs a
// gemv, an add, the loss (5 element-wise kernels + a mean) and their backward (another ~8) over [B, 1] tensors; here
// one pass computes z = h.w + b + extra, the per-sample loss and dz = (sigmoid(z) - y) / B, a second one (backward)
// gh = g dz w and the reductions gw = g sum_b dz h, gb = g sum_b dz.  Same thread geometry as the tower kernels:
// a row of h is covered by CG = H / 4 lanes (a power of two <= 32), reduced with xor shuffles.
struct HeadArgs {
    const float *h;
    int64_t ldh;
    const float *w;        // [H]
    const float *bias;     // [1] or null
    const float *extra;    // [B, extra_stride] column 0, or null
    int64_t extra_stride;
    const float *labels;   // [B, label_stride] column 0
    int64_t label_stride;
    float *logits;         // [B] or null
    float *dz;             // [B]
    float inv_B;
    const float *xe;       // optional second linear term over raw features (DeepFM's Linear(Nd, 1) on the dense block):
    int64_t ldxe;          //   z += xe[r, :ne] . we + be
    const float *we;       // [ne], ne <= 32
    const float *be;       // [1] or null
    int ne;
};

__global__ void __launch_bounds__(kTowerThreads)
    head_fwd_kernel(const HeadArgs a, const TowerGeom g, float *__restrict__ partial) {
    __shared__ float red[kTowerThreads / 32];
    const int cg = threadIdx.x % g.CG, rl = threadIdx.x / g.CG;
    float loss_acc = 0.f;
    if (rl < g.RL) {
        const float4 w4 = ld4(a.w + 4 * cg);
        const float b0 = a.bias != nullptr ? __ldg(a.bias) : 0.f;
        const int r0 = blockIdx.x * g.rows_per_block;
        const int r1 = min(r0 + g.rows_per_block, g.B);
        const int iters = (r1 - r0 + g.RL - 1) / g.RL;       // the same trip count for every lane of a row team
        // four rows per trip: every load of the four rows (h, the second term's columns, the extra term, the label) is issued
        // before the first shuffle, so a thread has 4 x 16 bytes in flight instead of one dependent chain per row
        for (int it = 0; it < iters; it += 4) {
            float4 v[4];
            float d[4], ex[4], y[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + rl + (it + u) * g.RL;
                ok[u] = it + u < iters && r < r1;
                v[u] = ok[u] ? ld4(a.h + (size_t)r * a.ldh + 4 * cg) : make_float4(0.f, 0.f, 0.f, 0.f);
                ex[u] = 0.f;
                y[u] = 0.f;
                if (ok[u] && cg == 0) {
                    if (a.extra != nullptr) ex[u] = __ldg(a.extra + (size_t)r * a.extra_stride);
                    y[u] = __ldg(a.labels + (size_t)r * a.label_stride);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + rl + (it + u) * g.RL;
                d[u] = v[u].x * w4.x + v[u].y * w4.y + v[u].z * w4.z + v[u].w * w4.w;
                // second linear term: the row team's lanes share its ne columns (coalesced), the shuffle sum below adds them up
                if (ok[u])
                    for (int j = cg; j < a.ne; j += g.CG) d[u] = fmaf(__ldg(a.xe + (size_t)r * a.ldxe + j), __ldg(a.we + j), d[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                for (int off = 1; off < g.CG; off <<= 1) d[u] += __shfl_xor_sync(kFull, d[u], off);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (ok[u] && cg == 0) {
                    const int r = r0 + rl + (it + u) * g.RL;
                    float z = d[u] + b0;
                    if (a.extra != nullptr) z += ex[u];
                    if (a.xe != nullptr && a.be != nullptr) z += __ldg(a.be);
                    // max(z, 0) - z y + log1p(exp(-|z|)): torch's stable form
                    loss_acc += fmaxf(z, 0.f) - z * y[u] + log1pf(expf(-fabsf(z)));
                    const float sg = 1.f / (1.f + expf(-z));
                    a.dz[r] = (sg - y[u]) * a.inv_B;
                    if (a.logits != nullptr) a.logits[r] = z;
                }
            }
        }
    }
    for (int off = 16; off; off >>= 1) loss_acc += __shfl_xor_sync(kFull, loss_acc, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kTowerThreads / 32; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}

__global__ void head_loss_finalize_kernel(const float *__restrict__ partial, int blocks, float inv_B, float *__restrict__ loss) {
    __shared__ double sh[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < blocks; i += 256) t += (double)partial[i];
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int s = 128; s; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(sh[0] * (double)inv_B);
}

// gh[r, :] = g dz[r] w;  partial (2 quantities): sum_r dz[r] h[r, :], and sum_r dz[r] (column 0 of quantity 1)
// (xe, ldxe, ne, gwe_partial): the gradient of the optional second linear term: lane `cg` of a row team accumulates
// sum_r dz[r] xe[r, j] for the columns j = cg, cg + CG, ... < ne <= 32 (at most 8 per lane as CG >= 4... one float4 pair)
__global__ void __launch_bounds__(kTowerThreads)
    head_bwd_kernel(const float *__restrict__ h, int64_t ldh, const float *__restrict__ w, const float *__restrict__ dz,
                    const float *__restrict__ gscale, const TowerGeom g, float *__restrict__ gh, int64_t ldgh,
                    float *__restrict__ gextra, int64_t gextra_stride, float *__restrict__ partial,
                    const float *__restrict__ xe, int64_t ldxe, int ne, float *__restrict__ xe_partial) {
    __shared__ float4 red[kTowerThreads];
    const int cg = threadIdx.x % g.CG, rl = threadIdx.x / g.CG;
    float4 acc[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
    float xacc[4] = {0.f, 0.f, 0.f, 0.f};      // ne <= 32 columns over CG >= 8 lanes (H >= 32): at most 4 per lane
    if (rl < g.RL) {
        const float gs = __ldg(gscale);
        const float4 w4 = ld4(w + 4 * cg);
        const int r0 = blockIdx.x * g.rows_per_block;
        const int r1 = min(r0 + g.rows_per_block, g.B);
#pragma unroll 4
        for (int r = r0 + rl; r < r1; r += g.RL) {
            const float d = __ldg(dz + r);
            const float4 v = ld4(h + (size_t)r * ldh + 4 * cg);
            acc[0].x = fmaf(d, v.x, acc[0].x); acc[0].y = fmaf(d, v.y, acc[0].y);
            acc[0].z = fmaf(d, v.z, acc[0].z); acc[0].w = fmaf(d, v.w, acc[0].w);
            const float gd = gs * d;
            if (gh != nullptr)
                *reinterpret_cast<float4 *>(gh + (size_t)r * ldgh + 4 * cg) = make_float4(gd * w4.x, gd * w4.y, gd * w4.z, gd * w4.w);
            if (cg == 0) {
                acc[1].x += d;
                if (gextra != nullptr) gextra[(size_t)r * gextra_stride] = gd;
            }
            if (xe != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int j = cg + k * g.CG;
                    if (j < ne) xacc[k] = fmaf(d, __ldg(xe + (size_t)r * ldxe + j), xacc[k]);
                }
            }
        }
    }
    reduce_rows_and_store<2>(acc, g, cg, rl, red, partial);
    if (xe != nullptr) {      // per-block partial of the second term's weight gradient: [block][32], fixed order over the row lanes
        float *sx = reinterpret_cast<float *>(red);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k * g.CG >= ne) break;            // (block-uniform)
            const int j = cg + k * g.CG;
            __syncthreads();
            if (rl < g.RL && j < 32) sx[rl * 32 + j] = xacc[k];
            __syncthreads();
            if (rl == 0 && j < ne) {
                float t = 0.f;
                for (int r = 0; r < g.RL; ++r) t += sx[r * 32 + j];
                xe_partial[(size_t)blockIdx.x * 32 + j] = t;
            }
        }
    }
}

// gw = g sum_b dz h, gb = g sum_b dz from the two-quantity partials; the LAST block of the grid (when there is a second linear
// term) computes gwe[j] = g sum over blocks of xe_partial[block][j] instead: one warp per column, lane l adds the blocks
// l, l + 32, ... in double, then a shuffle tree -- a fixed order.  (gbe = g sum dz is the head's own bias gradient.)
__global__ void __launch_bounds__(kFinCols * kFinLanes)
    head_bwd_finalize_kernel(const float *__restrict__ partial, int blocks, int H, const float *__restrict__ gscale,
                             float *__restrict__ gw, float *__restrict__ gb, const float *__restrict__ xe_partial, int ne,
                             float *__restrict__ gwe) {
    if (ne > 0 && blockIdx.x == gridDim.x - 1) {
        const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (j >= ne) return;
        double t = 0.0;
#pragma unroll 8
        for (int b = lane; b < blocks; b += 32) t += (double)xe_partial[(size_t)b * 32 + j];
        for (int off = 16; off; off >>= 1) t += __shfl_xor_sync(kFull, t, off);
        if (lane == 0) gwe[j] = (float)((double)__ldg(gscale) * t);
        return;
    }
    double tot[2];
    int n;
    if (!column_totals<2>(partial, blocks, H, tot, &n)) return;
    const double gs = (double)__ldg(gscale);
    gw[n] = (float)(gs * tot[0]);
    if (n == 0 && gb != nullptr) *gb = (float)(gs * tot[1]);
}

// ---- dense Adagrad over a list of small tensors in one launch (the replicated tower parameters) -------------------
constexpr int kMaxDenseTensors = 48;
struct DenseAdagradArgs {
    float *p[kMaxDenseTensors];
    const float *g[kMaxDenseTensors];
    float *s[kMaxDenseTensors];
    long long n[kMaxDenseTensors];
    int count;
    float lr, eps;
};

__global__ void __launch_bounds__(256) dense_adagrad_kernel(const __grid_constant__ DenseAdagradArgs a) {
    const int t = blockIdx.y;
    const long long n = a.n[t];
    float *p = a.p[t];
    const float *gr = a.g[t];
    float *st = a.s[t];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = gr[i];
        const float si = st[i] + gi * gi;
        st[i] = si;
        p[i] = p[i] - a.lr * __fdiv_rn(gi, __fsqrt_rn(si) + a.eps);      // torch.optim.Adagrad: addcdiv_(grad, sqrt(sum) + eps, -clr)
    }
}

static int check_tower_shape(int B, int N, int64_t ld) {
    CTR_REQUIRE(B >= 1 && N >= 4 && N % 4 == 0 && N <= 1024, "tower block: B=%d, N=%d unsupported (N: multiple of 4 up to 1024)", B, N);
    CTR_REQUIRE(ld >= N && ld % 4 == 0, "row pitch %lld must be a multiple of 4 and >= N", (long long)ld);
    return CTR_OK;
}
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace ctr

using namespace ctr;

extern "C" int64_t ctr_tower_workspace_bytes(int32_t N) {
    // partials: up to kTowerMaxBlocks x 2 quantities x N floats, + 2 N floats (c1, c2), + kTowerMaxBlocks x 32 floats (the
    // logit head's second linear term)
    return ((int64_t)kTowerMaxBlocks * 2 * N + 2 * N + (int64_t)kTowerMaxBlocks * 32) * (int64_t)sizeof(float) + 256;
}

extern "C" int ctr_bn_stats(const float *z, int64_t ldz, int32_t B, int32_t N, float eps, float momentum, float *mean,
                            float *rstd, float *running_mean, float *running_var, int64_t *num_batches_tracked, void *workspace,
                            void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_tower_shape(B, N, ldz);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(z != nullptr && mean != nullptr && rstd != nullptr && workspace != nullptr, "null pointer");
    CTR_REQUIRE(aligned16(z) && aligned16(workspace), "z / workspace must be 16-byte aligned");
    CTR_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "running_mean and running_var go together");
    const TowerGeom g = tower_geom(B, N);
    float *partial = static_cast<float *>(workspace);
    note_launch(), col_stats_kernel<<<g.blocks, kTowerThreads, 0, stream>>>(z, ldz, g, partial);
    note_launch(), bn_finalize_fwd_kernel<kFinLanes><<<(N + kFinColsFast - 1) / kFinColsFast, kFinColsFast * kFinLanes, 0, stream>>>(partial, g.blocks, B, N, eps, momentum, mean, rstd,
                                                                              running_mean, running_var,
                                                                              reinterpret_cast<long long *>(num_batches_tracked));
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

// the same finalisation on partial sums that came out of the GEMM epilogue (ctr_linear_fwd_stats): [blocks][2][N]
extern "C" int ctr_bn_stats_from_partials(const float *partial, int32_t blocks, int32_t B, int32_t N, float eps, float momentum,
                                          float *mean, float *rstd, float *running_mean, float *running_var,
                                          int64_t *num_batches_tracked, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CTR_REQUIRE(partial && mean && rstd && blocks >= 1 && B >= 1 && N >= 1, "bad arguments");
    CTR_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "running_mean and running_var go together");
    // lanes per column: the GEMM epilogue leaves B / 32 partial rows (2048 for a 65536-row batch), i.e. a chain of 32 dependent-free
    // but latency-bound loads per quantity and lane at 64 lanes; more lanes shorten it (CTR_FIN_LANES = 64 | 128 | 256, tuning knob)
    static const int lanes_env = getenv("CTR_FIN_LANES") ? atoi(getenv("CTR_FIN_LANES")) : 64;
    const int grid = (N + kFinColsFast - 1) / kFinColsFast;
    long long *nbt = reinterpret_cast<long long *>(num_batches_tracked);
    note_launch();
    if (lanes_env >= 256 && blocks >= 1024)
        bn_finalize_fwd_kernel<256><<<grid, kFinColsFast * 256, 0, stream>>>(partial, blocks, B, N, eps, momentum, mean, rstd, running_mean, running_var, nbt);
    else if (lanes_env >= 128 && blocks >= 512)
        bn_finalize_fwd_kernel<128><<<grid, kFinColsFast * 128, 0, stream>>>(partial, blocks, B, N, eps, momentum, mean, rstd, running_mean, running_var, nbt);
    else
        bn_finalize_fwd_kernel<64><<<grid, kFinColsFast * 64, 0, stream>>>(partial, blocks, B, N, eps, momentum, mean, rstd, running_mean, running_var, nbt);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_bn_relu_dropout_fwd(const float *z, int64_t ldz, int32_t B, int32_t N, const float *mean, const float *rstd,
                                       const float *gamma, const float *beta, float p_drop, const uint64_t *seed_dev,
                                       uint64_t seed_offset, float *y, int64_t ldy, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_tower_shape(B, N, ldz);
    if (rc != CTR_OK) return rc;
    rc = check_tower_shape(B, N, ldy);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(z && mean && rstd && gamma && beta && y, "null pointer");
    CTR_REQUIRE(aligned16(z) && aligned16(y) && aligned16(mean) && aligned16(rstd) && aligned16(gamma) && aligned16(beta),
                "operands must be 16-byte aligned");
    CTR_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "p_drop=%f outside [0, 1)", p_drop);
    const TowerGeom g = tower_geom(B, N);
    ActArgs a{mean, rstd, gamma, beta, reinterpret_cast<const unsigned long long *>(seed_dev), seed_offset, p_drop};
    note_launch(), bn_act_fwd_kernel<<<g.blocks, kTowerThreads, 0, stream>>>(z, ldz, g, a, y, ldy);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_bn_relu_dropout_bwd(const float *gy, int64_t ldgy, const float *z, int64_t ldz, int32_t B, int32_t N,
                                       const float *mean, const float *rstd, const float *gamma, const float *beta, float p_drop,
                                       const uint64_t *seed_dev, uint64_t seed_offset, float *gz, int64_t ldgz, float *dgamma,
                                       float *dbeta, float *dbias, void *workspace, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_tower_shape(B, N, ldz);
    if (rc == CTR_OK) rc = check_tower_shape(B, N, ldgy);
    if (rc == CTR_OK) rc = check_tower_shape(B, N, ldgz);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(gy && z && mean && rstd && gamma && beta && gz && dgamma && dbeta && workspace, "null pointer");
    CTR_REQUIRE(aligned16(gy) && aligned16(z) && aligned16(gz) && aligned16(mean) && aligned16(rstd) && aligned16(gamma) &&
                    aligned16(beta) && aligned16(workspace),
                "operands must be 16-byte aligned");
    CTR_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "p_drop=%f outside [0, 1)", p_drop);
    const TowerGeom g = tower_geom(B, N);
    ActArgs a{mean, rstd, gamma, beta, reinterpret_cast<const unsigned long long *>(seed_dev), seed_offset, p_drop};
    float *partial = static_cast<float *>(workspace);
    float *c1 = partial + (size_t)kTowerMaxBlocks * 2 * N;
    float *c2 = c1 + N;
    note_launch(), bn_act_bwd_reduce_kernel<<<g.blocks, kTowerThreads, 0, stream>>>(gy, ldgy, z, ldz, g, a, partial);
    note_launch(), bn_finalize_bwd_kernel<<<(N + kFinColsFast - 1) / kFinColsFast, kFinColsFast * kFinLanes, 0, stream>>>(partial, g.blocks, B, N, dgamma, dbeta, c1, c2);
    note_launch(), bn_act_bwd_apply_kernel<<<g.blocks, kTowerThreads, 0, stream>>>(gy, ldgy, z, ldz, g, a, c1, c2, gz, ldgz, partial);
    if (dbias != nullptr)
        note_launch(), col_sum_finalize_kernel<<<(N + kFinCols - 1) / kFinCols, kFinCols * kFinLanes, 0, stream>>>(partial, g.blocks, N, dbias);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

// the bias gradient left out of a ctr_bn_relu_dropout_bwd(..., dbias = NULL, workspace) call: column sums of gz from the
// per-block partials that call left in `workspace`.  Nothing on the path to dL/dx needs it, so the caller may issue it on
// another stream (ordered after that call, and before the next call that uses the same workspace).
extern "C" int ctr_bn_bias_grad_from_partials(const void *workspace, int32_t B, int32_t N, float *dbias, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_tower_shape(B, N, N);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(workspace != nullptr && dbias != nullptr, "null pointer");
    const TowerGeom g = tower_geom(B, N);
    note_launch(), col_sum_finalize_kernel<<<(N + kFinCols - 1) / kFinCols, kFinCols * kFinLanes, 0, stream>>>(
        static_cast<const float *>(workspace), g.blocks, N, dbias);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_logit_bce_fwd(const float *h, int64_t ldh, int32_t B, int32_t H, const float *w, const float *bias,
                                 const float *extra, int64_t extra_stride, const float *labels, int64_t label_stride,
                                 float *logits, float *dz, float *loss, void *workspace, void *stream_) {
    return ctr_logit_bce_fwd_ex(h, ldh, B, H, w, bias, extra, extra_stride, nullptr, 0, 0, nullptr, nullptr, labels, label_stride,
                                logits, dz, loss, workspace, stream_);
}

extern "C" int ctr_logit_bce_fwd_ex(const float *h, int64_t ldh, int32_t B, int32_t H, const float *w, const float *bias,
                                    const float *extra, int64_t extra_stride, const float *xe, int64_t ldxe, int32_t ne,
                                    const float *we, const float *be, const float *labels, int64_t label_stride, float *logits,
                                    float *dz, float *loss, void *workspace, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CTR_REQUIRE(xe == nullptr || (we != nullptr && ne >= 1 && ne <= 32 && ldxe >= ne && H >= 32),
                "second linear term: 1 <= ne <= 32, H >= 32, we required");
    CTR_REQUIRE(B >= 1 && H >= 4 && H <= 128 && (H & (H - 1)) == 0, "head: B=%d, H=%d unsupported (H: power of two in [4, 128])", B, H);
    CTR_REQUIRE(ldh >= H && ldh % 4 == 0, "row pitch %lld must be a multiple of 4 and >= H", (long long)ldh);
    CTR_REQUIRE(h && w && labels && dz && loss && workspace, "null pointer");
    CTR_REQUIRE(aligned16(h) && aligned16(w), "h and w must be 16-byte aligned");
    const TowerGeom g = tower_geom(B, H);
    HeadArgs a{h, ldh, w, bias, extra, extra_stride, labels, label_stride, logits, dz, 1.0f / (float)B, xe, ldxe, we, be, xe ? ne : 0};
    float *partial = static_cast<float *>(workspace);
    note_launch(), head_fwd_kernel<<<g.blocks, kTowerThreads, 0, stream>>>(a, g, partial);
    note_launch(), head_loss_finalize_kernel<<<1, 256, 0, stream>>>(partial, g.blocks, 1.0f / (float)B, loss);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_logit_bce_bwd(const float *h, int64_t ldh, int32_t B, int32_t H, const float *w, const float *dz,
                                 const float *gscale, float *gh, int64_t ldgh, float *gw, float *gb, float *gextra,
                                 int64_t gextra_stride, void *workspace, void *stream_) {
    return ctr_logit_bce_bwd_ex(h, ldh, B, H, w, dz, gscale, gh, ldgh, gw, gb, gextra, gextra_stride, nullptr, 0, 0, nullptr, workspace,
                                stream_);
}

extern "C" int ctr_logit_bce_bwd_ex(const float *h, int64_t ldh, int32_t B, int32_t H, const float *w, const float *dz,
                                    const float *gscale, float *gh, int64_t ldgh, float *gw, float *gb, float *gextra,
                                    int64_t gextra_stride, const float *xe, int64_t ldxe, int32_t ne, float *gwe, void *workspace,
                                    void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CTR_REQUIRE(xe == nullptr || ((gwe != nullptr || gw == nullptr) && ne >= 1 && ne <= 32 && ldxe >= ne && H >= 32),
                "second linear term: 1 <= ne <= 32, H >= 32 (the block's row lanes stage [RL][32] partials), gwe required");
    CTR_REQUIRE(B >= 1 && H >= 4 && H <= 128 && (H & (H - 1)) == 0, "head: B=%d, H=%d unsupported (H: power of two in [4, 128])", B, H);
    CTR_REQUIRE(ldh >= H && ldh % 4 == 0 && (gh == nullptr || (ldgh >= H && ldgh % 4 == 0)), "row pitches must be multiples of 4 and >= H");
    CTR_REQUIRE(h && w && dz && gscale && workspace, "null pointer");
    CTR_REQUIRE(aligned16(h) && aligned16(w) && aligned16(gh) && aligned16(workspace), "operands must be 16-byte aligned");
    const TowerGeom g = tower_geom(B, H);
    float *partial = static_cast<float *>(workspace);
    float *xe_partial = partial + (size_t)kTowerMaxBlocks * 2 * H;       // [blocks][32], behind the two-quantity partials
    note_launch(), head_bwd_kernel<<<g.blocks, kTowerThreads, 0, stream>>>(h, ldh, w, dz, gscale, g, gh, ldgh, gextra, gextra_stride, partial,
                                                                           xe, ldxe, xe ? ne : 0, xe_partial);
    CTR_CUDA_OK(cudaGetLastError());
    if (gw == nullptr) return CTR_OK;        // the parameter gradients are finalised later: ctr_logit_bce_bwd_finalize
    return ctr_logit_bce_bwd_finalize(B, H, gscale, gw, gb, xe != nullptr ? ne : 0, gwe, workspace, stream_);
}

// gw[H], gb[1] (may be NULL) and gwe[ne] (ne = 0: none) from the partials a ctr_logit_bce_bwd(_ex)(..., gw = NULL, ...) call left
// in `workspace`.  dL/dh does not depend on them, so the caller may issue this on another stream (ordered after that call,
// and before the next call that uses the same workspace).
extern "C" int ctr_logit_bce_bwd_finalize(int32_t B, int32_t H, const float *gscale, float *gw, float *gb, int32_t ne, float *gwe,
                                          const void *workspace, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CTR_REQUIRE(B >= 1 && H >= 4 && H <= 128 && (H & (H - 1)) == 0, "head: B=%d, H=%d unsupported (H: power of two in [4, 128])", B, H);
    CTR_REQUIRE(gscale && gw && workspace, "null pointer");
    CTR_REQUIRE(ne >= 0 && ne <= 32 && (ne == 0 || gwe != nullptr), "second linear term: 0 <= ne <= 32, gwe required");
    const TowerGeom g = tower_geom(B, H);
    const float *partial = static_cast<const float *>(workspace);
    const float *xe_partial = partial + (size_t)kTowerMaxBlocks * 2 * H;
    note_launch(), head_bwd_finalize_kernel<<<(H + kFinCols - 1) / kFinCols + (ne > 0 ? 1 : 0), kFinCols * kFinLanes, 0, stream>>>(
        partial, g.blocks, H, gscale, gw, gb, xe_partial, ne, gwe);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_dense_adagrad(int32_t count, float *const *params, const float *const *grads, float *const *sums,
                                 const int64_t *sizes, float lr, float eps, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CTR_REQUIRE(count >= 0 && count <= kMaxDenseTensors, "count=%d outside [0, %d]", count, kMaxDenseTensors);
    if (count == 0) return CTR_OK;
    CTR_REQUIRE(params && grads && sums && sizes, "null pointer");
    DenseAdagradArgs a{};
    long long biggest = 1;
    for (int i = 0; i < count; ++i) {
        CTR_REQUIRE(params[i] && grads[i] && sums[i] && sizes[i] >= 0, "tensor %d: null pointer or negative size", i);
        a.p[i] = params[i]; a.g[i] = grads[i]; a.s[i] = sums[i]; a.n[i] = sizes[i];
        if (sizes[i] > biggest) biggest = sizes[i];
    }
    a.count = count; a.lr = lr; a.eps = eps;
    long long bx = (biggest + 1023) / 1024;
    if (bx > kNumSMs * 2) bx = kNumSMs * 2;
    note_launch(), dense_adagrad_kernel<<<dim3((unsigned)bx, count), 256, 0, stream>>>(a);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}
