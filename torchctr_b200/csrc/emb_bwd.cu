// K2: backward of the pooled lookup fused with the optimizer step.
//
// Replaces autograd's aten::embedding_dense_backward behind torchctr/models/dnn.py:57-58 and
// the whole-table optimizer.step() of torchctr/trainer.py:303.  No dense [V, D] gradient is
// ever built:
//   plan   keygen (row key = table base + mapped row, value = slot) -> stable radix sort ->
//          runs of equal keys (= unique rows) listed by a head-flag scan;
//   apply  one team of lanes per unique row sums coef * grad_out[bag] over the run in slot
//          order (deterministic), then applies SGD / Adagrad / row-wise Adagrad / lazy Adam
//          to that row in place.  Runs longer than kLongRun (hot Zipf rows) are queued, cut
//          into chunks of kChunk positions that one block each reduces with a fixed-order
//          shared-memory tree, and finished by a team that adds the chunk partials in order.
#include "sort.cuh"

namespace ctr {

constexpr int kTeamRun = 8;    // runs up to this length: one team of lanes (short-run kernel)
constexpr int kLongRun = 256;  // up to this length: one warp per run; longer: chunked block reduction
constexpr int kApplyThreads = 256;

// ---- workspace layout ---------------------------------------------------------------------
struct PlanLayout {
    int64_t S;           // id slots in the group
    int key_bits;
    int sorted_in_b;     // which ping-pong buffer holds the sorted pairs
    // offsets in bytes from the workspace base
    int64_t counters, keys_a, keys_b, vals_a, vals_b, counts, spine, run_start;
    // apply-time scratch (rebuilt by every apply; sized by the group being applied)
    int64_t med_list, long_list, long_cbase, chunk_q, long_done, partials, max_med, max_long, max_chunks, row_floats, total;
};
// counters (u32): [0] runs, [1] unique valid rows, [2] long-run queue, [3] chunk queue, [4] medium-run queue
constexpr int kNumCounters = 8;
constexpr int kChunk = 1024;   // sorted positions reduced by one block

static int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

static PlanLayout plan_layout(const DevGroup &g) {
    PlanLayout p{};
    int64_t S = 0;
    uint64_t rows = 0;
    for (int i = 0; i < g.num_features; ++i) {
        S += (int64_t)g.B * g.f[i].L;
        rows += g.f[i].num_rows;
    }
    p.S = S;
    int bits = 0;
    while ((rows >> bits) != 0) ++bits;  // bit_length(rows): 2^bits - 1 > every valid key
    p.key_bits = bits < 1 ? 1 : bits;
    const int passes = (p.key_bits + kRadixBits - 1) / kRadixBits;
    p.sorted_in_b = passes & 1;
    int64_t off = 0;
    p.counters = off; off = align256(off + kNumCounters * 4);
    p.keys_a = off; off = align256(off + S * 4);
    p.keys_b = off; off = align256(off + S * 4);
    p.vals_a = off; off = align256(off + S * 4);
    p.vals_b = off; off = align256(off + S * 4);
    const int64_t counts = sort_counts_elems(S);
    p.counts = off; off = align256(off + counts * 4);
    const int64_t spine = scan_spine_elems(counts > S ? counts : S) + 8;
    p.spine = off; off = align256(off + spine * 4);
    p.run_start = off; off = align256(off + (S + 2) * 4);
    int row_floats = 4;  // one float4 slot per lane of the widest row, whatever the lane width
    for (int i = 0; i < g.num_features; ++i) row_floats = max(row_floats, g.f[i].G * 4);
    p.row_floats = row_floats;
    p.max_med = S / (kTeamRun + 1) + 2;
    p.max_long = S / (kLongRun + 1) + 2;
    p.max_chunks = p.max_long + S / kChunk + 2;
    p.med_list = off; off = align256(off + p.max_med * 4);
    p.long_list = off; off = align256(off + p.max_long * 4);
    p.long_cbase = off; off = align256(off + p.max_long * 4);
    p.chunk_q = off; off = align256(off + p.max_chunks * 4);
    p.long_done = off; off = align256(off + p.max_long * 4);
    p.partials = off; off = align256(off + p.max_chunks * row_floats * 4);
    p.total = off;
    return p;
}

// ---- keygen -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    emb_keygen_kernel(const __grid_constant__ DevGroup g, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int fi = blockIdx.y;
    const DevFeature &f = g.f[fi];
    int64_t slot_base = 0;
    for (int i = 0; i < fi; ++i) slot_base += (int64_t)g.B * g.f[i].L;
    const int64_t n = (int64_t)g.B * f.L;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int32_t row = map_index(f, __ldg(f.ids + j));
        keys[slot_base + j] = row >= 0 ? f.row_base + (uint32_t)row : 0xffffffffu;
        vals[slot_base + j] = (uint32_t)j;
    }
}

__global__ void reset_counters_kernel(uint32_t *counters) {
    if (threadIdx.x < kNumCounters) counters[threadIdx.x] = 0;
}

// Plan-time classification of the runs by length: <= kTeamRun stay with the short-run kernel, up to
// kLongRun go to the warp-per-run queue, longer ones are cut into kChunk-position chunks.  Queue order is
// scheduling-dependent, results are not (every run is reduced in slot order by whoever takes it).
__global__ void __launch_bounds__(256)
    classify_runs_kernel(const uint32_t *__restrict__ run_start, uint32_t *counters, uint32_t *__restrict__ med_list,
                         uint32_t *__restrict__ long_list, uint32_t *__restrict__ long_cbase, uint32_t *__restrict__ chunk_q) {
    const uint32_t num_runs = counters[1];
    for (uint32_t run = blockIdx.x * blockDim.x + threadIdx.x; run < num_runs; run += gridDim.x * blockDim.x) {
        const uint32_t len = run_start[run + 1] - run_start[run];
        if (len <= (uint32_t)kTeamRun) continue;
        if (len <= (uint32_t)kLongRun) {
            med_list[atomicAdd(&counters[4], 1u)] = run;
        } else {
            const uint32_t nch = (len + kChunk - 1) / kChunk;
            const uint32_t q = atomicAdd(&counters[2], 1u);
            const uint32_t cbase = atomicAdd(&counters[3], nch);
            long_list[q] = run;
            long_cbase[q] = cbase;
            for (uint32_t c = 0; c < nch; ++c) chunk_q[cbase + c] = q;
        }
    }
}

// ---- apply ----------------------------------------------------------------------------------
struct ApplyArgs {
    const uint32_t *keys;       // sorted
    const uint32_t *vals;       // sorted with the keys: slot inside the feature (bag * L + l)
    const uint32_t *run_start;
    uint32_t *counters;
    uint32_t *med_list;         // [q] -> run (medium runs, one warp each)
    uint32_t *long_list;        // [q] -> run
    uint32_t *long_cbase;       // [q] -> first chunk of the run
    uint32_t *long_done;        // [q] -> chunks finished so far (zeroed by every apply)
    uint32_t *chunk_q;          // [chunk] -> q
    float *partials;            // [chunk, row_floats]
    int row_floats;
    int32_t *uniq_feature;
    int32_t *uniq_row;
    float *row_grad;
    int64_t row_grad_stride;
    int64_t *num_unique;
    int kind;
    ctr_hyper_t h;              // used when hyper_dev == nullptr
    const ctr_hyper_t *hyper_dev;  // device copy read at run time (CUDA-graph replays see new values)
    int team;                   // lanes per run in the short-run kernel (max G of the group)
};

__device__ __forceinline__ int find_feature(const DevGroup &g, uint32_t key) {
    int lo = 0, hi = g.num_features - 1;
    while (lo < hi) {  // last feature whose row_base <= key
        const int mid = (lo + hi + 1) >> 1;
        if (g.f[mid].row_base <= key) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// coefficient of one slot's bag gradient: per-id weight x mean scale
__device__ __forceinline__ float slot_coef(const DevFeature &f, uint32_t slot, uint32_t bag) {
    float c = 1.f;
    if (f.id_weight != nullptr) c = __ldg(f.id_weight + slot);
    if (f.pooling == CTR_POOL_MEAN) c *= __ldg(f.bag_scale + bag);
    return c;
}

__device__ __forceinline__ float4 load_grad_part(const DevGroup &g, const DevFeature &f, uint32_t bag, int g_lane) {
    const float *src = g.out + (int64_t)bag * g.out_stride + f.out_col;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f.vec == 4) {
        if (f.aligned) {
            v = __ldg(reinterpret_cast<const float4 *>(src) + g_lane);
        } else {
            v.x = __ldg(src + 4 * g_lane + 0);
            v.y = __ldg(src + 4 * g_lane + 1);
            v.z = __ldg(src + 4 * g_lane + 2);
            v.w = __ldg(src + 4 * g_lane + 3);
        }
    } else {
        v.x = __ldg(src + g_lane);
    }
    return v;
}

__device__ __forceinline__ float adagrad_elem(float w, float &s, float gr, float lr, float eps) {
    s = s + gr * gr;
    return w - lr * __fdiv_rn(gr, __fsqrt_rn(s) + eps);
}
__device__ __forceinline__ float adam_elem(float w, float &m, float &v, float gr, const ApplyArgs &a) {
    m = m + (gr - m) * a.h.one_minus_beta1;
    v = v + (gr * gr - v) * a.h.one_minus_beta2;
    return w - a.h.adam_step_size * __fdiv_rn(m, __fsqrt_rn(v) + a.h.eps);
}

// Lanes [0, G) of a team hold the summed gradient of (feature f, row); mask names the team.
__device__ __forceinline__ void update_row(const DevGroup &g, const DevFeature &f, const ApplyArgs &a, uint32_t run,
                                           int fi, uint32_t row, int g_lane, bool col_ok, float4 gr, unsigned mask,
                                           int team_lanes) {
    if (a.uniq_row != nullptr && g_lane == 0) {
        a.uniq_row[run] = (int32_t)row;
        if (a.uniq_feature != nullptr) a.uniq_feature[run] = fi;
    }
    if (a.row_grad != nullptr && col_ok) {
        float *dst = a.row_grad + (int64_t)run * a.row_grad_stride;
        if (f.vec == 4) {
            dst[4 * g_lane + 0] = gr.x; dst[4 * g_lane + 1] = gr.y;
            dst[4 * g_lane + 2] = gr.z; dst[4 * g_lane + 3] = gr.w;
        } else {
            dst[g_lane] = gr.x;
        }
    }
    if (a.kind == CTR_OPT_NONE) return;
    float rowwise_denominator = 0.f;
    if (a.kind == CTR_OPT_ROWWISE_ADAGRAD) {  // whole team takes part in the reduction
        float sq = col_ok ? (f.vec == 4 ? gr.x * gr.x + gr.y * gr.y + gr.z * gr.z + gr.w * gr.w : gr.x * gr.x) : 0.f;
        for (int off = 1; off < team_lanes; off <<= 1) sq += __shfl_xor_sync(mask, sq, off);
        const float acc = f.state0[row] + sq / (float)f.D;
        rowwise_denominator = __fsqrt_rn(acc) + a.h.eps;
        __syncwarp(mask);
        if (g_lane == 0) f.state0[row] = acc;
    }
    if (!col_ok) return;
    const size_t off = (size_t)row * f.D + (size_t)g_lane * f.vec;
    if (f.vec == 4) {
        float4 w = *reinterpret_cast<float4 *>(f.table + off);
        if (a.kind == CTR_OPT_SGD) {
            w.x -= a.h.lr * gr.x; w.y -= a.h.lr * gr.y; w.z -= a.h.lr * gr.z; w.w -= a.h.lr * gr.w;
        } else if (a.kind == CTR_OPT_ADAGRAD) {
            float4 s = *reinterpret_cast<float4 *>(f.state0 + off);
            w.x = adagrad_elem(w.x, s.x, gr.x, a.h.lr, a.h.eps);
            w.y = adagrad_elem(w.y, s.y, gr.y, a.h.lr, a.h.eps);
            w.z = adagrad_elem(w.z, s.z, gr.z, a.h.lr, a.h.eps);
            w.w = adagrad_elem(w.w, s.w, gr.w, a.h.lr, a.h.eps);
            *reinterpret_cast<float4 *>(f.state0 + off) = s;
        } else if (a.kind == CTR_OPT_ROWWISE_ADAGRAD) {
            w.x -= a.h.lr * __fdiv_rn(gr.x, rowwise_denominator);
            w.y -= a.h.lr * __fdiv_rn(gr.y, rowwise_denominator);
            w.z -= a.h.lr * __fdiv_rn(gr.z, rowwise_denominator);
            w.w -= a.h.lr * __fdiv_rn(gr.w, rowwise_denominator);
        } else {  // CTR_OPT_ADAM
            float4 m = *reinterpret_cast<float4 *>(f.state0 + off);
            float4 v = *reinterpret_cast<float4 *>(f.state1 + off);
            w.x = adam_elem(w.x, m.x, v.x, gr.x, a);
            w.y = adam_elem(w.y, m.y, v.y, gr.y, a);
            w.z = adam_elem(w.z, m.z, v.z, gr.z, a);
            w.w = adam_elem(w.w, m.w, v.w, gr.w, a);
            *reinterpret_cast<float4 *>(f.state0 + off) = m;
            *reinterpret_cast<float4 *>(f.state1 + off) = v;
        }
        *reinterpret_cast<float4 *>(f.table + off) = w;
    } else {
        float w = f.table[off];
        if (a.kind == CTR_OPT_SGD) {
            w -= a.h.lr * gr.x;
        } else if (a.kind == CTR_OPT_ADAGRAD) {
            float s = f.state0[off];
            w = adagrad_elem(w, s, gr.x, a.h.lr, a.h.eps);
            f.state0[off] = s;
        } else if (a.kind == CTR_OPT_ROWWISE_ADAGRAD) {
            w -= a.h.lr * __fdiv_rn(gr.x, rowwise_denominator);
        } else {
            float m = f.state0[off], v = f.state1[off];
            w = adam_elem(w, m, v, gr.x, a);
            f.state0[off] = m;
            f.state1[off] = v;
        }
        f.table[off] = w;
    }
}

// ---- the three run-length tiers as device functions ------------------------------------------------
// accumulate coef * grad_out[bag] over positions p0, p0 + stride, ... < e, four loads in flight
__device__ __forceinline__ float4 strided_run_sum(const DevGroup &g, const DevFeature &f, const ApplyArgs &a, uint32_t p_first,
                                                  uint32_t e, uint32_t stride, int g_lane, bool col_ok) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t p0 = p_first; p0 < e; p0 += 4 * stride) {
        uint32_t bag[4];
        float coef[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t p = p0 + j * stride;
            bag[j] = 0; coef[j] = 0.f;
            if (p < e) {
                const uint32_t slot = a.vals[p];
                bag[j] = slot / (uint32_t)f.L;
                coef[j] = slot_coef(f, slot, bag[j]);
            }
        }
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col_ok && p0 + j * stride < e) v[j] = load_grad_part(g, f, bag[j], g_lane);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc.x = fmaf(coef[j], v[j].x, acc.x);
            acc.y = fmaf(coef[j], v[j].y, acc.y);
            acc.z = fmaf(coef[j], v[j].z, acc.z);
            acc.w = fmaf(coef[j], v[j].w, acc.w);
        }
    }
    return acc;
}

// Short run (<= kTeamRun): one team of TG lanes, slot order.
__device__ __forceinline__ void short_run(const DevGroup &g, const ApplyArgs &a, uint32_t run, uint32_t s, uint32_t e, int TG,
                                          int t, int team_in_warp, unsigned mask) {
    const uint32_t key = a.keys[s];
    const int fi = find_feature(g, key);
    const DevFeature &f = g.f[fi];
    const bool col_ok = t < f.G && t * f.vec < f.D;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t base = s; base < e; base += TG) {
        uint32_t bag = 0;
        float coef = 0.f;
        if (base + t < e) {
            const uint32_t slot = a.vals[base + t];
            bag = slot / (uint32_t)f.L;
            coef = slot_coef(f, slot, bag);
        }
        const int m = (int)min((uint32_t)TG, e - base);
        for (int k = 0; k < m; ++k) {
            const uint32_t bk = __shfl_sync(mask, bag, team_in_warp * TG + k);
            const float ck = __shfl_sync(mask, coef, team_in_warp * TG + k);
            if (col_ok) {
                const float4 v = load_grad_part(g, f, bk, t);
                acc.x = fmaf(ck, v.x, acc.x);
                acc.y = fmaf(ck, v.y, acc.y);
                acc.z = fmaf(ck, v.z, acc.z);
                acc.w = fmaf(ck, v.w, acc.w);
            }
        }
    }
    update_row(g, f, a, run, fi, key - f.row_base, t, col_ok, acc, mask, TG);
}

// Medium run (<= kLongRun): one warp; 32 / G row slots stride the run, xor-shuffles fold the slots.
__device__ __forceinline__ void medium_run(const DevGroup &g, const ApplyArgs &a, uint32_t run, int lane) {
    const uint32_t s = a.run_start[run], e = a.run_start[run + 1];
    const uint32_t key = a.keys[s];
    const int fi = find_feature(g, key);
    const DevFeature &f = g.f[fi];
    const int G = f.G;
    const int g_lane = lane & (G - 1);
    const bool col_ok = g_lane * f.vec < f.D;
    float4 acc = strided_run_sum(g, f, a, s + lane / G, e, kWarp / G, g_lane, col_ok);
    for (int off = G; off < kWarp; off <<= 1) {
        acc.x += __shfl_xor_sync(kFull, acc.x, off);
        acc.y += __shfl_xor_sync(kFull, acc.y, off);
        acc.z += __shfl_xor_sync(kFull, acc.z, off);
        acc.w += __shfl_xor_sync(kFull, acc.w, off);
    }
    update_row(g, f, a, run, fi, key - f.row_base, lane, lane < G && col_ok, acc, kFull, G);
}

// One chunk of a long run: the whole block; 256 / G row slots, fixed-order shared-memory tree -> partials[ci].
__device__ __forceinline__ void chunk_task(const DevGroup &g, const ApplyArgs &a, uint32_t ci, float4 *red) {
    const uint32_t q = a.chunk_q[ci];
    const uint32_t run = a.long_list[q];
    const uint32_t c = ci - a.long_cbase[q];
    const uint32_t s = a.run_start[run] + c * kChunk;
    const uint32_t e = min(a.run_start[run + 1], s + (uint32_t)kChunk);
    const DevFeature &f = g.f[find_feature(g, a.keys[s])];
    const int G = f.G;
    const int g_lane = threadIdx.x & (G - 1);
    const bool col_ok = g_lane * f.vec < f.D;
    red[threadIdx.x] = strided_run_sum(g, f, a, s + threadIdx.x / G, e, kApplyThreads / G, g_lane, col_ok);
    __syncthreads();
    for (int stride = kApplyThreads / 2; stride >= G; stride >>= 1) {
        if ((int)threadIdx.x < stride) {
            const float4 o = red[threadIdx.x + stride];
            float4 m = red[threadIdx.x];
            m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
            red[threadIdx.x] = m;
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < G) reinterpret_cast<float4 *>(a.partials + (size_t)ci * a.row_floats)[threadIdx.x] = red[threadIdx.x];
}

// Finish a long run whose chunks are all done: the first warp adds the partials in chunk order and updates.
__device__ __forceinline__ void long_finish(const DevGroup &g, const ApplyArgs &a, uint32_t q, int lane) {
    const uint32_t run = a.long_list[q];
    const uint32_t s = a.run_start[run], e = a.run_start[run + 1];
    const uint32_t nch = (e - s + kChunk - 1) / kChunk;
    const uint32_t cbase = a.long_cbase[q];
    const uint32_t key = a.keys[s];
    const int fi = find_feature(g, key);
    const DevFeature &f = g.f[fi];
    const bool col_ok = lane < f.G && lane * f.vec < f.D;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < f.G) {
        for (uint32_t c = 0; c < nch; ++c) {   // .cg: partials were written by other blocks
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(a.partials + (size_t)(cbase + c) * a.row_floats) + lane);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    update_row(g, f, a, run, fi, key - f.row_base, lane, col_ok, acc, kFull, f.G);
}

// One launch for all tiers.  Work is handed out dynamically (atomic tickets in the plan's counter block),
// longest tasks first: chunks of hot rows, then warp-sized runs, then the bulk of short runs -- so the
// latency-bound tails of the tiers overlap instead of running back to back.  counters: [5] chunk ticket,
// [6] medium ticket, [7] short ticket; long_done[q] counts finished chunks of long run q.
__global__ void __launch_bounds__(kApplyThreads)
    emb_bwd_apply_kernel(const __grid_constant__ DevGroup g, const __grid_constant__ ApplyArgs a_in) {
    __shared__ float4 red[kApplyThreads];
    __shared__ uint32_t ticket;
    __shared__ uint32_t finish_q;
    ApplyArgs a = a_in;
    if (a.hyper_dev != nullptr) a.h = *a.hyper_dev;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t num_runs = a.counters[1], nlong_chunks = a.counters[3], nmed = a.counters[4];
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.num_unique != nullptr) *a.num_unique = (int64_t)num_runs;

    // tier 1: chunks of long runs (whole block per chunk)
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) ticket = atomicAdd(&a.counters[5], 1u);
        __syncthreads();
        const uint32_t ci = ticket;
        if (ci >= nlong_chunks) break;
        chunk_task(g, a, ci, red);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t q = a.chunk_q[ci];
            const uint32_t run = a.long_list[q];
            const uint32_t nch = (a.run_start[run + 1] - a.run_start[run] + kChunk - 1) / kChunk;
            finish_q = (atomicAdd(&a.long_done[q], 1u) == nch - 1) ? q : 0xffffffffu;
        }
        __syncthreads();
        if (finish_q != 0xffffffffu && warp == 0) {
            __threadfence();
            long_finish(g, a, finish_q, lane);
        }
    }
    // tier 2: medium runs, one per warp, eight per ticket
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) ticket = atomicAdd(&a.counters[6], (uint32_t)(kApplyThreads / kWarp));
        __syncthreads();
        const uint32_t base = ticket;
        if (base >= nmed) break;
        if (base + warp < nmed) medium_run(g, a, a.med_list[base + warp], lane);
    }
    // tier 3: short runs, one per team; a ticket covers kShortBatch rounds of the block's teams
    const int TG = a.team;
    const int t = lane & (TG - 1);
    const int team_in_warp = lane / TG;
    const unsigned mask = TG == 32 ? kFull : (((1u << TG) - 1u) << (team_in_warp * TG));
    const uint32_t teams = kApplyThreads / TG;
    constexpr uint32_t kShortBatch = 4;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) ticket = atomicAdd(&a.counters[7], teams * kShortBatch);
        __syncthreads();
        const uint32_t base = ticket;
        if (base >= num_runs) break;
#pragma unroll 1
        for (uint32_t r = 0; r < kShortBatch; ++r) {
            const uint32_t run = base + r * teams + threadIdx.x / TG;
            if (run < num_runs) {
                const uint32_t s = a.run_start[run], e = a.run_start[run + 1];
                if (e - s <= (uint32_t)kTeamRun) short_run(g, a, run, s, e, TG, t, team_in_warp, mask);
            }
        }
    }
}

}  // namespace ctr

using namespace ctr;

extern "C" void ctr_opt_hyper(const ctr_opt_t *opt, ctr_hyper_t *out) {
    out->lr = (float)opt->lr;
    out->eps = (float)opt->eps;
    out->one_minus_beta1 = (float)(1.0 - opt->beta1);   // torch forms 1 - beta in double, then rounds
    out->one_minus_beta2 = (float)(1.0 - opt->beta2);
    out->adam_step_size = 0.f;
    if (opt->kind == CTR_OPT_ADAM && opt->step >= 1) {
        const double bc1 = 1.0 - pow(opt->beta1, (double)opt->step);
        const double bc2 = 1.0 - pow(opt->beta2, (double)opt->step);
        out->adam_step_size = (float)(opt->lr * sqrt(bc2) / bc1);
    }
}

extern "C" int64_t ctr_emb_bwd_workspace_bytes(const ctr_group_t *group) {
    static thread_local DevGroup dg;
    int rc = lower_group(group, &dg, /*need_tables=*/false, /*need_out=*/false);
    if (rc != CTR_OK) return rc;
    return plan_layout(dg).total;
}

extern "C" int ctr_emb_bwd_plan(const ctr_group_t *group, void *workspace, int64_t workspace_bytes, void *stream_) {
    static thread_local DevGroup dg;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = lower_group(group, &dg, /*need_tables=*/false, /*need_out=*/false);
    if (rc != CTR_OK) return rc;
    const PlanLayout p = plan_layout(dg);
    CTR_REQUIRE(p.S < (1ll << 31), "group has %lld id slots; must stay below 2^31", (long long)p.S);
    CTR_REQUIRE(workspace != nullptr, "workspace is null");
    if (workspace_bytes < p.total) {
        set_error("workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)p.total);
        return CTR_E_WORKSPACE;
    }
    CTR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
    char *ws = static_cast<char *>(workspace);
    uint32_t *counters = reinterpret_cast<uint32_t *>(ws + p.counters);
    uint32_t *keys_a = reinterpret_cast<uint32_t *>(ws + p.keys_a), *keys_b = reinterpret_cast<uint32_t *>(ws + p.keys_b);
    uint32_t *vals_a = reinterpret_cast<uint32_t *>(ws + p.vals_a), *vals_b = reinterpret_cast<uint32_t *>(ws + p.vals_b);
    note_launch(), reset_counters_kernel<<<1, 32, 0, stream>>>(counters);
    if (p.S > 0) {
        int64_t max_slots = 0;
        for (int i = 0; i < dg.num_features; ++i) {
            const int64_t n = (int64_t)dg.B * dg.f[i].L;
            if (n > max_slots) max_slots = n;
        }
        int64_t bx = (max_slots + 1023) / 1024;  // 4 slots per thread
        if (bx > kNumSMs * 8) bx = kNumSMs * 8;
        if (bx < 1) bx = 1;
        note_launch(), emb_keygen_kernel<<<dim3((unsigned)bx, dg.num_features), 256, 0, stream>>>(dg, keys_a, vals_a);
        CTR_CUDA_OK(cudaGetLastError());
    }
    rc = radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, p.S, p.key_bits, reinterpret_cast<uint32_t *>(ws + p.counts),
                          reinterpret_cast<uint32_t *>(ws + p.spine), stream);
    if (rc < 0) return rc;
    if (p.S > 0 && rc != p.sorted_in_b) {
        set_error("internal: sort parity mismatch");
        return CTR_E_CUDA;
    }
    const uint32_t *sorted_keys = p.sorted_in_b ? keys_b : keys_a;
    rc = find_runs(sorted_keys, p.S, reinterpret_cast<uint32_t *>(ws + p.run_start), counters,
                   reinterpret_cast<uint32_t *>(ws + p.spine), stream);
    if (rc != CTR_OK || p.S == 0) return rc;
    int64_t cb = (p.S + 255) / 256;
    if (cb > kNumSMs * 8) cb = kNumSMs * 8;
    note_launch(), classify_runs_kernel<<<(unsigned)cb, 256, 0, stream>>>(
        reinterpret_cast<const uint32_t *>(ws + p.run_start), counters, reinterpret_cast<uint32_t *>(ws + p.med_list),
        reinterpret_cast<uint32_t *>(ws + p.long_list), reinterpret_cast<uint32_t *>(ws + p.long_cbase),
        reinterpret_cast<uint32_t *>(ws + p.chunk_q));
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_emb_bwd_apply(const ctr_group_t *group, void *workspace, const ctr_opt_t *opt,
                                 int32_t *uniq_feature, int32_t *uniq_row, float *row_grad, int64_t row_grad_stride,
                                 int64_t *num_unique, void *stream_) {
    static thread_local DevGroup dg;
    cudaStream_t stream = (cudaStream_t)stream_;
    CTR_REQUIRE(opt != nullptr, "opt is null");
    const bool updates = opt->kind != CTR_OPT_NONE;
    int rc = lower_group(group, &dg, /*need_tables=*/updates, /*need_out=*/true);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(workspace != nullptr, "workspace is null");
    CTR_REQUIRE(opt->kind >= CTR_OPT_NONE && opt->kind <= CTR_OPT_ADAM, "bad optimizer kind %d", opt->kind);
    int team = 1;
    for (int i = 0; i < dg.num_features; ++i) {
        const DevFeature &f = dg.f[i];
        if (f.G > team) team = f.G;
        if (opt->kind == CTR_OPT_ADAGRAD || opt->kind == CTR_OPT_ROWWISE_ADAGRAD || opt->kind == CTR_OPT_ADAM)
            CTR_REQUIRE(f.state0 != nullptr, "feature %d: optimizer state0 is null", i);
        if (opt->kind == CTR_OPT_ADAM) CTR_REQUIRE(f.state1 != nullptr, "feature %d: optimizer state1 is null", i);
        if (f.vec == 4 && updates) {
            CTR_REQUIRE(f.state0 == nullptr || opt->kind == CTR_OPT_ROWWISE_ADAGRAD ||
                            (reinterpret_cast<uintptr_t>(f.state0) & 15u) == 0,
                        "feature %d: state0 not 16-byte aligned", i);
            CTR_REQUIRE(f.state1 == nullptr || (reinterpret_cast<uintptr_t>(f.state1) & 15u) == 0,
                        "feature %d: state1 not 16-byte aligned", i);
        }
    }
    CTR_REQUIRE(row_grad == nullptr || uniq_row != nullptr, "row_grad needs uniq_row");
    const PlanLayout p = plan_layout(dg);
    char *ws = static_cast<char *>(workspace);
    ApplyArgs a{};
    a.keys = reinterpret_cast<const uint32_t *>(ws + (p.sorted_in_b ? p.keys_b : p.keys_a));
    a.vals = reinterpret_cast<const uint32_t *>(ws + (p.sorted_in_b ? p.vals_b : p.vals_a));
    a.run_start = reinterpret_cast<const uint32_t *>(ws + p.run_start);
    a.counters = reinterpret_cast<uint32_t *>(ws + p.counters);
    a.med_list = reinterpret_cast<uint32_t *>(ws + p.med_list);
    a.long_list = reinterpret_cast<uint32_t *>(ws + p.long_list);
    a.long_cbase = reinterpret_cast<uint32_t *>(ws + p.long_cbase);
    a.long_done = reinterpret_cast<uint32_t *>(ws + p.long_done);
    a.chunk_q = reinterpret_cast<uint32_t *>(ws + p.chunk_q);
    a.partials = reinterpret_cast<float *>(ws + p.partials);
    a.row_floats = (int)p.row_floats;
    a.uniq_feature = uniq_feature;
    a.uniq_row = uniq_row;
    a.row_grad = row_grad;
    a.row_grad_stride = row_grad_stride;
    a.num_unique = num_unique;
    a.kind = opt->kind;
    ctr_opt_hyper(opt, &a.h);
    a.hyper_dev = opt->device_hyper;
    a.team = team;
    if (opt->kind == CTR_OPT_ADAM) CTR_REQUIRE(opt->step >= 1, "Adam step must be >= 1");
    if (p.S == 0) {
        if (num_unique != nullptr) CTR_CUDA_OK(cudaMemsetAsync(num_unique, 0, sizeof(int64_t), stream));
        return CTR_OK;
    }
    const int teams_per_block = kApplyThreads / team;
    int64_t blocks = (p.S + teams_per_block - 1) / teams_per_block;  // upper bound: one run per slot
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    // tickets and per-long-run completion counters are rebuilt by every apply
    CTR_CUDA_OK(cudaMemsetAsync(a.counters + 5, 0, 3 * sizeof(uint32_t), stream));
    CTR_CUDA_OK(cudaMemsetAsync(a.long_done, 0, (size_t)p.max_long * sizeof(uint32_t), stream));
    note_launch(), emb_bwd_apply_kernel<<<kNumSMs * 5, kApplyThreads, 0, stream>>>(dg, a);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}
