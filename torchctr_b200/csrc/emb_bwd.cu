// K2: backward of the pooled lookup fused with the optimizer step.
//
// Replaces autograd's aten::embedding_dense_backward behind torchctr/models/dnn.py:57-58 and
// the whole-table optimizer.step() of torchctr/trainer.py:303.  No dense [V, D] gradient is
// ever built:
//   plan   keygen (row key = table base + mapped row, value = slot) -> stable radix sort ->
//          (optionally) runs of equal keys listed by a head-flag scan;
//   apply  a position-parallel segmented sweep over the sorted pairs.  Every warp owns a range
//          of consecutive sorted positions; a team of G lanes gathers grad_out[bag] of one
//          position (all gathers independent, four in flight per lane), teams combine equal
//          keys with a segmented shuffle scan, a register carry joins the sub-steps, and the
//          team that holds the last position of a run applies SGD / Adagrad / row-wise
//          Adagrad / lazy Adam to that row in place.  Work per warp does not depend on the run
//          lengths, so Zipf-hot rows cost the same per position as cold ones.  Runs that cross
//          a range boundary leave per-range partial sums; a small fix-up kernel adds them in
//          range order and updates those rows.  Summation order is fixed by the layout, so the
//          result is deterministic.
#include <stdlib.h>
#include <string.h>

#include "sort.cuh"

namespace ctr {

constexpr int kApplyThreads = 256;
constexpr int kApplyWarps = kApplyThreads / kWarp;
constexpr uint32_t kInvalidKey = 0xffffffffu;
constexpr int kMinRange = 32;   // sorted positions per warp range: at least this many

// ---- workspace layout ---------------------------------------------------------------------
struct PlanLayout {
    int64_t S;           // id slots in the group
    int key_bits;
    int sorted_in_b;     // which ping-pong buffer holds the sorted pairs
    // offsets in bytes from the workspace base
    int64_t counters, keys_a, keys_b, vals_a, vals_b, counts, spine, run_start, run_of_pos;
    // apply-time scratch (rebuilt by every apply; sized by the group being applied)
    int64_t range_flags, tail_part, max_ranges, row_floats, total;
};
// counters (u32): [0] runs, [1] unique valid rows, [2] 1 when run_start was built by the plan, [3] number of
// pairs when it is only known on the device (sharded owner side); [8..] per-peer offsets of that gather
constexpr int kNumCounters = 8;
constexpr int kCounterWords = 64;
constexpr int kMetaSrc = 8, kMetaDst = 24, kMetaCnt = 40;   // u32 [CTR_MAX_WORLD] each, inside the counter block

static int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

// scale > 1: owner side of the sharded backward, where up to `scale` ranks' slots can land on this rank;
// capacity > 0: the pair capacity is given outright (owner side of the de-duplicated exchange)
static PlanLayout plan_layout(const DevGroup &g, int scale = 1, int64_t capacity = 0) {
    PlanLayout p{};
    int64_t S = 0;
    uint64_t rows = 0;
    for (int i = 0; i < g.num_features; ++i) {
        S += (int64_t)g.B * g.f[i].L;
        rows += g.f[i].num_rows;
    }
    S *= scale;
    if (capacity > 0) S = capacity;
    p.S = S;
    int bits = 0;
    while ((rows >> bits) != 0) ++bits;  // bit_length(rows): 2^bits - 1 > every valid key
    p.key_bits = bits < 1 ? 1 : bits;
    const int passes = sort_num_passes(p.key_bits);
    const bool segmented = scale == 1 && capacity == 0;   // the single-GPU plan: one independent sort per table
    p.sorted_in_b = segmented ? 1 : (passes & 1);
    int64_t off = 0;
    p.counters = off; off = align256(off + kCounterWords * 4);
    p.keys_a = off; off = align256(off + (S + 1) * 4);
    p.keys_b = off; off = align256(off + (S + 1) * 4);
    p.vals_a = off; off = align256(off + S * 4);
    p.vals_b = off; off = align256(off + S * 4);
    int64_t counts = sort_counts_elems(S);
    if (segmented) {
        int64_t tiles = 0;
        for (int i = 0; i < g.num_features; ++i) tiles += sort_num_tiles((int64_t)g.B * g.f[i].L);
        const int64_t seg = seg_counts_elems(tiles, g.num_features);
        if (seg > counts) counts = seg;
    }
    p.counts = off; off = align256(off + counts * 4);
    const int64_t spine = scan_spine_elems(counts > S ? counts : S) + 8;
    p.spine = off; off = align256(off + spine * 4);
    p.run_start = off; off = align256(off + (S + 2) * 4);
    p.run_of_pos = off; off = align256(off + (S + 1) * 4);
    p.row_floats = 128;   // one float4 per lane of a full warp: any team width fits
    p.max_ranges = S / kMinRange + 2;
    p.range_flags = off; off = align256(off + (p.max_ranges + 64) * 4);   // 64 words of header: [0] = range ticket
    // partial sums of runs that cross a range boundary: [range][team lane] float4
    int row_floats = 4;
    for (int i = 0; i < g.num_features; ++i) row_floats = max(row_floats, g.f[i].G * 4);
    row_floats += 4;      // one more float4 slot: the summed per-bag coefficient of the fused DeepFM terms
    p.row_floats = row_floats;
    p.tail_part = off; off = align256(off + p.max_ranges * row_floats * 4);
    p.total = off;
    return p;
}

// ---- keygen -------------------------------------------------------------------------------
// Also builds the digit histograms of every pass of the table's sort (the sort then needs no pass of its own over the
// keys): hist[table][pass][digit], digits taken from the table-local row (sort.cuh, segmented variant).
__global__ void __launch_bounds__(256)
    emb_keygen_kernel(const __grid_constant__ DevGroup g, const __grid_constant__ SegSortDesc sd, uint32_t *__restrict__ keys,
                      uint32_t *__restrict__ vals, uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh[kSegMaxPasses][kMaxRadix];
    const int fi = blockIdx.y;
    const DevFeature &f = g.f[fi];
    const int passes = sd.passes[fi], bits = sd.digit_bits[fi], kb = sd.key_bits[fi];
    const int radix = 1 << bits;
    for (int i = threadIdx.x; i < passes * kMaxRadix; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int64_t slot_base = sd.slot_base[fi];
    const int64_t n = (int64_t)g.B * f.L;
    const int lane = threadIdx.x & 31;
    const uint32_t dmask = (uint32_t)radix - 1u;
    const int64_t nround = (n + 31) / 32 * 32;   // whole warps iterate together (match.any below)
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += (int64_t)gridDim.x * blockDim.x) {
        const bool valid = j < n;
        uint32_t key = kInvalidKey;
        if (valid) {
            const int32_t row = map_index(f, __ldg(f.ids + j));
            if (row >= 0) key = f.row_base + (uint32_t)row;
            keys[slot_base + j] = key;
            vals[slot_base + j] = (uint32_t)j;
        }
        const uint32_t local = seg_local_key(key, f.row_base, kb);
        for (int p = 0; p < passes; ++p) {
            const uint32_t d = (local >> (p * bits)) & dmask;
            const uint32_t peers = __match_any_sync(kFull, valid ? d : (uint32_t)(kMaxRadix + lane));
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&sh[p][d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    uint32_t *dst = hist + (size_t)fi * kSegMaxPasses * kMaxRadix;
    for (int i = threadIdx.x; i < passes * radix; i += blockDim.x) {
        const uint32_t c = sh[i / radix][i % radix];
        if (c) atomicAdd(&dst[(i / radix) * kMaxRadix + (i % radix)], c);
    }
}

// the per-table sort layout of a group (host)
static void seg_desc(const DevGroup &g, SegSortDesc *sd) {
    memset(sd, 0, sizeof(*sd));
    sd->num_tables = g.num_features;
    uint32_t tile = 0, slot = 0;
    int maxp = 1;
    for (int i = 0; i < g.num_features; ++i) {
        const int64_t n = (int64_t)g.B * g.f[i].L;
        sd->tile_base[i] = tile;
        sd->slot_base[i] = slot;
        sd->row_base[i] = g.f[i].row_base;
        int bits = 0;
        while (((uint64_t)g.f[i].num_rows >> bits) != 0) ++bits;       // bit_length(V): 2^bits - 1 >= V > every valid row
        if (bits < 1) bits = 1;
        const int passes = seg_num_passes(bits);
        sd->key_bits[i] = (uint8_t)bits;
        sd->passes[i] = (uint8_t)passes;
        sd->digit_bits[i] = (uint8_t)((bits + passes - 1) / passes);
        if (passes > maxp) maxp = passes;
        tile += (uint32_t)sort_num_tiles(n);
        slot += (uint32_t)n;
    }
    sd->tile_base[g.num_features] = tile;
    sd->slot_base[g.num_features] = slot;
    sd->max_passes = maxp;
}

__global__ void reset_counters_kernel(uint32_t *counters, uint32_t runs_built) {
    if (threadIdx.x < kNumCounters) counters[threadIdx.x] = threadIdx.x == 2 ? runs_built : 0u;
}

// ---- apply ----------------------------------------------------------------------------------
struct ApplyArgs {
    const uint32_t *keys;       // sorted
    const uint32_t *vals;       // sorted with the keys: slot inside the feature (bag * L + l)
    const uint32_t *run_start;  // only read when uniq outputs are requested
    uint32_t *counters;
    uint32_t *ticket;           // next range to hand out (zeroed by every apply)
    uint32_t *range_flags;      // [range] 0: not finished yet, 1: finished, 3: finished and tail_part holds the open run
    float *tail_part;           // [range, row_floats]
    int row_floats;
    uint32_t S;                 // sorted positions (upper bound when S_dev is set)
    const uint32_t *S_dev;      // device-side count of sorted positions (sharded owner side), or null
    const float *peer_grads[CTR_MAX_WORLD];   // sharded owner side: grad_out of every rank (slot's top 4 bits = rank)
    int p2p;
    uint32_t range;             // positions per warp range (multiple of the positions per outer iteration)
    uint32_t num_ranges;
    int32_t *uniq_feature;
    int32_t *uniq_row;
    float *row_grad;
    int64_t row_grad_stride;
    unsigned long long *num_unique;
    int kind;
    ctr_hyper_t h;              // used when hyper_dev == nullptr
    const ctr_hyper_t *hyper_dev;  // device copy read at run time (CUDA-graph replays see new values)
    int team;                   // lanes per sorted position (max G of the group), power of two
    uint32_t *status;           // group status word (may be null)
    const float *extra_grad;    // [B] dL/d extra[bag] (twin tables + FM term), or null
    const float *fm_sum;        // [B, D] sum over the fields of the pooled vectors, or null
    const float *peer_extra[CTR_MAX_WORLD];    // sharded owner side: extra_grad / fm_sum of the rank that sent the slot
    const float *peer_fm_sum[CTR_MAX_WORLD];
    int fm_row_only;            // FM term with the "c * fm_sum" part already folded into the gradients (ctr_fm_pack_grads):
                                // only "- row * sum c" is left to do here
    int gather_prefetch;        // single-id sweep: pull the next outer iteration's gradient slices into L2 one iteration ahead
};

// last feature whose row_base <= key.  sf is the kernel's parameter copy of the features: sorted positions
// next to each other almost always belong to one table, so these constant-bank reads are warp-uniform.
__device__ __forceinline__ int find_feature(const DevFeature *sf, int nf, uint32_t key) {
    int lo = 0, hi = nf - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sf[mid].row_base <= key) lo = mid; else hi = mid - 1;
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

__device__ __forceinline__ float4 load_grad_part(const float *gout, int64_t gstride, const DevFeature &f, uint32_t bag,
                                                 int g_lane) {
    const float *src = gout + (int64_t)bag * gstride + f.out_col;
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

// index of the run that contains sorted position p (uniq outputs only): last r with run_start[r] <= p
__device__ __forceinline__ uint32_t run_of_position(const ApplyArgs &a, uint32_t p) {
    uint32_t lo = 0, hi = a.counters[0];
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.run_start[mid] <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// Lanes [0, G) of a team hold the summed gradient of (feature f, row); mask names the team; pos is any sorted
// position of the run (names the run when the unique-row outputs are requested).
__device__ __forceinline__ void update_row(const DevFeature &f, const ApplyArgs &a, uint32_t pos, int fi, uint32_t row,
                                           int g_lane, bool col_ok, float4 gr, unsigned mask, int team_lanes) {
    if (a.uniq_row != nullptr && a.counters[2] == 0u) {   // plan was built without the run list
        if (a.status != nullptr && g_lane == 0) atomicOr(a.status, CTR_STATUS_NO_RUNS);
    } else if (a.uniq_row != nullptr) {
        const uint32_t run = run_of_position(a, pos);
        if (g_lane == 0) {
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
    }
    if (a.kind == CTR_OPT_NONE) return;
    if (a.kind == CTR_OPT_GRAD_OUT) {          // replicated table: the summed gradient goes to the dense buffer behind state0
        if (col_ok) {
            float *dst = f.state0 + (size_t)row * f.D + (size_t)g_lane * f.vec;
            if (f.vec == 4) *reinterpret_cast<float4 *>(dst) = gr;
            else *dst = gr.x;
        }
        return;
    }
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

// one shared copy of the update code for the general (many-option) kernels
__device__ __noinline__ void update_row_call(const DevFeature &f, const ApplyArgs &a, uint32_t pos, int fi, uint32_t row,
                                             int g_lane, bool col_ok, float4 gr, unsigned mask, int team_lanes) {
    update_row(f, a, pos, fi, row, g_lane, col_ok, gr, mask, team_lanes);
}

template <int VEC>
__device__ __forceinline__ float4 shfl_up4(float4 v, int delta) {
    float4 o;
    o.x = __shfl_up_sync(kFull, v.x, delta);
    if (VEC == 4) {
        o.y = __shfl_up_sync(kFull, v.y, delta);
        o.z = __shfl_up_sync(kFull, v.z, delta);
        o.w = __shfl_up_sync(kFull, v.w, delta);
    } else {
        o.y = o.z = o.w = 0.f;
    }
    return o;
}
template <int VEC>
__device__ __forceinline__ float4 shfl_idx4(float4 v, int src) {
    float4 o;
    o.x = __shfl_sync(kFull, v.x, src);
    if (VEC == 4) {
        o.y = __shfl_sync(kFull, v.y, src);
        o.z = __shfl_sync(kFull, v.z, src);
        o.w = __shfl_sync(kFull, v.w, src);
    } else {
        o.y = o.z = o.w = 0.f;
    }
    return o;
}
__device__ __forceinline__ void add4(float4 &x, const float4 &o) { x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w; }

// range flags: plain relaxed accesses at gpu scope; the (rare) publisher of a partial sum fences before the store
// and its reader fences after the spin, so the common case pays for no fence at all
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// The sweep.  A warp is P = 32 / TG teams of TG lanes.  An outer iteration covers Q = P * U consecutive sorted
// positions, U = min(4, 32 / P) per team: lane i < Q loads the key / slot of position p0 + i, team t owns positions
// p0 + t * U .. + U - 1.  Where runs end is known from one ballot (key != next key), so the reduction needs no key
// traffic: a team adds up its own positions, the teams' trailing sums are combined with a segmented shuffle scan
// whose segment flags come from the ballot, and every team then walks its positions once more, starting from what
// flowed in from the left, and updates the row wherever a run ends.  `carry` joins the outer iterations.
// Ranges are handed out by ticket so that a warp may wait for ranges before its own: the run that was open when
// the range started is finished by the warp that sees its end, which adds the partial sums the earlier ranges
// published (in range order) -- no second kernel.
// VEC = 4: lanes hold float4 pieces of the rows (some feature has D % 4 == 0); VEC = 1: every feature is scalar-laned.
// (Single-id groups of one width take emb_bwd_sweep_l1_kernel below instead.)
template <int VEC>
__global__ void __launch_bounds__(kApplyThreads, 3)
    emb_bwd_sweep_kernel(const __grid_constant__ DevGroup g, const __grid_constant__ ApplyArgs a_in) {
    __shared__ ApplyArgs sa;
    __shared__ uint32_t s_first_range;
    if (threadIdx.x == 0) {
        sa = a_in;
        if (a_in.hyper_dev != nullptr) sa.h = *a_in.hyper_dev;
        s_first_range = atomicAdd(a_in.ticket, (uint32_t)kApplyWarps);
    }
    __syncthreads();
    const ApplyArgs &a = sa;
    const DevFeature *sf = g.f;
    const int lane = threadIdx.x & 31;
    const uint32_t w = s_first_range + (threadIdx.x >> 5);
    if (w >= a.num_ranges) return;
    const int nf = g.num_features;
    const float *gout_local = g.out;
    const int64_t gstride = g.out_stride;
    const int TG = a.team;
    const int P = kWarp / TG;
    const int U = P >= 8 ? (kWarp / P < 4 ? kWarp / P : 4) : 4;
    const int Q = P * U;
    const int team = lane / TG;
    const int t = lane & (TG - 1);
    const unsigned tmask = TG == kWarp ? kFull : (((1u << TG) - 1u) << (team * TG));
    const unsigned umask = (1u << U) - 1u;
    const uint32_t S = a.S_dev != nullptr ? min(*a.S_dev, a.S) : a.S;
    const uint32_t first = w * a.range;
    if (first >= S) {                      // beyond the device-side count: nothing to do, nobody waits for this range
        if (lane == 0) st_relaxed_u32(a.range_flags + w, 1u);
        return;
    }
    const uint32_t end = min(first + a.range, S);
    const uint32_t first_key = a.keys[first];
    // the run at the start of the range began in an earlier range: its end (if inside) is finished by look-back
    bool head_pending = first > 0 && first_key != kInvalidKey && a.keys[first - 1] == first_key;
    bool have_head = false;                 // this team holds the sum of that run's part inside this range
    float4 head_acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool carry_open = false;                // the position before p0 (inside this range) did not end a run
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
    int fi = 0;                             // per-team cache of the table the current key belongs to
    uint32_t f_lo = 1, f_hi = 0;
    uint32_t closed = 0;

    uint32_t k = kInvalidKey, nk = kInvalidKey, slot = 0;
    if (lane < Q && first + lane < end) {
        k = a.keys[first + lane];
        slot = a.vals[first + lane];
        if (first + lane + 1 < S) nk = a.keys[first + lane + 1];
    }
    for (uint32_t p0 = first; p0 < end; p0 += Q) {
        // the next outer iteration's keys are requested before this one's gradients (one latency off the chain)
        uint32_t k_next = kInvalidKey, nk_next = kInvalidKey, slot_next = 0;
        const uint32_t pn = p0 + Q + lane;
        if (lane < Q && pn < end) {
            k_next = a.keys[pn];
            slot_next = a.vals[pn];
            if (pn + 1 < S) nk_next = a.keys[pn + 1];
        }
        const bool is_tail = k != kInvalidKey && nk != k;
        const unsigned tails = __ballot_sync(kFull, is_tail);
        if (is_tail && a.kind != CTR_OPT_NONE) {
            // this lane's position ends a run: pull the row (and its optimizer state) towards L2 now, so that the
            // update at the end of the iteration does not wait for DRAM after the gradients already did
            if (k < f_lo || k >= f_hi) {
                fi = find_feature(sf, nf, k);
                f_lo = sf[fi].row_base;
                f_hi = f_lo + sf[fi].num_rows;
            }
            const DevFeature &f = sf[fi];
            const size_t off = (size_t)(k - f_lo) * f.D;
            for (int b = 0; b < f.D; b += 32) {
                prefetch_l2(f.table + off + b);
                if (a.kind == CTR_OPT_ADAGRAD || a.kind == CTR_OPT_ADAM) prefetch_l2(f.state0 + off + b);
                if (a.kind == CTR_OPT_ADAM) prefetch_l2(f.state1 + off + b);
            }
        }
        float4 v[4];
        uint32_t kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            kk[u] = kInvalidKey;
            if (u < U) {
                const int src = team * U + u;
                kk[u] = __shfl_sync(kFull, k, src);
                uint32_t sl = __shfl_sync(kFull, slot, src);
                const float *gout = gout_local;
                if (a.p2p) {               // the gradient lives on the rank that sent this slot
                    gout = a.peer_grads[sl >> 28];
                    sl &= 0x0fffffffu;
                }
                if (kk[u] != kInvalidKey) {
                    if (kk[u] < f_lo || kk[u] >= f_hi) {
                        fi = find_feature(sf, nf, kk[u]);
                        f_lo = sf[fi].row_base;
                        f_hi = f_lo + sf[fi].num_rows;
                    }
                    const DevFeature &f = sf[fi];
                    if (t < f.G && t * f.vec < f.D) {
                        const uint32_t bag = f.L == 1 ? sl : sl / (uint32_t)f.L;
                        const float4 x = load_grad_part(gout, gstride, f, bag, t);
                        if (f.id_weight != nullptr || f.pooling == CTR_POOL_MEAN) {
                            const float c = slot_coef(f, sl, bag);
                            v[u] = make_float4(c * x.x, c * x.y, c * x.z, c * x.w);
                        } else {
                            v[u] = x;
                        }
                    }
                }
            }
        }
        // ends of runs among this team's positions; bit U - 1 = the team's last position
        const unsigned m = (tails >> (team * U)) & umask;
        const bool single = (m & (umask >> 1)) == 0u;          // no run ends strictly inside the team
        const bool prev_open = team > 0 ? ((tails >> (team * U - 1)) & 1u) == 0u : carry_open;
        // sum of the team's trailing run
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (u < U) {
                add4(out, v[u]);
                if (u < U - 1 && ((m >> u) & 1u)) out = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (team == 0 && single && carry_open) add4(out, carry);
        // segmented inclusive scan over the teams; a segment starts at a team unless it is one run joined to its left
        const unsigned starts = __ballot_sync(kFull, t == 0 && !(single && prev_open));
        for (int d = 1; d < P; d <<= 1) {
            const float4 o = shfl_up4<VEC>(out, d * TG);
            // lanes of teams team - d + 1 .. team
            const unsigned span = team >= d ? ((d * TG == kWarp ? kFull : ((1u << (d * TG)) - 1u)) << ((team - d + 1) * TG)) : kFull;
            if (team >= d && (starts & span) == 0u) add4(out, o);
        }
        float4 acc = shfl_up4<VEC>(out, TG);                   // what flows in from the left
        if (team == 0) acc = carry;
        if (!prev_open) acc = make_float4(0.f, 0.f, 0.f, 0.f);
        carry = shfl_idx4<VEC>(out, (P - 1) * TG + t);
        carry_open = ((tails >> (Q - 1)) & 1u) == 0u;
        // second walk: close the runs that end at this team's positions
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (u < U) {
                add4(acc, v[u]);
                if ((m >> u) & 1u) {
                    const int idx = team * U + u;
                    if (head_pending && (tails & ((1u << idx) - 1u)) == 0u) {   // first end of a run in this range
                        have_head = true;
                        head_acc = acc;
                    } else {
                        const uint32_t key = kk[u];
                        if (key < f_lo || key >= f_hi) {
                            fi = find_feature(sf, nf, key);
                            f_lo = sf[fi].row_base;
                            f_hi = f_lo + sf[fi].num_rows;
                        }
                        const DevFeature &f = sf[fi];
                        update_row_call(f, a, p0 + idx, fi, key - f.row_base, t, t < f.G && t * f.vec < f.D, acc, tmask, TG);
                    }
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        if (tails != 0u) head_pending = false;
        closed += (uint32_t)__popc(tails);
        k = k_next; nk = nk_next; slot = slot_next;
    }
    // publish: the last run of the range continues in the next one -> leave its partial sum for whoever ends it
    uint32_t flag = 1u;
    if (end < S) {
        const uint32_t kl = a.keys[end - 1];
        if (kl != kInvalidKey && a.keys[end] == kl) {
            if (team == 0) reinterpret_cast<float4 *>(a.tail_part + (size_t)w * a.row_floats)[t] = carry;
            flag = 3u;
            __threadfence();
            __syncwarp();
        }
    }
    if (lane == 0) {
        st_relaxed_u32(a.range_flags + w, flag);
        if (a.num_unique != nullptr && closed != 0) atomicAdd(a.num_unique, (unsigned long long)closed);
    }
    // finish the run that was open when this range started (at most one per range)
    const unsigned head_lanes = __ballot_sync(kFull, have_head);
    if (head_lanes != 0u) {
        const int src_team = (__ffs(head_lanes) - 1) / TG;
        const float4 hv = shfl_idx4<4>(head_acc, src_team * TG + t);
        // first range of the run: smallest j whose last key reaches first_key (keys are sorted; j = w - 1 qualifies)
        uint32_t lo = 0, hi = w - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            // padding keys (0xffffffff) sit at the end of every table's segment: "+ 1" wraps them to 0, below every valid key
            if (a.keys[(mid + 1) * a.range - 1] + 1u >= first_key + 1u) hi = mid; else lo = mid + 1;
        }
        // wait until every range the run passed through has published, fence ONCE, then read the partial sums
        for (uint32_t j = lo + lane; j < w; j += kWarp)
            while (ld_relaxed_u32(a.range_flags + j) == 0u) { }
        __syncwarp();
        __threadfence();
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t j = lo + team; j < w; j += P)
            add4(sum, __ldcg(reinterpret_cast<const float4 *>(a.tail_part + (size_t)j * a.row_floats) + t));
        for (int off = TG; off < kWarp; off <<= 1) {
            float4 o;
            o.x = __shfl_xor_sync(kFull, sum.x, off); o.y = __shfl_xor_sync(kFull, sum.y, off);
            o.z = __shfl_xor_sync(kFull, sum.z, off); o.w = __shfl_xor_sync(kFull, sum.w, off);
            add4(sum, o);
        }
        add4(sum, hv);
        if (team == 0) {
            const int hf = find_feature(sf, nf, first_key);
            const DevFeature &f = sf[hf];
            update_row_call(f, a, first, hf, first_key - f.row_base, t, t < f.G && t * f.vec < f.D, sum,
                            TG == kWarp ? kFull : ((1u << TG) - 1u), TG);
        }
    }
}


// ---- the sweep for Criteo-shaped groups: single-id bags, one width D = 4 * TG, aligned slices, sum pooling ----------
// Same algorithm as emb_bwd_sweep_kernel with everything about the lane layout known at compile time (P = 32 / TG teams,
// U = 4 positions per team, Q = P * U positions per outer iteration: no local-memory arrays, no runtime divisions), and,
// with EXTRA, the DeepFM terms folded in: every position also carries c = dL/d extra[bag]; the gradient of its row is
//     grad_out[bag, cols] + c * (fm_sum[bag, :] - row)        (FM second order; the "- row * sum c" part is applied once
// per run, when the row has been loaded for its update anyway) and the gradient of its twin (first-order weight) is
// sum c.  That removes the FM backward pass over [B, F * D] and the whole D = 1 launch group of DeepFM.
template <bool EXTRA>
__device__ __forceinline__ void update_row_l1(const DevFeature &f, const ApplyArgs &a, uint32_t pos, int fi, uint32_t row, int t,
                                              float4 gr, float cs, bool fm, unsigned mask, int team_lanes) {
    if (!EXTRA) {
        update_row(f, a, pos, fi, row, t, true, gr, mask, team_lanes);
        return;
    }
    const size_t off = (size_t)row * f.D + (size_t)t * 4;
    float4 w = *reinterpret_cast<float4 *>(f.table + off);
    if (fm) {
        gr.x = fmaf(-cs, w.x, gr.x); gr.y = fmaf(-cs, w.y, gr.y); gr.z = fmaf(-cs, w.z, gr.z); gr.w = fmaf(-cs, w.w, gr.w);
    }
    if (a.kind == CTR_OPT_GRAD_OUT) {          // replicated table: gradients out, no update (see ctr_b200.h)
        *reinterpret_cast<float4 *>(f.state0 + off) = gr;
        if (t == 0 && f.twin_table != nullptr) f.twin_state0[row] = cs;
        return;
    }
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
        float sq = gr.x * gr.x + gr.y * gr.y + gr.z * gr.z + gr.w * gr.w;
        for (int o = 1; o < team_lanes; o <<= 1) sq += __shfl_xor_sync(mask, sq, o);
        const float acc = f.state0[row] + sq / (float)f.D;
        const float den = __fsqrt_rn(acc) + a.h.eps;
        __syncwarp(mask);
        if (t == 0) f.state0[row] = acc;
        w.x -= a.h.lr * __fdiv_rn(gr.x, den); w.y -= a.h.lr * __fdiv_rn(gr.y, den);
        w.z -= a.h.lr * __fdiv_rn(gr.z, den); w.w -= a.h.lr * __fdiv_rn(gr.w, den);
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
    if (t == 0 && f.twin_table != nullptr) {     // the first-order weight of the same id: gradient = sum of c over the run
        float tw = f.twin_table[row];
        if (a.kind == CTR_OPT_SGD) {
            tw -= a.h.lr * cs;
        } else if (a.kind == CTR_OPT_ADAGRAD || a.kind == CTR_OPT_ROWWISE_ADAGRAD) {   // one column: the two coincide
            float s = f.twin_state0[row];
            tw = adagrad_elem(tw, s, cs, a.h.lr, a.h.eps);
            f.twin_state0[row] = s;
        } else {
            float m = f.twin_state0[row], v = f.twin_state1[row];
            tw = adam_elem(tw, m, v, cs, a);
            f.twin_state0[row] = m;
            f.twin_state1[row] = v;
        }
        f.twin_table[row] = tw;
    }
}

template <int TG, bool EXTRA, int MINB>
__global__ void __launch_bounds__(kApplyThreads, MINB)
    emb_bwd_sweep_l1_kernel(const __grid_constant__ DevGroup g, const __grid_constant__ ApplyArgs a_in) {
    constexpr int P = kWarp / TG, U = 4, Q = P * U;
    constexpr int D = 4 * TG;
    constexpr unsigned umask = (1u << U) - 1u;
    __shared__ ApplyArgs sa;
    __shared__ uint32_t s_first_range;
    if (threadIdx.x == 0) {
        sa = a_in;
        if (a_in.hyper_dev != nullptr) sa.h = *a_in.hyper_dev;
        s_first_range = atomicAdd(a_in.ticket, (uint32_t)kApplyWarps);
    }
    __syncthreads();
    const ApplyArgs &a = sa;
    const DevFeature *sf = g.f;
    const int lane = threadIdx.x & 31;
    const uint32_t w = s_first_range + (threadIdx.x >> 5);
    if (w >= a.num_ranges) return;
    const int nf = g.num_features;
    const float *gout_local = g.out;
    const int64_t gstride = g.out_stride;
    const int team = lane / TG;
    const int t = lane % TG;
    const unsigned tmask = TG == kWarp ? kFull : (((1u << TG) - 1u) << (team * TG));
    const bool fm_gather = EXTRA && a.fm_sum != nullptr;          // add c * fm_sum[bag] to every slot's gradient
    const bool fm = fm_gather || (EXTRA && a.fm_row_only != 0);    // subtract row * sum c once per row
    const uint32_t S = a.S_dev != nullptr ? min(*a.S_dev, a.S) : a.S;
    const uint32_t first = w * a.range;
    if (first >= S) {
        if (lane == 0) st_relaxed_u32(a.range_flags + w, 1u);
        return;
    }
    const uint32_t end = min(first + a.range, S);
    const uint32_t first_key = a.keys[first];
    bool head_pending = first > 0 && first_key != kInvalidKey && a.keys[first - 1] == first_key;
    bool have_head = false;
    float4 head_acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float head_c = 0.f;
    bool carry_open = false;
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
    float carry_c = 0.f;
    int fi = 0;
    uint32_t f_lo = 1, f_hi = 0;
    uint32_t closed = 0;

    // keys / slots are loaded two outer iterations ahead: the pairs of the NEXT iteration are in registers when this one starts,
    // so its gradient slices (random 64-byte reads over [B, stride], the latency this kernel waits on) can be pulled into L2 now
    uint32_t k = kInvalidKey, nk = kInvalidKey, slot = 0;
    uint32_t k1 = kInvalidKey, nk1 = kInvalidKey, slot1 = 0;
    if (lane < Q && first + lane < end) {
        k = a.keys[first + lane];
        slot = a.vals[first + lane];
        if (first + lane + 1 < S) nk = a.keys[first + lane + 1];
    }
    if (lane < Q && first + Q + lane < end) {
        k1 = a.keys[first + Q + lane];
        slot1 = a.vals[first + Q + lane];
        if (first + Q + lane + 1 < S) nk1 = a.keys[first + Q + lane + 1];
    }
    const bool gather_ahead = a.gather_prefetch != 0 && !a.p2p;
    for (uint32_t p0 = first; p0 < end; p0 += Q) {
        uint32_t k_next = kInvalidKey, nk_next = kInvalidKey, slot_next = 0;
        const uint32_t pn = p0 + 2 * Q + lane;
        if (lane < Q && pn < end) {
            k_next = a.keys[pn];
            slot_next = a.vals[pn];
            if (pn + 1 < S) nk_next = a.keys[pn + 1];
        }
        if (gather_ahead && k1 != kInvalidKey) {
            if (k1 < f_lo || k1 >= f_hi) {
                fi = find_feature(sf, nf, k1);
                f_lo = sf[fi].row_base;
                f_hi = f_lo + sf[fi].num_rows;
            }
            const float *gp = gout_local + (int64_t)slot1 * gstride + sf[fi].out_col;
#pragma unroll
            for (int b = 0; b < D; b += 8) prefetch_l2(gp + b);
        }
        const bool is_tail = k != kInvalidKey && nk != k;
        const unsigned tails = __ballot_sync(kFull, is_tail);
        if (is_tail && a.kind != CTR_OPT_NONE) {     // pull the row (and its optimizer state) towards L2 now
            if (k < f_lo || k >= f_hi) {
                fi = find_feature(sf, nf, k);
                f_lo = sf[fi].row_base;
                f_hi = f_lo + sf[fi].num_rows;
            }
            const DevFeature &f = sf[fi];
            const size_t off = (size_t)(k - f_lo) * D;
#pragma unroll
            for (int b = 0; b < D; b += 32) {
                prefetch_l2(f.table + off + b);
                if (a.kind == CTR_OPT_ADAGRAD || a.kind == CTR_OPT_ADAM) prefetch_l2(f.state0 + off + b);
                if (a.kind == CTR_OPT_ADAM) prefetch_l2(f.state1 + off + b);
            }
            if (EXTRA && f.twin_table != nullptr) {
                prefetch_l2(f.twin_table + (k - f_lo));
                if (a.kind != CTR_OPT_SGD) prefetch_l2(f.twin_state0 + (k - f_lo));
            }
        }
        float4 v[U];
        float c[U];
        uint32_t kk[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            c[u] = 0.f;
            const int src = team * U + u;
            kk[u] = __shfl_sync(kFull, k, src);
            uint32_t sl = __shfl_sync(kFull, slot, src);
            const float *gout = gout_local;
            const float *ex = a.extra_grad, *fs = a.fm_sum;
            if (a.p2p) {               // the gradient lives on the rank that sent this slot
                const uint32_t src_rank = sl >> 28;
                gout = a.peer_grads[src_rank];
                if (EXTRA) {
                    ex = a.peer_extra[src_rank];
                    fs = a.peer_fm_sum[src_rank];
                }
                sl &= 0x0fffffffu;
            }
            if (kk[u] != kInvalidKey) {
                if (kk[u] < f_lo || kk[u] >= f_hi) {
                    fi = find_feature(sf, nf, kk[u]);
                    f_lo = sf[fi].row_base;
                    f_hi = f_lo + sf[fi].num_rows;
                }
                v[u] = __ldg(reinterpret_cast<const float4 *>(gout + (int64_t)sl * gstride + sf[fi].out_col) + t);
                if (EXTRA) {
                    c[u] = __ldg(ex + sl);
                    if (fm_gather) {
                        const float4 s4 = __ldg(reinterpret_cast<const float4 *>(fs + (size_t)sl * D) + t);
                        v[u].x = fmaf(c[u], s4.x, v[u].x); v[u].y = fmaf(c[u], s4.y, v[u].y);
                        v[u].z = fmaf(c[u], s4.z, v[u].z); v[u].w = fmaf(c[u], s4.w, v[u].w);
                    }
                }
            }
        }
        const unsigned m = (tails >> (team * U)) & umask;
        const bool single = (m & (umask >> 1)) == 0u;
        const bool prev_open = team > 0 ? ((tails >> (team * U - 1)) & 1u) == 0u : carry_open;
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        float oc = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            add4(out, v[u]);
            oc += c[u];
            if (u < U - 1 && ((m >> u) & 1u)) {
                out = make_float4(0.f, 0.f, 0.f, 0.f);
                oc = 0.f;
            }
        }
        if (team == 0 && single && carry_open) {
            add4(out, carry);
            oc += carry_c;
        }
        const unsigned starts = __ballot_sync(kFull, t == 0 && !(single && prev_open));
#pragma unroll
        for (int d = 1; d < P; d <<= 1) {
            const float4 o = shfl_up4<4>(out, d * TG);
            const float o_c = EXTRA ? __shfl_up_sync(kFull, oc, d * TG) : 0.f;
            const unsigned span = team >= d ? ((d * TG == kWarp ? kFull : ((1u << (d * TG)) - 1u)) << ((team - d + 1) * TG)) : kFull;
            if (team >= d && (starts & span) == 0u) {
                add4(out, o);
                oc += o_c;
            }
        }
        float4 acc = shfl_up4<4>(out, TG);
        float acc_c = EXTRA ? __shfl_up_sync(kFull, oc, TG) : 0.f;
        if (team == 0) {
            acc = carry;
            acc_c = carry_c;
        }
        if (!prev_open) {
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
            acc_c = 0.f;
        }
        carry = shfl_idx4<4>(out, (P - 1) * TG + t);
        carry_c = EXTRA ? __shfl_sync(kFull, oc, (P - 1) * TG + t) : 0.f;
        carry_open = ((tails >> (Q - 1)) & 1u) == 0u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            add4(acc, v[u]);
            acc_c += c[u];
            if ((m >> u) & 1u) {
                const int idx = team * U + u;
                if (head_pending && (tails & ((1u << idx) - 1u)) == 0u) {   // first end of a run in this range
                    have_head = true;
                    head_acc = acc;
                    head_c = acc_c;
                } else {
                    const uint32_t key = kk[u];
                    if (key < f_lo || key >= f_hi) {
                        fi = find_feature(sf, nf, key);
                        f_lo = sf[fi].row_base;
                        f_hi = f_lo + sf[fi].num_rows;
                    }
                    update_row_l1<EXTRA>(sf[fi], a, p0 + idx, fi, key - f_lo, t, acc, acc_c, fm, tmask, TG);
                }
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
                acc_c = 0.f;
            }
        }
        if (tails != 0u) head_pending = false;
        closed += (uint32_t)__popc(tails);
        k = k1; nk = nk1; slot = slot1;
        k1 = k_next; nk1 = nk_next; slot1 = slot_next;
    }
    uint32_t flag = 1u;
    if (end < S) {
        const uint32_t kl = a.keys[end - 1];
        if (kl != kInvalidKey && a.keys[end] == kl) {
            if (team == 0) {
                reinterpret_cast<float4 *>(a.tail_part + (size_t)w * a.row_floats)[t] = carry;
                if (EXTRA && t == 0) a.tail_part[(size_t)w * a.row_floats + D] = carry_c;
            }
            flag = 3u;
            __threadfence();
            __syncwarp();
        }
    }
    if (lane == 0) {
        st_relaxed_u32(a.range_flags + w, flag);
        if (a.num_unique != nullptr && closed != 0) atomicAdd(a.num_unique, (unsigned long long)closed);
    }
    const unsigned head_lanes = __ballot_sync(kFull, have_head);
    if (head_lanes != 0u) {
        const int src_team = (__ffs(head_lanes) - 1) / TG;
        const float4 hv = shfl_idx4<4>(head_acc, src_team * TG + t);
        const float hc = __shfl_sync(kFull, head_c, src_team * TG + t);
        uint32_t lo = 0, hi = w - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (a.keys[(mid + 1) * a.range - 1] + 1u >= first_key + 1u) hi = mid; else lo = mid + 1;
        }
        for (uint32_t j = lo + lane; j < w; j += kWarp)
            while (ld_relaxed_u32(a.range_flags + j) == 0u) { }
        __syncwarp();
        __threadfence();
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        float sum_c = 0.f;
        for (uint32_t j = lo + team; j < w; j += P) {
            add4(sum, __ldcg(reinterpret_cast<const float4 *>(a.tail_part + (size_t)j * a.row_floats) + t));
            if (EXTRA) sum_c += __ldcg(a.tail_part + (size_t)j * a.row_floats + D);
        }
#pragma unroll
        for (int off = TG; off < kWarp; off <<= 1) {
            float4 o;
            o.x = __shfl_xor_sync(kFull, sum.x, off); o.y = __shfl_xor_sync(kFull, sum.y, off);
            o.z = __shfl_xor_sync(kFull, sum.z, off); o.w = __shfl_xor_sync(kFull, sum.w, off);
            add4(sum, o);
            if (EXTRA) sum_c += __shfl_xor_sync(kFull, sum_c, off);
        }
        add4(sum, hv);
        sum_c += hc;
        if (team == 0) {
            const int hf = find_feature(sf, nf, first_key);
            update_row_l1<EXTRA>(sf[hf], a, first, hf, first_key - sf[hf].row_base, t, sum, sum_c, fm,
                                 TG == kWarp ? kFull : ((1u << TG) - 1u), TG);
        }
    }
}

// ---- hybrid placement, requester side: per-slot gradients of the SHARDED features, FM part folded in, packed densely ------
// packed[b, j * D .. (j + 1) * D) = gx[b, cols[j] .. + D) + c[b] * fm_sum[b, :]: the owner then pulls ONE 64-byte (D = 16) piece
// per slot over NVLink instead of the gradient slice plus the bag's fm_sum row.
struct PackCols {
    int32_t col[CTR_MAX_FEATURES];
};

template <int G>
__global__ void __launch_bounds__(256)
    fm_pack_grads_kernel(const float *__restrict__ gx, int64_t gx_stride, const float *__restrict__ c, const float *__restrict__ fm_sum,
                         int B, int J, const __grid_constant__ PackCols pc, float *__restrict__ packed) {
    constexpr int D = 4 * G;
    const int t = threadIdx.x % G;
    const long long total = (long long)B * J;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G; i < total; i += (long long)gridDim.x * blockDim.x / G) {
        const int b = (int)(i / J), j = (int)(i - (long long)b * J);
        float4 v = __ldg(reinterpret_cast<const float4 *>(gx + (size_t)b * gx_stride + pc.col[j]) + t);
        if (fm_sum != nullptr) {
            const float cb = __ldg(c + b);
            const float4 s4 = __ldg(reinterpret_cast<const float4 *>(fm_sum + (size_t)b * D) + t);
            v.x = fmaf(cb, s4.x, v.x); v.y = fmaf(cb, s4.y, v.y); v.z = fmaf(cb, s4.z, v.z); v.w = fmaf(cb, s4.w, v.w);
        }
        reinterpret_cast<float4 *>(packed + (size_t)i * D)[t] = v;
    }
}

// ---- replicated tables: dense update from an (all-reduced) gradient buffer -------------------------------------------------
// p / g / s hold n4 float4 each.  A float4 whose gradient is all zero is skipped (no read of p / s, no write): untouched rows
// do not move, exactly as the fused sparse update leaves them (sgd / adagrad with g = 0 are the identity anyway).  The
// gradient is cleared behind the read, so the buffer is zero again for the next step's CTR_OPT_GRAD_OUT sweep.
__global__ void __launch_bounds__(256)
    rows_dense_apply_kernel(float4 *__restrict__ p, float4 *__restrict__ g, float4 *__restrict__ s, long long n4, int kind,
                            ctr_hyper_t h, const ctr_hyper_t *__restrict__ hd, int clear) {
    if (hd != nullptr) h = *hd;
    constexpr int U = 4;                       // gradient loads in flight per thread
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += U * stride) {
        float4 gr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            gr[u] = i < n4 ? __ldcs(g + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (gr[u].x == 0.f && gr[u].y == 0.f && gr[u].z == 0.f && gr[u].w == 0.f) continue;
            float4 w = p[i];
            if (kind == CTR_OPT_SGD) {
                w.x -= h.lr * gr[u].x; w.y -= h.lr * gr[u].y; w.z -= h.lr * gr[u].z; w.w -= h.lr * gr[u].w;
            } else {
                float4 st = s[i];
                w.x = adagrad_elem(w.x, st.x, gr[u].x, h.lr, h.eps);
                w.y = adagrad_elem(w.y, st.y, gr[u].y, h.lr, h.eps);
                w.z = adagrad_elem(w.z, st.z, gr[u].z, h.lr, h.eps);
                w.w = adagrad_elem(w.w, st.w, gr[u].w, h.lr, h.eps);
                s[i] = st;
            }
            p[i] = w;
            if (clear) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
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

extern "C" int ctr_fm_pack_grads(const float *gx, int64_t gx_stride, const float *extra_grad, const float *fm_sum, int32_t B,
                                 int32_t D, const int32_t *cols, int32_t J, float *packed, void *stream_) {
    CTR_REQUIRE(gx != nullptr && packed != nullptr && cols != nullptr, "null pointer");
    CTR_REQUIRE(fm_sum == nullptr || extra_grad != nullptr, "fm_sum needs extra_grad");
    CTR_REQUIRE(B >= 0 && J >= 1 && J <= CTR_MAX_FEATURES && (D == 16 || D == 32 || D == 64), "B=%d J=%d D=%d unsupported", B, J, D);
    CTR_REQUIRE(gx_stride % 4 == 0 && ((reinterpret_cast<uintptr_t>(gx) | reinterpret_cast<uintptr_t>(packed) |
                                        reinterpret_cast<uintptr_t>(fm_sum)) & 15u) == 0, "operands must be 16-byte aligned");
    if (B == 0) return CTR_OK;
    PackCols pc{};
    for (int j = 0; j < J; ++j) {
        CTR_REQUIRE(cols[j] >= 0 && cols[j] % 4 == 0 && cols[j] + D <= gx_stride, "cols[%d]=%d outside the gradient row", j, cols[j]);
        pc.col[j] = cols[j];
    }
    const int G = D / 4;
    long long blocks = ((long long)B * J * G + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    cudaStream_t stream = (cudaStream_t)stream_;
    note_launch();
    if (G == 4) fm_pack_grads_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(gx, gx_stride, extra_grad, fm_sum, B, J, pc, packed);
    else if (G == 8) fm_pack_grads_kernel<8><<<(unsigned)blocks, 256, 0, stream>>>(gx, gx_stride, extra_grad, fm_sum, B, J, pc, packed);
    else fm_pack_grads_kernel<16><<<(unsigned)blocks, 256, 0, stream>>>(gx, gx_stride, extra_grad, fm_sum, B, J, pc, packed);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_rows_dense_apply(const ctr_opt_t *opt, float *params, float *grads, float *state0, int64_t n,
                                    int32_t clear_grads, void *stream_) {
    CTR_REQUIRE(opt != nullptr && params != nullptr && grads != nullptr, "null pointer");
    CTR_REQUIRE(opt->kind == CTR_OPT_SGD || opt->kind == CTR_OPT_ADAGRAD,
                "ctr_rows_dense_apply: sgd or element-wise adagrad (kind %d given)", opt->kind);
    CTR_REQUIRE(opt->kind == CTR_OPT_SGD || state0 != nullptr, "adagrad needs state0");
    CTR_REQUIRE(n >= 0 && n % 4 == 0, "n=%lld must be a multiple of 4", (long long)n);
    CTR_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(state0)) & 15u) == 0,
                "params / grads / state0 must be 16-byte aligned");
    if (n == 0) return CTR_OK;
    ctr_hyper_t h;
    ctr_opt_hyper(opt, &h);
    const long long n4 = n / 4;
    long long bx = (n4 + 1023) / 1024;         // 4 float4 per thread
    if (bx > kNumSMs * 8) bx = kNumSMs * 8;
    note_launch(), rows_dense_apply_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)stream_>>>(
        reinterpret_cast<float4 *>(params), reinterpret_cast<float4 *>(grads), reinterpret_cast<float4 *>(state0), n4, opt->kind, h,
        opt->device_hyper, clear_grads);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int64_t ctr_emb_bwd_workspace_bytes(const ctr_group_t *group) {
    static thread_local DevGroup dg;
    int rc = lower_group(group, &dg, /*need_tables=*/false, /*need_out=*/false);
    if (rc != CTR_OK) return rc;
    return plan_layout(dg).total;
}

extern "C" int ctr_emb_bwd_plan_ex(const ctr_group_t *group, void *workspace, int64_t workspace_bytes, uint32_t flags,
                                   void *stream_) {
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
    const bool want_runs = (flags & CTR_PLAN_NO_RUNS) == 0;
    note_launch(), reset_counters_kernel<<<1, 32, 0, stream>>>(counters, want_runs ? 1u : 0u);
    if (p.S > 0) {
        int64_t max_slots = 0;
        for (int i = 0; i < dg.num_features; ++i) {
            const int64_t n = (int64_t)dg.B * dg.f[i].L;
            if (n > max_slots) max_slots = n;
        }
        // ~4 blocks per SM over all features; every block folds its keys into the radix histograms
        int64_t bx = (max_slots + 1023) / 1024;  // at least 4 slots per thread
        const int64_t per_feature = ((int64_t)kNumSMs * 4 + dg.num_features - 1) / dg.num_features;
        if (bx > per_feature) bx = per_feature;
        if (bx < 1) bx = 1;
        uint32_t *scratch = reinterpret_cast<uint32_t *>(ws + p.counts);
        static thread_local SegSortDesc sd;
        seg_desc(dg, &sd);
        rc = seg_sort_prepare(scratch, sd, stream);
        if (rc != CTR_OK) return rc;
        note_launch(), emb_keygen_kernel<<<dim3((unsigned)bx, dg.num_features), 256, 0, stream>>>(dg, sd, keys_a, vals_a,
                                                                                                  seg_hist(scratch));
        CTR_CUDA_OK(cudaGetLastError());
        rc = seg_sort_pairs(sd, keys_a, vals_a, keys_b, vals_b, scratch, stream);
        if (rc < 0) return rc;
    }
    if (!want_runs) return CTR_OK;
    const uint32_t *sorted_keys = p.sorted_in_b ? keys_b : keys_a;
    return find_runs(sorted_keys, p.S, reinterpret_cast<uint32_t *>(ws + p.run_start), counters,
                     reinterpret_cast<uint32_t *>(ws + p.spine), stream, reinterpret_cast<uint32_t *>(ws + p.run_of_pos));
}

extern "C" int ctr_emb_bwd_plan(const ctr_group_t *group, void *workspace, int64_t workspace_bytes, void *stream) {
    return ctr_emb_bwd_plan_ex(group, workspace, workspace_bytes, 0u, stream);
}

// shared by the single-GPU apply and the sharded owner-side apply
static int apply_impl(DevGroup &dg, const PlanLayout &p, void *workspace, const ctr_opt_t *opt, int32_t *uniq_feature,
                      int32_t *uniq_row, float *row_grad, int64_t row_grad_stride, int64_t *num_unique,
                      const float *const *peer_grads, int world, cudaStream_t stream,
                      const float *const *peer_extra = nullptr, const float *const *peer_fm_sum = nullptr,
                      bool fm_row_only = false) {
    const bool updates = opt->kind != CTR_OPT_NONE;
    CTR_REQUIRE(workspace != nullptr, "workspace is null");
    CTR_REQUIRE(opt->kind >= CTR_OPT_NONE && opt->kind <= CTR_OPT_GRAD_OUT, "bad optimizer kind %d", opt->kind);
    CTR_REQUIRE(opt->kind != CTR_OPT_GRAD_OUT || (uniq_row == nullptr && peer_grads == nullptr),
                "CTR_OPT_GRAD_OUT: local (unsharded) apply without unique-row outputs");
    int team = 1;
    for (int i = 0; i < dg.num_features; ++i) {
        const DevFeature &f = dg.f[i];
        if (f.G > team) team = f.G;
        if (opt->kind == CTR_OPT_ADAGRAD || opt->kind == CTR_OPT_ROWWISE_ADAGRAD || opt->kind == CTR_OPT_ADAM ||
            opt->kind == CTR_OPT_GRAD_OUT)
            CTR_REQUIRE(f.state0 != nullptr, "feature %d: optimizer state0 is null", i);
        if (opt->kind == CTR_OPT_ADAM) CTR_REQUIRE(f.state1 != nullptr, "feature %d: optimizer state1 is null", i);
        if (f.vec == 4 && updates) {
            CTR_REQUIRE(f.state0 == nullptr || opt->kind == CTR_OPT_ROWWISE_ADAGRAD || opt->kind == CTR_OPT_SGD ||
                            (reinterpret_cast<uintptr_t>(f.state0) & 15u) == 0,
                        "feature %d: state0 not 16-byte aligned", i);
            CTR_REQUIRE(f.state1 == nullptr || (reinterpret_cast<uintptr_t>(f.state1) & 15u) == 0,
                        "feature %d: state1 not 16-byte aligned", i);
        }
    }
    CTR_REQUIRE(row_grad == nullptr || uniq_row != nullptr, "row_grad needs uniq_row");
    char *ws = static_cast<char *>(workspace);
    ApplyArgs a{};
    a.keys = reinterpret_cast<const uint32_t *>(ws + (p.sorted_in_b ? p.keys_b : p.keys_a));
    a.vals = reinterpret_cast<const uint32_t *>(ws + (p.sorted_in_b ? p.vals_b : p.vals_a));
    a.run_start = reinterpret_cast<const uint32_t *>(ws + p.run_start);
    a.counters = reinterpret_cast<uint32_t *>(ws + p.counters);
    a.ticket = reinterpret_cast<uint32_t *>(ws + p.range_flags);
    a.range_flags = a.ticket + 64;
    a.tail_part = reinterpret_cast<float *>(ws + p.tail_part);
    a.row_floats = (int)p.row_floats;
    a.uniq_feature = uniq_feature;
    a.uniq_row = uniq_row;
    a.row_grad = row_grad;
    a.row_grad_stride = row_grad_stride;
    a.num_unique = reinterpret_cast<unsigned long long *>(num_unique);
    a.status = dg.status;
    a.kind = opt->kind;
    ctr_opt_hyper(opt, &a.h);
    a.hyper_dev = opt->device_hyper;
    a.team = team;
    if (opt->kind == CTR_OPT_ADAM) CTR_REQUIRE(opt->step >= 1, "Adam step must be >= 1");
    if (p.S == 0) {
        if (num_unique != nullptr) CTR_CUDA_OK(cudaMemsetAsync(num_unique, 0, sizeof(int64_t), stream));
        return CTR_OK;
    }
    int64_t expected = p.S;
    if (peer_grads != nullptr) {
        a.p2p = 1;
        a.S_dev = a.counters + 3;
        for (int r = 0; r < world; ++r) {
            CTR_REQUIRE(peer_grads[r] != nullptr, "peer_grads[%d] is null", r);
            a.peer_grads[r] = peer_grads[r];
        }
        expected = p.S / world;            // the capacity is world x that; ranges are sized for the balanced case
    }
    // range length: a multiple of the positions one outer iteration covers, sized for ~3 waves of resident warps
    const int P = kWarp / team;
    const int U = P >= 8 ? (kWarp / P < 4 ? kWarp / P : 4) : 4;
    const int Q = P * U;
    const int64_t target_ranges = (int64_t)kNumSMs * 24 * 2;   // ~2 waves of resident warps: longer ranges = fewer boundary fences
    int64_t range = (expected + target_ranges - 1) / target_ranges;
    range = (range + Q - 1) / Q * Q;
    const int64_t lo = kMinRange > Q ? kMinRange : Q;
    if (range < lo) range = lo;
    if (range > 512) range = 512;
    static const int range_env = getenv("CTR_SWEEP_RANGE") ? atoi(getenv("CTR_SWEEP_RANGE")) : 0;   // tuning knob
    static const int prefetch_env = getenv("CTR_SWEEP_PREFETCH") ? atoi(getenv("CTR_SWEEP_PREFETCH")) : 1;   // 164 vs 170 us (cfg2)
    a.gather_prefetch = prefetch_env;
    if (range_env > 0) range = (range_env + Q - 1) / Q * Q;
    a.S = (uint32_t)p.S;
    a.range = (uint32_t)range;
    a.num_ranges = (uint32_t)((p.S + range - 1) / range);
    bool any_vec4 = false, l1 = dg.num_features > 0;
    for (int i = 0; i < dg.num_features; ++i) {
        const DevFeature &f = dg.f[i];
        any_vec4 |= f.vec == 4;
        l1 &= f.vec == 4 && f.aligned && f.L == 1 && f.id_weight == nullptr && f.pooling == CTR_POOL_SUM && f.D == dg.f[0].D &&
              (f.D == 16 || f.D == 32 || f.D == 64);
    }
    const bool extra = dg.extra != nullptr;
    if (extra) {
        if (!l1 || (peer_grads != nullptr && peer_extra == nullptr) || !updates || uniq_row != nullptr) {
            set_error("group->extra (twin tables / FM term) needs a single-id group of one width (D = 16, 32 or 64), sum pooling, "
                      "a fused optimizer and no unique-row outputs (sharded owner side: ctr_emb_bwd_apply_p2p_ex)");
            return CTR_E_UNSUPPORTED;
        }
        if (peer_grads != nullptr) {       // owner side: every slot reads the scalars of the rank that sent it
            for (int r = 0; r < world; ++r) {
                CTR_REQUIRE(peer_extra[r] != nullptr, "peer_extra[%d] is null", r);
                a.peer_extra[r] = peer_extra[r];
                if (peer_fm_sum != nullptr) {
                    CTR_REQUIRE(peer_fm_sum[r] != nullptr && (reinterpret_cast<uintptr_t>(peer_fm_sum[r]) & 15u) == 0,
                                "peer_fm_sum[%d] is null or not 16-byte aligned", r);
                    a.peer_fm_sum[r] = peer_fm_sum[r];
                }
            }
            a.extra_grad = peer_extra[0];
            a.fm_sum = peer_fm_sum != nullptr ? peer_fm_sum[0] : nullptr;
            a.fm_row_only = (fm_row_only && peer_fm_sum == nullptr) ? 1 : 0;
        } else {
            a.extra_grad = dg.extra;
            a.fm_sum = dg.fm ? dg.fm_sum : nullptr;
            if (dg.fm) CTR_REQUIRE((reinterpret_cast<uintptr_t>(dg.fm_sum) & 15u) == 0, "fm_sum must be 16-byte aligned");
        }
        for (int i = 0; i < dg.num_features; ++i) {
            const DevFeature &f = dg.f[i];
            if (f.twin_table == nullptr) continue;
            if (opt->kind != CTR_OPT_SGD) CTR_REQUIRE(f.twin_state0 != nullptr, "feature %d: twin_state0 is null", i);
            if (opt->kind == CTR_OPT_ADAM) CTR_REQUIRE(f.twin_state1 != nullptr, "feature %d: twin_state1 is null", i);
        }
    }
    if (num_unique != nullptr) CTR_CUDA_OK(cudaMemsetAsync(num_unique, 0, sizeof(int64_t), stream));
    CTR_CUDA_OK(cudaMemsetAsync(a.ticket, 0, (64 + (size_t)a.num_ranges) * sizeof(uint32_t), stream));
    const unsigned blocks = (a.num_ranges + kApplyWarps - 1) / kApplyWarps;
    note_launch();
    if (l1) {
        const int tg = dg.f[0].D / 4;
        // resident blocks per SM the kernel is compiled for: 3 (80 registers, a few spilled words) or 2 (no spills)
        static const int occ = getenv("CTR_SWEEP_OCC") ? atoi(getenv("CTR_SWEEP_OCC")) : 3;
#define CTR_L1_SWEEP(TG_, EX_)                                                                                    \
    do {                                                                                                          \
        if (occ == 2) emb_bwd_sweep_l1_kernel<TG_, EX_, 2><<<blocks, kApplyThreads, 0, stream>>>(dg, a);          \
        else emb_bwd_sweep_l1_kernel<TG_, EX_, 3><<<blocks, kApplyThreads, 0, stream>>>(dg, a);                   \
    } while (0)
        if (tg == 4 && extra) CTR_L1_SWEEP(4, true);
        else if (tg == 4) CTR_L1_SWEEP(4, false);
        else if (tg == 8 && extra) CTR_L1_SWEEP(8, true);
        else if (tg == 8) CTR_L1_SWEEP(8, false);
        else if (extra) CTR_L1_SWEEP(16, true);
        else CTR_L1_SWEEP(16, false);
#undef CTR_L1_SWEEP
    }
    else if (any_vec4) emb_bwd_sweep_kernel<4><<<blocks, kApplyThreads, 0, stream>>>(dg, a);
    else emb_bwd_sweep_kernel<1><<<blocks, kApplyThreads, 0, stream>>>(dg, a);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_emb_bwd_apply(const ctr_group_t *group, void *workspace, const ctr_opt_t *opt,
                                 int32_t *uniq_feature, int32_t *uniq_row, float *row_grad, int64_t row_grad_stride,
                                 int64_t *num_unique, void *stream_) {
    static thread_local DevGroup dg;
    CTR_REQUIRE(opt != nullptr, "opt is null");
    int rc = lower_group(group, &dg, /*need_tables=*/opt->kind != CTR_OPT_NONE, /*need_out=*/true);
    if (rc != CTR_OK) return rc;
    return apply_impl(dg, plan_layout(dg), workspace, opt, uniq_feature, uniq_row, row_grad, row_grad_stride, num_unique,
                      nullptr, 1, (cudaStream_t)stream_);
}

// ---- sharded owner side: pull the routed (key, slot) lists from the peers, sort, apply -----------------------
namespace ctr {

struct PeerLists {
    int world, rank;
    const uint32_t *counts[CTR_MAX_WORLD];
    const uint32_t *keys[CTR_MAX_WORLD];
    const uint32_t *slots[CTR_MAX_WORLD];
};

// one warp: where each peer's list for this rank starts (in the peer's buffer and in the gathered array)
__global__ void p2p_offsets_kernel(const __grid_constant__ PeerLists pl, uint32_t *counters) {
    const int r = threadIdx.x;
    uint32_t src = 0, cnt = 0;
    if (r < pl.world) {
        for (int o = 0; o < pl.rank; ++o) src += pl.counts[r][o];
        cnt = pl.counts[r][pl.rank];
    }
    uint32_t incl = cnt;
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, incl, off);
        if (r >= off) incl += v;
    }
    if (r < pl.world) {
        counters[kMetaSrc + r] = src;
        counters[kMetaDst + r] = incl - cnt;
        counters[kMetaCnt + r] = cnt;
    }
    if (r == 31) counters[3] = incl;
}

// grid (bx, world): copies peer r's list for this rank into the sort input, tags the slots with r and builds the
// radix histograms of the keys
__global__ void __launch_bounds__(256)
    p2p_gather_kernel(const __grid_constant__ PeerLists pl, const uint32_t *__restrict__ counters, uint32_t *__restrict__ keys,
                      uint32_t *__restrict__ vals, uint32_t *__restrict__ hist, int passes, int bits) {
    __shared__ uint32_t sh[kMaxPasses][kMaxRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kMaxRadix; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int r = blockIdx.y;
    const uint32_t n = counters[kMetaCnt + r];
    const uint32_t *src_k = pl.keys[r] + counters[kMetaSrc + r];
    const uint32_t *src_s = pl.slots[r] + counters[kMetaSrc + r];
    const uint32_t dst = counters[kMetaDst + r];
    const int lane = threadIdx.x & 31;
    const uint32_t dmask = (1u << bits) - 1u;
    const uint32_t nround = (n + 31u) / 32u * 32u;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += gridDim.x * blockDim.x) {
        const bool valid = j < n;
        uint32_t key = 0;
        if (valid) {
            key = src_k[j];
            keys[dst + j] = key;
            vals[dst + j] = src_s[j] | ((uint32_t)r << 28);
        }
        for (int p = 0; p < passes; ++p) {
            const uint32_t d = (key >> (p * bits)) & dmask;
            const uint32_t peers = __match_any_sync(kFull, valid ? d : (uint32_t)(kMaxRadix + lane));
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&sh[p][d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    const int radix = 1 << bits;
    for (int i = threadIdx.x; i < passes * radix; i += blockDim.x) {
        const uint32_t c = sh[i / radix][i % radix];
        if (c) atomicAdd(&hist[(i / radix) * kMaxRadix + (i % radix)], c);
    }
}

static int check_owner_group(const DevGroup &dg, const ctr_shard_t *shard) {
    CTR_REQUIRE(shard != nullptr, "shard is null");
    CTR_REQUIRE(shard->world >= 1 && shard->world <= CTR_MAX_WORLD, "world=%d outside [1, %d]", shard->world, CTR_MAX_WORLD);
    CTR_REQUIRE(shard->rank >= 0 && shard->rank < shard->world, "rank=%d outside [0, %d)", shard->rank, shard->world);
    for (int i = 0; i < dg.num_features; ++i)
        CTR_REQUIRE((int64_t)dg.B * dg.f[i].L < (1ll << 28), "feature %d: B*L must stay below 2^28 on the sharded path", i);
    return CTR_OK;
}

}  // namespace ctr

extern "C" int64_t ctr_emb_bwd_p2p_workspace_bytes(const ctr_group_t *group, int32_t world) {
    static thread_local DevGroup dg;
    int rc = lower_group(group, &dg, false, false);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(world >= 1 && world <= CTR_MAX_WORLD, "world=%d outside [1, %d]", world, CTR_MAX_WORLD);
    return plan_layout(dg, world).total;
}

extern "C" int ctr_emb_bwd_plan_p2p(const ctr_group_t *group, const ctr_shard_t *shard, const uint32_t *const *peer_counts,
                                    const uint32_t *const *peer_keys, const uint32_t *const *peer_slots, void *workspace,
                                    int64_t workspace_bytes, void *stream_) {
    static thread_local DevGroup dg;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = lower_group(group, &dg, false, false);
    if (rc != CTR_OK) return rc;
    rc = check_owner_group(dg, shard);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(peer_counts != nullptr && peer_keys != nullptr && peer_slots != nullptr, "null peer list arrays");
    const PlanLayout p = plan_layout(dg, shard->world);
    CTR_REQUIRE(p.S < (1ll << 30), "capacity of %lld pairs; must stay below 2^30", (long long)p.S);
    CTR_REQUIRE(workspace != nullptr, "workspace is null");
    if (workspace_bytes < p.total) {
        set_error("workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)p.total);
        return CTR_E_WORKSPACE;
    }
    CTR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
    PeerLists pl{};
    pl.world = shard->world;
    pl.rank = shard->rank;
    for (int r = 0; r < shard->world; ++r) {
        CTR_REQUIRE(peer_counts[r] != nullptr && peer_keys[r] != nullptr && peer_slots[r] != nullptr, "peer %d: null list", r);
        pl.counts[r] = peer_counts[r];
        pl.keys[r] = peer_keys[r];
        pl.slots[r] = peer_slots[r];
    }
    char *ws = static_cast<char *>(workspace);
    uint32_t *counters = reinterpret_cast<uint32_t *>(ws + p.counters);
    uint32_t *keys_a = reinterpret_cast<uint32_t *>(ws + p.keys_a), *keys_b = reinterpret_cast<uint32_t *>(ws + p.keys_b);
    uint32_t *vals_a = reinterpret_cast<uint32_t *>(ws + p.vals_a), *vals_b = reinterpret_cast<uint32_t *>(ws + p.vals_b);
    note_launch(), reset_counters_kernel<<<1, 32, 0, stream>>>(counters, 0u);
    if (p.S == 0) return CTR_OK;
    note_launch(), p2p_offsets_kernel<<<1, 32, 0, stream>>>(pl, counters);
    uint32_t *scratch = reinterpret_cast<uint32_t *>(ws + p.counts);
    rc = radix_sort_prepare(scratch, p.S, p.key_bits, stream);
    if (rc != CTR_OK) return rc;
    int64_t bx = (p.S / shard->world / shard->world + 2047) / 2048;   // a peer's list for this rank, balanced case
    const int64_t per_peer = ((int64_t)kNumSMs * 4 + shard->world - 1) / shard->world;
    if (bx > per_peer) bx = per_peer;
    if (bx < 1) bx = 1;
    note_launch(), p2p_gather_kernel<<<dim3((unsigned)bx, shard->world), 256, 0, stream>>>(
        pl, counters, keys_a, vals_a, sort_hist(scratch), sort_num_passes(p.key_bits), sort_digit_bits(p.key_bits));
    CTR_CUDA_OK(cudaGetLastError());
    rc = radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, p.S, p.key_bits, scratch, reinterpret_cast<uint32_t *>(ws + p.spine),
                          stream, /*hist_ready=*/true, /*n_dev=*/counters + 3);
    if (rc < 0) return rc;
    if (rc != p.sorted_in_b) {
        set_error("internal: sort parity mismatch");
        return CTR_E_CUDA;
    }
    return CTR_OK;
}

extern "C" int ctr_emb_bwd_apply_p2p(const ctr_group_t *group, const ctr_shard_t *shard, void *workspace, const ctr_opt_t *opt,
                                     const float *const *peer_grads, int64_t *num_unique, void *stream_) {
    static thread_local DevGroup dg;
    CTR_REQUIRE(opt != nullptr && peer_grads != nullptr, "null pointer");
    int rc = lower_group(group, &dg, /*need_tables=*/opt->kind != CTR_OPT_NONE, /*need_out=*/false);
    if (rc != CTR_OK) return rc;
    rc = check_owner_group(dg, shard);
    if (rc != CTR_OK) return rc;
    // `aligned` was derived from group->out, which is not used here: look at the peers' matrices instead
    for (int i = 0; i < dg.num_features; ++i) {
        DevFeature &f = dg.f[i];
        bool al = f.vec == 4 && f.out_col % 4 == 0 && dg.out_stride % 4 == 0;
        for (int r = 0; r < shard->world && al; ++r) al = (reinterpret_cast<uintptr_t>(peer_grads[r]) & 15u) == 0;
        f.aligned = al ? 1 : 0;
    }
    return apply_impl(dg, plan_layout(dg, shard->world), workspace, opt, nullptr, nullptr, nullptr, 0, num_unique, peer_grads,
                      shard->world, (cudaStream_t)stream_);
}

extern "C" int ctr_emb_bwd_apply_p2p_ex(const ctr_group_t *group, const ctr_shard_t *shard, void *workspace, const ctr_opt_t *opt,
                                        const float *const *peer_grads, const float *const *peer_extra,
                                        const float *const *peer_fm_sum, int64_t *num_unique, void *stream_) {
    static thread_local DevGroup dg;
    CTR_REQUIRE(opt != nullptr && peer_grads != nullptr && peer_extra != nullptr, "null pointer");
    CTR_REQUIRE(group != nullptr && group->extra != nullptr, "group->extra must be set (marks the fused terms; it is not read)");
    ctr_group_t gcopy = *group;
    const bool fm_row_only = group->fm != 0 && peer_fm_sum == nullptr;   // gradients pre-combined by ctr_fm_pack_grads
    gcopy.fm = 0;                 // the FM sums come from the peers (peer_fm_sum), not from group->fm_sum
    gcopy.fm_sum = nullptr;
    int rc = lower_group(&gcopy, &dg, /*need_tables=*/opt->kind != CTR_OPT_NONE, /*need_out=*/false);
    if (rc != CTR_OK) return rc;
    rc = check_owner_group(dg, shard);
    if (rc != CTR_OK) return rc;
    for (int i = 0; i < dg.num_features; ++i) {
        DevFeature &f = dg.f[i];
        bool al = f.vec == 4 && f.out_col % 4 == 0 && dg.out_stride % 4 == 0;
        for (int r = 0; r < shard->world && al; ++r) al = (reinterpret_cast<uintptr_t>(peer_grads[r]) & 15u) == 0;
        f.aligned = al ? 1 : 0;
    }
    return apply_impl(dg, plan_layout(dg, shard->world), workspace, opt, nullptr, nullptr, nullptr, 0, num_unique, peer_grads,
                      shard->world, (cudaStream_t)stream_, peer_extra, peer_fm_sum, fm_row_only);
}

// ---- de-duplicated exchange: every rank fetches / sends each distinct row of its batch once ------------------------
// Zipf ids repeat: 1.7 M slots of a Criteo batch touch 0.28 M distinct rows, and peer reads are not cached in L2, so
// the plain sharded lookup pays NVLink for every repeat.  Here the requester sorts its slots first (the same plan the
// single-GPU backward uses), fetches every distinct row ONCE into a local staging matrix and pools from there;
// backward, it reduces its own duplicates locally (the sweep in NONE mode) and the owners pull one gradient row per
// (rank, distinct row) instead of one per slot.
namespace ctr {

struct FetchArgs {
    const uint32_t *keys, *vals, *run_start, *run_of_pos, *counters;
    float *staging;       // [S, D]
    long long *uidx;      // [S] or null: slot (feature-major, bag * L + l) -> run
    uint32_t S;
};

// one team of G lanes per run: the row of that run, read from its owner's shard; four runs in flight per team (peer
// loads take ~2 us, the only way to fill NVLink is to have many of them outstanding)
__global__ void __launch_bounds__(256) uniq_fetch_kernel(const __grid_constant__ DevGroup g, const FetchArgs a) {
    const DevFeature &f0 = g.f[0];
    const int G = f0.G, vec = f0.vec, D = f0.D;
    const int t = threadIdx.x % G;
    const uint32_t teams = gridDim.x * (blockDim.x / G);
    const uint32_t U = a.counters[1];                      // runs with a valid key
    if (t * vec >= D) return;
    for (uint32_t r0 = (blockIdx.x * (blockDim.x / G) + threadIdx.x / G) * 4u; r0 < U; r0 += teams * 4u) {
        const float *src[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            src[j] = nullptr;
            if (r0 + j < U) {
                const uint32_t key = a.keys[a.run_start[r0 + j]];
                int lo = 0, hi = g.num_features - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (g.f[mid].row_base <= key) lo = mid; else hi = mid - 1;
                }
                src[j] = table_row(g, g.f[lo], lo, (int32_t)(key - g.f[lo].row_base));
            }
        }
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src[j] != nullptr) {
                if (vec == 4) v[j] = __ldg(reinterpret_cast<const float4 *>(src[j]) + t);
                else v[j].x = __ldg(src[j] + t);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (src[j] != nullptr) {
                float *dst = a.staging + (size_t)(r0 + j) * D;
                if (vec == 4) reinterpret_cast<float4 *>(dst)[t] = v[j];
                else dst[t] = v[j].x;
            }
        }
    }
}

// slot -> run, for every valid sorted position (uidx was pre-filled with -1 = padding)
__global__ void __launch_bounds__(256) slot_uidx_kernel(const __grid_constant__ DevGroup g, const FetchArgs a) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < a.S; p += gridDim.x * blockDim.x) {
        const uint32_t key = a.keys[p];
        if (key == kInvalidKey) continue;
        int lo = 0, hi = g.num_features - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (g.f[mid].row_base <= key) lo = mid; else hi = mid - 1;
        }
        long long slot_base = 0;
        for (int i = 0; i < lo; ++i) slot_base += (long long)g.B * g.f[i].L;
        a.uidx[slot_base + a.vals[p]] = (long long)a.run_of_pos[p];
    }
}

struct PeerUniq {
    int world, rank, num_features;
    const long long *num_unique[CTR_MAX_WORLD];   // [1] on every rank
    const int32_t *uniq_feature[CTR_MAX_WORLD];   // [U_r]
    const int32_t *uniq_row[CTR_MAX_WORLD];       // [U_r]
    const int64_t *adj;
};

__global__ void p2p_uniq_offsets_kernel(const __grid_constant__ PeerUniq pu, uint32_t *counters) {
    const int r = threadIdx.x;
    uint32_t cnt = r < pu.world ? (uint32_t)*pu.num_unique[r] : 0u;
    uint32_t incl = cnt;
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, incl, off);
        if (r >= off) incl += v;
    }
    if (r < pu.world) {
        counters[kMetaSrc + r] = 0;
        counters[kMetaDst + r] = incl - cnt;
        counters[kMetaCnt + r] = cnt;
    }
    if (r == 31) counters[3] = incl;
}

// grid (bx, world): every distinct (table, row) of peer r becomes a pair; the ones this rank does not own get the
// invalid key (they sort behind everything and the sweep skips them), so the input order -- and with it the order
// of every sum -- is fixed by (rank, index).  (An ordered compaction of the owned pairs was tried: its two extra
// passes over the peers' lists cost more than sorting the invalid keys along, 232 vs 154 us at 8 GPUs.)
__global__ void __launch_bounds__(256)
    p2p_uniq_gather_kernel(const __grid_constant__ PeerUniq pu, const uint32_t *__restrict__ counters, uint32_t *__restrict__ keys,
                           uint32_t *__restrict__ vals, uint32_t *__restrict__ hist, int passes, int bits) {
    __shared__ uint32_t sh[kMaxPasses][kMaxRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kMaxRadix; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int r = blockIdx.y;
    const uint32_t n = counters[kMetaCnt + r];
    const uint32_t dst = counters[kMetaDst + r];
    const int lane = threadIdx.x & 31;
    const uint32_t dmask = (1u << bits) - 1u;
    const uint32_t nround = (n + 31u) / 32u * 32u;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += gridDim.x * blockDim.x) {
        const bool valid = j < n;
        uint32_t key = kInvalidKey;
        if (valid) {
            const uint32_t fi = (uint32_t)pu.uniq_feature[r][j];
            const uint32_t rr = (uint32_t)pu.uniq_row[r][j] + fi;
            if ((int)(rr % (uint32_t)pu.world) == pu.rank)
                key = (uint32_t)(__ldg(pu.adj + (size_t)pu.rank * pu.num_features + fi) + (long long)(rr / (uint32_t)pu.world));
            keys[dst + j] = key;
            vals[dst + j] = j | ((uint32_t)r << 28);
        }
        for (int p = 0; p < passes; ++p) {
            const uint32_t d = (key >> (p * bits)) & dmask;
            const uint32_t peers = __match_any_sync(kFull, valid ? d : (uint32_t)(kMaxRadix + lane));
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&sh[p][d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    const int radix = 1 << bits;
    for (int i = threadIdx.x; i < passes * radix; i += blockDim.x) {
        const uint32_t c = sh[i / radix][i % radix];
        if (c) atomicAdd(&hist[(i / radix) * kMaxRadix + (i % radix)], c);
    }
}

}  // namespace ctr

extern "C" int ctr_unique_fetch(const ctr_group_t *group, const ctr_shard_t *shard, const float *const *tables,
                                void *plan_workspace, float *staging, int64_t *uidx, void *stream_) {
    static thread_local DevGroup dg;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = lower_group(group, &dg, false, false);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(tables != nullptr && plan_workspace != nullptr && staging != nullptr, "null pointer");
    rc = attach_shard(&dg, shard, tables);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE((reinterpret_cast<uintptr_t>(staging) & 15u) == 0, "staging must be 16-byte aligned");
    const PlanLayout p = plan_layout(dg);
    if (p.S == 0) return CTR_OK;
    char *ws = static_cast<char *>(plan_workspace);
    FetchArgs a{};
    a.keys = reinterpret_cast<const uint32_t *>(ws + (p.sorted_in_b ? p.keys_b : p.keys_a));
    a.vals = reinterpret_cast<const uint32_t *>(ws + (p.sorted_in_b ? p.vals_b : p.vals_a));
    a.run_start = reinterpret_cast<const uint32_t *>(ws + p.run_start);
    a.run_of_pos = reinterpret_cast<const uint32_t *>(ws + p.run_of_pos);
    a.counters = reinterpret_cast<const uint32_t *>(ws + p.counters);
    a.staging = staging;
    a.uidx = reinterpret_cast<long long *>(uidx);
    a.S = (uint32_t)p.S;
    const int teams = 256 / dg.f[0].G;
    int64_t blocks = (p.S / 4 + teams - 1) / teams;        // ~ a quarter of the slots are distinct rows; the loop strides
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    if (blocks < 1) blocks = 1;
    note_launch(), uniq_fetch_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dg, a);
    if (uidx != nullptr) {
        CTR_CUDA_OK(cudaMemsetAsync(uidx, 0xff, (size_t)p.S * sizeof(int64_t), stream));
        int64_t b2 = (p.S + 1023) / 1024;
        if (b2 > kNumSMs * 8) b2 = kNumSMs * 8;
        note_launch(), slot_uidx_kernel<<<(unsigned)b2, 256, 0, stream>>>(dg, a);
    }
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int64_t ctr_emb_bwd_p2p_unique_workspace_bytes(const ctr_group_t *group, int64_t capacity) {
    static thread_local DevGroup dg;
    int rc = lower_group(group, &dg, false, false);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(capacity >= 1, "capacity must be positive");
    return plan_layout(dg, 1, capacity).total;
}

extern "C" int ctr_emb_bwd_plan_p2p_unique(const ctr_group_t *group, const ctr_shard_t *shard, const int64_t *const *peer_num_unique,
                                           const int32_t *const *peer_uniq_feature, const int32_t *const *peer_uniq_row,
                                           int64_t capacity, void *workspace, int64_t workspace_bytes, void *stream_) {
    static thread_local DevGroup dg;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = lower_group(group, &dg, false, false);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(shard != nullptr && shard->adj != nullptr, "shard is null");
    CTR_REQUIRE(shard->world >= 1 && shard->world <= CTR_MAX_WORLD && shard->rank >= 0 && shard->rank < shard->world, "bad world / rank");
    CTR_REQUIRE(peer_num_unique && peer_uniq_feature && peer_uniq_row && workspace, "null pointer");
    CTR_REQUIRE(capacity >= 1 && capacity < (1ll << 30) && capacity / shard->world < (1ll << 28), "capacity %lld out of range", (long long)capacity);
    const PlanLayout p = plan_layout(dg, 1, capacity);
    if (workspace_bytes < p.total) {
        set_error("workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)p.total);
        return CTR_E_WORKSPACE;
    }
    CTR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
    PeerUniq pu{};
    pu.world = shard->world; pu.rank = shard->rank; pu.num_features = dg.num_features; pu.adj = shard->adj;
    for (int r = 0; r < shard->world; ++r) {
        CTR_REQUIRE(peer_num_unique[r] && peer_uniq_feature[r] && peer_uniq_row[r], "peer %d: null list", r);
        pu.num_unique[r] = reinterpret_cast<const long long *>(peer_num_unique[r]);
        pu.uniq_feature[r] = peer_uniq_feature[r];
        pu.uniq_row[r] = peer_uniq_row[r];
    }
    char *ws = static_cast<char *>(workspace);
    uint32_t *counters = reinterpret_cast<uint32_t *>(ws + p.counters);
    uint32_t *keys_a = reinterpret_cast<uint32_t *>(ws + p.keys_a), *keys_b = reinterpret_cast<uint32_t *>(ws + p.keys_b);
    uint32_t *vals_a = reinterpret_cast<uint32_t *>(ws + p.vals_a), *vals_b = reinterpret_cast<uint32_t *>(ws + p.vals_b);
    note_launch(), reset_counters_kernel<<<1, 32, 0, stream>>>(counters, 0u);
    note_launch(), p2p_uniq_offsets_kernel<<<1, 32, 0, stream>>>(pu, counters);
    uint32_t *scratch = reinterpret_cast<uint32_t *>(ws + p.counts);
    rc = radix_sort_prepare(scratch, p.S, p.key_bits, stream);
    if (rc != CTR_OK) return rc;
    int64_t bx = (capacity / shard->world / 4 + 2047) / 2048;
    const int64_t per_peer = ((int64_t)kNumSMs * 4 + shard->world - 1) / shard->world;
    if (bx > per_peer) bx = per_peer;
    if (bx < 1) bx = 1;
    note_launch(), p2p_uniq_gather_kernel<<<dim3((unsigned)bx, shard->world), 256, 0, stream>>>(
        pu, counters, keys_a, vals_a, sort_hist(scratch), sort_num_passes(p.key_bits), sort_digit_bits(p.key_bits));
    CTR_CUDA_OK(cudaGetLastError());
    rc = radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, p.S, p.key_bits, scratch, reinterpret_cast<uint32_t *>(ws + p.spine),
                          stream, /*hist_ready=*/true, /*n_dev=*/counters + 3);
    if (rc < 0) return rc;
    if (rc != p.sorted_in_b) {
        set_error("internal: sort parity mismatch");
        return CTR_E_CUDA;
    }
    return CTR_OK;
}

extern "C" int ctr_emb_bwd_apply_p2p_unique(const ctr_group_t *group, const ctr_shard_t *shard, void *workspace, const ctr_opt_t *opt,
                                            const float *const *peer_row_grads, int64_t capacity, int64_t *num_unique, void *stream_) {
    static thread_local DevGroup dg;
    CTR_REQUIRE(opt != nullptr && peer_row_grads != nullptr && shard != nullptr, "null pointer");
    int rc = lower_group(group, &dg, /*need_tables=*/opt->kind != CTR_OPT_NONE, /*need_out=*/false);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(shard->world >= 1 && shard->world <= CTR_MAX_WORLD, "bad world");
    // the "gradient matrix" of a peer is its [distinct rows, D] buffer: row index = the slot, no column offset
    for (int i = 0; i < dg.num_features; ++i) {
        DevFeature &f = dg.f[i];
        CTR_REQUIRE(f.L == 1 && f.out_col == 0, "feature %d: the de-duplicated owner group wants L = 1 and out_col = 0", i);
        bool al = f.vec == 4 && dg.out_stride % 4 == 0;
        for (int r = 0; r < shard->world && al; ++r) al = (reinterpret_cast<uintptr_t>(peer_row_grads[r]) & 15u) == 0;
        f.aligned = al ? 1 : 0;
    }
    return apply_impl(dg, plan_layout(dg, 1, capacity), workspace, opt, nullptr, nullptr, nullptr, 0, num_unique, peer_row_grads,
                      shard->world, (cudaStream_t)stream_);
}
