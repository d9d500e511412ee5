// K1: fused index-map (direct | murmur3 hash | vocabulary remap) + row gather + segment pool.
//
// Replaces the per-feature loop of torchctr/models/dnn.py:53-59 (mask, nn.Embedding gather,
// mask-multiply, sum over the sequence axis) and the concat of dnn.py:61-67: every feature of
// a launch group is pooled straight into its column range of the tower-input matrix, the
// dense block is copied beside it, and no [B, L, D] intermediate exists.
//
// Lane layout.  A row of D floats is covered by G lanes (float4 each when D % 4 == 0, one
// float each otherwise).  A bag is owned by a team of T = clamp(pow2(L), G, 32) lanes =
// R = T / G row slots.  Per chunk of T id slots every lane loads and maps ONE id (coalesced
// 8-byte loads, one murmur3 / map probe per id), the mapped rows are exchanged by warp
// shuffle, and each row slot issues its G row loads back to back (G independent 16-byte
// loads in flight per lane).  Row slots are summed with xor-shuffles.  For single-id bags
// (L == 1) a team takes K = min(G, 4) consecutive bags per trip instead, to keep the same
// number of loads in flight.
#include "common.cuh"

namespace ctr {

constexpr int kFwdThreads = 256;
constexpr int kFwdWarps = kFwdThreads / kWarp;

__device__ __forceinline__ void flag_status(uint32_t *status, uint32_t bit) {
    if (status != nullptr) atomicOr(status, bit);
}

__device__ __forceinline__ float4 load_row_part(const DevGroup &g, const DevFeature &f, int fi, int32_t row, int g_lane) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *src = table_row(g, f, fi, row);
    if (f.vec == 4) {
        v = __ldg(reinterpret_cast<const float4 *>(src) + g_lane);
    } else {
        v.x = __ldg(src + g_lane);
    }
    return v;
}

__device__ __forceinline__ void store_out_part(const DevGroup &g, const DevFeature &f, int64_t bag, int g_lane,
                                               float4 v) {
    float *dst = g.out + bag * g.out_stride + f.out_col;
    if (f.vec == 4) {
        if (f.aligned) {
            reinterpret_cast<float4 *>(dst)[g_lane] = v;
        } else {
            dst[4 * g_lane + 0] = v.x;
            dst[4 * g_lane + 1] = v.y;
            dst[4 * g_lane + 2] = v.z;
            dst[4 * g_lane + 3] = v.w;
        }
    } else {
        dst[g_lane] = v.x;
    }
}

__global__ void __launch_bounds__(kFwdThreads) emb_pool_fwd_kernel(const __grid_constant__ DevGroup g) {
    const int fi = blockIdx.y;
    const int B = g.B;
    if (fi == g.num_features) {  // dense block (dnn.py:61-67 concat) + zero padding columns
        const int zero_w = g.zero_from >= 0 ? (int)(g.out_stride - g.zero_from) : 0;
        const int w = g.dense_width + zero_w;
        const int64_t total = (int64_t)B * w;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t b = i / w;
            const int c = (int)(i - b * w);
            if (c < g.dense_width)
                g.out[b * g.out_stride + g.dense_col + c] = __ldg(g.dense + b * g.dense_width + c);
            else
                g.out[b * g.out_stride + g.zero_from + (c - g.dense_width)] = 0.f;
        }
        return;
    }
    const DevFeature &f = g.f[fi];
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * kFwdWarps + (threadIdx.x >> 5);
    const int64_t total_warps = (int64_t)gridDim.x * kFwdWarps;
    const int G = f.G;
    const int L = f.L;
    const bool mean = f.pooling == CTR_POOL_MEAN;

    if (L == 1) {
        const int K = G < 4 ? G : 4;
        const int g_lane = lane & (G - 1);
        const int team = lane / G;
        const int bags_per_warp = (kWarp / G) * K;
        const bool col_ok = g_lane * f.vec < f.D;
        for (int64_t b0 = warp_global * bags_per_warp; b0 < B; b0 += total_warps * bags_per_warp) {
            const int64_t team_bag0 = b0 + (int64_t)team * K;
            int32_t row = -1;
            float w = 1.f;
            if (g_lane < K && team_bag0 + g_lane < B) {
                const int64_t id = __ldg(f.ids + team_bag0 + g_lane);
                row = map_index(f, id);
                if (row == -2) { flag_status(g.status, CTR_STATUS_INDEX_OOB); row = -1; }
                if (f.id_weight != nullptr && row >= 0) w = __ldg(f.id_weight + team_bag0 + g_lane);
            }
            float4 acc[4];
            int32_t rk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int src = team * G + (k < K ? k : 0);
                rk[k] = __shfl_sync(kFull, row, src);
                const float wk = __shfl_sync(kFull, w, src);
                acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < K && rk[k] >= 0 && col_ok) {
                    const float4 v = load_row_part(g, f, fi, rk[k], g_lane);
                    acc[k] = make_float4(v.x * wk, v.y * wk, v.z * wk, v.w * wk);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t bag = team_bag0 + k;
                if (k < K && bag < B) {
                    if (col_ok) store_out_part(g, f, bag, g_lane, acc[k]);
                    if (mean && g_lane == 0) f.bag_scale[bag] = 1.f;
                }
            }
        }
        return;
    }

    int T = pow2_ceil(L);
    T = T < G ? G : (T > kWarp ? kWarp : T);
    const int t = lane & (T - 1);
    const int team = lane / T;
    const int teams_per_warp = kWarp / T;
    const int slot = t / G;
    const int g_lane = t & (G - 1);
    const bool col_ok = g_lane * f.vec < f.D;
    for (int64_t b0 = warp_global * teams_per_warp; b0 < B; b0 += total_warps * teams_per_warp) {
        const int64_t bag = b0 + team;
        const bool bag_ok = bag < B;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int cnt = 0;
        for (int base = 0; base < L; base += T) {
            const int pos = base + t;
            int32_t row = -1;
            float w = 1.f;
            if (bag_ok && pos < L) {
                const int64_t id = __ldg(f.ids + bag * L + pos);
                row = map_index(f, id);
                if (row == -2) { flag_status(g.status, CTR_STATUS_INDEX_OOB); row = -1; }
                if (row >= 0) {
                    ++cnt;
                    if (f.id_weight != nullptr) w = __ldg(f.id_weight + bag * L + pos);
                }
            }
            const int src0 = team * T + slot * G;
#pragma unroll 4
            for (int k = 0; k < G; ++k) {
                const int32_t r = __shfl_sync(kFull, row, src0 + k);
                const float wk = __shfl_sync(kFull, w, src0 + k);
                if (r >= 0 && col_ok) {
                    const float4 v = load_row_part(g, f, fi, r, g_lane);
                    acc.x = fmaf(wk, v.x, acc.x);
                    acc.y = fmaf(wk, v.y, acc.y);
                    acc.z = fmaf(wk, v.z, acc.z);
                    acc.w = fmaf(wk, v.w, acc.w);
                }
            }
        }
        for (int off = G; off < T; off <<= 1) {
            acc.x += __shfl_xor_sync(kFull, acc.x, off);
            acc.y += __shfl_xor_sync(kFull, acc.y, off);
            acc.z += __shfl_xor_sync(kFull, acc.z, off);
            acc.w += __shfl_xor_sync(kFull, acc.w, off);
        }
        for (int off = 1; off < T; off <<= 1) cnt += __shfl_xor_sync(kFull, cnt, off);
        if (bag_ok) {
            float scale = 1.f;
            if (mean) {
                scale = 1.f / (float)(cnt > 1 ? cnt : 1);
                if (t == 0) f.bag_scale[bag] = scale;
                acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
            }
            if (slot == 0 && col_ok) store_out_part(g, f, bag, g_lane, acc);
        }
    }
}

// ---- fast path: single-id bags, uniform width ------------------------------------------------------------
// Every table of the group has L == 1, the same D = G * VEC, 16-byte aligned output slices (VEC == 4) and no
// per-id weights (Criteo-shaped groups: BASELINE configs 2-4).  Bag-major: a team of G lanes owns one bag and
// walks the tables four at a time -- four id loads, then four independent row loads, then four stores -- so
// there is no per-feature block row, no shuffle and ~2 instructions per id instead of ~20; the dense block and
// the zero padding of the same bag are written by the same team.
// EXTRA (DeepFM): the team holds every field of its bag, so the FM second-order term 0.5 * sum_d[(sum_f v)^2 - sum_f v^2]
// and the sum of the one-column twin tables (first-order weights) cost a few FMAs and one 4-byte load per field here,
// instead of two more passes over the [B, F * D] matrix: extra[bag] = sum_f twin_f[row] + FM, fm_sum[bag, :] = sum_f v.
template <int G, int VEC, bool SHARDED, bool EXTRA>
__global__ void __launch_bounds__(kFwdThreads) emb_pool_fwd_l1_kernel(const __grid_constant__ DevGroup g) {
    constexpr int kTeams = kFwdThreads / G;
    constexpr int D = G * VEC;
    const int t = threadIdx.x % G;
    const int F = g.num_features;
    const unsigned team_mask = G >= 32 ? kFull : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    for (int bag = blockIdx.x * kTeams + threadIdx.x / G; bag < g.B; bag += gridDim.x * kTeams) {
        float *out_row = g.out + (size_t)bag * g.out_stride;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), q4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float twin_sum = 0.f;
        for (int f0 = 0; f0 < F; f0 += 4) {
            int32_t row[4];
            bool writes[4];          // false: a WINDOW part whose id belongs to the table's other part (which writes the slice)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                row[j] = -1;
                writes[j] = f0 + j < F;
                if (f0 + j < F) {
                    const DevFeature &f = g.f[f0 + j];
                    const int64_t id = __ldg(f.ids + bag);
                    row[j] = map_index(f, id);
                    if (row[j] == -2) flag_status(g.status, CTR_STATUS_INDEX_OOB);
                    if (f.index_kind == CTR_INDEX_WINDOW && row[j] < 0) writes[j] = id < 0 && f.hash_seed == 0u;
                }
            }
            float4 v[4];
            float tw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                tw[j] = 0.f;
                if (row[j] >= 0) {
                    const DevFeature &ft = g.f[f0 + j];
                    const float *src, *tsrc = nullptr;
                    if (!SHARDED || ft.table != nullptr) {      // the local table (a sharded launch: a replicated feature)
                        src = ft.table + (size_t)(uint32_t)row[j] * D + t * VEC;
                        if (EXTRA && ft.twin_table != nullptr) tsrc = ft.twin_table + (uint32_t)row[j];
                    } else {                                    // the owner's shard through its peer mapping
                        uint32_t o;
                        const int64_t vr = shard_vrow(g, f0 + j, (uint32_t)row[j], &o);
                        src = g.peer_tables[o] + vr * D + t * VEC;
                        if (EXTRA && g.peer_twins[o] != nullptr) tsrc = g.peer_twins[o] + vr;
                    }
                    if (VEC == 4) v[j] = __ldg(reinterpret_cast<const float4 *>(src));
                    else v[j].x = __ldg(src);
                    if (EXTRA && t == 0 && tsrc != nullptr) tw[j] = __ldg(tsrc);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (writes[j]) {
                    float *dst = out_row + g.f[f0 + j].out_col + t * VEC;
                    if (VEC == 4) *reinterpret_cast<float4 *>(dst) = v[j];
                    else *dst = v[j].x;
                    if (t == 0 && g.f[f0 + j].pooling == CTR_POOL_MEAN) g.f[f0 + j].bag_scale[bag] = 1.f;
                    if (EXTRA) {
                        s4.x += v[j].x; s4.y += v[j].y; s4.z += v[j].z; s4.w += v[j].w;
                        q4.x = fmaf(v[j].x, v[j].x, q4.x); q4.y = fmaf(v[j].y, v[j].y, q4.y);
                        q4.z = fmaf(v[j].z, v[j].z, q4.z); q4.w = fmaf(v[j].w, v[j].w, q4.w);
                        twin_sum += tw[j];
                    }
                }
            }
        }
        for (int c = t; c < g.dense_width; c += G) out_row[g.dense_col + c] = __ldg(g.dense + (size_t)bag * g.dense_width + c);
        if (g.zero_from >= 0)
            for (int c = g.zero_from + t; c < (int)g.out_stride; c += G) out_row[c] = 0.f;
        if (EXTRA) {
            float e = 0.f;
            if (g.fm) {
                *reinterpret_cast<float4 *>(g.fm_sum + (size_t)bag * D + t * 4) = s4;
                e = 0.5f * ((s4.x * s4.x - q4.x) + (s4.y * s4.y - q4.y) + (s4.z * s4.z - q4.z) + (s4.w * s4.w - q4.w));
#pragma unroll
                for (int off = 1; off < G; off <<= 1) e += __shfl_xor_sync(team_mask, e, off);
            }
            if (t == 0) g.extra[bag] = e + twin_sum;
        }
    }
}

// 0 when the group does not qualify, else the lane count G (VEC reported through *vec)
static int l1_fast_path(const DevGroup &dg, int *vec) {
    if (dg.num_features == 0) return 0;
    const DevFeature &f0 = dg.f[0];
    for (int i = 0; i < dg.num_features; ++i) {
        const DevFeature &f = dg.f[i];
        if (f.L != 1 || f.D != f0.D || f.id_weight != nullptr) return 0;
        if (f.vec == 4 && !f.aligned) return 0;
    }
    *vec = f0.vec;
    if (f0.vec == 4 && (f0.G == 4 || f0.G == 8 || f0.G == 16)) return f0.G;
    if (f0.vec == 1 && f0.D == 1) return 1;
    return 0;
}

// ---- standalone index kernels ----------------------------------------------------------

__global__ void hash_bucket_kernel(const int64_t *__restrict__ ids, int64_t n, uint32_t buckets, uint32_t seed,
                                   int32_t *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int32_t)(murmur3_decimal(ids[i], seed) % buckets);
}

// MurmurHash3_x86_32 over arbitrary byte strings (string categories: after torchctr/transformer.py:367-401 a category that
// is not an int32 number keeps its own characters, e.g. Amazon's "B001NPEBGU"): string i = data[offsets[i], offsets[i + 1]).
__global__ void hash_bucket_bytes_kernel(const uint8_t *__restrict__ data, const int64_t *__restrict__ offsets, int64_t n,
                                         uint32_t buckets, uint32_t seed, int32_t *__restrict__ out) {
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *p = data + offsets[i];
        const int64_t len = offsets[i + 1] - offsets[i];
        uint32_t h = seed;
        int64_t j = 0;
        for (; j + 4 <= len; j += 4) {
            uint32_t k = (uint32_t)p[j] | ((uint32_t)p[j + 1] << 8) | ((uint32_t)p[j + 2] << 16) | ((uint32_t)p[j + 3] << 24);
            k *= c1; k = rotl32(k, 15); k *= c2;
            h ^= k; h = rotl32(h, 13); h = h * 5u + 0xe6546b64u;
        }
        uint32_t k = 0;
        for (int t = 0; j + t < len; ++t) k |= (uint32_t)p[j + t] << (8 * t);
        if (j < len) { k *= c1; k = rotl32(k, 15); k *= c2; h ^= k; }
        h ^= (uint32_t)len;
        h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
        out[i] = (int32_t)(h % buckets);
    }
}

__global__ void rows_gather_kernel(const int64_t *__restrict__ ids, int64_t n, const float *__restrict__ table,
                                   int64_t num_rows, int D, float *__restrict__ out, uint32_t *status) {
    // one thread per (row, 4-float piece) when D % 4 == 0, else per element
    const int pieces = (D % 4 == 0) ? D / 4 : D;
    const int64_t total = n * pieces;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / pieces;
        const int p = (int)(i - r * pieces);
        const int64_t id = __ldg(ids + r);
        const bool ok = id >= 0 && id < num_rows;
        if (!ok && p == 0) flag_status(status, CTR_STATUS_INDEX_OOB);
        if (D % 4 == 0) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) v = __ldg(reinterpret_cast<const float4 *>(table + id * D) + p);
            reinterpret_cast<float4 *>(out + r * D)[p] = v;
        } else {
            out[r * D + p] = ok ? __ldg(table + id * D + p) : 0.f;
        }
    }
}

// Counter-based normal generator: Philox-free, two rounds of mix64 -> Box-Muller.  Row/col
// addressed so that growing a table in several steps gives the same rows as growing it once.
// (row_first, row_stride): the GLOBAL row that local row i stands for is row_first + i * row_stride -- a rank that owns every
// P-th row of a table draws exactly the values the unsharded table holds in those rows.
__global__ void normal_fill_rows_kernel(float *table, int64_t row0, int64_t n, int D, float mean, float stdv,
                                        uint64_t seed, int64_t row_first, int64_t row_stride) {
    const int64_t total = n * D;
    table += row0 * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row_first + (i / D) * row_stride;
        const int c = (int)(i % D);
        const uint64_t ctr = mix64(seed ^ mix64((uint64_t)r * 0x9e3779b97f4a7c15ull + (uint64_t)c + 1ull));
        const uint32_t a = (uint32_t)(ctr >> 32), b = (uint32_t)ctr;
        const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
        const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
        table[i] = mean + stdv * z;
    }
}

__global__ void ids_minmax_kernel(const int64_t *__restrict__ ids, int64_t n, int64_t *out) {
    long long lo = INT64_MAX, hi = INT64_MIN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long v = ids[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    for (int off = 16; off; off >>= 1) {
        const long long l2 = __shfl_xor_sync(kFull, lo, off), h2 = __shfl_xor_sync(kFull, hi, off);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(reinterpret_cast<long long *>(out), lo);
        atomicMax(reinterpret_cast<long long *>(out) + 1, hi);
    }
}

__global__ void minmax_init_kernel(int64_t *out) {
    out[0] = INT64_MAX;
    out[1] = INT64_MIN;
}

static int grid_for(int64_t work_items, int threads, int cap) {
    int64_t blocks = (work_items + threads - 1) / threads;
    if (blocks < 1) blocks = 1;
    if (blocks > cap) blocks = cap;
    return (int)blocks;
}

}  // namespace ctr

using namespace ctr;

static int launch_pool_fwd(const DevGroup &dg, void *stream);

extern "C" int ctr_emb_pool_fwd(const ctr_group_t *group, void *stream) {
    static thread_local DevGroup dg;
    CTR_REQUIRE(group == nullptr || !group->grad_blocked, "grad_blocked describes a gradient layout: backward (ctr_emb_bwd_apply) only");
    int rc = lower_group(group, &dg, /*need_tables=*/true, /*need_out=*/true);
    if (rc != CTR_OK) return rc;
    return launch_pool_fwd(dg, stream);
}

extern "C" int ctr_emb_pool_fwd_sharded(const ctr_group_t *group, const ctr_shard_t *shard, const float *const *tables,
                                        void *stream) {
    static thread_local DevGroup dg;
    CTR_REQUIRE(group == nullptr || !group->grad_blocked, "grad_blocked describes a gradient layout: backward (ctr_emb_bwd_apply) only");
    int rc = lower_group(group, &dg, /*need_tables=*/false, /*need_out=*/true);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(tables != nullptr, "tables is null");
    rc = attach_shard(&dg, shard, tables);
    if (rc != CTR_OK) return rc;
    return launch_pool_fwd(dg, stream);
}

extern "C" int ctr_emb_pool_fwd_sharded_ex(const ctr_group_t *group, const ctr_shard_t *shard, const float *const *tables,
                                           const float *const *twin_tables, void *stream) {
    static thread_local DevGroup dg;
    CTR_REQUIRE(group == nullptr || !group->grad_blocked, "grad_blocked describes a gradient layout: backward (ctr_emb_bwd_apply) only");
    int rc = lower_group(group, &dg, /*need_tables=*/false, /*need_out=*/true);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(tables != nullptr, "tables is null");
    rc = attach_shard(&dg, shard, tables);
    if (rc != CTR_OK) return rc;
    if (twin_tables != nullptr) {
        CTR_REQUIRE(dg.extra != nullptr, "twin shards need group->extra [B]");
        for (int o = 0; o < shard->world; ++o) {
            CTR_REQUIRE(twin_tables[o] != nullptr, "twin_tables[%d] is null", o);
            dg.peer_twins[o] = twin_tables[o];
        }
    }
    return launch_pool_fwd(dg, stream);
}

static int launch_pool_fwd(const DevGroup &dg, void *stream) {
    const int extra_w = dg.dense_width + (dg.zero_from >= 0 ? (int)(dg.out_stride - dg.zero_from) : 0);
    if (dg.B == 0 || (dg.num_features == 0 && extra_w == 0)) return CTR_OK;
    int vec = 0;
    const int fastG = l1_fast_path(dg, &vec);
    if (fastG > 0) {
        const int teams = kFwdThreads / fastG;
        const int blocks = grid_for(dg.B, teams, kNumSMs * 16);
        cudaStream_t st = (cudaStream_t)stream;
        note_launch();
        const bool sh = dg.world > 1;
        const bool ex = dg.extra != nullptr;
        if (ex) {
            if (vec != 4) {
                set_error("group->extra (twin tables / FM term) needs a single-id group of one width, D %% 4 == 0");
                return CTR_E_UNSUPPORTED;
            }
            if (dg.fm) CTR_REQUIRE((reinterpret_cast<uintptr_t>(dg.fm_sum) & 15u) == 0, "fm_sum must be 16-byte aligned");
            if (fastG == 4 && sh) emb_pool_fwd_l1_kernel<4, 4, true, true><<<blocks, kFwdThreads, 0, st>>>(dg);
            else if (fastG == 4) emb_pool_fwd_l1_kernel<4, 4, false, true><<<blocks, kFwdThreads, 0, st>>>(dg);
            else if (fastG == 8 && sh) emb_pool_fwd_l1_kernel<8, 4, true, true><<<blocks, kFwdThreads, 0, st>>>(dg);
            else if (fastG == 8) emb_pool_fwd_l1_kernel<8, 4, false, true><<<blocks, kFwdThreads, 0, st>>>(dg);
            else if (sh) emb_pool_fwd_l1_kernel<16, 4, true, true><<<blocks, kFwdThreads, 0, st>>>(dg);
            else emb_pool_fwd_l1_kernel<16, 4, false, true><<<blocks, kFwdThreads, 0, st>>>(dg);
        }
        else if (vec == 1 && sh) emb_pool_fwd_l1_kernel<1, 1, true, false><<<blocks, kFwdThreads, 0, st>>>(dg);
        else if (vec == 1) emb_pool_fwd_l1_kernel<1, 1, false, false><<<blocks, kFwdThreads, 0, st>>>(dg);
        else if (fastG == 4 && sh) emb_pool_fwd_l1_kernel<4, 4, true, false><<<blocks, kFwdThreads, 0, st>>>(dg);
        else if (fastG == 4) emb_pool_fwd_l1_kernel<4, 4, false, false><<<blocks, kFwdThreads, 0, st>>>(dg);
        else if (fastG == 8 && sh) emb_pool_fwd_l1_kernel<8, 4, true, false><<<blocks, kFwdThreads, 0, st>>>(dg);
        else if (fastG == 8) emb_pool_fwd_l1_kernel<8, 4, false, false><<<blocks, kFwdThreads, 0, st>>>(dg);
        else if (sh) emb_pool_fwd_l1_kernel<16, 4, true, false><<<blocks, kFwdThreads, 0, st>>>(dg);
        else emb_pool_fwd_l1_kernel<16, 4, false, false><<<blocks, kFwdThreads, 0, st>>>(dg);
        CTR_CUDA_OK(cudaGetLastError());
        return CTR_OK;
    }
    if (dg.extra != nullptr) {
        set_error("group->extra (twin tables / FM term) needs a single-id group of one width, D %% 4 == 0, no per-id weights");
        return CTR_E_UNSUPPORTED;
    }
    int64_t max_warps = 1;
    for (int i = 0; i < dg.num_features; ++i) {
        const DevFeature &f = dg.f[i];
        int bags_per_warp;
        if (f.L == 1) {
            bags_per_warp = (kWarp / f.G) * (f.G < 4 ? f.G : 4);
        } else {
            int T = pow2_ceil(f.L);
            T = T < f.G ? f.G : (T > kWarp ? kWarp : T);
            bags_per_warp = kWarp / T;
        }
        const int64_t warps = (dg.B + bags_per_warp - 1) / bags_per_warp;
        if (warps > max_warps) max_warps = warps;
    }
    if (extra_w > 0) {
        const int64_t warps = ((int64_t)dg.B * extra_w + 4 * kWarp - 1) / (4 * kWarp);
        if (warps > max_warps) max_warps = warps;
    }
    // enough blocks for every feature row of the grid to fill the machine a few times over
    const int cap = kNumSMs * 8;
    dim3 grid(grid_for(max_warps, kFwdWarps, cap), dg.num_features + (extra_w > 0 ? 1 : 0));
    note_launch(), emb_pool_fwd_kernel<<<grid, kFwdThreads, 0, (cudaStream_t)stream>>>(dg);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_hash_bucket_i64(const int64_t *ids, int64_t n, uint32_t buckets, uint32_t seed, int32_t *out,
                                   void *stream) {
    CTR_REQUIRE(n >= 0 && buckets > 0, "n=%lld buckets=%u", (long long)n, buckets);
    if (n == 0) return CTR_OK;
    CTR_REQUIRE(ids != nullptr && out != nullptr, "null pointer");
    note_launch(), hash_bucket_kernel<<<grid_for(n, 256, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(ids, n, buckets, seed, out);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_hash_bucket_bytes(const uint8_t *data, const int64_t *offsets, int64_t n, uint32_t buckets, uint32_t seed,
                                     int32_t *out, void *stream) {
    CTR_REQUIRE(n >= 0 && buckets > 0, "n=%lld buckets=%u", (long long)n, buckets);
    if (n == 0) return CTR_OK;
    CTR_REQUIRE(offsets != nullptr && out != nullptr, "null pointer");
    note_launch(), hash_bucket_bytes_kernel<<<grid_for(n, 256, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(data, offsets, n, buckets, seed, out);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_rows_gather(const int64_t *ids, int64_t n, const float *table, int64_t num_rows, int32_t D,
                               float *out, uint32_t *status, void *stream) {
    CTR_REQUIRE(n >= 0 && D >= 1 && num_rows >= 0, "bad sizes");
    if (n == 0) return CTR_OK;
    CTR_REQUIRE(ids != nullptr && table != nullptr && out != nullptr, "null pointer");
    const int pieces = (D % 4 == 0) ? D / 4 : D;
    note_launch(), rows_gather_kernel<<<grid_for(n * pieces, 256, kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(ids, n, table, num_rows,
                                                                                              D, out, status);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_normal_fill_rows(float *table, int64_t row0, int64_t n, int32_t D, float mean, float stdv,
                                    uint64_t seed, void *stream) {
    CTR_REQUIRE(n >= 0 && D >= 1 && row0 >= 0, "bad sizes");
    if (n == 0) return CTR_OK;
    CTR_REQUIRE(table != nullptr, "null pointer");
    note_launch(), normal_fill_rows_kernel<<<grid_for(n * D, 256, kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(table, row0, n, D, mean,
                                                                                              stdv, seed, row0, 1);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_normal_fill_rows_strided(float *table, int64_t row0, int64_t n, int32_t D, float mean, float stdv, uint64_t seed,
                                            int64_t global_first, int64_t global_stride, void *stream) {
    CTR_REQUIRE(n >= 0 && D >= 1 && row0 >= 0 && global_first >= 0 && global_stride >= 1, "bad sizes");
    if (n == 0) return CTR_OK;
    CTR_REQUIRE(table != nullptr, "null pointer");
    note_launch(), normal_fill_rows_kernel<<<grid_for(n * D, 256, kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
        table, row0, n, D, mean, stdv, seed, global_first, global_stride);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_ids_minmax(const int64_t *ids, int64_t n, int64_t *out, void *stream) {
    CTR_REQUIRE(n >= 0 && out != nullptr, "bad args");
    note_launch(), minmax_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out);
    if (n > 0) {
        CTR_REQUIRE(ids != nullptr, "null ids");
        note_launch(), ids_minmax_kernel<<<grid_for(n, 256, kNumSMs * 4), 256, 0, (cudaStream_t)stream>>>(ids, n, out);
    }
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}
