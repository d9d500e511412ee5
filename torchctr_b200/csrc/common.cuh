// Shared device helpers for libctr_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ctr_b200.h"

namespace ctr {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNumSMs = 148;  // B200

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void note_launch();

#define CTR_CUDA_OK(expr)                                     \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return ctr::cuda_fail(_e, #expr); \
    } while (0)

#define CTR_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            ctr::set_error(__VA_ARGS__); \
            return CTR_E_BADARG;        \
        }                               \
    } while (0)

// Device-side view of one feature (ctr_feature_t with the vocabulary map inlined and the
// lane layout precomputed on the host).
struct DevFeature {
    const int64_t *ids;
    const float *id_weight;
    float *table;
    float *state0;
    float *state1;
    float *bag_scale;
    const int64_t *map_keys;
    const int32_t *map_rows;
    int64_t map_mask;  // capacity - 1
    uint32_t num_rows;
    uint32_t row_base;  // first key of this table in the group's key space (backward)
    int32_t L;
    int32_t D;
    int32_t out_col;
    int32_t pooling;
    int32_t index_kind;
    uint32_t hash_seed;
    int32_t vec;       // 4: float4 lanes, 1: scalar lanes
    int32_t G;         // lanes that cover one row (power of two)
    int32_t aligned;   // 1 when float4 access to the pooled / grad matrix is 16-byte aligned
    float *twin_table;   // one-column twin table (DeepFM first-order weights) or null
    float *twin_state0;
    float *twin_state1;
};

struct DevGroup {
    DevFeature f[CTR_MAX_FEATURES];
    int32_t num_features;
    int32_t B;
    float *out;
    int64_t out_stride;
    const float *dense;
    int32_t dense_width;
    int32_t dense_col;
    int32_t zero_from;
    uint32_t *status;
    float *extra;        // [B] per-bag scalar: sum of twins (+ FM term) forward, its gradient backward; or null
    float *fm_sum;       // [B, D] sum over the features of the pooled vectors (FM), or null
    int32_t fm;
    int32_t has_twin;    // some feature has a twin table
    // row-sharded tables (world > 1): row r of table f lives on rank (r + f) % world, at row
    // shard_adj[owner * num_features + f] + (r + f) / world of that rank's fused shard peer_tables[owner]
    int32_t world;
    int32_t rank;
    const int64_t *shard_adj;
    const float *peer_tables[CTR_MAX_WORLD];
    // one-column twin shards (DeepFM first-order weights) of the sharded features, same geometry; null = none
    const float *peer_twins[CTR_MAX_WORLD];
};

// Validates a ctr_shard_t and attaches it to a lowered group.
int attach_shard(DevGroup *g, const ctr_shard_t *shard, const float *const *tables);

// (table f, row) -> owner rank and row inside the owner's fused shard
__device__ __forceinline__ int64_t shard_vrow(const DevGroup &g, int fi, uint32_t row, uint32_t *owner) {
    const uint32_t r = row + (uint32_t)fi;
    const uint32_t o = r % (uint32_t)g.world;
    *owner = o;
    return __ldg(g.shard_adj + (size_t)o * g.num_features + fi) + (int64_t)(r / (uint32_t)g.world);
}

// address of (table f, row): the local table, or the owner's shard through its peer mapping.  In a sharded launch a
// feature that carries its own `table` pointer is REPLICATED (every rank holds all of it): read locally.
__device__ __forceinline__ const float *table_row(const DevGroup &g, const DevFeature &f, int fi, int32_t row) {
    if (g.world <= 1 || f.table != nullptr) return f.table + (size_t)(uint32_t)row * f.D;
    uint32_t o;
    const int64_t v = shard_vrow(g, fi, (uint32_t)row, &o);
    return g.peer_tables[o] + v * f.D;
}

// Validates a ctr_group_t and lowers it to the device view.  need_tables: table pointers must
// be non-null.  Returns CTR_OK or a negative code.
int lower_group(const ctr_group_t *g, DevGroup *out, bool need_tables, bool need_out);

// ---------------------------------------------------------------------------------------
// MurmurHash3_x86_32 over the decimal ASCII of a signed 64-bit integer: what
// torchctr/utils.py:113 computes for a numeric category (transformer.py:371-381 turns the
// value into its decimal string first).  Bytes are produced most-significant digit first and
// consumed as little-endian 32-bit blocks.
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

__device__ __forceinline__ uint32_t murmur3_decimal(int64_t v, uint32_t seed) {
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    const bool neg = v < 0;
    uint64_t u = neg ? (uint64_t)(-(v + 1)) + 1ull : (uint64_t)v;
    // digits, least significant first
    uint8_t dig[20];
    int nd = 0;
    if (u <= 0xffffffffull) {
        uint32_t w = (uint32_t)u;
        do { dig[nd++] = (uint8_t)('0' + w % 10u); w /= 10u; } while (w);
    } else {
        do { dig[nd++] = (uint8_t)('0' + (uint32_t)(u % 10ull)); u /= 10ull; } while (u);
    }
    const int len = nd + (neg ? 1 : 0);
    uint32_t h = seed;
    uint32_t k = 0;
    int inblock = 0;
    for (int p = 0; p < len; ++p) {
        const uint32_t ch = (neg && p == 0) ? (uint32_t)'-' : (uint32_t)dig[len - 1 - p];
        k |= ch << (8 * inblock);
        if (++inblock == 4) {
            k *= c1; k = rotl32(k, 15); k *= c2;
            h ^= k; h = rotl32(h, 13); h = h * 5u + 0xe6546b64u;
            k = 0; inblock = 0;
        }
    }
    if (inblock) {
        k *= c1; k = rotl32(k, 15); k *= c2; h ^= k;
    }
    h ^= (uint32_t)len;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

// 64-bit mix used to pick the first probe slot of the vocabulary map (splitmix64 finaliser).
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

// Linear-probing lookup; returns the stored row or `miss`.
__device__ __forceinline__ int32_t map_find(const int64_t *__restrict__ keys, const int32_t *__restrict__ rows,
                                            int64_t mask, int64_t key, int32_t miss) {
    uint64_t slot = mix64((uint64_t)key) & (uint64_t)mask;
    for (int64_t probe = 0; probe <= mask; ++probe) {
        const int64_t k = keys[slot];
        if (k == key) return rows[slot];
        if (k == CTR_VOCAB_EMPTY) return miss;
        slot = (slot + 1) & (uint64_t)mask;
    }
    return miss;
}

// Raw id -> table row of a feature.  Returns -1 for padding; -2 for an out-of-range row.
__device__ __forceinline__ int32_t map_index(const DevFeature &f, int64_t id) {
    if (id < 0) return -1;
    if (f.index_kind == CTR_INDEX_DIRECT) {
        return id < (int64_t)f.num_rows ? (int32_t)id : -2;
    } else if (f.index_kind == CTR_INDEX_HASH) {
        return (int32_t)(murmur3_decimal(id, f.hash_seed) % f.num_rows);
    } else if (f.index_kind == CTR_INDEX_WINDOW) {      // a part [first, first + num_rows) of a table; ids outside are not ours
        const int64_t r = id - (int64_t)f.hash_seed;
        return (r >= 0 && r < (int64_t)f.num_rows) ? (int32_t)r : -1;
    } else {
        const int32_t r = map_find(f.map_keys, f.map_rows, f.map_mask, id, 0);
        return (uint32_t)r < f.num_rows ? r : -2;
    }
}

__device__ __forceinline__ float4 ld_row4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__host__ __device__ __forceinline__ int pow2_ceil(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}
__host__ __device__ __forceinline__ int pow2_floor(int x) {
    int p = 1;
    while (2 * p <= x) p <<= 1;
    return p;
}

}  // namespace ctr
