// Error reporting and argument lowering shared by every entry point of libctr_b200.
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace ctr {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return CTR_E_CUDA;
}

int lower_group(const ctr_group_t *g, DevGroup *out, bool need_tables, bool need_out) {
    CTR_REQUIRE(g != nullptr && out != nullptr, "group is null");
    CTR_REQUIRE(g->num_features >= 0 && g->num_features <= CTR_MAX_FEATURES,
                "num_features=%d outside [0, %d]", g->num_features, CTR_MAX_FEATURES);
    CTR_REQUIRE(g->num_features == 0 || g->features != nullptr, "features is null");
    CTR_REQUIRE(g->B >= 0, "B=%d is negative", g->B);
    CTR_REQUIRE(!need_out || g->out != nullptr || g->B == 0, "out is null");
    CTR_REQUIRE(g->dense_width >= 0, "dense_width is negative");
    CTR_REQUIRE(g->dense_width == 0 || g->dense != nullptr, "dense is null but dense_width=%d", g->dense_width);
    memset(out, 0, sizeof(*out));
    out->num_features = g->num_features;
    out->B = g->B;
    out->out = g->out;
    out->out_stride = (g->grad_blocked && g->num_features > 0) ? g->features[0].D : g->out_stride;
    out->dense = g->dense;
    out->dense_width = g->dense_width;
    out->dense_col = g->dense_col;
    out->status = g->status;
    out->extra = g->extra;
    out->fm_sum = g->fm_sum;
    out->fm = g->fm ? 1 : 0;
    CTR_REQUIRE(!g->fm || (g->extra != nullptr && g->fm_sum != nullptr), "fm needs extra [B] and fm_sum [B, D]");
    out->zero_from = (g->zero_from >= 0 && g->zero_from < g->out_stride) ? g->zero_from : -1;
    CTR_REQUIRE(g->dense_width == 0 || (g->dense_col >= 0 && g->dense_col + (int64_t)g->dense_width <= g->out_stride),
                "dense block [%d, %d) outside out_stride=%lld", g->dense_col, g->dense_col + g->dense_width,
                (long long)g->out_stride);
    uint64_t base = 0;
    for (int i = 0; i < g->num_features; ++i) {
        const ctr_feature_t &s = g->features[i];
        DevFeature &d = out->f[i];
        CTR_REQUIRE(s.ids != nullptr || g->B == 0, "feature %d: ids is null", i);
        CTR_REQUIRE(!need_tables || s.table != nullptr, "feature %d: table is null", i);
        CTR_REQUIRE(s.num_rows > 0 && s.num_rows < (1ll << 31), "feature %d: num_rows=%lld outside [1, 2^31)", i,
                    (long long)s.num_rows);
        CTR_REQUIRE(s.L >= 1, "feature %d: L=%d must be >= 1", i, s.L);
        CTR_REQUIRE((int64_t)s.L * g->B < (1ll << 32), "feature %d: B*L does not fit 32 bits", i);
        CTR_REQUIRE(s.D >= 1 && ((s.D % 4 == 0 && s.D <= 128) || s.D <= 32),
                    "feature %d: D=%d unsupported (multiple of 4 up to 128, or any D <= 32)", i, s.D);
        CTR_REQUIRE(s.pooling == CTR_POOL_SUM || s.pooling == CTR_POOL_MEAN, "feature %d: bad pooling %d", i, s.pooling);
        CTR_REQUIRE(s.out_col >= 0 && s.out_col + (int64_t)s.D <= g->out_stride,
                    "feature %d: columns [%d, %d) outside out_stride=%lld", i, s.out_col, s.out_col + s.D,
                    (long long)g->out_stride);
        if (g->grad_blocked) {
            // backward only: the columns of feature i are their own contiguous [B, D] matrix at out + out_col * B (what
            // ctr_linear_fwd_blocked writes with block_cols = D, block_stride = B * D): lowered to row pitch D and a
            // per-feature offset, which is all the sweep kernels address the gradient by
            CTR_REQUIRE(s.L == 1 && s.D % 4 == 0 && s.D == g->features[0].D && s.out_col % s.D == 0 && s.id_weight == nullptr &&
                            s.pooling == CTR_POOL_SUM,
                        "feature %d: grad_blocked needs single-id, sum-pooled features of one width D %% 4 == 0 at columns that are "
                        "multiples of D", i);
            CTR_REQUIRE((int64_t)s.out_col * g->B + (int64_t)g->B * s.D < (1ll << 31), "feature %d: blocked gradient offset does not fit 31 bits", i);
        }
        d.ids = s.ids;
        d.id_weight = s.id_weight;
        d.table = s.table;
        d.state0 = s.state0;
        d.state1 = s.state1;
        d.bag_scale = s.bag_scale;
        d.num_rows = (uint32_t)s.num_rows;
        d.L = s.L;
        d.D = s.D;
        d.out_col = g->grad_blocked ? (int32_t)((int64_t)s.out_col * g->B) : s.out_col;
        d.pooling = s.pooling;
        d.index_kind = s.index_kind;
        d.hash_seed = s.hash_seed;
        d.twin_table = s.twin_table;
        d.twin_state0 = s.twin_state0;
        d.twin_state1 = s.twin_state1;
        if (s.twin_table != nullptr) {
            CTR_REQUIRE(g->extra != nullptr, "feature %d: a twin table needs group->extra [B]", i);
            out->has_twin = 1;
        }
        if (s.index_kind == CTR_INDEX_REMAP) {
            CTR_REQUIRE(s.map != nullptr && s.map->keys != nullptr && s.map->rows != nullptr,
                        "feature %d: REMAP needs a vocabulary map", i);
            CTR_REQUIRE(s.map->capacity > 0 && (s.map->capacity & (s.map->capacity - 1)) == 0,
                        "feature %d: map capacity must be a power of two", i);
            d.map_keys = s.map->keys;
            d.map_rows = s.map->rows;
            d.map_mask = s.map->capacity - 1;
        } else {
            CTR_REQUIRE(s.index_kind == CTR_INDEX_DIRECT || s.index_kind == CTR_INDEX_HASH || s.index_kind == CTR_INDEX_WINDOW,
                        "feature %d: bad index_kind %d", i, s.index_kind);
        }
        CTR_REQUIRE(s.pooling != CTR_POOL_MEAN || s.bag_scale != nullptr || g->B == 0,
                    "feature %d: mean pooling needs bag_scale [B]", i);
        d.vec = (s.D % 4 == 0) ? 4 : 1;
        d.G = pow2_ceil(s.D / d.vec);
        d.aligned = (d.vec == 4 && s.out_col % 4 == 0 && out->out_stride % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(g->out) & 15u) == 0)
                        ? 1
                        : 0;
        if (d.vec == 4 && s.table != nullptr)
            CTR_REQUIRE((reinterpret_cast<uintptr_t>(s.table) & 15u) == 0, "feature %d: table not 16-byte aligned", i);
        d.row_base = (uint32_t)base;
        base += (uint64_t)s.num_rows;
        CTR_REQUIRE(base < 0xffffffffull, "group key space (sum of num_rows) must stay below 2^32-1; split the group");
    }
    return CTR_OK;
}

int attach_shard(DevGroup *g, const ctr_shard_t *shard, const float *const *tables) {
    CTR_REQUIRE(shard != nullptr, "shard is null");
    CTR_REQUIRE(shard->world >= 1 && shard->world <= CTR_MAX_WORLD, "world=%d outside [1, %d]", shard->world, CTR_MAX_WORLD);
    CTR_REQUIRE(shard->rank >= 0 && shard->rank < shard->world, "rank=%d outside [0, %d)", shard->rank, shard->world);
    CTR_REQUIRE(shard->adj != nullptr, "shard->adj is null");
    for (int i = 1; i < g->num_features; ++i)
        CTR_REQUIRE(g->f[i].D == g->f[0].D, "sharded groups need one embedding dim (feature %d: %d vs %d)", i, g->f[i].D,
                    g->f[0].D);
    g->world = shard->world;
    g->rank = shard->rank;
    g->shard_adj = shard->adj;
    if (tables != nullptr) {
        for (int o = 0; o < shard->world; ++o) {
            CTR_REQUIRE(tables[o] != nullptr, "tables[%d] is null", o);
            CTR_REQUIRE((reinterpret_cast<uintptr_t>(tables[o]) & 15u) == 0, "tables[%d] not 16-byte aligned", o);
            g->peer_tables[o] = tables[o];
        }
    }
    return CTR_OK;
}

}  // namespace ctr

extern "C" const char *ctr_last_error_string(void) { return ctr::g_err; }
extern "C" int ctr_abi_version(void) { return CTR_B200_ABI_VERSION; }

// every kernel launch of the library passes through note_launch(): bench.py reports the count
namespace ctr {
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace ctr
extern "C" int64_t ctr_kernel_launches(void) { return (int64_t)ctr::g_launches.load(); }
