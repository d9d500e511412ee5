// K3: vocabulary fit / transform on the device (dynamic-growth tables).
//
// Restates torchctr/transformer.py:451-498 for integer keys:
//   fit        count the keys of the batch (value_counts, :452-453), keep those with
//              count >= min_freq (:456-460), give every admitted key that is not yet in the
//              vocabulary the next free row (:462-472), add the batch count to known keys (:473-474);
//   transform  key -> row, unknown -> OOV row (:492-498).
// The reference walks polars' value_counts rows, whose order is unspecified; here new keys
// take their rows in order of FIRST OCCURRENCE in the batch, which is deterministic: a scratch
// table records (count, first position) per distinct key with atomics whose results do not
// depend on scheduling (add / min), the representatives are flagged, and an exclusive scan
// over batch positions hands out the rows.
#include "sort.cuh"

namespace ctr {

struct FitLayout {
    int64_t cap;  // scratch table capacity (power of two >= 2n)
    int64_t skeys, scount, sfirst, pslot, flag, rank, spine, total;
};

static FitLayout fit_layout(int64_t n) {
    FitLayout l{};
    int64_t cap = 64;
    while (cap < 2 * n) cap <<= 1;
    l.cap = cap;
    auto al = [](int64_t x) { return (x + 255) & ~int64_t(255); };
    int64_t off = 0;
    l.skeys = off; off = al(off + cap * 8);
    l.scount = off; off = al(off + cap * 4);
    l.sfirst = off; off = al(off + cap * 4);
    l.pslot = off; off = al(off + n * 4);
    l.flag = off; off = al(off + n * 4);
    l.rank = off; off = al(off + n * 4);
    l.spine = off; off = al(off + (scan_spine_elems(n) + 8) * 4);
    l.total = off;
    return l;
}

__global__ void fit_init_kernel(long long *skeys, uint32_t *scount, uint32_t *sfirst, int64_t cap) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
        skeys[i] = CTR_VOCAB_EMPTY;
        scount[i] = 0;
        sfirst[i] = 0xffffffffu;
    }
}

// scratch insert: distinct keys of the batch with their count and first position
__global__ void fit_count_kernel(const int64_t *__restrict__ keys, int64_t n, long long *skeys, uint32_t *scount,
                                 uint32_t *sfirst, int64_t mask, uint32_t *__restrict__ pslot) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const long long key = keys[p];
        if (key < 0) { pslot[p] = 0xffffffffu; continue; }
        uint64_t slot = mix64((uint64_t)key) & (uint64_t)mask;
        while (true) {  // capacity >= 2n: an empty slot always exists
            long long cur = skeys[slot];
            if (cur == CTR_VOCAB_EMPTY)
                cur = (long long)atomicCAS(reinterpret_cast<unsigned long long *>(skeys + slot),
                                           (unsigned long long)CTR_VOCAB_EMPTY, (unsigned long long)key);
            if (cur == CTR_VOCAB_EMPTY || cur == key) break;
            slot = (slot + 1) & (uint64_t)mask;
        }
        atomicAdd(scount + slot, 1u);
        atomicMin(sfirst + slot, (uint32_t)p);
        pslot[p] = (uint32_t)slot;
    }
}

// representative positions: known key -> add count; unknown and frequent enough -> flag as new
__global__ void fit_flag_kernel(const int64_t *__restrict__ keys, int64_t n, const uint32_t *__restrict__ scount,
                                const uint32_t *__restrict__ sfirst, const uint32_t *__restrict__ pslot,
                                const int64_t *__restrict__ mkeys, const int32_t *__restrict__ mrows, int64_t mmask,
                                int min_freq, long long *counts, uint32_t *__restrict__ flag) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        uint32_t fl = 0;
        const uint32_t slot = pslot[p];
        if (slot != 0xffffffffu && sfirst[slot] == (uint32_t)p) {
            const uint32_t c = scount[slot];
            if (min_freq <= 0 || c >= (uint32_t)min_freq) {  // transformer.py:456-460 filters before the vocab is consulted
                const int32_t row = map_find(mkeys, mrows, mmask, keys[p], -1);
                if (row >= 0) {
                    if (counts != nullptr) atomicAdd(reinterpret_cast<unsigned long long *>(counts + row), (unsigned long long)c);
                } else {
                    fl = 1;
                }
            }
        }
        flag[p] = fl;
    }
}

__global__ void fit_insert_kernel(const int64_t *__restrict__ keys, int64_t n, const uint32_t *__restrict__ flag,
                                  const uint32_t *__restrict__ rank, const uint32_t *__restrict__ scount,
                                  const uint32_t *__restrict__ pslot, long long *mkeys, int32_t *mrows, int64_t mmask,
                                  const int64_t *__restrict__ next_row, long long *counts, uint32_t *status) {
    const int64_t base = next_row[0];
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        if (!flag[p]) continue;
        const long long key = keys[p];
        const int64_t row = base + rank[p];
        uint64_t slot = mix64((uint64_t)key) & (uint64_t)mmask;
        bool placed = false;
        for (int64_t probe = 0; probe <= mmask; ++probe) {
            if (mkeys[slot] == CTR_VOCAB_EMPTY &&
                atomicCAS(reinterpret_cast<unsigned long long *>(mkeys + slot), (unsigned long long)CTR_VOCAB_EMPTY,
                          (unsigned long long)key) == (unsigned long long)CTR_VOCAB_EMPTY) {
                mrows[slot] = (int32_t)row;
                placed = true;
                break;
            }
            slot = (slot + 1) & (uint64_t)mmask;
        }
        if (!placed && status != nullptr) atomicOr(status, CTR_STATUS_MAP_FULL);
        if (placed && counts != nullptr) counts[row] = scount[pslot[p]];
    }
}

__global__ void fit_advance_kernel(int64_t *next_row, const uint32_t *total) { next_row[0] += (int64_t)total[0]; }

__global__ void vocab_transform_kernel(const int64_t *__restrict__ keys, int64_t n, const int64_t *__restrict__ mkeys,
                                       const int32_t *__restrict__ mrows, int64_t mmask, int32_t oov,
                                       int32_t *__restrict__ rows) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t key = keys[p];
        rows[p] = key < 0 ? oov : map_find(mkeys, mrows, mmask, key, oov);
    }
}

// insert distinct (key, row) pairs that are not in the map yet (rebuild after a capacity change)
__global__ void vocab_insert_kernel(const int64_t *__restrict__ keys, const int32_t *__restrict__ rows, int64_t n,
                                    long long *mkeys, int32_t *mrows, int64_t mmask, uint32_t *status) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const long long key = keys[p];
        if (key < 0) continue;
        uint64_t slot = mix64((uint64_t)key) & (uint64_t)mmask;
        bool placed = false;
        for (int64_t probe = 0; probe <= mmask; ++probe) {
            if (mkeys[slot] == CTR_VOCAB_EMPTY &&
                atomicCAS(reinterpret_cast<unsigned long long *>(mkeys + slot), (unsigned long long)CTR_VOCAB_EMPTY,
                          (unsigned long long)key) == (unsigned long long)CTR_VOCAB_EMPTY) {
                mrows[slot] = rows[p];
                placed = true;
                break;
            }
            slot = (slot + 1) & (uint64_t)mmask;
        }
        if (!placed && status != nullptr) atomicOr(status, CTR_STATUS_MAP_FULL);
    }
}

__global__ void vocab_clear_kernel(long long *mkeys, int32_t *mrows, int64_t cap) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
        mkeys[i] = CTR_VOCAB_EMPTY;
        mrows[i] = 0;
    }
}

static int check_map(const ctr_vocab_map_t *m) {
    CTR_REQUIRE(m != nullptr && m->keys != nullptr && m->rows != nullptr, "vocabulary map is null");
    CTR_REQUIRE(m->capacity > 0 && (m->capacity & (m->capacity - 1)) == 0, "map capacity must be a power of two");
    return CTR_OK;
}

static unsigned grid1d(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > kNumSMs * 16) b = kNumSMs * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace ctr

using namespace ctr;

extern "C" int64_t ctr_vocab_fit_workspace_bytes(int64_t n) { return n < 0 ? CTR_E_BADARG : fit_layout(n).total; }

extern "C" int ctr_vocab_fit(const ctr_vocab_map_t *map, const int64_t *keys, int64_t n, int32_t min_freq,
                             int64_t *next_row, int64_t *counts, uint32_t *status, void *workspace,
                             int64_t workspace_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_map(map);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(n >= 0 && n < (1ll << 31), "n=%lld outside [0, 2^31)", (long long)n);
    CTR_REQUIRE(next_row != nullptr, "next_row is null");
    if (n == 0) return CTR_OK;
    CTR_REQUIRE(keys != nullptr && workspace != nullptr, "null pointer");
    const FitLayout l = fit_layout(n);
    if (workspace_bytes < l.total) {
        set_error("workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)l.total);
        return CTR_E_WORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    long long *skeys = reinterpret_cast<long long *>(ws + l.skeys);
    uint32_t *scount = reinterpret_cast<uint32_t *>(ws + l.scount), *sfirst = reinterpret_cast<uint32_t *>(ws + l.sfirst);
    uint32_t *pslot = reinterpret_cast<uint32_t *>(ws + l.pslot), *flag = reinterpret_cast<uint32_t *>(ws + l.flag);
    uint32_t *rank = reinterpret_cast<uint32_t *>(ws + l.rank), *spine = reinterpret_cast<uint32_t *>(ws + l.spine);
    note_launch(), fit_init_kernel<<<grid1d(l.cap), 256, 0, stream>>>(skeys, scount, sfirst, l.cap);
    note_launch(), fit_count_kernel<<<grid1d(n), 256, 0, stream>>>(keys, n, skeys, scount, sfirst, l.cap - 1, pslot);
    note_launch(), fit_flag_kernel<<<grid1d(n), 256, 0, stream>>>(keys, n, scount, sfirst, pslot, map->keys, map->rows, map->capacity - 1,
                                                  min_freq, reinterpret_cast<long long *>(counts), flag);
    rc = exclusive_scan_u32_to(flag, rank, n, spine, stream);
    if (rc != CTR_OK) return rc;
    note_launch(), fit_insert_kernel<<<grid1d(n), 256, 0, stream>>>(keys, n, flag, rank, scount, pslot,
                                                    reinterpret_cast<long long *>(map->keys), map->rows, map->capacity - 1,
                                                    next_row, reinterpret_cast<long long *>(counts), status);
    note_launch(), fit_advance_kernel<<<1, 1, 0, stream>>>(next_row, spine + scan_num_blocks(n));
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_vocab_transform(const ctr_vocab_map_t *map, const int64_t *keys, int64_t n, int32_t oov_row,
                                   int32_t *rows, void *stream) {
    int rc = check_map(map);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(n >= 0, "n is negative");
    if (n == 0) return CTR_OK;
    CTR_REQUIRE(keys != nullptr && rows != nullptr, "null pointer");
    note_launch(), vocab_transform_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(keys, n, map->keys, map->rows, map->capacity - 1,
                                                                       oov_row, rows);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_vocab_insert(const ctr_vocab_map_t *map, const int64_t *keys, const int32_t *rows, int64_t n,
                                uint32_t *status, void *stream) {
    int rc = check_map(map);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(n >= 0, "n is negative");
    if (n == 0) return CTR_OK;
    CTR_REQUIRE(keys != nullptr && rows != nullptr, "null pointer");
    note_launch(), vocab_insert_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(keys, rows, n, reinterpret_cast<long long *>(map->keys),
                                                                    map->rows, map->capacity - 1, status);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_vocab_clear(const ctr_vocab_map_t *map, void *stream) {
    int rc = check_map(map);
    if (rc != CTR_OK) return rc;
    note_launch(), vocab_clear_kernel<<<grid1d(map->capacity), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long *>(map->keys),
                                                                               map->rows, map->capacity);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}
