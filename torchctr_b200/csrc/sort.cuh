// Device-wide primitives the backward plan is built from: a stable LSD radix sort of
// (u32 key, u32 value) pairs and a head-flag scan that lists the runs of equal keys.
// All scratch comes from the caller; nothing here allocates or synchronises with the host.
#pragma once
#include "common.cuh"

namespace ctr {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // keys per block per pass
constexpr int kMaxRadixBits = 9;
constexpr int kMaxRadix = 1 << kMaxRadixBits;
constexpr int kMaxPasses = 4;
constexpr int kScanTile = 4096;                        // elements per block of the device scan

inline int64_t sort_num_tiles(int64_t n) { return (n + kSortTile - 1) / kSortTile; }
inline int64_t scan_num_blocks(int64_t m) { return (m + kScanTile - 1) / kScanTile; }

// digit width of the LSD passes: 9 bits when that saves a pass (e.g. 26-bit keys: 3 passes instead of 4), else 8
inline int sort_digit_bits(int key_bits) {
    const int p8 = (key_bits + 7) / 8, p9 = (key_bits + 8) / 9;
    return p9 < p8 ? 9 : 8;
}
inline int sort_num_passes(int key_bits) {
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    const int b = sort_digit_bits(key_bits);
    return (key_bits + b - 1) / b;
}

// scratch sizes, in u32 elements: [tickets 8][global histograms passes x radix][tile status passes x tiles x radix]
inline int64_t sort_counts_elems(int64_t n) {
    return 8 + (int64_t)kMaxPasses * kMaxRadix + (int64_t)kMaxPasses * sort_num_tiles(n) * kMaxRadix;
}
inline int64_t scan_spine_elems(int64_t m) { return scan_num_blocks(m) + 1; }

// Sorts n pairs by the low `key_bits` bits of the key: stable LSD radix sort, one kernel per pass (tile histograms
// are chained with a decoupled look-back instead of a separate device scan), plus one kernel that builds the
// digit histograms of every pass up front.  Ping-pongs between (keys_a, vals_a) and (keys_b, vals_b); returns 0
// when the result ends in the a buffers (sort_num_passes(key_bits) even), 1 when in the b buffers, negative on a
// launch error.  counts: sort_counts_elems(n) u32 of scratch; n < 2^30.
// hist_ready: the caller zeroed the scratch with radix_sort_prepare() and then filled the digit histograms
// (sort_hist(counts), [pass * kMaxRadix + digit]) itself, e.g. inside the kernel that produced the keys.
int radix_sort_pairs(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n,
                     int key_bits, uint32_t *counts, uint32_t *spine, cudaStream_t stream, bool hist_ready = false,
                     const uint32_t *n_dev = nullptr);   // n_dev: device-side pair count (<= n), needs hist_ready
int radix_sort_prepare(uint32_t *counts, int64_t n, int key_bits, cudaStream_t stream);
inline uint32_t *sort_hist(uint32_t *counts) { return counts + 8; }

// ---- segmented variant: one independent sort per table of a launch group --------------------------------------------
// The id slots of a group are laid out table after table, so the pairs are ALREADY partitioned by table: what is left
// is to sort each table's segment by row.  Doing that as independent sorts (instead of one global 26-bit sort) gives
//   * digits sized to the table: ceil(bit_length(V) / passes) bits, 1 pass for V < 512, 3 up to 2^27 rows, 5 beyond
//     (always odd, and a table's passes are the LAST launches, so every table ends in the b buffers);
//   * look-back chains that span one table's tiles (16 for a 65536-bag batch) instead of all ~400 tiles of the launch:
//     the chain, not bandwidth, is what a single-wave onesweep pass costs.
// Keys stay global (row_base + row, or 0xffffffff for padding, which sorts to the end of ITS table); digits are taken
// from the table-local row.
constexpr int kSegMaxPasses = 5;
struct SegSortDesc {
    int32_t num_tables;
    int32_t max_passes;                                // launches: the largest passes[] (odd)
    uint32_t tile_base[CTR_MAX_FEATURES + 1];          // first tile of every table, [num_tables] = total tiles
    uint32_t slot_base[CTR_MAX_FEATURES + 1];          // first slot (sorted position) of every table
    uint32_t row_base[CTR_MAX_FEATURES];
    uint8_t key_bits[CTR_MAX_FEATURES];                // bit_length(num_rows): 2^bits - 1 > every valid local row
    uint8_t digit_bits[CTR_MAX_FEATURES];
    uint8_t passes[CTR_MAX_FEATURES];
};
inline int seg_num_passes(int key_bits) { return key_bits <= kMaxRadixBits ? 1 : (key_bits <= 3 * kMaxRadixBits ? 3 : 5); }
// scratch, in u32 elements: [tickets 8][histograms tables x kSegMaxPasses x kMaxRadix][tile status launches x tiles x kMaxRadix]
inline int64_t seg_hist_elems(int num_tables) { return (int64_t)num_tables * kSegMaxPasses * kMaxRadix; }
inline int64_t seg_counts_elems(int64_t total_tiles, int num_tables) {
    return 8 + seg_hist_elems(num_tables) + (int64_t)kSegMaxPasses * total_tiles * kMaxRadix;
}
inline uint32_t *seg_hist(uint32_t *counts) { return counts + 8; }
// zeroes the scratch (tickets, histograms, the tile status of the launches that will run)
int seg_sort_prepare(uint32_t *counts, const SegSortDesc &sd, cudaStream_t stream);
// histograms must have been filled by the caller ([table][pass][digit], pass 0 = least significant digit); the
// sorted pairs end in (keys_b, vals_b).
int seg_sort_pairs(const SegSortDesc &sd, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b,
                   uint32_t *counts, cudaStream_t stream);
// table-local sort key of a global key
__device__ __forceinline__ uint32_t seg_local_key(uint32_t key, uint32_t row_base, int key_bits) {
    return key == 0xffffffffu ? ((1u << key_bits) - 1u) : key - row_base;
}

// Given sorted keys: run_start[r] = first position of the r-th run of equal keys, run_start[R] = n,
// counters[0] = R (all runs), counters[1] = number of runs whose key != 0xffffffff.
// spine: scan_spine_elems(n) u32.
// run_of_pos (optional, u32 [n]): the run index of every sorted position.
int find_runs(const uint32_t *sorted_keys, int64_t n, uint32_t *run_start, uint32_t *counters, uint32_t *spine,
              cudaStream_t stream, uint32_t *run_of_pos = nullptr);

// Exclusive prefix of `x` over the 256 threads of a block (warp shuffles + one smem hop).
// scratch: 33 u32.  Returns the exclusive prefix; *total receives the block sum.
__device__ __forceinline__ uint32_t block_exclusive_256(uint32_t x, uint32_t *scratch, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = x;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < (int)(blockDim.x >> 5) ? scratch[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_up_sync(kFull, wi, off);
            if (lane >= off) wi += v;
        }
        scratch[lane] = wi - w;
        if (lane == 31) scratch[32] = wi;
    }
    __syncthreads();
    const uint32_t excl = incl - x + scratch[warp];
    *total = scratch[32];
    __syncthreads();
    return excl;
}

// In-place exclusive scan of a u32 array (used on the radix counts; exposed for the tests).
int exclusive_scan_u32(uint32_t *data, int64_t m, uint32_t *spine, cudaStream_t stream);
// Out-of-place variant; the grand total is left in spine[scan_num_blocks(m)].
int exclusive_scan_u32_to(const uint32_t *in, uint32_t *out, int64_t m, uint32_t *spine, cudaStream_t stream);

}  // namespace ctr
