// Device-wide primitives the backward plan is built from: a stable LSD radix sort of
// (u32 key, u32 value) pairs and a head-flag scan that lists the runs of equal keys.
// All scratch comes from the caller; nothing here allocates or synchronises with the host.
#pragma once
#include "common.cuh"

namespace ctr {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // keys per block per pass
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kScanTile = 4096;                        // elements per block of the device scan

inline int64_t sort_num_tiles(int64_t n) { return (n + kSortTile - 1) / kSortTile; }
inline int64_t scan_num_blocks(int64_t m) { return (m + kScanTile - 1) / kScanTile; }

// scratch sizes, in u32 elements
inline int64_t sort_counts_elems(int64_t n) { return (int64_t)kRadix * sort_num_tiles(n); }
inline int64_t scan_spine_elems(int64_t m) { return scan_num_blocks(m) + 1; }

// Sorts n pairs by the low `key_bits` bits of the key.  Ping-pongs between (keys_a, vals_a) and
// (keys_b, vals_b); returns 0 when the result ends in the a buffers, 1 when in the b buffers,
// negative on a launch error.  counts: sort_counts_elems(n) u32; spine: scan_spine_elems(counts) u32.
int radix_sort_pairs(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n,
                     int key_bits, uint32_t *counts, uint32_t *spine, cudaStream_t stream);

// Given sorted keys: run_start[r] = first position of the r-th run of equal keys, run_start[R] = n,
// counters[0] = R (all runs), counters[1] = number of runs whose key != 0xffffffff.
// spine: scan_spine_elems(n) u32.
int find_runs(const uint32_t *sorted_keys, int64_t n, uint32_t *run_start, uint32_t *counters, uint32_t *spine,
              cudaStream_t stream);

// In-place exclusive scan of a u32 array (used on the radix counts; exposed for the tests).
int exclusive_scan_u32(uint32_t *data, int64_t m, uint32_t *spine, cudaStream_t stream);
// Out-of-place variant; the grand total is left in spine[scan_num_blocks(m)].
int exclusive_scan_u32_to(const uint32_t *in, uint32_t *out, int64_t m, uint32_t *spine, cudaStream_t stream);

}  // namespace ctr
