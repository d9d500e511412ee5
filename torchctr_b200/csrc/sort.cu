// Stable LSD radix sort of (u32 key, u32 value) pairs + run detection, hand-written for the
// backward plan (K2).  8- or 9-bit digits; the digit histograms of all passes come from one read
// of the keys, then each pass is ONE kernel: tile ranking is warp-synchronous (match.any, no
// per-key atomics) and the per-digit tile offsets are chained with a decoupled look-back.
#include "sort.cuh"

namespace ctr {

// ---- block scan -----------------------------------------------------------------------

struct ArrayIn {
    const uint32_t *p;
    __device__ __forceinline__ uint32_t operator()(int64_t i) const { return p[i]; }
};
struct HeadIn {  // 1 where a run of equal VALID keys starts (padding, 0xffffffff, may sit at the end of every table)
    const uint32_t *keys;
    __device__ __forceinline__ uint32_t operator()(int64_t i) const {
        const uint32_t k = keys[i];
        return (k != 0xffffffffu && (i == 0 || k != keys[i - 1])) ? 1u : 0u;
    }
};
struct ArrayOut {
    uint32_t *p;
    __device__ __forceinline__ void operator()(int64_t i, uint32_t prefix, uint32_t) const { p[i] = prefix; }
    __device__ __forceinline__ void finish(uint32_t) const {}
};
struct RunsOut {
    uint32_t *run_start;
    uint32_t *counters;
    const uint32_t *keys;
    int64_t n;
    uint32_t *run_of_pos;   // optional: index of the run every position belongs to
    __device__ __forceinline__ void operator()(int64_t i, uint32_t prefix, uint32_t head) const {
        if (head) run_start[prefix] = (uint32_t)i;
        if (run_of_pos != nullptr) run_of_pos[i] = prefix + head - 1u;
    }
    __device__ __forceinline__ void finish(uint32_t total) const {
        run_start[total] = (uint32_t)n;
        counters[0] = total;
        counters[1] = total;
    }
};

constexpr int kScanItems = kScanTile / 256;

template <class In>
__global__ void __launch_bounds__(256) scan_reduce_kernel(In in, int64_t m, uint32_t *spine) {
    __shared__ uint32_t scratch[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
        if (base + j < m) s += in(base + j);
    uint32_t total;
    block_exclusive_256(s, scratch, &total);
    if (threadIdx.x == 0) spine[blockIdx.x] = total;
}

// one block: exclusive scan of spine[0..nblk), total to spine[nblk]
__global__ void __launch_bounds__(256) scan_spine_kernel(uint32_t *spine, int64_t nblk) {
    __shared__ uint32_t scratch[33];
    uint32_t carry = 0;
    for (int64_t base = 0; base < nblk; base += 256) {
        const int64_t i = base + threadIdx.x;
        const uint32_t x = i < nblk ? spine[i] : 0u;
        uint32_t total;
        const uint32_t excl = block_exclusive_256(x, scratch, &total);
        if (i < nblk) spine[i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0) spine[nblk] = carry;
}

template <class In, class Out>
__global__ void __launch_bounds__(256) scan_apply_kernel(In in, int64_t m, const uint32_t *spine, int64_t nblk, Out out) {
    __shared__ uint32_t scratch[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t x[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        x[j] = base + j < m ? in(base + j) : 0u;
        s += x[j];
    }
    uint32_t total;
    uint32_t prefix = block_exclusive_256(s, scratch, &total) + spine[blockIdx.x];
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        if (base + j < m) out(base + j, prefix, x[j]);
        prefix += x[j];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out.finish(spine[nblk]);
}

template <class In, class Out>
static int device_scan(In in, Out out, int64_t m, uint32_t *spine, cudaStream_t stream) {
    const int64_t nblk = scan_num_blocks(m);
    note_launch(), scan_reduce_kernel<In><<<(unsigned)nblk, 256, 0, stream>>>(in, m, spine);
    note_launch(), scan_spine_kernel<<<1, 256, 0, stream>>>(spine, nblk);
    note_launch(), scan_apply_kernel<In, Out><<<(unsigned)nblk, 256, 0, stream>>>(in, m, spine, nblk, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "device_scan");
    return CTR_OK;
}

int exclusive_scan_u32(uint32_t *data, int64_t m, uint32_t *spine, cudaStream_t stream) {
    if (m <= 0) return CTR_OK;
    return device_scan(ArrayIn{data}, ArrayOut{data}, m, spine, stream);
}

int exclusive_scan_u32_to(const uint32_t *in, uint32_t *out, int64_t m, uint32_t *spine, cudaStream_t stream) {
    if (m <= 0) return CTR_OK;
    return device_scan(ArrayIn{in}, ArrayOut{out}, m, spine, stream);
}

__global__ void empty_runs_kernel(uint32_t *run_start, uint32_t *counters) {
    run_start[0] = 0;
    counters[0] = 0;
    counters[1] = 0;
}

int find_runs(const uint32_t *sorted_keys, int64_t n, uint32_t *run_start, uint32_t *counters, uint32_t *spine,
              cudaStream_t stream, uint32_t *run_of_pos) {
    if (n <= 0) {
        note_launch(), empty_runs_kernel<<<1, 1, 0, stream>>>(run_start, counters);
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? CTR_OK : cuda_fail(e, "empty_runs_kernel");
    }
    return device_scan(HeadIn{sorted_keys}, RunsOut{run_start, counters, sorted_keys, n, run_of_pos}, n, spine, stream);
}

// ---- radix passes -----------------------------------------------------------------------

// Digit histograms of every pass in one read of the keys: hist[pass][digit] += ...
template <int BITS>
__global__ void __launch_bounds__(kSortThreads)
    radix_hist_all_kernel(const uint32_t *__restrict__ keys, int64_t n, int passes, uint32_t *__restrict__ hist) {
    constexpr int RADIX = 1 << BITS;
    __shared__ uint32_t sh[kMaxPasses][RADIX];
    for (int i = threadIdx.x; i < kMaxPasses * RADIX; i += kSortThreads) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * kSortThreads; base < n; base += (int64_t)gridDim.x * kSortThreads) {
        const int64_t i = base + threadIdx.x;
        const bool valid = i < n;
        const uint32_t k = valid ? keys[i] : 0u;
        for (int p = 0; p < passes; ++p) {
            const uint32_t d = (k >> (p * BITS)) & (RADIX - 1);
            const uint32_t peers = __match_any_sync(kFull, valid ? d : (uint32_t)(RADIX + lane));
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&sh[p][d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += kSortThreads) {
        const uint32_t c = sh[i / RADIX][i % RADIX];
        if (c) atomicAdd(&hist[(i / RADIX) * kMaxRadix + (i % RADIX)], c);
    }
}

constexpr uint32_t kFlagAgg = 1u << 30;   // tile status: count of this tile only
constexpr uint32_t kFlagInc = 1u << 31;   // tile status: count of this tile and every tile before it
constexpr uint32_t kValMask = kFlagAgg - 1u;

__device__ __forceinline__ uint32_t ld_status(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One LSD pass in one kernel.  Tiles are taken in ticket order, so a tile only ever waits for tiles that
// started before it.  Ranking inside a tile is warp-synchronous: lanes that hold the same digit find each
// other with match.any and the lowest one bumps the warp's shared-memory counter for all of them (no per-key
// atomics); warp w owns a contiguous slice of the tile, so (warp, round, lane) order is memory order and the
// sort is stable.  Per digit, the tile publishes its count, looks back over earlier tiles until it meets an
// inclusive prefix, and publishes its own inclusive prefix.
template <int BITS>
__global__ void __launch_bounds__(kSortThreads, 3)
    radix_onesweep_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                          uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n_cap,
                          const uint32_t *__restrict__ n_dev, int shift, const uint32_t *__restrict__ ghist,
                          uint32_t *status, uint32_t *ticket) {
    constexpr int RADIX = 1 << BITS;
    constexpr int DPT = RADIX / kSortThreads;   // digits per thread
    constexpr int kWarps = kSortThreads / kWarp;
    __shared__ uint32_t wh[kWarps][RADIX];
    __shared__ uint32_t base[RADIX];
    __shared__ uint32_t scratch[33];
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < kWarps * RADIX; i += kSortThreads) (&wh[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    // the pair count may live on the device (owner side of the sharded backward): tiles past it have nothing to do
    const int64_t n = n_dev != nullptr ? ((int64_t)*n_dev < n_cap ? (int64_t)*n_dev : n_cap) : n_cap;
    if ((int64_t)tile * kSortTile >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp_base = (int64_t)tile * kSortTile + (int64_t)warp * (kWarp * kSortItems);
    uint32_t k[kSortItems], v[kSortItems];
    uint16_t rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t i = warp_base + r * kWarp + lane;
        k[r] = i < n ? keys_in[i] : 0xffffffffu;
        v[r] = i < n ? vals_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const bool valid = warp_base + r * kWarp + lane < n;
        const uint32_t d = (k[r] >> shift) & (RADIX - 1);
        const uint32_t peers = __match_any_sync(kFull, valid ? d : (uint32_t)(RADIX + lane));
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && lane == leader) {
            pre = wh[warp][d];
            wh[warp][d] = pre + (uint32_t)__popc(peers);
        }
        pre = __shfl_sync(kFull, pre, leader);
        rank[r] = (uint16_t)(pre + (uint32_t)__popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    __syncthreads();
    // thread t owns digits t * DPT .. t * DPT + DPT - 1
    uint32_t gsum = 0, gh[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        gh[j] = ghist[threadIdx.x * DPT + j];
        gsum += gh[j];
    }
    uint32_t total;
    uint32_t gbase = block_exclusive_256(gsum, scratch, &total);   // first output position of the thread's digits
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const int d = threadIdx.x * DPT + j;
        uint32_t count = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = wh[w][d];
            wh[w][d] = count;          // exclusive over the warps of this tile
            count += c;
        }
        uint32_t *mine = status + (size_t)tile * RADIX + d;
        uint32_t prefix = 0;
        if (tile == 0) {
            st_status(mine, count | kFlagInc);
        } else {
            st_status(mine, count | kFlagAgg);
            // look back over earlier tiles, kLook of them per round trip, until an inclusive prefix turns up
            constexpr int kLook = 8;
            int64_t prev = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t s[kLook];
#pragma unroll
                for (int i = 0; i < kLook; ++i)
                    s[i] = prev - i >= 0 ? ld_status(status + (size_t)(prev - i) * RADIX + d) : kFlagInc;
#pragma unroll
                for (int i = 0; i < kLook; ++i) {
                    if (!done) {
                        while ((s[i] & (kFlagAgg | kFlagInc)) == 0u) s[i] = ld_status(status + (size_t)(prev - i) * RADIX + d);
                        prefix += s[i] & kValMask;
                        done = (s[i] & kFlagInc) != 0u;
                    }
                }
                prev -= kLook;
            }
            st_status(mine, (prefix + count) | kFlagInc);
        }
        base[d] = gbase + prefix;
        gbase += gh[j];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        if (warp_base + r * kWarp + lane < n) {
            const uint32_t d = (k[r] >> shift) & (RADIX - 1);
            const uint32_t dst = base[d] + wh[warp][d] + rank[r];
            keys_out[dst] = k[r];
            vals_out[dst] = v[r];
        }
    }
}

// One launch of the segmented sort (see sort.cuh): tile -> table by the static tile layout, digits from the table-local
// row, look-back only over the tiles of the same table.  Tables whose passes have not started yet leave at once.
__global__ void __launch_bounds__(kSortThreads, 3)
    radix_seg_kernel(const __grid_constant__ SegSortDesc sd, const uint32_t *__restrict__ keys_in,
                     const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                     int launch, const uint32_t *__restrict__ hist, uint32_t *status, uint32_t *ticket) {
    constexpr int RADIX = kMaxRadix;
    constexpr int DPT = RADIX / kSortThreads;   // digits per thread
    constexpr int kWarps = kSortThreads / kWarp;
    __shared__ uint32_t wh[kWarps][RADIX];
    __shared__ uint32_t base[RADIX];
    __shared__ uint32_t scratch[33];
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    int f = 0;
    {
        int lo = 0, hi = sd.num_tables - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (sd.tile_base[mid] <= tile) lo = mid; else hi = mid - 1;
        }
        f = lo;
    }
    const int first_launch = sd.max_passes - (int)sd.passes[f];
    if (launch < first_launch) return;                       // block-uniform
    const int pass = launch - first_launch;
    const int w = sd.digit_bits[f], kb = sd.key_bits[f];
    const int shift = pass * w;
    const uint32_t dmask = (1u << w) - 1u;
    const uint32_t rbase = sd.row_base[f];
    const uint32_t tile0 = sd.tile_base[f];
    const uint32_t seg_begin = sd.slot_base[f], seg_end = sd.slot_base[f + 1];
    for (int i = threadIdx.x; i < kWarps * RADIX; i += kSortThreads) (&wh[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warp_base = seg_begin + (tile - tile0) * (uint32_t)kSortTile + (uint32_t)warp * (kWarp * kSortItems);
    uint32_t k[kSortItems], v[kSortItems];
    uint16_t rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t i = warp_base + r * kWarp + lane;
        k[r] = i < seg_end ? keys_in[i] : 0xffffffffu;
        v[r] = i < seg_end ? vals_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const bool valid = warp_base + r * kWarp + lane < seg_end;
        const uint32_t d = (seg_local_key(k[r], rbase, kb) >> shift) & dmask;
        const uint32_t peers = __match_any_sync(kFull, valid ? d : (uint32_t)(RADIX + lane));
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && lane == leader) {
            pre = wh[warp][d];
            wh[warp][d] = pre + (uint32_t)__popc(peers);
        }
        pre = __shfl_sync(kFull, pre, leader);
        rank[r] = (uint16_t)(pre + (uint32_t)__popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    __syncthreads();
    const uint32_t *ghist = hist + ((size_t)f * kSegMaxPasses + pass) * RADIX;
    uint32_t *st = status + (size_t)launch * sd.tile_base[sd.num_tables] * RADIX;
    uint32_t gsum = 0, gh[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        gh[j] = ghist[threadIdx.x * DPT + j];
        gsum += gh[j];
    }
    uint32_t total;
    uint32_t gbase = seg_begin + block_exclusive_256(gsum, scratch, &total);
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const int d = threadIdx.x * DPT + j;
        uint32_t count = 0;
#pragma unroll
        for (int ww = 0; ww < kWarps; ++ww) {
            const uint32_t c = wh[ww][d];
            wh[ww][d] = count;          // exclusive over the warps of this tile
            count += c;
        }
        uint32_t prefix = 0;
        if ((uint32_t)d <= dmask) {     // digits this table does not have are never looked up
            uint32_t *mine = st + (size_t)tile * RADIX + d;
            if (tile == tile0) {
                st_status(mine, count | kFlagInc);
            } else {
                st_status(mine, count | kFlagAgg);
                constexpr int kLook = 8;
                int64_t prev = (int64_t)tile - 1;
                bool done = false;
                while (!done) {
                    uint32_t sv[kLook];
#pragma unroll
                    for (int i = 0; i < kLook; ++i)
                        sv[i] = prev - i >= (int64_t)tile0 ? ld_status(st + (size_t)(prev - i) * RADIX + d) : kFlagInc;
#pragma unroll
                    for (int i = 0; i < kLook; ++i) {
                        if (!done) {
                            while ((sv[i] & (kFlagAgg | kFlagInc)) == 0u) sv[i] = ld_status(st + (size_t)(prev - i) * RADIX + d);
                            prefix += sv[i] & kValMask;
                            done = (sv[i] & kFlagInc) != 0u;
                        }
                    }
                    prev -= kLook;
                }
                st_status(mine, (prefix + count) | kFlagInc);
            }
        }
        base[d] = gbase + prefix;
        gbase += gh[j];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        if (warp_base + r * kWarp + lane < seg_end) {
            const uint32_t d = (seg_local_key(k[r], rbase, kb) >> shift) & dmask;
            const uint32_t dst = base[d] + wh[warp][d] + rank[r];
            keys_out[dst] = k[r];
            vals_out[dst] = v[r];
        }
    }
}

int seg_sort_prepare(uint32_t *counts, const SegSortDesc &sd, cudaStream_t stream) {
    const size_t total_tiles = sd.tile_base[sd.num_tables];
    const size_t zero_elems = 8 + (size_t)seg_hist_elems(sd.num_tables) + (size_t)sd.max_passes * total_tiles * kMaxRadix;
    cudaError_t e = cudaMemsetAsync(counts, 0, zero_elems * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return cuda_fail(e, "segmented sort scratch memset");
    return CTR_OK;
}

int seg_sort_pairs(const SegSortDesc &sd, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b,
                   uint32_t *counts, cudaStream_t stream) {
    const uint32_t total_tiles = sd.tile_base[sd.num_tables];
    if (total_tiles == 0) return CTR_OK;
    uint32_t *tickets = counts;
    const uint32_t *hist = seg_hist(counts);
    uint32_t *status = counts + 8 + seg_hist_elems(sd.num_tables);
    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    for (int l = 0; l < sd.max_passes; ++l) {
        note_launch(), radix_seg_kernel<<<total_tiles, kSortThreads, 0, stream>>>(sd, kin, vin, kout, vout, l, hist, status,
                                                                                  tickets + l);
        uint32_t *t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "seg_sort_pairs");
    return CTR_OK;
}

template <int BITS>
static int radix_sort_impl(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n, int passes,
                           uint32_t *scratch, cudaStream_t stream, bool hist_ready, const uint32_t *n_dev) {
    constexpr int RADIX = 1 << BITS;
    const int64_t ntiles = sort_num_tiles(n);
    uint32_t *tickets = scratch;
    uint32_t *hist = scratch + 8;
    uint32_t *status = hist + kMaxPasses * kMaxRadix;
    cudaError_t e = cudaSuccess;
    if (!hist_ready) {
        if (n_dev != nullptr) {
            set_error("radix_sort_pairs: a device-side count needs the histograms from the caller");
            return CTR_E_BADARG;
        }
        int rc = radix_sort_prepare(scratch, n, passes * BITS, stream);
        if (rc != CTR_OK) return rc;
        int64_t hb = (n + kSortThreads * 8 - 1) / (kSortThreads * 8);
        if (hb > kNumSMs * 4) hb = kNumSMs * 4;
        note_launch(), radix_hist_all_kernel<BITS><<<(unsigned)hb, kSortThreads, 0, stream>>>(keys_a, n, passes, hist);
    }
    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    for (int p = 0; p < passes; ++p) {
        note_launch(), radix_onesweep_kernel<BITS><<<(unsigned)ntiles, kSortThreads, 0, stream>>>(
            kin, vin, kout, vout, n, n_dev, p * BITS, hist + p * kMaxRadix, status + (size_t)p * ntiles * RADIX, tickets + p);
        uint32_t *t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "radix_sort_pairs");
    return passes & 1;
}

int radix_sort_prepare(uint32_t *scratch, int64_t n, int key_bits, cudaStream_t stream) {
    if (n <= 0) return CTR_OK;
    const int passes = sort_num_passes(key_bits);
    const size_t zero_elems = 8 + (size_t)kMaxPasses * kMaxRadix + (size_t)passes * sort_num_tiles(n) * (1u << sort_digit_bits(key_bits));
    cudaError_t e = cudaMemsetAsync(scratch, 0, zero_elems * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return cuda_fail(e, "radix sort scratch memset");
    return CTR_OK;
}

int radix_sort_pairs(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n, int key_bits,
                     uint32_t *counts, uint32_t *spine, cudaStream_t stream, bool hist_ready, const uint32_t *n_dev) {
    (void)spine;
    if (n <= 0) return 0;
    if (n >= (1ll << 30)) {
        set_error("radix_sort_pairs: n=%lld must stay below 2^30", (long long)n);
        return CTR_E_BADARG;
    }
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    const int passes = sort_num_passes(key_bits);
    if (sort_digit_bits(key_bits) == 9) return radix_sort_impl<9>(keys_a, vals_a, keys_b, vals_b, n, passes, counts, stream, hist_ready, n_dev);
    return radix_sort_impl<8>(keys_a, vals_a, keys_b, vals_b, n, passes, counts, stream, hist_ready, n_dev);
}

}  // namespace ctr
