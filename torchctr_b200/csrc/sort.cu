// Stable LSD radix sort of (u32 key, u32 value) pairs + run detection, hand-written for the
// backward plan (K2).  8-bit digits; per pass: per-tile digit histogram -> device scan of the
// digit-major count matrix -> stable scatter.  Ranking inside a tile is warp-synchronous:
// lanes that hold the same digit find each other with match.any, the lowest one bumps the
// warp's shared-memory counter for all of them, so no per-key atomics and no sorting network.
#include "sort.cuh"

namespace ctr {

// ---- block scan -----------------------------------------------------------------------

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

struct ArrayIn {
    const uint32_t *p;
    __device__ __forceinline__ uint32_t operator()(int64_t i) const { return p[i]; }
};
struct HeadIn {  // 1 where a run of equal keys starts
    const uint32_t *keys;
    __device__ __forceinline__ uint32_t operator()(int64_t i) const {
        return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
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
    __device__ __forceinline__ void operator()(int64_t i, uint32_t prefix, uint32_t head) const {
        if (head) run_start[prefix] = (uint32_t)i;
    }
    __device__ __forceinline__ void finish(uint32_t total) const {
        run_start[total] = (uint32_t)n;
        counters[0] = total;
        counters[1] = total - ((n > 0 && keys[n - 1] == 0xffffffffu) ? 1u : 0u);
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
              cudaStream_t stream) {
    if (n <= 0) {
        note_launch(), empty_runs_kernel<<<1, 1, 0, stream>>>(run_start, counters);
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? CTR_OK : cuda_fail(e, "empty_runs_kernel");
    }
    return device_scan(HeadIn{sorted_keys}, RunsOut{run_start, counters, sorted_keys, n}, n, spine, stream);
}

// ---- radix passes -----------------------------------------------------------------------

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, int shift,
                                                                uint32_t *__restrict__ counts, int64_t ntiles) {
    __shared__ uint32_t hist[kRadix];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int j = 0; j < kSortItems; ++j) {
        const int64_t i = base + (int64_t)j * kSortThreads + threadIdx.x;
        const bool valid = i < n;
        const uint32_t d = valid ? ((keys[i] >> shift) & (kRadix - 1)) : (uint32_t)(kRadix + lane);
        const uint32_t peers = __match_any_sync(kFull, d);
        if (valid && lane == __ffs(peers) - 1) atomicAdd(&hist[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    counts[(int64_t)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads)
    radix_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                         uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, int shift,
                         const uint32_t *__restrict__ offsets, int64_t ntiles) {
    constexpr int kWarps = kSortThreads / kWarp;
    __shared__ uint32_t wh[kWarps][kRadix];
    for (int i = threadIdx.x; i < kWarps * kRadix; i += kSortThreads) (&wh[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // warp w owns the contiguous slice [w * 32 * items, (w + 1) * 32 * items) of the tile so that
    // (warp, round, lane) order is memory order: the sort stays stable.
    const int64_t warp_base = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (kWarp * kSortItems);
    uint32_t k[kSortItems], v[kSortItems], rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t i = warp_base + r * kWarp + lane;
        k[r] = i < n ? keys_in[i] : 0xffffffffu;
        v[r] = i < n ? vals_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const bool valid = warp_base + r * kWarp + lane < n;
        const uint32_t d = (k[r] >> shift) & (kRadix - 1);
        const uint32_t peers = __match_any_sync(kFull, valid ? d : (uint32_t)(kRadix + lane));
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && lane == leader) {
            pre = wh[warp][d];
            wh[warp][d] = pre + (uint32_t)__popc(peers);
        }
        pre = __shfl_sync(kFull, pre, leader);
        rank[r] = pre + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = threadIdx.x;  // one digit per thread (kSortThreads == kRadix)
        uint32_t run = offsets[(int64_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = wh[w][d];
            wh[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        if (warp_base + r * kWarp + lane < n) {
            const uint32_t d = (k[r] >> shift) & (kRadix - 1);
            const uint32_t dst = wh[warp][d] + rank[r];
            keys_out[dst] = k[r];
            vals_out[dst] = v[r];
        }
    }
}

static_assert(kSortThreads == kRadix, "scatter kernel maps one digit per thread");

int radix_sort_pairs(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n, int key_bits,
                     uint32_t *counts, uint32_t *spine, cudaStream_t stream) {
    if (n <= 0) return 0;
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    const int passes = (key_bits + kRadixBits - 1) / kRadixBits;
    const int64_t ntiles = sort_num_tiles(n);
    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * kRadixBits;
        note_launch(), radix_hist_kernel<<<(unsigned)ntiles, kSortThreads, 0, stream>>>(kin, n, shift, counts, ntiles);
        int rc = exclusive_scan_u32(counts, (int64_t)kRadix * ntiles, spine, stream);
        if (rc != CTR_OK) return rc;
        note_launch(), radix_scatter_kernel<<<(unsigned)ntiles, kSortThreads, 0, stream>>>(kin, vin, kout, vout, n, shift, counts, ntiles);
        uint32_t *t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "radix_sort_pairs");
    return passes & 1;
}

}  // namespace ctr
