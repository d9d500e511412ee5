// K7: pack / unpack kernels around the NCCL all-to-all of row-sharded tables.
//
// The reference is replicas-only (accelerate/DDP, torchctr/trainer.py:128-130) and would all-reduce
// dense [V, D] gradients.  Here tables are row-sharded: owner(row) = row mod P, and on the owner all
// tables of a group live in ONE fused shard whose row is  base[owner][table] + row div P  ("virtual
// row").  Per step a rank
//   route   maps every id slot of the group to (owner, virtual row), buckets the slots by owner with
//           the stable radix sort (deterministic order), builds the send list of virtual rows and
//           the inverse permutation slot -> position in the exchanged buffer;
//   ...     all-to-all of rows, owner-side gather (K1 on the fused shard), all-to-all of vectors,
//           local pooling (K1 again, over the received vectors) -- issued from Python;
//   grads   gathers grad_out[bag] per sent position for the way back; the owner then runs K2
//           (dedup across all requesters + fused update) on the received (row, gradient) pairs.
#include "sort.cuh"

namespace ctr {

struct RouteLayout {
    int64_t S;
    int key_bits, sorted_in_b;
    int64_t keys_a, keys_b, vals_a, vals_b, counts, spine, vrow, total;
};

static int64_t al256(int64_t x) { return (x + 255) & ~int64_t(255); }

static RouteLayout route_layout(int64_t S, int world) {
    RouteLayout r{};
    r.S = S;
    int bits = 0;
    while ((world >> bits) != 0) ++bits;  // keys are 0 .. world (world = padding bucket)
    r.key_bits = bits < 1 ? 1 : bits;
    r.sorted_in_b = sort_num_passes(r.key_bits) & 1;
    int64_t off = 0;
    r.keys_a = off; off = al256(off + S * 4);
    r.keys_b = off; off = al256(off + S * 4);
    r.vals_a = off; off = al256(off + S * 4);
    r.vals_b = off; off = al256(off + S * 4);
    const int64_t counts = sort_counts_elems(S);
    r.counts = off; off = al256(off + counts * 4);
    r.spine = off; off = al256(off + (scan_spine_elems(counts > S ? counts : S) + 8) * 4);
    r.vrow = off; off = al256(off + S * 4);
    r.total = off;
    return r;
}

// one thread per slot: owner key, virtual row, per-owner counts (warp-aggregated), inv = -1
__global__ void __launch_bounds__(256)
    route_build_kernel(const __grid_constant__ DevGroup g, int world, const int64_t *__restrict__ base,
                       uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, int32_t *__restrict__ vrow,
                       long long *__restrict__ counts, long long *__restrict__ inv) {
    const int fi = blockIdx.y;
    const DevFeature &f = g.f[fi];
    int64_t slot_base = 0;
    for (int i = 0; i < fi; ++i) slot_base += (int64_t)g.B * g.f[i].L;
    const int64_t n = (int64_t)g.B * f.L;
    const int lane = threadIdx.x & 31;
    const int64_t nround = (n + 31) / 32 * 32;  // whole warps iterate together (match.any below)
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += (int64_t)gridDim.x * blockDim.x) {
        uint32_t owner = (uint32_t)world;
        int32_t vr = -1;
        if (j < n) {
            const int32_t row = map_index(f, __ldg(f.ids + j));
            if (row >= 0) {
                owner = (uint32_t)row % (uint32_t)world;
                vr = (int32_t)(base[(int64_t)owner * g.num_features + fi] + (int64_t)((uint32_t)row / (uint32_t)world));
            } else if (row == -2 && g.status != nullptr) {
                atomicOr(g.status, CTR_STATUS_INDEX_OOB);
            }
            keys[slot_base + j] = owner;
            vals[slot_base + j] = (uint32_t)(slot_base + j);
            vrow[slot_base + j] = vr;
            inv[slot_base + j] = -1;
        }
        const uint32_t peers = __match_any_sync(kFull, j < n ? owner : 0xffffffffu);
        if (j < n && lane == __ffs(peers) - 1)
            atomicAdd(reinterpret_cast<unsigned long long *>(counts + owner), (unsigned long long)__popc(peers));
    }
}

// sorted position k -> send_rows[k] = virtual row of that slot, inv[slot] = k   (k < n_valid)
__global__ void __launch_bounds__(256)
    route_finish_kernel(const uint32_t *__restrict__ sorted_slots, const int32_t *__restrict__ vrow, int64_t S,
                        long long *__restrict__ send_rows, long long *__restrict__ inv) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < S; k += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t slot = sorted_slots[k];
        const int32_t vr = vrow[slot];
        if (vr >= 0) {           // valid slots sort before the padding bucket, so k is their exchange position
            send_rows[k] = vr;
            inv[slot] = k;
        }
    }
}

// g_send[k, :] = grad_out[bag(slot_k), cols(feature(slot_k))]: the gradient of the exchanged vector k
__global__ void __launch_bounds__(256)
    slot_grad_gather_kernel(const __grid_constant__ DevGroup g, const uint32_t *__restrict__ sorted_slots, int64_t n,
                            int D, float *__restrict__ g_send) {
    const int pieces = (D % 4 == 0) ? D / 4 : D;
    const int64_t total = n * pieces;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / pieces;
        const int p = (int)(i - k * pieces);
        int64_t slot = sorted_slots[k];
        int fi = 0;
        while (fi + 1 < g.num_features && slot >= (int64_t)g.B * g.f[fi].L) {
            slot -= (int64_t)g.B * g.f[fi].L;
            ++fi;
        }
        const DevFeature &f = g.f[fi];
        const int64_t bag = slot / f.L;
        const float *src = g.out + bag * g.out_stride + f.out_col;
        if (D % 4 == 0) {
            float4 v;
            if (f.aligned) {
                v = __ldg(reinterpret_cast<const float4 *>(src) + p);
            } else {
                v = make_float4(__ldg(src + 4 * p), __ldg(src + 4 * p + 1), __ldg(src + 4 * p + 2), __ldg(src + 4 * p + 3));
            }
            reinterpret_cast<float4 *>(g_send + k * D)[p] = v;
        } else {
            g_send[k * D + p] = __ldg(src + p);
        }
    }
}

}  // namespace ctr

using namespace ctr;

extern "C" int64_t ctr_route_workspace_bytes(const ctr_group_t *group, int32_t world) {
    static thread_local DevGroup dg;
    int rc = lower_group(group, &dg, false, false);
    if (rc != CTR_OK) return rc;
    int64_t S = 0;
    for (int i = 0; i < dg.num_features; ++i) S += (int64_t)dg.B * dg.f[i].L;
    return route_layout(S, world).total;
}

extern "C" int ctr_route_build(const ctr_group_t *group, int32_t world, const int64_t *base, int64_t *counts,
                               int64_t *send_rows, int64_t *inv, void *workspace, int64_t workspace_bytes,
                               void *stream_) {
    static thread_local DevGroup dg;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = lower_group(group, &dg, false, false);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(world >= 1 && world <= 4096, "world=%d outside [1, 4096]", world);
    CTR_REQUIRE(base != nullptr && counts != nullptr && send_rows != nullptr && inv != nullptr && workspace != nullptr,
                "null pointer");
    int64_t S = 0, max_slots = 0;
    for (int i = 0; i < dg.num_features; ++i) {
        const int64_t n = (int64_t)dg.B * dg.f[i].L;
        S += n;
        if (n > max_slots) max_slots = n;
    }
    CTR_REQUIRE(S < (1ll << 31), "group has %lld id slots; must stay below 2^31", (long long)S);
    const RouteLayout r = route_layout(S, world);
    if (workspace_bytes < r.total) {
        set_error("workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)r.total);
        return CTR_E_WORKSPACE;
    }
    CTR_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * (world + 1), stream));
    if (S == 0) return CTR_OK;
    char *ws = static_cast<char *>(workspace);
    uint32_t *keys_a = reinterpret_cast<uint32_t *>(ws + r.keys_a), *keys_b = reinterpret_cast<uint32_t *>(ws + r.keys_b);
    uint32_t *vals_a = reinterpret_cast<uint32_t *>(ws + r.vals_a), *vals_b = reinterpret_cast<uint32_t *>(ws + r.vals_b);
    int32_t *vrow = reinterpret_cast<int32_t *>(ws + r.vrow);
    int64_t bx = (max_slots + 1023) / 1024;
    if (bx > kNumSMs * 8) bx = kNumSMs * 8;
    if (bx < 1) bx = 1;
    note_launch(), route_build_kernel<<<dim3((unsigned)bx, dg.num_features), 256, 0, stream>>>(
        dg, world, base, keys_a, vals_a, vrow, reinterpret_cast<long long *>(counts), reinterpret_cast<long long *>(inv));
    CTR_CUDA_OK(cudaGetLastError());
    rc = radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, S, r.key_bits, reinterpret_cast<uint32_t *>(ws + r.counts),
                          reinterpret_cast<uint32_t *>(ws + r.spine), stream);
    if (rc < 0) return rc;
    const uint32_t *sorted_slots = rc ? vals_b : vals_a;
    int64_t gb = (S + 255) / 256;
    if (gb > kNumSMs * 16) gb = kNumSMs * 16;
    note_launch(), route_finish_kernel<<<(unsigned)gb, 256, 0, stream>>>(sorted_slots, vrow, S,
                                                                        reinterpret_cast<long long *>(send_rows),
                                                                        reinterpret_cast<long long *>(inv));
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

extern "C" int ctr_route_grad_gather(const ctr_group_t *group, int32_t world, const void *workspace, int64_t n,
                                     int32_t D, float *g_send, void *stream_) {
    static thread_local DevGroup dg;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = lower_group(group, &dg, false, true);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(n >= 0 && D >= 1, "bad sizes");
    if (n == 0) return CTR_OK;
    CTR_REQUIRE(workspace != nullptr && g_send != nullptr, "null pointer");
    int64_t S = 0;
    for (int i = 0; i < dg.num_features; ++i) {
        S += (int64_t)dg.B * dg.f[i].L;
        CTR_REQUIRE(dg.f[i].D == D, "feature %d: D=%d differs from the exchanged width %d", i, dg.f[i].D, D);
    }
    CTR_REQUIRE(n <= S, "n exceeds the slot count");
    const RouteLayout r = route_layout(S, world);
    const char *ws = static_cast<const char *>(workspace);
    const uint32_t *sorted_slots = reinterpret_cast<const uint32_t *>(ws + (r.sorted_in_b ? r.vals_b : r.vals_a));
    const int pieces = (D % 4 == 0) ? D / 4 : D;
    int64_t gb = (n * pieces + 255) / 256;
    if (gb > kNumSMs * 16) gb = kNumSMs * 16;
    note_launch(), slot_grad_gather_kernel<<<(unsigned)gb, 256, 0, stream>>>(dg, sorted_slots, n, D, g_send);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}

// ---- peer-memory variant: routing lists the OWNERS read over NVLink (no all-to-all) -----------------------------
namespace ctr {

// one thread per slot: owner, virtual row on the owner, per-owner counts (warp-aggregated)
__global__ void __launch_bounds__(256)
    route_p2p_build_kernel(const __grid_constant__ DevGroup g, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
                           uint32_t *__restrict__ vrow, uint32_t *__restrict__ counts) {
    const int fi = blockIdx.y;
    const DevFeature &f = g.f[fi];
    int64_t slot_base = 0;
    for (int i = 0; i < fi; ++i) slot_base += (int64_t)g.B * g.f[i].L;
    const int64_t n = (int64_t)g.B * f.L;
    const int lane = threadIdx.x & 31;
    const int64_t nround = (n + 31) / 32 * 32;  // whole warps iterate together (match.any below)
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += (int64_t)gridDim.x * blockDim.x) {
        uint32_t owner = (uint32_t)g.world;       // padding / invalid ids sort behind every rank
        if (j < n) {
            const int32_t row = map_index(f, __ldg(f.ids + j));
            uint32_t vr = 0xffffffffu;
            if (row >= 0) {
                vr = (uint32_t)shard_vrow(g, fi, (uint32_t)row, &owner);
            } else if (row == -2 && g.status != nullptr) {
                atomicOr(g.status, CTR_STATUS_INDEX_OOB);
            }
            keys[slot_base + j] = owner;
            vals[slot_base + j] = (uint32_t)(slot_base + j);
            vrow[slot_base + j] = vr;
        }
        const uint32_t peers = __match_any_sync(kFull, j < n ? owner : 0xffffffffu);
        if (j < n && lane == __ffs(peers) - 1) atomicAdd(counts + owner, (uint32_t)__popc(peers));
    }
}

// owner-major position k -> keys_out[k] = virtual row, slots_out[k] = slot inside its feature (bag * L + l)
__global__ void __launch_bounds__(256)
    route_p2p_finish_kernel(const __grid_constant__ DevGroup g, const uint32_t *__restrict__ sorted_slots,
                            const uint32_t *__restrict__ vrow, int64_t S, uint32_t *__restrict__ keys_out,
                            uint32_t *__restrict__ slots_out) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < S; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t slot = sorted_slots[k];
        const uint32_t vr = vrow[slot];
        int fi = 0;
        while (fi + 1 < g.num_features && slot >= (int64_t)g.B * g.f[fi].L) {
            slot -= (int64_t)g.B * g.f[fi].L;
            ++fi;
        }
        keys_out[k] = vr;                 // 0xffffffff for the padding bucket, which no owner reads
        slots_out[k] = (uint32_t)slot;
    }
}

}  // namespace ctr

extern "C" int64_t ctr_route_p2p_workspace_bytes(const ctr_group_t *group) {
    static thread_local DevGroup dg;
    int rc = lower_group(group, &dg, false, false);
    if (rc != CTR_OK) return rc;
    int64_t S = 0;
    for (int i = 0; i < dg.num_features; ++i) S += (int64_t)dg.B * dg.f[i].L;
    return route_layout(S, CTR_MAX_WORLD).total;
}

extern "C" int ctr_route_p2p_build(const ctr_group_t *group, const ctr_shard_t *shard, uint32_t *counts, uint32_t *keys,
                                   uint32_t *slots, void *workspace, int64_t workspace_bytes, void *stream_) {
    static thread_local DevGroup dg;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = lower_group(group, &dg, false, false);
    if (rc != CTR_OK) return rc;
    rc = attach_shard(&dg, shard, nullptr);
    if (rc != CTR_OK) return rc;
    CTR_REQUIRE(counts != nullptr && workspace != nullptr, "null pointer");
    int64_t S = 0, max_slots = 0;
    for (int i = 0; i < dg.num_features; ++i) {
        const int64_t n = (int64_t)dg.B * dg.f[i].L;
        CTR_REQUIRE(n < (1ll << 28), "feature %d: B*L must stay below 2^28 on the sharded path", i);
        S += n;
        if (n > max_slots) max_slots = n;
    }
    CTR_REQUIRE(S < (1ll << 30), "group has %lld id slots; must stay below 2^30", (long long)S);
    const RouteLayout r = route_layout(S, CTR_MAX_WORLD);   // key bits for any world <= CTR_MAX_WORLD
    if (workspace_bytes < r.total) {
        set_error("workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)r.total);
        return CTR_E_WORKSPACE;
    }
    CTR_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (CTR_MAX_WORLD + 1), stream));
    if (S == 0) return CTR_OK;
    CTR_REQUIRE(keys != nullptr && slots != nullptr, "null pointer");
    char *ws = static_cast<char *>(workspace);
    uint32_t *keys_a = reinterpret_cast<uint32_t *>(ws + r.keys_a), *keys_b = reinterpret_cast<uint32_t *>(ws + r.keys_b);
    uint32_t *vals_a = reinterpret_cast<uint32_t *>(ws + r.vals_a), *vals_b = reinterpret_cast<uint32_t *>(ws + r.vals_b);
    uint32_t *vrow = reinterpret_cast<uint32_t *>(ws + r.vrow);
    int64_t bx = (max_slots + 1023) / 1024;
    const int64_t per_feature = ((int64_t)kNumSMs * 4 + dg.num_features - 1) / dg.num_features;
    if (bx > per_feature) bx = per_feature;
    if (bx < 1) bx = 1;
    note_launch(), route_p2p_build_kernel<<<dim3((unsigned)bx, dg.num_features), 256, 0, stream>>>(dg, keys_a, vals_a, vrow, counts);
    CTR_CUDA_OK(cudaGetLastError());
    rc = radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, S, r.key_bits, reinterpret_cast<uint32_t *>(ws + r.counts),
                          reinterpret_cast<uint32_t *>(ws + r.spine), stream);
    if (rc < 0) return rc;
    const uint32_t *sorted_slots = rc ? vals_b : vals_a;
    int64_t gb = (S + 255) / 256;
    if (gb > kNumSMs * 16) gb = kNumSMs * 16;
    note_launch(), route_p2p_finish_kernel<<<(unsigned)gb, 256, 0, stream>>>(dg, sorted_slots, vrow, S, keys, slots);
    CTR_CUDA_OK(cudaGetLastError());
    return CTR_OK;
}
