// Micro-benchmark: how many random row reads (and read-modify-writes) per second does a B200 sustain from HBM,
// as a function of the row size?  Sets the real ceiling for the embedding lookup / update kernels, whose rows
// are 64 B (D = 16) .. 256 B (D = 64).   nvcc -O3 -arch=sm_100a random_access.cu -o random_access
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31; return x;
}

// each group of G lanes reads one random row of G float4; K independent rows in flight per lane
template <int G, int K, bool RMW>
__global__ void gather_kernel(float4 *table, uint64_t rows, uint64_t n, float4 *sink, uint64_t seed) {
    const uint64_t team = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) / G;
    const int t = threadIdx.x % G;
    const uint64_t teams = (uint64_t)gridDim.x * blockDim.x / G;
    float4 acc = make_float4(0, 0, 0, 0);
    for (uint64_t i = team * K; i < n; i += teams * K) {
        float4 v[K];
        uint64_t r[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            r[k] = mix64(seed + i + k) % rows;
            v[k] = table[r[k] * G + t];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (RMW) {
                v[k].x += 1.f;
                table[r[k] * G + t] = v[k];
            } else {
                acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w;
            }
        }
    }
    if (acc.x == 123.456f) sink[0] = acc;
}

template <int G, int K, bool RMW>
void run(float4 *table, uint64_t bytes, uint64_t n, float4 *sink, const char *name) {
    const uint64_t rows = bytes / (16 * G);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        gather_kernel<G, K, RMW><<<148 * 8, 256>>>(table, rows, n, sink, 1234567ull * (it + 1));
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
    }
    printf("%-28s row %4d B  n %8llu  %8.1f us  %7.2f G rows/s  %7.1f GB/s%s\n", name, 16 * G, (unsigned long long)n,
           best * 1e3, n / (best * 1e-3) / 1e9, n * 16.0 * G * (RMW ? 2 : 1) / (best * 1e-3) / 1e9, RMW ? " (r+w)" : "");
}

int main() {
    const uint64_t bytes = 8ull << 30;   // 8 GiB table: far beyond L2 and the TLB reach
    float4 *table, *sink;
    cudaMalloc(&table, bytes); cudaMalloc(&sink, 64);
    cudaMemset(table, 0, bytes);
    const uint64_t n = 4u << 20;
    run<2, 4, false>(table, bytes, n, sink, "read  K=4");
    run<4, 1, false>(table, bytes, n, sink, "read  K=1");
    run<4, 4, false>(table, bytes, n, sink, "read  K=4");
    run<4, 8, false>(table, bytes, n, sink, "read  K=8");
    run<8, 4, false>(table, bytes, n, sink, "read  K=4");
    run<16, 4, false>(table, bytes, n, sink, "read  K=4");
    run<32, 2, false>(table, bytes, n, sink, "read  K=2");
    run<4, 4, true>(table, bytes, n, sink, "rmw   K=4");
    run<8, 4, true>(table, bytes, n, sink, "rmw   K=4");
    run<16, 4, true>(table, bytes, n, sink, "rmw   K=4");
    // smaller footprints: inside the TLB reach / inside L2
    run<4, 4, false>(table, 256ull << 20, n, sink, "read  K=4 256MiB");
    run<4, 4, false>(table, 1ull << 30, n, sink, "read  K=4 1GiB");
    run<4, 4, false>(table, 64ull << 20, n, sink, "read  K=4 64MiB(L2)");
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
