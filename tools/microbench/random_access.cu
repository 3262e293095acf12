// Micro-benchmark behind the bucket-index design (DESIGN.md "large databases"): what does a random 16 / 32 / 64 / 128-byte
// read cost on B200, in time and in DRAM bytes, from a table larger than L2 (1 GiB) and from one of L2's size (128 MiB)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o random_access random_access.cu && ./random_access
// Each thread reads `bytes` from a random slot per step (steps are dependent); 148 x 1024 threads, 64 steps.
// Load kinds: 0 = 16-byte evict-first loads (ld.global.cs), 1 = 16-byte read-only loads (ld.global.nc),
//             2 = 32-byte read-only loads (ld.global.nc.v8.u32, sm_100+).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t ld32(const uint4 *p)
{
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int kVec, int kKind>
__global__ void __launch_bounds__(1024, 1) probe(const uint4 *table, uint32_t slots, int slot_vec, int steps, uint32_t *out)
{
    uint32_t acc = 0, rng = mix(blockIdx.x * 1024u + threadIdx.x + 1u);
    for (int s = 0; s < steps; ++s) {
        rng = mix(rng + s);
        const uint4 *p = table + (size_t)(rng % slots) * slot_vec;
        if constexpr (kKind == 2) {
            uint32_t q[kVec / 2 > 0 ? kVec / 2 : 1];
#pragma unroll
            for (int i = 0; i < kVec / 2; ++i) q[i] = ld32(p + 2 * i);
#pragma unroll
            for (int i = 0; i < kVec / 2; ++i) acc += q[i];
        } else {
            uint4 q[kVec];
#pragma unroll
            for (int i = 0; i < kVec; ++i) q[i] = kKind == 0 ? __ldcs(p + i) : __ldg(p + i);
#pragma unroll
            for (int i = 0; i < kVec; ++i) acc += q[i].x ^ q[i].y ^ q[i].z ^ q[i].w;
        }
        rng += acc & 1u;
    }
    out[blockIdx.x * 1024 + threadIdx.x] = acc;
}

template <int kVec, int kKind> void run(const char *name, const uint4 *table, size_t table_bytes, int slot_bytes, uint32_t *out)
{
    const int steps = 64, blocks = 148;
    const uint32_t slots = (uint32_t)(table_bytes / slot_bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(a);
        probe<kVec, kKind><<<blocks, 1024>>>(table, slots, slot_bytes / 16, steps, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r && ms < best) best = ms;
    }
    const double n = (double)blocks * 1024 * steps;
    printf("%4zu MiB  %-28s %8.3f ms  %6.1f G accesses/s  %7.1f GB/s useful\n", table_bytes >> 20, name, best,
           n / best / 1e6, n * kVec * 16 / best / 1e6);
}

int main()
{
    uint4 *table; uint32_t *out;
    cudaMalloc(&table, (size_t)1 << 30);
    cudaMemset(table, 0, (size_t)1 << 30);
    cudaMalloc(&out, 148 * 1024 * 4);
    for (size_t bytes : {(size_t)1 << 30, (size_t)128 << 20, (size_t)64 << 20}) {
        run<1, 0>("16 B of 64 B slot, cs", table, bytes, 64, out);
        run<1, 1>("16 B of 64 B slot, nc", table, bytes, 64, out);
        run<2, 1>("32 B slot, 2x16 nc", table, bytes, 32, out);
        run<2, 2>("32 B slot, 1x32 nc", table, bytes, 32, out);
        run<4, 0>("64 B slot, 4x16 cs", table, bytes, 64, out);
        run<4, 1>("64 B slot, 4x16 nc", table, bytes, 64, out);
        run<4, 2>("64 B slot, 2x32 nc", table, bytes, 64, out);
        run<8, 1>("128 B slot, 8x16 nc", table, bytes, 128, out);
        run<8, 2>("128 B slot, 4x32 nc", table, bytes, 128, out);
    }
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
