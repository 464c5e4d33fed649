// Development microbenchmark: do scattered TMA bulk reductions (UBLKRED) overlap with LSU REDs?
// Per "particle": 1 LSU RED to a cube-like array + map update done either as 2 LSU REDs or as one
// 16-byte TMA reduce; NT of every 4 particles use the TMA path.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int NT>
__global__ void __launch_bounds__(256) k_mix(double* cube, double* maps, uint32_t cmask, uint32_t mmask, int iters) {
    __shared__ __align__(16) double slots[256 * 4 * 2];
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    double* mine = slots + threadIdx.x * 8;
    for (int it = 0; it < iters; ++it) {
        if (NT > 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
            for (int k = 0; k < NT; ++k) { mine[k * 2] = 1.0; mine[k * 2 + 1] = 2.0; }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t c = lcg(s) & cmask, m = lcg(s) & mmask;
            atomicAdd(cube + c, 1.0);
            double* dst = maps + (size_t)m * 4 + 1;     // slots 1,2 of a 32-byte pixel record
            if (k < NT) {
                // 16-byte aligned destination: use slots 2,3 instead
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(mine + k * 2);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 16;"
                             :: "l"(maps + (size_t)m * 4 + 2), "r"(src) : "memory");
            } else {
                atomicAdd(dst, 1.0);
                atomicAdd(dst + 1, 2.0);
            }
        }
        if (NT > 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (NT > 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    const uint32_t ncube = 1u << 22, nmap = 1u << 20;   // 32 MB cube, 32 MB of map records
    double *cube, *maps;
    cudaMalloc(&cube, (size_t)ncube * 8); cudaMemset(cube, 0, (size_t)ncube * 8);
    cudaMalloc(&maps, (size_t)nmap * 32); cudaMemset(maps, 0, (size_t)nmap * 32);
    const int grid = 148 * 4, iters = 256;
    const double parts = (double)grid * 256 * iters * 4;
    float t;
#define RUN(NT) t = timeit([&] { k_mix<NT><<<grid, 256>>>(cube, maps, ncube - 1, nmap - 1, iters); }); \
    printf("TMA share %d/4 : %7.3f ms  -> %6.3f ms per 1e8 particles (%s)\n", NT, t, t * 1e8 / parts, cudaGetErrorString(cudaGetLastError()));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
