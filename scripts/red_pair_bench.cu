// Development microbenchmark: does a warp-level RED whose lanes hit the SAME 32-byte sector cost L2 one request
// or one per lane?  N lane-reductions (f64) issued as
//   0  every lane its own random sector
//   1  lanes (2j, 2j+1) adjacent 8-byte slots of one random sector
//   2  lanes (4j .. 4j+3) the four slots of one random sector
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int GROUP>
__global__ void __launch_bounds__(256) k_pair(double* maps, uint32_t mmask, int iters) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t m = lcg(s) & mmask;
            if (GROUP > 1) m = __shfl_sync(0xffffffffu, m, lane & ~(GROUP - 1));
            atomicAdd(maps + (size_t)m * 4 + (GROUP > 1 ? (lane & (GROUP - 1)) : 0), 1.0);
        }
    }
}

template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    const uint32_t nmap = 1u << 21;                      // 64 MB of 32-byte records
    double* maps; cudaMalloc(&maps, (size_t)nmap * 32); cudaMemset(maps, 0, (size_t)nmap * 32);
    const int grid = 148 * 8, iters = 64;
    const double ops = (double)grid * 256 * iters * 8;
#define RUN(G) { float t = timeit([&] { k_pair<G><<<grid, 256>>>(maps, nmap - 1, iters); }); \
    printf("lanes per sector %d : %6.3f ms per 1e8 lane-reductions = %.3g lane-ops/s (%s)\n", G, t * 1e8 / ops, ops / (t * 1e-3), cudaGetErrorString(cudaGetLastError())); }
    RUN(1) RUN(2) RUN(4)
    return 0;
}
