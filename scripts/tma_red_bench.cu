// Development microbenchmark (not part of the product): throughput of scattered 16/32-byte
// TMA bulk reductions (cp.reduce.async.bulk ... add.f64) vs plain RED.F64, random 32-byte slots.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int BYTES, int PER>
__global__ void __launch_bounds__(256) k_tma(double* acc, uint32_t mask, int iters) {
    __shared__ __align__(16) double slots[256 * PER * 4];
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    double* mine = slots + threadIdx.x * PER * 4;
    for (int it = 0; it < iters; ++it) {
        // make sure the previous group has finished READING our slots
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            mine[k * 4 + 0] = 1.0; mine[k * 4 + 1] = 2.0;
            if (BYTES == 32) { mine[k * 4 + 2] = 3.0; mine[k * 4 + 3] = 0.0; }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const uint32_t idx = lcg(s) & mask;
            double* dst = acc + (size_t)idx * 4;
            const uint32_t src = (uint32_t)__cvta_generic_to_shared(mine + k * 4);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                         :: "l"(dst), "r"(src), "n"(BYTES) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int NRED, int PER>
__global__ void __launch_bounds__(256) k_red(double* acc, uint32_t mask, int iters) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const uint32_t idx = lcg(s) & mask;
            double* dst = acc + (size_t)idx * 4;
            atomicAdd(dst, 1.0);
            if (NRED > 1) atomicAdd(dst + 1, 2.0);
            if (NRED > 2) atomicAdd(dst + 2, 3.0);
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
    const uint32_t nslots = 1u << 20;            // 32 MB of 32-byte slots (L2 resident)
    double* acc; cudaMalloc(&acc, (size_t)nslots * 32); cudaMemset(acc, 0, (size_t)nslots * 32);
    const int grid = 148 * 4, iters = 256;
    constexpr int PER = 4;
    const double ops = (double)grid * 256 * iters * PER;
    float t;
    t = timeit([&] { k_red<1, PER><<<grid, 256>>>(acc, nslots - 1, iters); });
    printf("RED.F64 x1        : %7.3f ms  %6.1f Gslot/s\n", t, ops / t / 1e6);
    t = timeit([&] { k_red<2, PER><<<grid, 256>>>(acc, nslots - 1, iters); });
    printf("RED.F64 x2        : %7.3f ms  %6.1f Gslot/s\n", t, ops / t / 1e6);
    t = timeit([&] { k_red<3, PER><<<grid, 256>>>(acc, nslots - 1, iters); });
    printf("RED.F64 x3        : %7.3f ms  %6.1f Gslot/s\n", t, ops / t / 1e6);
    t = timeit([&] { k_tma<16, PER><<<grid, 256>>>(acc, nslots - 1, iters); });
    printf("TMA reduce 16 B   : %7.3f ms  %6.1f Gslot/s  (%s)\n", t, ops / t / 1e6, cudaGetErrorString(cudaGetLastError()));
    t = timeit([&] { k_tma<32, PER><<<grid, 256>>>(acc, nslots - 1, iters); });
    printf("TMA reduce 32 B   : %7.3f ms  %6.1f Gslot/s  (%s)\n", t, ops / t / 1e6, cudaGetErrorString(cudaGetLastError()));
    double h[4]; cudaMemcpy(h, acc, 32, cudaMemcpyDeviceToHost);
    printf("slot0 = %g %g %g %g  err=%s\n", h[0], h[1], h[2], h[3], cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
