// Development microbenchmark: how should a thread hand its (S1, S2) records to the TMA unit?
// 8 "particles" per thread per tile (like k_project_bin), 1 cube RED each, NT of the 8 map updates as 16-byte
// bulk reductions, the rest as 2 REDs.  Variants:
//   0  burst: records written while the tile is processed, one fence, NT reductions issued at the end of the
//      tile, double-buffered staging with wait_group.read 1 (what the kernel does)
//   1  interleaved: record written, fence, reduction issued right after its particle (single buffer per slot,
//      wait_group.read 0 at the start of the tile)
//   2  two bursts per tile (after particles 4 and 8), double-buffered per half
// usage: tma_handoff_bench <dynamic smem KB to cap CTAs per SM>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_red(double* dst, const double* src) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 16;"
                 :: "l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src)) : "memory");
}

template <int NT, int VARIANT>
__global__ void __launch_bounds__(256, 2) k_handoff(double* cube, double* maps, uint32_t cmask, uint32_t mmask, int iters) {
    extern __shared__ __align__(16) unsigned char raw[];
    // [stage 2][slot 8][thread 256] records of 16 B  (64 KB) + destination pixel per record
    double* recs = reinterpret_cast<double*>(raw);
    uint32_t* dsts = reinterpret_cast<uint32_t*>(raw + 2 * 8 * 256 * 16);
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    int parity = 0;
    for (int it = 0; it < iters; ++it) {
        if (NT > 0) { if (VARIANT == 1) wait_read_0(); else wait_read_1(); }
        double* my = recs + ((size_t)parity * 8 * 256 + threadIdx.x) * 2;
        uint32_t* myd = dsts + (size_t)parity * 8 * 256 + threadIdx.x;
        int used = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t c = lcg(s) & cmask, m = lcg(s) & mmask;
            const double a = (double)(c & 7) + 0.5, b = a * a;
            atomicAdd(cube + c, 1.0);
            if (k < NT) {
                my[(size_t)used * 256 * 2] = a; my[(size_t)used * 256 * 2 + 1] = b;
                myd[used * 256] = m;
                ++used;
                if (VARIANT == 1) { fence_async(); bulk_red(maps + (size_t)m * 4 + 2, my + (size_t)(used - 1) * 256 * 2); }
            } else {
                atomicAdd(maps + (size_t)m * 4 + 2, a);
                atomicAdd(maps + (size_t)m * 4 + 3, b);
            }
            if (VARIANT == 2 && NT > 0 && k == 3) {
                fence_async();
                for (int j = 0; j < used; ++j) bulk_red(maps + (size_t)myd[j * 256] * 4 + 2, my + (size_t)j * 256 * 2);
                commit();
                parity ^= 1;
                wait_read_1();
                my = recs + ((size_t)parity * 8 * 256 + threadIdx.x) * 2;
                myd = dsts + (size_t)parity * 8 * 256 + threadIdx.x;
                used = 0;
            }
        }
        if (NT > 0) {
            if (VARIANT != 1) {
                fence_async();
                for (int j = 0; j < used; ++j) bulk_red(maps + (size_t)myd[j * 256] * 4 + 2, my + (size_t)j * 256 * 2);
            }
            commit();
            if (VARIANT != 1) parity ^= 1;
        }
    }
    if (NT > 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main(int argc, char** argv) {
    const int pad_kb = argc > 1 ? atoi(argv[1]) : 0;
    const uint32_t ncube = 1u << 22, nmap = 1u << 20;
    double *cube, *maps;
    cudaMalloc(&cube, (size_t)ncube * 8); cudaMemset(cube, 0, (size_t)ncube * 8);
    cudaMalloc(&maps, (size_t)nmap * 32); cudaMemset(maps, 0, (size_t)nmap * 32);
    const int grid = 148 * 8, iters = 64;
    const size_t smem = 2 * 8 * 256 * 16 + 2 * 8 * 256 * 4 + (size_t)pad_kb * 1024;
    const double parts = (double)grid * 256 * iters * 8;
    float t;
#define RUN(NT, V) { cudaFuncSetAttribute(k_handoff<NT, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_handoff<NT, V>, 256, smem); \
    t = timeit([&] { k_handoff<NT, V><<<grid, 256, smem>>>(cube, maps, ncube - 1, nmap - 1, iters); }); \
    printf("variant %d  TMA %d/8  %d CTAs/SM : %6.3f ms per 1e8 particles (%s)\n", V, NT, nb, t * 1e8 / parts, cudaGetErrorString(cudaGetLastError())); }
    RUN(0, 0) RUN(3, 0) RUN(4, 0) RUN(5, 0) RUN(6, 0) RUN(8, 0)
    RUN(3, 1) RUN(4, 1) RUN(6, 1) RUN(8, 1)
    RUN(4, 2) RUN(6, 2) RUN(8, 2)
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
