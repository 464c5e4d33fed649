// Development microbenchmark (round 2): WHERE is the limit of scattered f64 reductions -- in the SM (LSU / L1TEX /
// the SM's port into the crossbar) or in L2 -- and do streaming loads share it?
//   A  RED.F64 with L = 1, 2, 3, 4 lanes per random 32-byte sector, on all SMs and on half of them (one CTA per SM,
//      forced by its shared-memory size): an SM-side limit halves the rate with half the SMs, an L2-side limit does not.
//   B  independent warps of one kernel stream 2.8 GB with LDG.256 (evict-first) while others issue 3-lane reductions:
//      alone, alone, together -- no latency coupling between the two, so any slowdown is a shared resource.
//   C  shared-memory atomics (32-bit integer, f32, f64) on spread addresses: lane-operations per clock per SM
//      (the reason there is no shared-memory privatisation in k_project_bin / k_project_cube).
// usage: red_limits_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
__device__ __forceinline__ void ldg8(const float* p, float* o) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7]) : "l"(p));
}

// ---- A: `iters` x 8 RED instructions per thread; quads of lanes share a sector, L of the 4 lanes are active
template <int L>
__global__ void __launch_bounds__(1024, 1) k_red(double* acc, uint32_t mask, int iters, unsigned* smids) {
    extern __shared__ unsigned char pad[];
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const int lane = threadIdx.x & 31, q = lane & 3;
    if (threadIdx.x == 0) { unsigned id; asm("mov.u32 %0, %%smid;" : "=r"(id)); smids[blockIdx.x] = id; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t m = lcg(s) & mask;
            m = __shfl_sync(0xffffffffu, m, lane & ~3);
            if (q < L) atomicAdd(acc + (size_t)m * 4 + q, 1.0);
        }
    }
}

// ---- B: warps [0, nload) stream, warps [nload, nwarps) reduce (3 lanes per sector).  Work is split evenly inside
// each role: `nvec` LDG.256 per array over 7 arrays, `nred` sector reductions in total.
__global__ void __launch_bounds__(512, 1) k_mix(const float* src, int64_t nvec, double* acc, uint32_t mask, int64_t nred,
                                               int nload, float* sink) {
    extern __shared__ unsigned char pad[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (warp < nload) {
        const int64_t t0 = ((int64_t)blockIdx.x * nload + warp) * 32 + lane, stride = (int64_t)gridDim.x * nload * 32;
        float keep = 0.f;
        for (int64_t i = t0; i < nvec; i += stride) {
            float v[7][8];
#pragma unroll
            for (int k = 0; k < 7; ++k) ldg8(src + (k * nvec + i) * 8, v[k]);
#pragma unroll
            for (int k = 0; k < 7; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j) keep += v[k][j];
        }
        if (keep == 123.456f) *sink = keep;
    } else {
        const int nr = nw - nload;
        if (nr <= 0 || nred <= 0) return;
        const int64_t per_warp = nred / ((int64_t)gridDim.x * nr) / 8;    // 8 sectors per RED instruction
        uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
        const int q = lane & 3;
        for (int64_t it = 0; it < per_warp; ++it) {
            uint32_t m = lcg(s) & mask;
            m = __shfl_sync(0xffffffffu, m, lane & ~3);
            if (q != 1) atomicAdd(acc + (size_t)m * 4 + q, 1.0);
        }
    }
}

// ---- C: shared-memory atomics on spread addresses
template <typename T>
__global__ void __launch_bounds__(1024, 1) k_smem_atomic(int iters, T* out) {
    __shared__ T buf[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = T(0);
    __syncthreads();
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const uint32_t r = lcg(s); atomicAdd(buf + (r & 4095), T((r >> 13) & 7)); }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = buf[0];
}

template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const double clk = prop.clockRate * 1e3;
    printf("%s, %d SMs, %.0f MHz\n", prop.name, prop.multiProcessorCount, clk / 1e6);
    const uint32_t nrec = 1u << 21;                       // 64 MB of 32-byte records (the interleaved cube is 67 MB)
    double* acc; cudaMalloc(&acc, (size_t)nrec * 32); cudaMemset(acc, 0, (size_t)nrec * 32);
    unsigned* smids; cudaMalloc(&smids, 4096 * 4);
    const size_t big = 200 * 1024;                        // one CTA per SM
    // ---- A
    {
        const int iters = 64;
#define RUN_A(L, G) { cudaFuncSetAttribute(k_red<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big); \
        float t = timeit([&] { k_red<L><<<G, 1024, big>>>(acc, nrec - 1, iters, smids); }); \
        unsigned h[1024]; cudaMemcpy(h, smids, G * 4, cudaMemcpyDeviceToHost); int distinct = 0; bool seen[256] = {}; \
        for (int i = 0; i < G; ++i) if (!seen[h[i] & 255]) { seen[h[i] & 255] = true; ++distinct; } \
        const double sectors = (double)G * 1024 * iters * 8 / 4; \
        printf("A  L=%d lanes/sector, %3d CTAs on %3d SMs: %7.3f ms  %.3g sectors/s  %.3g lanes/s  %.2f clk per sector per SM  (%s)\n", \
               L, G, distinct, t, sectors / (t * 1e-3), sectors * L / (t * 1e-3), t * 1e-3 * clk * distinct / sectors, \
               cudaGetErrorString(cudaGetLastError())); }
        RUN_A(1, 148) RUN_A(1, 74) RUN_A(2, 148) RUN_A(2, 74) RUN_A(3, 148) RUN_A(3, 74) RUN_A(3, 37) RUN_A(4, 148) RUN_A(4, 74)
    }
    // ---- B
    {
        const int64_t n = 100'000'000 / 256 * 256, nvec = n / 8;
        float *src, *sink; cudaMalloc(&src, 7 * n * sizeof(float)); cudaMemset(src, 0x3c, 7 * n * sizeof(float)); cudaMalloc(&sink, 4);
        cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big);
        const int64_t nred = 100'000'000;
#define RUN_B(name, NV, NR, NLOAD, THREADS) { \
        float t = timeit([&] { k_mix<<<148, THREADS, big>>>(src, NV, acc, nrec - 1, NR, NLOAD, sink); }); \
        printf("B  %-58s %7.3f ms  (%s)\n", name, t, cudaGetErrorString(cudaGetLastError())); }
        RUN_B("loads only, 8 load warps per SM (2.8 GB)", nvec, 0, 8, 256)
        RUN_B("loads only, 4 load warps per SM", nvec, 0, 4, 128)
        RUN_B("loads only, 16 load warps per SM", nvec, 0, 16, 512)
        RUN_B("reductions only, 8 red warps per SM (1e8 x 3 lanes)", 0, nred, 0, 256)
        RUN_B("reductions only, 4 red warps per SM", 0, nred, 0, 128)
        RUN_B("reductions only, 12 red warps per SM", 0, nred, 0, 384)
        RUN_B("8 load warps + 8 red warps", nvec, nred, 8, 512)
        RUN_B("4 load warps + 12 red warps", nvec, nred, 4, 512)
        RUN_B("4 load warps + 4 red warps", nvec, nred, 4, 256)
        RUN_B("8 load warps + 8 red warps, half the reductions", nvec, nred / 2, 8, 512)
    }
    // ---- C
    {
        const int iters = 256;
        void* out; cudaMalloc(&out, 148 * 8);
        const double ops = 148.0 * 1024 * iters * 8;
#define RUN_C(T, name) { float t = timeit([&] { k_smem_atomic<T><<<148, 1024>>>(iters, (T*)out); }); \
        printf("C  shared atomicAdd %-8s spread over 4096 slots: %7.3f ms  %.3g lane-ops/s chip-wide  %.2f clk per lane-op per SM  (%s)\n", \
               name, t, ops / (t * 1e-3), t * 1e-3 * clk * 148 / ops, cudaGetErrorString(cudaGetLastError())); }
        RUN_C(int, "int32") RUN_C(float, "float") RUN_C(double, "double") RUN_C(unsigned long long, "uint64")
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
