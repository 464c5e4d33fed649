// Development microbenchmark: do the particle LOADS compete with the scattered REDs for the LSU?
// Every "particle" = 7 floats streamed from HBM (SoA arrays, like k_project_bin) + 3 scattered f64 REDs whose
// addresses depend on the loaded values.  LOAD variants:
//   0  no loads (addresses from an LCG only)               -> cost of the reductions alone
//   1  7 x ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 per 8 particles (what the kernel does)
//   2  7 x 1 KB cp.async.bulk (TMA) per warp-tile into a per-warp double-buffered shared-memory ring, completion
//      on an mbarrier, then 14 x LDS.128 -- the LSU sees no global loads
//   3  loads only (variant 1) without reductions                -> streaming cost alone
// usage: tma_load_bench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

__device__ __forceinline__ void ldg8(const float* p, float* o) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7]) : "l"(p));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
                 :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

struct Arrays { const float* a[7]; };

template <int LOAD>
__global__ void __launch_bounds__(256, 2) k_load(Arrays arr, int64_t ntiles, double* cube, double* maps, uint32_t cmask,
                                                 uint32_t mmask, float* sink) {
    extern __shared__ __align__(128) unsigned char raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ring = reinterpret_cast<float*>(raw) + (size_t)warp * 2 * 7 * 256;       // [stage][array][256 floats]
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw + (size_t)(blockDim.x >> 5) * 2 * 7 * 1024) + warp * 2;
    const int64_t gwarp = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp, nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float keep = 0.f;
    if (LOAD == 2) {
        if (lane == 0) {
            mbar_init(bars, 1); mbar_init(bars + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            for (int st = 0; st < 2; ++st) {
                const int64_t t = gwarp + st * nwarps;
                if (t < ntiles) {
                    mbar_expect(bars + st, 7 * 1024);
                    for (int k = 0; k < 7; ++k) bulk_load(ring + (st * 7 + k) * 256, arr.a[k] + t * 256, 1024, bars + st);
                }
            }
        }
        __syncwarp();
    }
    int it = 0;
    for (int64_t t = gwarp; t < ntiles; t += nwarps, ++it) {
        float v[7][8];
        if (LOAD == 1 || LOAD == 3) {
#pragma unroll
            for (int k = 0; k < 7; ++k) ldg8(arr.a[k] + t * 256 + lane * 8, v[k]);
        } else if (LOAD == 2) {
            const int st = it & 1;
            mbar_wait(bars + st, (it >> 1) & 1);
            const float4* src = reinterpret_cast<const float4*>(ring + st * 7 * 256);
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const float4 lo = src[k * 64 + lane], hi = src[k * 64 + 32 + lane];   // conflict-free: 16 B per lane, stride 16 B
                v[k][0] = lo.x; v[k][1] = lo.y; v[k][2] = lo.z; v[k][3] = lo.w;
                v[k][4] = hi.x; v[k][5] = hi.y; v[k][6] = hi.z; v[k][7] = hi.w;
            }
            __syncwarp();                                                             // everyone has read the stage
            const int64_t tn = t + 2 * nwarps;
            if (lane == 0 && tn < ntiles) {
                mbar_expect(bars + st, 7 * 1024);
                for (int k = 0; k < 7; ++k) bulk_load(ring + (st * 7 + k) * 256, arr.a[k] + tn * 256, 1024, bars + st);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 7; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float mix = 0.f;
#pragma unroll
            for (int k = 0; k < 7; ++k) mix += v[k][j];
            if (LOAD == 3) { keep += mix; continue; }
            const uint32_t h = __float_as_uint(mix) >> 3;
            const uint32_t c = (lcg(s) ^ h) & cmask, m = (lcg(s) ^ h) & mmask;
            const double a = (double)(c & 7) + 0.5;
            atomicAdd(cube + c, 1.0);
            atomicAdd(maps + (size_t)m * 4 + 2, a);
            atomicAdd(maps + (size_t)m * 4 + 3, a * a);
        }
    }
    if (keep == 123.456f) *sink = keep;
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
    const int64_t n = 100'000'000 / 256 * 256, ntiles = n / 256;
    Arrays arr;
    for (int k = 0; k < 7; ++k) {
        float* p; cudaMalloc(&p, n * sizeof(float)); cudaMemset(p, 0x3c + k, n * sizeof(float)); arr.a[k] = p;
    }
    const uint32_t ncube = 1u << 22, nmap = 1u << 20;
    double *cube, *maps; float* sink;
    cudaMalloc(&cube, (size_t)ncube * 8); cudaMemset(cube, 0, (size_t)ncube * 8);
    cudaMalloc(&maps, (size_t)nmap * 32); cudaMemset(maps, 0, (size_t)nmap * 32);
    cudaMalloc(&sink, 4);
    const int grid = 148 * 2;
    const size_t smem = 8 * 2 * 7 * 1024 + 8 * 2 * 8;
#define RUN(L, name) { cudaFuncSetAttribute(k_load<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_load<L>, 256, smem); \
    float t = timeit([&] { k_load<L><<<grid, 256, smem>>>(arr, ntiles, cube, maps, ncube - 1, nmap - 1, sink); }); \
    printf("%-44s %d CTAs/SM : %6.3f ms per 1e8 particles (%s)\n", name, nb, t * 1e8 / (double)n, cudaGetErrorString(cudaGetLastError())); }
    RUN(0, "reductions only (no loads)")
    RUN(3, "LDG.256 loads only (no reductions)")
    RUN(1, "LDG.256 loads + reductions")
    RUN(2, "TMA bulk loads -> smem -> LDS + reductions")
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
