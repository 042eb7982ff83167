// Microbenchmark: random 16-byte gathers from a large table with different load flavours, to see
// how many DRAM bytes each costs on B200 (run under ncu for dram__bytes_read.sum) and how fast it goes.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31);
}
template <int V> __device__ __forceinline__ uint4 ld(const uint4* p) {
    uint4 v;
    if (V == 0) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (V == 1) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (V == 2) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (V == 3) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (V == 4) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (V == 5) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (V == 6) { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol)); }
    if (V == 7) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (V == 8) { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol)); }
    if (V == 9) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int V> __global__ void __launch_bounds__(256) gather(const uint4* t, uint64_t n, int iters, uint32_t* out) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
        uint64_t r = mix(tid * 0x9e3779b97f4a7c15ULL + i);
        uint64_t idx = __umul64hi(r, n);
        uint4 v = ld<V>(t + idx);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678) out[0] = acc;
}
template <int V> void run(const uint4* t, uint64_t n, uint32_t* out, const char* name) {
    const int grid = 148 * 8, iters = 256;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather<V><<<grid, 256>>>(t, n, 8, out);
    cudaEventRecord(a);
    gather<V><<<grid, 256>>>(t, n, iters, out);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    double loads = (double)grid * 256 * iters;
    printf("variant %d %-40s: %.2f G loads/s (%.3f ms)\n", V, name, loads / ms / 1e6, ms);
}
int main(int argc, char** argv) {
    size_t gran = argc > 1 ? atoi(argv[1]) : 0;
    double gb = argc > 2 ? atof(argv[2]) : 13.0;
    if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("L2 fetch granularity limit = %zu, table %.1f GB\n", g, gb);
    uint64_t n = (uint64_t)(gb * 1e9 / 16);
    uint4* t; CK(cudaMalloc(&t, n * 16)); CK(cudaMemset(t, 1, n * 16));
    uint32_t* out; CK(cudaMalloc(&out, 4));
    run<0>(t, n, out, "ld.global");
    run<1>(t, n, out, "ld.global.nc");
    run<2>(t, n, out, "ld.global.nc.L1::no_allocate");
    run<3>(t, n, out, "ld.global.cg");
    run<4>(t, n, out, "ld.global.cs");
    run<5>(t, n, out, "ld.global.cv");
    run<6>(t, n, out, "ld.nc.L1::no_allocate.L2 evict_first");
    run<7>(t, n, out, "ld.global.L1::no_allocate");
    run<8>(t, n, out, "ld.L1::no_allocate.L2 evict_normal hint");
    run<9>(t, n, out, "ld.relaxed.gpu");
    return 0;
}
