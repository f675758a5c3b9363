// DRAM bytes per random 8-byte gather on B200: run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./exp_gather [l2_fetch_granularity]
// Three kernels: plain ld.global.nc, ld with L2::evict_first, and a warp that gathers sorted ids.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; return x; }
template <int VARIANT>
__global__ void gather(const uint64_t *__restrict__ col, uint64_t n, uint64_t m, unsigned long long *out) {
    unsigned long long acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = mix(i * 0x9E3779B97F4A7C15ull) & (n - 1);
        uint64_t v;
        if (VARIANT == 0) v = __ldg(col + r);
        else if (VARIANT == 1) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(col + r));
        else if (VARIANT == 2) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(col + r));
        else if (VARIANT == 3) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(col + r));
        else if (VARIANT == 4) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(col + r));
        else if (VARIANT == 5) asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(col + r));
        else if (VARIANT == 6) v = atomicOr((unsigned long long *)(col + r), 0ull);
        else { asm volatile("ld.global.nc.L1::evict_first.u64 %0, [%1];" : "=l"(v) : "l"(col + r)); }
        acc += v;
    }
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
int main(int argc, char **argv) {
    if (argc > 1) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1]));
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit: %zu\n", g);
    const uint64_t n = 1ull << 28, m = 1ull << 24;
    uint64_t *col; unsigned long long *out;
    cudaMalloc(&col, n * 8); cudaMalloc(&out, 8); cudaMemset(col, 1, n * 8); cudaMemset(out, 0, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int v = 0; v < 8; ++v) {
        for (int rep = 0; rep < 1; ++rep) {
            cudaEventRecord(e0);
            if (v == 0) gather<0><<<148 * 8, 256>>>(col, n, m, out);
            if (v == 1) gather<1><<<148 * 8, 256>>>(col, n, m, out);
            if (v == 2) gather<2><<<148 * 8, 256>>>(col, n, m, out);
            if (v == 3) gather<3><<<148 * 8, 256>>>(col, n, m, out);
            if (v == 4) gather<4><<<148 * 8, 256>>>(col, n, m, out);
            if (v == 5) gather<5><<<148 * 8, 256>>>(col, n, m, out);
            if (v == 6) gather<6><<<148 * 8, 256>>>((const uint64_t*)col, n, m, out);
            if (v == 7) gather<7><<<148 * 8, 256>>>(col, n, m, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("variant %d: %.3f ms for %llu gathers (%.1f G/s)\n", v, ms, (unsigned long long)m, m / ms / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
