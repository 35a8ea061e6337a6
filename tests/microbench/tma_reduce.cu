// Development microbenchmark: throughput of TMA bulk reduce-adds (cp.reduce.async.bulk.global.shared::cta .add.u64) from shared memory
// into global memory, the primitive a fused mod-up + key inner product would accumulate its products with.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_reduce tma_reduce.cu && ./tma_reduce
// Each CTA owns a 64 KiB tile; per iteration each of its 8 warps reduce-adds its 8 KiB block into the accumulator region of
// "its ciphertext / output modulus"; `share` CTAs (the digits) accumulate into the same region; regions cover `footprint` bytes.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_reduce_add_u64(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

template <int MODE> __global__ void __launch_bounds__(256, 3) k(uint64_t *acc, uint64_t regions, uint32_t share, int iters) {
    extern __shared__ __align__(128) uint64_t sm[];
    for (int i = threadIdx.x; i < 8192; i += 256) sm[i] = i + blockIdx.x;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int it = 0; it < iters; it++) {
        // unit = (iteration, CTA): `share` consecutive units hit the same region
        const uint64_t unit = (uint64_t)it * gridDim.x + blockIdx.x;
        const uint64_t region = (unit / share) % regions;
        uint64_t *dst = acc + region * 8192 + warp * 1024;
        if (lane == 0) {
            if (MODE == 0) bulk_reduce_add_u64(dst, sm + warp * 1024, 8192);
            else bulk_store(dst, sm + warp * 1024, 8192);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const int grid = 148 * 3, iters = 400;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (uint64_t footprint_mb : {64ull, 600ull, 2400ull}) {
        const uint64_t regions = footprint_mb * (1ull << 20) / 65536;
        uint64_t *acc;
        cudaMalloc(&acc, regions * 65536);
        cudaMemset(acc, 0, regions * 65536);
        for (uint32_t share : {1u, 8u}) {
            for (int mode = 0; mode < 2; mode++) {
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                for (int rep = 0; rep < 2; rep++) {
                    cudaEventRecord(e0);
                    if (mode == 0) k<0><<<grid, 256, 65536>>>(acc, regions, share, iters);
                    else k<1><<<grid, 256, 65536>>>(acc, regions, share, iters);
                    cudaEventRecord(e1);
                    cudaDeviceSynchronize();
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = (double)grid * iters * 65536;
                printf("{\"op\": \"%s\", \"footprint_mb\": %llu, \"ctas_per_region\": %u, \"gbs\": %.1f, \"err\": \"%s\"}\n", mode == 0 ? "bulk_reduce_add_u64" : "bulk_store",
                       (unsigned long long)footprint_mb, share, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
        }
        cudaFree(acc);
    }
    return 0;
}
