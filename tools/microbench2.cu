// tools/microbench2.cu -- ceiling of the "finished tile -> HBM" store path on B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// MODE 0: zero smem tile + fence + one bulk store per tile, wait (single buffer)
// MODE 1: as 0 but no zeroing (store path only)
// MODE 2: smem -> registers -> st.global.v4
// MODE 3: registers -> st.global.v4 (no smem)
// MODE 4: two half-size buffers, bulk stores kept one in flight (wait_group.read 1)
template <int MODE>
__global__ void __launch_bounds__(256) store_tiles(uint8_t *out, uint32_t n_tiles, uint32_t tile_bytes) {
    extern __shared__ __align__(128) uint8_t sm[];
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
        uint8_t *dst = out + (uint64_t)t * tile_bytes;
        if (MODE == 0 || MODE == 4) {
            const uint32_t off = (MODE == 4) ? (it & 1) * tile_bytes : 0;
            uint4 *tv = (uint4 *)(sm + off);
            for (uint32_t i = threadIdx.x; i < tile_bytes / 16; i += 256) tv[i] = make_uint4(t, 0, 0, 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr + off), "r"(tile_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (MODE == 4) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncthreads();
        } else if (MODE == 1) {
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "r"(tile_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncthreads();
        } else if (MODE == 2) {
            const uint4 *tv = (const uint4 *)sm;
            for (uint32_t i = threadIdx.x; i < tile_bytes / 16; i += 256) ((uint4 *)dst)[i] = tv[i];
        } else {
            for (uint32_t i = threadIdx.x; i < tile_bytes / 16; i += 256) ((uint4 *)dst)[i] = make_uint4(t, i, 0, 0);
        }
    }
    if (MODE == 4 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const uint64_t total = 8ull << 30;
    uint8_t *out; CK(cudaMalloc(&out, total));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
#define RUN(MODE, TILE, PER_SM, SMEM) { \
        CK(cudaFuncSetAttribute(store_tiles<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); \
        uint32_t nt = (uint32_t)(total / (TILE)); \
        store_tiles<MODE><<<sms * (PER_SM), 256, SMEM>>>(out, nt, TILE); \
        CK(cudaEventRecord(e0)); \
        store_tiles<MODE><<<sms * (PER_SM), 256, SMEM>>>(out, nt, TILE); \
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); \
        printf("mode %d tile %6d B  %d CTA/SM smem %6d: %7.1f GB/s\n", MODE, (int)(TILE), PER_SM, (int)(SMEM), total / ms / 1e6); }
    RUN(0, 65536, 3, 65536) RUN(0, 65536, 2, 65536) RUN(0, 65536, 1, 65536)
    RUN(0, 32768, 6, 32768) RUN(0, 32768, 4, 32768) RUN(0, 16384, 8, 16384)
    RUN(0, 131072, 1, 131072)
    RUN(1, 65536, 3, 65536) RUN(1, 65536, 1, 65536) RUN(1, 32768, 6, 32768)
    RUN(2, 65536, 3, 65536) RUN(2, 32768, 6, 32768)
    RUN(3, 65536, 3, 0) RUN(3, 65536, 8, 0)
    RUN(4, 32768, 3, 65536) RUN(4, 65536, 1, 131072) RUN(4, 32768, 1, 65536) RUN(4, 32768, 2, 65536)
    CK(cudaMemset(out, 0, total));
    CK(cudaEventRecord(e0)); CK(cudaMemsetAsync(out, 0, total)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("cudaMemset 8 GiB: %.1f GB/s\n", total / ms / 1e6);
    printf("done\n");
    return 0;
}
