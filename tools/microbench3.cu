// tools/microbench3.cu -- shared-memory scatter throughput on B200: the unit both K1 roles lean on.
//   MODE 0: atom.shared.add.u32 with return (rank in a bucket)      MODE 1: red.shared.or.b32 (bit set, no return)
//   MODE 2: st.shared.u32 scatter                                   MODE 3: ld.shared.u32 gather
// Addresses are pre-computed random words (one LDG-free xorshift step per op is too ALU heavy, so they are
// generated once into registers and rotated), SPAN words of shared memory, THREADS per CTA, 1 CTA per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int MODE>
__global__ void scatter(uint32_t span_mask, uint32_t iters, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    for (uint32_t i = threadIdx.x; i <= span_mask; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t a[8];
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int j = 0; j < 8; j++) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; a[j] = x; }
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t addr = base + ((a[j] & span_mask) << 2);
            if (MODE == 0) { uint32_t r; asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"(addr) : "memory"); acc += r; }
            if (MODE == 1) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(a[j]) : "memory");
            if (MODE == 2) asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(a[j]) : "memory");
            if (MODE == 3) { uint32_t r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory"); acc += r; }
            a[j] = a[j] * 1664525u + 1013904223u;  // 1 IMAD per op
        }
    }
    if (acc == 0xdeadbeef) sink[0] = acc;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t *sink; CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const uint32_t iters = 2000;
#define RUN(MODE, THREADS, SPANW) { \
        CK(cudaFuncSetAttribute(scatter<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (SPANW) * 4)); \
        scatter<MODE><<<sms, THREADS, (SPANW) * 4>>>((SPANW) - 1, 10, sink); \
        CK(cudaEventRecord(e0)); \
        scatter<MODE><<<sms, THREADS, (SPANW) * 4>>>((SPANW) - 1, iters, sink); \
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); \
        double ops = (double)sms * THREADS * iters * 8; \
        printf("mode %d threads %4d span %6d words: %8.1f Gop/s chip  %6.2f lane-ops/clk/SM (1.965 GHz)  %5.1f clk per warp-op\n", MODE, THREADS, (int)(SPANW), ops / ms / 1e6, ops / ms / 1e6 / sms / 1.965, 32.0 / (ops / ms / 1e6 / sms / 1.965)); }
    for (int t = 0; t < 1; t++) {
        RUN(0, 256, 1024) RUN(0, 768, 1024) RUN(0, 1024, 1024) RUN(0, 1024, 16384)
        RUN(1, 256, 16384) RUN(1, 768, 16384) RUN(1, 1024, 16384) RUN(1, 1024, 1024)
        RUN(2, 256, 16384) RUN(2, 768, 16384) RUN(2, 1024, 16384)
        RUN(3, 256, 16384) RUN(3, 1024, 16384)
    }
    printf("done\n");
    return 0;
}
