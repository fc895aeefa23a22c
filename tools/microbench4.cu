// microbench4 -- does it matter WHEN the 256-byte pieces of one 1,280-byte row are requested?
// The K4 sliced probe reads, per position, one 1,280-byte row (10,000 leaves) as five 256-byte warp loads.  Today the five
// pieces are fetched by five different CTAs (grid.y = column group) at unrelated times.  (a) "scattered": every warp reads a
// 256-byte piece of its OWN random row; (b) "coherent": the five warps of a CTA read the five pieces of the SAME random row at
// the same time; (c) "coherent8": one warp reads 8-byte/lane x 5 = the whole row itself.  Same bytes, same randomness.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench4 tools/microbench4.cu && tools/microbench4
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }
constexpr int ROW_WORDS = 320;  // u32 words per row (1,280 B)
template <int MODE>
__global__ void __launch_bounds__(160) gather(const uint2 *__restrict__ m, uint64_t rows, uint32_t iters, uint64_t *out) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t acc = 0;
    const uint64_t cta = blockIdx.x;
    for (uint32_t it = 0; it < iters; it += 7) {
        uint2 v[7];
#pragma unroll
        for (int j = 0; j < 7; j++) {
            uint64_t r, piece;
            if (MODE == 0) { r = mix((cta * 5 + warp) * 1000003ull + it + j) % rows; piece = mix(r ^ warp) % 5; }       // own row, random piece
            else { r = mix(cta * 1000003ull + it + j) % rows; piece = warp; }                                             // shared row, piece = warp
            v[j] = __ldg(m + r * (ROW_WORDS / 2) + piece * 32 + lane);
        }
#pragma unroll
        for (int j = 0; j < 7; j++) acc += v[j].x ^ v[j].y;
    }
    if (acc == 0x1234567) out[0] = acc;
}
int main() {
    const uint64_t rows = (1ull << 25);  // 43 GB matrix, as the config-4 shard on one GPU
    uint2 *m; uint64_t *out;
    cudaMalloc(&m, rows * ROW_WORDS * 4); cudaMalloc(&out, 8);
    cudaMemset(m, 1, rows * ROW_WORDS * 4);
    const uint32_t iters = 7 * 64; const int grid = 148 * 12 * 8;
    for (int rep = 0; rep < 2; rep++) for (int mode = 0; mode < 2; mode++) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        if (mode == 0) gather<0><<<grid, 160>>>(m, rows, iters, out); else gather<1><<<grid, 160>>>(m, rows, iters, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double bytes = (double)grid * 5 * iters * 256;
        printf("%s: %.2f ms  %.0f GB/s (%.1f GB gathered in 256-byte pieces of 1,280-byte rows)\n", mode == 0 ? "scattered pieces" : "coherent rows   ", ms, bytes / ms / 1e6, bytes / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
