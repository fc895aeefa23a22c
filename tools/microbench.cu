// tools/microbench.cu -- design experiments for the k-mer -> Bloom scatter on B200.
// Not part of the product; results are summarised in profiles/ and DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}

// positions[i] = random bit index in [0, nbits)
__global__ void gen_positions(uint32_t *pos, uint64_t n, uint64_t nbits_mask, uint64_t salt) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        pos[i] = (uint32_t)(mix64(i * 0x9E3779B97F4A7C15ULL + salt) & nbits_mask);
}

// RED.OR of positions whose bit index lies in [lo, hi)
__global__ void __launch_bounds__(256) red_scatter(const uint32_t *__restrict__ pos, uint64_t n, uint32_t *__restrict__ out, uint64_t lo, uint64_t hi) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p = pos[i];
        if (p >= lo && p < hi) atomicOr(out + (p >> 5), 1u << (p & 31));
    }
}

// plain (racy, for throughput only) read-modify-write
__global__ void __launch_bounds__(256) rmw_scatter(const uint32_t *__restrict__ pos, uint64_t n, uint32_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p = pos[i];
        out[p >> 5] |= 1u << (p & 31);
    }
}

__global__ void __launch_bounds__(256) zero_fill(uint4 *out, uint64_t nvec) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = make_uint4(0, 0, 0, 0);
}

// shared-memory tile experiments: each CTA sets `per_cta` random bits in a 128 KiB smem tile
template <int MODE>  // 0 = atomicOr, 1 = plain ld/or/st (racy), 2 = byte store into a byte map
__global__ void __launch_bounds__(512) smem_scatter(uint32_t per_cta, uint32_t *sink) {
    extern __shared__ uint32_t tile[];
    const uint32_t words = 32768;  // 128 KiB
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) tile[i] = 0;
    __syncthreads();
    uint64_t s = blockIdx.x * 1315423911ull + threadIdx.x;
    for (uint32_t i = threadIdx.x; i < per_cta; i += blockDim.x) {
        uint32_t p = (uint32_t)mix64(s + i) & (words * 32 - 1);
        if (MODE == 0) atomicOr(tile + (p >> 5), 1u << (p & 31));
        else if (MODE == 1) tile[p >> 5] |= 1u << (p & 31);
        else ((uint8_t *)tile)[p >> 3] = 1;
    }
    __syncthreads();
    uint32_t acc = 0;
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) acc ^= tile[i];
    if (acc == 0x12345678) sink[0] = acc;
}

// hash-only ALU throughput: murmur3 finaliser chain as in the k-mer hash (6 multiplies)
__global__ void __launch_bounds__(256) hash_alu(uint64_t n, uint64_t *sink) {
    uint64_t acc = 0;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t k1 = i * c1; k1 = (k1 << 31) | (k1 >> 33); k1 *= c2;
        uint64_t h1 = k1 ^ 8, h2 = 8; h1 += h2; h2 += h1;
        acc ^= mix64(h1) + mix64(h2);
    }
    if (acc == 0x12345) sink[0] = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main(int argc, char **argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s SMs %d L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    const uint64_t NPOS = 5000000;                 // k-mers of one 5 Mbp genome
    const uint64_t FILTER_BYTES = 1ull << 27;      // 2^30 bits
    const int NF = 64;                             // rotate over 8 GiB of filters so nothing is L2-warm
    uint8_t *filters; CK(cudaMalloc(&filters, FILTER_BYTES * NF));
    uint32_t *pos; CK(cudaMalloc(&pos, NPOS * 4 * 8));
    uint32_t *sink; CK(cudaMalloc(&sink, 64));
    gen_positions<<<sms * 8, 256>>>(pos, NPOS * 8, (1ull << 30) - 1, 12345);
    CK(cudaDeviceSynchronize());

    // ---- E: memset / zero-fill bandwidth
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(e0));
        for (int f = 0; f < NF; f++) CK(cudaMemsetAsync(filters + f * FILTER_BYTES, 0, FILTER_BYTES));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1);
        printf("memset 64x128MiB: %.3f ms  %.1f GB/s  (%.2f us per filter)\n", ms, NF * FILTER_BYTES / ms / 1e6, ms * 1e3 / NF);
    }
    {
        CK(cudaEventRecord(e0));
        for (int f = 0; f < NF; f++) zero_fill<<<sms * 8, 256>>>((uint4 *)(filters + f * FILTER_BYTES), FILTER_BYTES / 16);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1);
        printf("zero_fill kernel 64x128MiB: %.3f ms  %.1f GB/s\n", ms, NF * FILTER_BYTES / ms / 1e6);
    }

    // ---- A: direct: memset whole filter, then RED all positions (per genome), 64 genomes
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(e0));
        for (int f = 0; f < NF; f++) {
            CK(cudaMemsetAsync(filters + f * FILTER_BYTES, 0, FILTER_BYTES));
            red_scatter<<<sms * 8, 256>>>(pos + (f % 2) * NPOS, NPOS, (uint32_t *)(filters + f * FILTER_BYTES), 0, 1ull << 30);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1);
        printf("direct memset+RED (2^30 bits, 5M pos): %.2f us per genome\n", ms * 1e3 / NF);
    }
    // RED only (filters cold in L2: previous loop touched 8 GiB)
    {
        CK(cudaEventRecord(e0));
        for (int f = 0; f < NF; f++)
            red_scatter<<<sms * 8, 256>>>(pos + (f % 2) * NPOS, NPOS, (uint32_t *)(filters + f * FILTER_BYTES), 0, 1ull << 30);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1);
        printf("RED only, cold 128MiB filters: %.2f us per genome  (%.2f Gatom/s)\n", ms * 1e3 / NF, NPOS * NF / ms / 1e6);
    }

    // ---- B: region-staged: for each region of the filter: memset region, RED positions in region
    for (int R = 1; R <= 16; R *= 2) {
        const uint64_t region_bits = (1ull << 30) / R;
        CK(cudaEventRecord(e0));
        for (int f = 0; f < NF; f++)
            for (int r = 0; r < R; r++) {
                CK(cudaMemsetAsync(filters + f * FILTER_BYTES + r * (FILTER_BYTES / R), 0, FILTER_BYTES / R));
                red_scatter<<<sms * 8, 256>>>(pos + (f % 2) * NPOS, NPOS, (uint32_t *)(filters + f * FILTER_BYTES), r * region_bits, (r + 1) * region_bits);
            }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1);
        printf("region-staged R=%2d (region %4llu MiB): %.2f us per genome\n", R, (unsigned long long)(FILTER_BYTES / R >> 20), ms * 1e3 / NF);
    }

    // ---- C: RED throughput on an L2-resident region of varying size (no memset in the timed part)
    for (uint64_t mib = 4; mib <= 1024; mib *= 2) {
        const uint64_t bits = mib << 23;
        gen_positions<<<sms * 8, 256>>>(pos, NPOS * 8, bits - 1, mib);
        CK(cudaMemsetAsync(filters, 0, mib << 20));
        red_scatter<<<sms * 8, 256>>>(pos, NPOS * 8, (uint32_t *)filters, 0, bits);  // warm
        CK(cudaEventRecord(e0));
        for (int it = 0; it < 4; it++) red_scatter<<<sms * 8, 256>>>(pos, NPOS * 8, (uint32_t *)filters, 0, bits);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1);
        CK(cudaEventRecord(e0));
        for (int it = 0; it < 4; it++) rmw_scatter<<<sms * 8, 256>>>(pos, NPOS * 8, (uint32_t *)filters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms2 = time_ms(e0, e1);
        printf("random scatter into %4llu MiB: RED %.2f Gop/s   plain RMW %.2f Gop/s\n", (unsigned long long)mib, 4.0 * NPOS * 8 / ms / 1e6, 4.0 * NPOS * 8 / ms2 / 1e6);
    }
    gen_positions<<<sms * 8, 256>>>(pos, NPOS * 8, (1ull << 30) - 1, 12345);

    // ---- D: shared-memory tile scatter
    {
        CK(cudaFuncSetAttribute(smem_scatter<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        CK(cudaFuncSetAttribute(smem_scatter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        CK(cudaFuncSetAttribute(smem_scatter<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        const uint32_t per = 1 << 20;
        for (int mode = 0; mode < 3; mode++) {
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0));
                if (mode == 0) smem_scatter<0><<<sms, 512, 131072>>>(per, sink);
                if (mode == 1) smem_scatter<1><<<sms, 512, 131072>>>(per, sink);
                if (mode == 2) smem_scatter<2><<<sms, 512, 131072>>>(per, sink);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                if (rep) printf("smem scatter mode %d (0=atomicOr 1=ld/or/st 2=byte store): %.2f Gop/s chip-wide (includes mix64 ALU)\n", mode,
                                (double)per * sms / time_ms(e0, e1) / 1e6);
            }
        }
    }

    // ---- F: hash ALU
    {
        uint64_t n = 1ull << 30;
        hash_alu<<<sms * 8, 256>>>(n, (uint64_t *)sink);
        CK(cudaEventRecord(e0));
        hash_alu<<<sms * 8, 256>>>(n, (uint64_t *)sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1);
        printf("murmur-like hash ALU only: %.2f Ghash/s  (%.2f us per 5M)\n", n / ms / 1e6, 5e6 / (n / ms / 1e3));
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
