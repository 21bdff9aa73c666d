// Probe (measurement only): cost per tcgen05.mma (kind::f16, cta_group::1, both operands in shared memory) for the tile
// shapes a 144-token GEMM can use, issued back to back by one thread (64 MMAs accumulating into one tile, one commit).
// Also prints where M = 64 puts accumulator rows in TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/probe_umma_rate.cu -o build_tmp/probe_umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../ai-generated-gtav_b200/csrc/common.cuh"

using namespace gtav;

__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int M, int N, int n_mma, int two_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                  // 3 chunks x [256 rows x 128 B]
    uint8_t* sB = smem + 3 * 32768;      // 3 chunks x [256 rows x 128 B]
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 6 * 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(M, N);
        // descriptors precomputed; the timed loop is 16 fully unrolled MMAs per iteration (no address arithmetic), so
        // that what is measured is the tensor pipe, not the issuing thread's scalar code
        uint64_t da[16], db[16];
        for (int i = 0; i < 16; ++i) {
            const int c = (i >> 2) % 3, k = i & 3;
            da[i] = umma_desc_sw128(smem_u32(sA + c * 32768)) + 2 * k;
            db[i] = umma_desc_sw128(smem_u32(sB + c * 32768)) + 2 * k;
        }
        const int n_acc = two_tiles < 1 ? 1 : two_tiles;      // independent accumulators, round robin
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int it = 0; it < n_mma / 16; ++it) {
#pragma unroll
                for (int i = 0; i < 16; ++i) umma_bf16_ss(tm + (i & (n_acc - 1)) * 32, da[i], db[i], idesc, (it | (i >= n_acc)) ? 1u : 0u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, rep & 1);
            out[rep] = clock64() - t0;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// where does M = 64 put row i?  A[i][0] = i + 1, B[j][0] = 1
__global__ void __launch_bounds__(128, 1) layout_kernel(float* out, int M) {
    __shared__ __align__(1024) uint8_t sA[128 * 128];
    __shared__ __align__(1024) uint8_t sB[16 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 128 * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(sA)[i] = 0;
    for (int i = threadIdx.x; i < 16 * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(sB)[i] = 0;
    __syncthreads();
    if (threadIdx.x < M) *reinterpret_cast<bf16*>(sA + threadIdx.x * 128 + (threadIdx.x % 8) * 16) = __float2bfloat16(float(threadIdx.x + 1));
    if (threadIdx.x < 16) *reinterpret_cast<bf16*>(sB + threadIdx.x * 128 + (threadIdx.x % 8) * 16) = __float2bfloat16(1.0f);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&slot, 32); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        umma_bf16_ss(tm, umma_desc_sw128(smem_u32(sA)), umma_desc_sw128(smem_u32(sB)), umma_idesc_bf16(128, 16), 0u);   // clear
        umma_bf16_ss(tm, umma_desc_sw128(smem_u32(sA)), umma_desc_sw128(smem_u32(sB)), umma_idesc_bf16(M, 16), 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tcgen05_fence_after();
    uint32_t v[16];
    tmem_ld_32x16(tm + ((threadIdx.x / 32 * 32u) << 16), v);
    tmem_ld_wait();
    out[threadIdx.x] = __uint_as_float(v[0]);
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 32);
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768 + 2048);
    struct Cfg { int M, N, n, two; };
    const Cfg cfgs[] = {{128, 32, 64, 1}, {128, 32, 64, 2}, {128, 32, 64, 4}, {128, 32, 64, 8}, {128, 16, 64, 8}, {64, 32, 64, 8},
                        {128, 144, 64, 1}, {128, 144, 64, 2}, {64, 144, 64, 1}, {64, 144, 64, 2}, {128, 256, 64, 1}, {128, 256, 64, 2},
                        {128, 128, 64, 1}, {128, 128, 64, 2}, {128, 128, 64, 4}, {128, 64, 64, 4}, {128, 32, 16, 1}, {128, 144, 16, 1}};
    for (const Cfg& c : cfgs) {
        rate_kernel<<<1, 128, 6 * 32768 + 2048>>>(d, c.M, c.N, c.n, c.two);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d N=%d failed: %s\n", c.M, c.N, cudaGetErrorString(e)); return 1; }
        long long h[3];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const int total = c.n;
        printf("M=%3d N=%3d %s: %3d MMAs (K=16 each) in %6lld cycles -> %6.1f cycles per MMA\n", c.M, c.N,
               c.two <= 1 ? "1 accum  " : c.two == 2 ? "2 accums " : c.two == 4 ? "4 accums " : c.two == 8 ? "8 accums " : "16 accums", total, h[2], double(h[2]) / total);
    }
    float* f;
    cudaMalloc(&f, 128 * 4);
    layout_kernel<<<1, 128>>>(f, 64);
    cudaDeviceSynchronize();
    float hf[128];
    cudaMemcpy(hf, f, sizeof(hf), cudaMemcpyDeviceToHost);
    printf("M=64 layout: TMEM lane -> D row + 1 (0 = untouched)");
    for (int l = 0; l < 128; ++l) printf("%s%3d:%4.0f", (l % 16) ? " " : "\n", l, hf[l]);
    printf("\n");
    return 0;
}
