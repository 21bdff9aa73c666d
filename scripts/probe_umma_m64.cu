// Probe (measurement only): where does tcgen05.mma cta_group::1 with M = 64 put accumulator row i in TMEM?
// A[i][0] = i + 1, B[j][0] = 1  =>  D[i][j] = i + 1.  Prints, for every TMEM lane, the value found in column 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/probe_umma_m64.cu -o scripts/build/probe_umma_m64
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../ai-generated-gtav_b200/csrc/common.cuh"

using namespace gtav;

__global__ void __launch_bounds__(128, 1) probe(float* out, int M) {
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
    // clear the accumulator columns first so untouched lanes read as 0: D = 0 * 0 with M = 128
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(M, 16);
        umma_bf16_ss(tm, umma_desc_sw128(smem_u32(sA)), umma_desc_sw128(smem_u32(sB)), idesc, 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tcgen05_fence_after();
    uint32_t v[16];
    tmem_ld_32x16(tm + ((threadIdx.x / 32 * 32u) << 16), v);
    tmem_ld_wait();
    out[threadIdx.x * 2] = __uint_as_float(v[0]);
    out[threadIdx.x * 2 + 1] = __uint_as_float(v[15]);
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 32);
}

int main() {
    float* d;
    cudaMalloc(&d, 256 * 4);
    for (int M : {128, 64}) {
        cudaMemset(d, 0xff, 256 * 4);
        probe<<<1, 128>>>(d, M);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d failed: %s\n", M, cudaGetErrorString(e)); return 1; }
        float h[256];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("M=%d: lane -> D row + 1 (column 0 | column 15)\n", M);
        for (int l = 0; l < 128; ++l) printf("%s%3d:%5.0f|%5.0f", (l % 8) ? "  " : "\n", l, h[2 * l], h[2 * l + 1]);
        printf("\n");
    }
    return 0;
}
