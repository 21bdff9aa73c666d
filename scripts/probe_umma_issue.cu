// Probe (measurement only): what the MMA-issuing thread itself pays - cycles to ISSUE n tcgen05.mma (M=128, N=96 / 192, K=16,
// descriptor = base + constant), cycles per tcgen05.commit, and the completion latency seen through the mbarrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/probe_umma_issue.cu -o build_tmp/probe_umma_issue
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../ai-generated-gtav_b200/csrc/common.cuh"

using namespace gtav;

template <int N, int NM>
__global__ void __launch_bounds__(128, 1) issue_kernel(long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                  // 128 rows x 128 B
    uint8_t* sB = smem + 16384;          // 256 rows x 128 B
    __shared__ uint64_t bar[8];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N);
        const uint64_t da = umma_desc_sw128(smem_u32(sA)), db = umma_desc_sw128(smem_u32(sB));
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < NM; ++i) umma_bf16_ss(tm + (i & 1) * 256, da + 2 * (i & 3), db + 2 * (i & 3), idesc, (i >= 2) ? 1u : 0u);
            const long long t1 = clock64();
            umma_commit(&bar[0]);
            const long long t2 = clock64();
            umma_commit(&bar[1]);
            umma_commit(&bar[2]);
            umma_commit(&bar[3]);
            const long long t3 = clock64();
            mbar_wait(&bar[0], rep & 1);
            const long long t4 = clock64();
            mbar_wait(&bar[1], rep & 1); mbar_wait(&bar[2], rep & 1); mbar_wait(&bar[3], rep & 1);
            // polling cost: 16 test_wait probes on a barrier that is not complete
            const long long t5 = clock64();
            uint32_t acc = 0;
#pragma unroll 1
            for (int i = 0; i < 16; ++i) acc += mbar_try_wait(&bar[4], 0) ? 1u : 0u;
            const long long t6 = clock64();
            out[rep * 8 + 0] = t1 - t0; out[rep * 8 + 1] = t2 - t1; out[rep * 8 + 2] = t3 - t2; out[rep * 8 + 3] = t4 - t3;
            out[rep * 8 + 4] = t4 - t0; out[rep * 8 + 5] = t6 - t5; out[rep * 8 + 6] = acc;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, int NM>
void run(long long* d) {
    cudaFuncSetAttribute(issue_kernel<N, NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 2048);
    issue_kernel<N, NM><<<1, 128, 16384 + 32768 + 2048>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return; }
    long long h[24];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%3d, %2d MMAs: issue %5lld cycles (%5.1f per MMA), first commit %4lld, 3 more commits %4lld, wait after issue %5lld, total %5lld (%5.1f per MMA); 16 failing try_wait probes %5lld\n",
           N, NM, h[16], double(h[16]) / NM, h[17], h[18], h[19], h[20], double(h[20]) / NM, h[21]);
}

int main() {
    long long* d;
    cudaMalloc(&d, 4096);
    run<96, 4>(d); run<96, 16>(d); run<192, 4>(d); run<192, 16>(d); run<64, 12>(d); run<64, 6>(d); run<144, 4>(d);
    return 0;
}
