// Probe (measurement only): (1) cost of the P V product's MMAs - tcgen05.mma M=128, N=64, K=16, A K-major from three
// 16 KB atoms, B either K-major or MN-major (bit 16 of the instruction descriptor) stepping 2 KB per K step - issued by
// one thread, 12 per group with a commit + wait per group (as attn_tc.cu does) and 48 back to back;
// (2) MUFU rate of ex2.approx.ftz.f32 against ex2.approx.ftz.bf16x2 (two results per instruction) with 8 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/probe_umma_pv.cu -o build_tmp/probe_umma_pv
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../ai-generated-gtav_b200/csrc/common.cuh"

using namespace gtav;

__global__ void __launch_bounds__(128, 1) pv_kernel(long long* out, int N, int b_mn, int group, int groups) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                  // 3 atoms x [128 rows x 128 B]
    uint8_t* sB = smem + 3 * 16384;      // 576 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (3 * 16384 + 576 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N) | (b_mn ? (1u << 16) : 0u);
        uint32_t ph = 0;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int g = 0; g < groups; ++g) {
                for (int ks = 0; ks < group; ++ks) {
                    const int k12 = ks % 12;
                    const uint64_t da = umma_desc_sw128(smem_u32(sA + (k12 >> 2) * 16384)) + 2 * (k12 & 3);
                    const uint64_t db = b_mn ? umma_desc_sw128(smem_u32(sB + (ks % 36) * 16 * 128))
                                             : umma_desc_sw128(smem_u32(sB + (k12 >> 2) * 16384)) + 2 * (k12 & 3);
                    umma_bf16_ss(tm, da, db, idesc, ks != 0 ? 1u : 0u);
                }
                umma_commit(&bar);
                mbar_wait(&bar, ph & 1);
                ++ph;
            }
            out[rep] = clock64() - t0;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) ex2_kernel(float* out, long long* cyc, int iters) {
    float a[8];
    uint32_t u[8];
    for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); u[i] = 0xbc00bc00u + threadIdx.x + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            else asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    long long* d;
    cudaMalloc(&d, 4096);
    const int smem = 3 * 16384 + 576 * 128 + 2048;
    cudaFuncSetAttribute(pv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    struct Cfg { int N, mn, group, groups; };
    const Cfg cfgs[] = {{64, 0, 12, 8}, {64, 1, 12, 8}, {64, 0, 48, 2}, {64, 1, 48, 2}, {128, 0, 12, 8}, {128, 1, 12, 8}, {192, 0, 4, 8},
                        {64, 0, 1, 16}, {64, 1, 1, 16}, {80, 1, 12, 8}, {64, 1, 36, 4}};
    for (const Cfg& c : cfgs) {
        pv_kernel<<<1, 128, smem>>>(d, c.N, c.mn, c.group, c.groups);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d mn=%d failed: %s\n", c.N, c.mn, cudaGetErrorString(e)); return 1; }
        long long h[3];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("M=128 N=%3d B %s: %d groups of %2d MMAs, commit+wait per group: %7lld cycles -> %6.1f per group, %6.1f per MMA\n", c.N,
               c.mn ? "MN-major" : "K-major ", c.groups, c.group, h[2], double(h[2]) / c.groups, double(h[2]) / (c.groups * c.group));
    }
    float* f;
    cudaMalloc(&f, 148 * 256 * 4);
    const int iters = 4096;
    for (int mode = 0; mode < 2; ++mode) {
        if (mode == 0) ex2_kernel<0><<<148, 256>>>(f, d, iters); else ex2_kernel<1><<<148, 256>>>(f, d, iters);
        cudaDeviceSynchronize();
        long long h[1];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const double instr = 8.0 * iters * 256;     // thread-instructions per SM
        printf("ex2 %s: %lld cycles for %.0f thread-instr per SM -> %.2f thread-instr / clk / SM (%s results / clk)\n", mode ? "bf16x2" : "f32   ",
               h[0], instr, instr / h[0], mode ? "x2" : "x1");
    }
    return 0;
}
