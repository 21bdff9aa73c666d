// Probe (measurement only): what the MMA-issuing thread of a K-chunked mainloop pays per 64-wide K chunk
// (4 x tcgen05.mma M x N x 16) on top of the tensor pipe's own time, for the loop shapes of gemm_fullk.cu:
//   mode 0  64 MMAs straight, descriptors = base + compile-time offsets, one commit at the end
//   mode 1  + one successful mbarrier.try_wait per chunk (barrier already complete)
//   mode 2  + one tcgen05.commit per chunk (to a spare barrier nobody waits on)
//   mode 3  both
//   mode 4  mode 0 + a second thread streaming 18 KB bulk copies into the B ring meanwhile
//   mode 8 + m: mode m with WARP-UNIFORM control flow (all lanes run the loop, elect.sync picks the issuer) and a lean wait
// Prints SM cycles (clock64) and ns (globaltimer) so that the SM clock during the run is visible too.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/probe_umma_chunks.cu -o build_tmp/probe_umma_chunks
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../ai-generated-gtav_b200/csrc/common.cuh"

using namespace gtav;

static constexpr int CHUNKS = 16, STAGES = 5;

__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mbarrier wait without the bounded-spin bookkeeping of common.cuh (whole warp may call it)
__device__ __forceinline__ void lean_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LD_%=;\n\t"
        "bra LW_%=;\n\t"
        "LD_%=:\n\t}\n"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int M, int N>
__global__ void __launch_bounds__(128, 1) chunk_kernel(long long* out, int mode, const uint8_t* gsrc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int A_CHUNK = M * 128, B_CHUNK = N * 128;
    uint8_t* sA = smem;
    uint8_t* sB = smem + CHUNKS * A_CHUNK;
    __shared__ uint64_t bar_done, bar_ready[STAGES], bar_spare[STAGES], bar_copy;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (CHUNKS * A_CHUNK + STAGES * B_CHUNK) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        mbar_init(&bar_done, 1);
        mbar_init(&bar_copy, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bar_ready[s], 1); mbar_init(&bar_spare[s], 1); }
        fence_barrier_init();
        for (int s = 0; s < STAGES; ++s) mbar_arrive(&bar_ready[s]);       // phase 0 complete: try_wait(parity 0) succeeds at once
    }
    if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 64 && mode == 4) {
        // background traffic into the last ring slot
        uint8_t* dst = sB + (STAGES - 1) * B_CHUNK;   // (races with the MMAs reading that slot: values do not matter here)
        for (int i = 0; i < 16; ++i) {
            mbar_arrive_expect_tx(&bar_copy, B_CHUNK);
            bulk_g2s(dst, gsrc + (i % 4) * B_CHUNK, B_CHUNK, &bar_copy);
            mbar_wait(&bar_copy, i & 1);
        }
    }
    if (threadIdx.x < 32 && (mode & 8)) {
        // warp-uniform control flow: all 32 lanes run the loop, one elected lane issues (the compiler can keep the
        // descriptors in uniform registers instead of moving them there per instruction)
        constexpr uint32_t idesc = umma_idesc_bf16(M, N);
        const uint64_t da0 = umma_desc_sw128(smem_u32(sA));
        const uint64_t db0 = umma_desc_sw128(smem_u32(sB));
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            const long long g0 = gtimer();
            uint64_t da = da0;
#pragma unroll 1
            for (int c0 = 0; c0 < CHUNKS; c0 += STAGES) {
#pragma unroll
                for (int s = 0; s < STAGES; ++s) {
                    const int c = c0 + s;
                    if (c < CHUNKS) {
                        if (mode & 1) lean_wait(&bar_ready[s], 0);
                        const uint64_t db = db0 + static_cast<uint64_t>(s * (B_CHUNK >> 4));
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, da + 2 * k, db + 2 * k, idesc, (c | k) != 0 ? 1u : 0u);
                            if (mode & 2) umma_commit(&bar_spare[s]);
                        }
                        __syncwarp();
                        da += A_CHUNK >> 4;
                    }
                }
            }
            const long long t1 = clock64();
            if (elect_one()) umma_commit(&bar_done);
            __syncwarp();
            lean_wait(&bar_done, rep & 1);
            if (threadIdx.x == 0) {
                out[rep * 3 + 0] = clock64() - t0;
                out[rep * 3 + 1] = gtimer() - g0;
                out[rep * 3 + 2] = t1 - t0;
            }
        }
    } else if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(M, N);
        const uint64_t da0 = umma_desc_sw128(smem_u32(sA));
        const uint64_t db0 = umma_desc_sw128(smem_u32(sB));
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            const long long g0 = gtimer();
            uint64_t da = da0;
#pragma unroll 1
            for (int c0 = 0; c0 < CHUNKS; c0 += STAGES) {
#pragma unroll
                for (int s = 0; s < STAGES; ++s) {
                    const int c = c0 + s;
                    if (c < CHUNKS) {
                        if (mode & 1) mbar_wait(&bar_ready[s], 0);
                        const uint64_t db = db0 + static_cast<uint64_t>(s * (B_CHUNK >> 4));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, da + 2 * k, db + 2 * k, idesc, (c | k) != 0 ? 1u : 0u);
                        da += A_CHUNK >> 4;
                        if (mode & 2) umma_commit(&bar_spare[s]);
                    }
                }
            }
            const long long t1 = clock64();
            umma_commit(&bar_done);
            mbar_wait(&bar_done, rep & 1);
            out[rep * 3 + 0] = clock64() - t0;
            out[rep * 3 + 1] = gtimer() - g0;
            out[rep * 3 + 2] = t1 - t0;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}

template <int M, int N>
static int run(long long* d, const uint8_t* gsrc) {
    const size_t smem = CHUNKS * M * 128 + STAGES * N * 128 + 1024;
    cudaFuncSetAttribute(chunk_kernel<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode : {0, 1, 2, 3, 8, 9, 10, 11}) {
        chunk_kernel<M, N><<<1, 128, smem>>>(d, mode, gsrc);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d N=%d mode %d failed: %s\n", M, N, mode, cudaGetErrorString(e)); return 1; }
        long long h[9];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("M=%3d N=%3d mode %d: 64 MMAs in %6lld cycles = %6lld ns (%.2f GHz) -> %6.1f cycles per MMA; the issuing thread was done after %6lld cycles\n",
               M, N, mode, h[6], h[7], double(h[6]) / double(h[7]), double(h[6]) / 64, h[8]);
    }
    return 0;
}

int main() {
    long long* d;
    cudaMalloc(&d, 128);
    uint8_t* g;
    cudaMalloc(&g, 4 * 144 * 128);
    cudaMemset(g, 0, 4 * 144 * 128);
    if (run<64, 144>(d, g)) return 1;
    if (run<64, 128>(d, g)) return 1;
    return 0;
}
