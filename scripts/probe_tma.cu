// Micro-benchmark (measurement only, not product code): per-SM and chip-wide TMA ingest bandwidth on B200 as a
// function of CTA count, pipeline depth and whether the source is L2-resident or streamed from HBM.
// The numbers size the weight-streaming GEMM of the last-frame DiT step ("skinny GEMM").
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/probe_tma.cu -o scripts/build/probe_tma -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../ai-generated-gtav_b200/csrc/common.cuh"

using namespace gtav;

// Each CTA streams `iters` boxes of (64 x rows) bf16 through a ring of `stages` slots; thread 0 produces,
// thread 32 consumes (waits for the bytes, releases the slot).  CTA c reads rows starting at
// (c * iters + i) * rows modulo total_rows, so with a small total_rows everything hits L2.
__global__ void __launch_bounds__(64, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm, int iters, int stages, int rows, int kcols, int total_rows,
             int same_tile, long tile_offset) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = rows * 128;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
    uint64_t* empty = full + stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_barrier_init();
        tma_prefetch_desc(&tm);
    }
    __syncthreads();
    const int kblocks = kcols / 64;
    if (threadIdx.x == 0) {
        for (int i = 0; i < iters; ++i) {
            const int s = i % stages;
            mbar_wait(&empty[s], ((i / stages) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[s], stage_bytes);
            long tile = (same_tile ? i : static_cast<long>(blockIdx.x) * iters + i) + tile_offset;
            const int kb = static_cast<int>(tile % kblocks);
            const long rb = (tile / kblocks) % (total_rows / rows);
            tma_load_2d(smem + s * stage_bytes, &tm, &full[s], kb * 64, static_cast<int>(rb * rows));
        }
    } else if (threadIdx.x == 32) {
        for (int i = 0; i < iters; ++i) {
            const int s = i % stages;
            mbar_wait(&full[s], (i / stages) & 1);
            mbar_arrive(&empty[s]);
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
    const int kcols = 1024;
    const long big_rows = 1 << 20;                     // 1 Mi rows x 1024 x 2 B = 2 GiB
    bf16* buf = nullptr;
    if (cudaMalloc(&buf, big_rows * kcols * 2) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, big_rows * kcols * 2);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%-8s %5s %5s %6s %6s %10s %10s\n", "source", "ctas", "rows", "stages", "KB/ring", "GB/s", "GB/s/SM");
    const int row_opts[] = {128, 256};
    const int cta_opts[] = {8, 16, 32, 48, 64, 96, 128, 148};
    for (int src = 0; src < 3; ++src) {                // 0: HBM stream, 1: L2-resident distinct tiles, 2: all CTAs same tiles (L2)
        for (int rows : row_opts) {
            const long total_rows = src == 0 ? big_rows : 4096;        // 4096 x 2 KB = 8 MB: L2 resident
            CUtensorMap tm;
            cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kcols), static_cast<cuuint64_t>(total_rows)};
            cuuint64_t gstr[1] = {static_cast<cuuint64_t>(kcols) * 2};
            cuuint32_t box[2] = {64, static_cast<cuuint32_t>(rows)};
            cuuint32_t estr[2] = {1, 1};
            if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                printf("encode failed\n");
                return 1;
            }
            for (int stages : {2, 4, 6}) {
                if (stages * rows * 128 > 200 * 1024) continue;
                for (int ctas : cta_opts) {
                    const int iters = 512 * 128 / rows;                // 8 MB per CTA
                    const size_t smem = stages * rows * 128 + 2 * stages * 8 + 1024 + 64;
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        probe_kernel<<<ctas, 64, smem>>>(tm, iters, stages, rows, kcols, static_cast<int>(total_rows), src == 2,
                                                           src == 0 ? static_cast<long>(rep) * ctas * iters : 0L);
                        cudaEventRecord(e1);
                        cudaError_t e = cudaEventSynchronize(e1);
                        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    float ms = 0;
                    cudaEventElapsedTime(&ms, e0, e1);
                    const double bytes = static_cast<double>(ctas) * iters * rows * 128;
                    const double gbs = bytes / (ms * 1e-3) / 1e9;
                    printf("%-8s %5d %5d %6d %6d %10.1f %10.1f\n", src == 0 ? "hbm" : (src == 1 ? "l2" : "l2same"), ctas, rows,
                           stages, stages * rows * 128 / 1024, gbs, gbs / ctas);
                }
            }
        }
    }
    return 0;
}
