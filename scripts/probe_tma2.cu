// Micro-benchmark #2 (measurement only): why does one CTA ingest only ~40 GB/s with 16 KB TMA boxes?
// Variants: number of producer threads, box size via a 3-D tensor map (up to 128 KB per instruction),
// tensor map in global memory vs kernel parameter, plain 1-D bulk copies of pre-tiled data.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/probe_tma2.cu -o scripts/build/probe_tma2 -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../ai-generated-gtav_b200/csrc/common.cuh"

using namespace gtav;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// mode 0: 2-D tensor loads (64 x rows);  mode 1: 3-D tensor loads (64 x rows x kc);  mode 2: 1-D bulk copies of
// stage_bytes;  tm_g != nullptr: descriptor read from global memory instead of the kernel parameter.
// np producer warps (lane 0 of warps 0..np-1), consumer = lane 0 of warp 4.
__global__ void __launch_bounds__(160, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm, const CUtensorMap* tm_g, const uint8_t* base, int mode, int np,
             int iters, int stages, int rows, int kc, long region_bytes, long tile_offset) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = rows * 128 * kc;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
    uint64_t* empty = full + stages;
    const CUtensorMap* tmp = tm_g ? tm_g : &tm;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_barrier_init();
        tma_prefetch_desc(tmp);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long tiles_in_region = region_bytes / stage_bytes;
    if (warp < np && lane == 0) {
        for (int i = warp; i < iters; i += np) {
            const int s = i % stages;
            mbar_wait(&empty[s], ((i / stages) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[s], stage_bytes);
            const long tile = (static_cast<long>(blockIdx.x) * iters + i + tile_offset) % tiles_in_region;
            if (mode == 2) {
                bulk_load_1d(smem + s * stage_bytes, base + tile * stage_bytes, stage_bytes, &full[s]);
            } else if (mode == 1) {
                // tensor viewed as [row][kchunk][64]: rows of 1024 bf16 = 16 chunks of 64
                const long t = tile * kc;                     // in units of (rows x 64) tiles
                tma_load_3d(smem + s * stage_bytes, tmp, &full[s], 0, static_cast<int>(t % 16), static_cast<int>((t / 16) * rows));
            } else {
                const long t = tile;
                tma_load_2d(smem + s * stage_bytes, tmp, &full[s], static_cast<int>(t % 16) * 64, static_cast<int>((t / 16) * rows));
            }
        }
    } else if (warp == 4 && lane == 0) {
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
    const long total_rows = 1 << 20;                   // x 1024 bf16 = 2 GiB
    uint8_t* buf = nullptr;
    if (cudaMalloc(&buf, total_rows * 2048) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, total_rows * 2048);
    CUtensorMap* tm_dev = nullptr;
    cudaMalloc(&tm_dev, sizeof(CUtensorMap));
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%-6s %-7s %4s %5s %3s %3s %6s %8s %9s %9s\n", "src", "mode", "ctas", "rows", "kc", "np", "stages", "KB/instr", "GB/s", "GB/s/SM");
    struct Cfg { int mode, rows, kc, np, stages, desc_global; };
    const Cfg cfgs[] = {
        {0, 128, 1, 1, 4, 0}, {0, 128, 1, 2, 4, 0}, {0, 128, 1, 4, 8, 0}, {0, 128, 1, 1, 4, 1}, {0, 256, 1, 1, 4, 0},
        {0, 256, 1, 2, 4, 0}, {0, 256, 1, 4, 6, 0}, {1, 128, 2, 1, 4, 0}, {1, 128, 4, 1, 3, 0}, {1, 256, 2, 1, 3, 0},
        {1, 128, 4, 2, 3, 0}, {1, 64, 4, 1, 4, 0},  {2, 128, 1, 1, 4, 0}, {2, 256, 1, 1, 4, 0}, {2, 256, 2, 1, 3, 0},
        {2, 256, 2, 2, 3, 0}, {2, 64, 1, 1, 8, 0},  {2, 64, 1, 4, 8, 0},
    };
    for (int src = 0; src < 2; ++src) {                // 0: HBM stream (2 GiB region), 1: L2-resident (8 MiB region)
        for (const Cfg& c : cfgs) {
            CUtensorMap tm;
            if (c.mode == 1) {
                cuuint64_t gdim[3] = {64, 16, static_cast<cuuint64_t>(total_rows)};
                cuuint64_t gstr[2] = {128, 2048};
                cuuint32_t box[3] = {64, static_cast<cuuint32_t>(c.kc), static_cast<cuuint32_t>(c.rows)};
                cuuint32_t estr[3] = {1, 1, 1};
                // smem image: [row][kchunk][64]?  no - box order is (inner 64, kchunk, row); fine for a bandwidth probe
                if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                    printf("encode3d failed\n");
                    continue;
                }
            } else {
                cuuint64_t gdim[2] = {1024, static_cast<cuuint64_t>(total_rows)};
                cuuint64_t gstr[1] = {2048};
                cuuint32_t box[2] = {64, static_cast<cuuint32_t>(c.rows)};
                cuuint32_t estr[2] = {1, 1};
                if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                    printf("encode2d failed\n");
                    continue;
                }
            }
            cudaMemcpy(tm_dev, &tm, sizeof(tm), cudaMemcpyHostToDevice);
            const int stage_bytes = c.rows * 128 * c.kc;
            if (c.stages * stage_bytes > 216 * 1024) { printf("skip (smem)\n"); continue; }
            for (int ctas : {16, 64, 148}) {
                const int iters = (8 << 20) / stage_bytes;             // 8 MB per CTA
                const size_t smem = c.stages * stage_bytes + 2 * c.stages * 8 + 1024 + 64;
                const long region = src == 0 ? total_rows * 2048 : (8L << 20);
                float best = 1e30f;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaEventRecord(e0);
                    probe_kernel<<<ctas, 160, smem>>>(tm, c.desc_global ? tm_dev : nullptr, buf, c.mode, c.np, iters, c.stages,
                                                       c.rows, c.kc, region, static_cast<long>(rep) * ctas * iters);
                    cudaEventRecord(e1);
                    cudaError_t e = cudaEventSynchronize(e1);
                    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
                    float ms = 0;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (rep > 0 && ms < best) best = ms;
                }
                const double bytes = static_cast<double>(ctas) * iters * stage_bytes;
                const double gbs = bytes / (best * 1e-3) / 1e9;
                printf("%-6s %-7s %4d %5d %3d %3d %6d %8d %9.1f %9.1f\n", src == 0 ? "hbm" : "l2",
                       c.mode == 0 ? (c.desc_global ? "2d-gdes" : "2d") : (c.mode == 1 ? "3d" : "bulk1d"), ctas, c.rows, c.kc, c.np,
                       c.stages, stage_bytes / 1024, gbs, gbs / ctas);
            }
        }
    }
    return 0;
}
