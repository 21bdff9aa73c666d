// Probe: how fast can every SM of a cluster ingest the SAME activation slab from L2 - unicast (each CTA loads all of it)
// vs TMA multicast (each CTA loads 1/CS of it and multicasts to the whole cluster)?  Decides whether a no-split-K
// weight-stationary GEMM (every CTA needs the whole [144, K] activation matrix) is viable at M = 144.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_tmp/probe_mcast scripts/probe_mcast.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <vector>

__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load_mcast(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}

// total: bytes every CTA must end up with; piece: bytes per copy instruction; issuers: threads issuing copies
__global__ void __launch_bounds__(128, 1) ingest_kernel(const uint8_t* src, long long* stamps, int CS, int mcast, int total, int piece,
                                                        int issuers, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + total);
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long best = 1ll << 60;
    for (int rep = 0; rep < reps; ++rep) {
        cluster_sync();
        const long long t0 = gtimer();
        if (threadIdx.x == 0) mbar_expect(bar, total);
        __syncthreads();
        cluster_sync();                       // every CTA's barrier is armed before anyone multicasts into it
        if (mcast) {
            const int share = total / CS;     // this CTA's part of the slab
            const int pieces = share / piece;
            for (int i = threadIdx.x; i < pieces; i += issuers) {
                if (threadIdx.x < issuers) {
                    const int off = rank * share + i * piece;
                    bulk_load_mcast(smem + off, src + off, piece, bar, static_cast<uint16_t>((1u << CS) - 1));
                }
            }
        } else {
            const int pieces = total / piece;
            for (int i = threadIdx.x; i < pieces; i += issuers)
                if (threadIdx.x < issuers) bulk_load(smem + i * piece, src + i * piece, piece, bar);
        }
        const long long t1 = gtimer();
        mbar_wait(bar, rep & 1);
        const long long t2 = gtimer();
        if (t2 - t1 < best) best = t2 - t1;
        (void)t0;
    }
    if (threadIdx.x == 0) stamps[blockIdx.x] = best;
}

static void run(const uint8_t* src, int CS, int mcast, int total, int piece, int issuers, int grid) {
    long long* stamps;
    cudaMalloc(&stamps, grid * sizeof(long long));
    const int smem = total + 64;
    cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (CS > 8) cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, ingest_kernel, src, stamps, CS, mcast, total, piece, issuers, 6);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("CS=%d mcast=%d total=%d piece=%d: %s\n", CS, mcast, total, piece, cudaGetErrorString(e));
        cudaGetLastError();
        return;
    }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), stamps, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const long long med = h[grid / 2], mx = h.back();
    printf("cluster %2d %-9s slab %3d KB, %5.1f KB/copy, %2d issuing threads, %3d CTAs: %5lld ns median %5lld max -> %6.1f GB/s per SM, "
           "%5.2f TB/s delivered, %5.2f TB/s from L2\n", CS, mcast ? "multicast" : "unicast", total / 1024, piece / 1024.0, issuers, grid,
           med, mx, total / (double)med, total / (double)med * grid / 1000.0, total / (double)med * grid / 1000.0 / (mcast ? CS : 1));
    cudaFree(stamps);
}

int main() {
    uint8_t* src;
    cudaMalloc(&src, 1 << 20);
    cudaMemset(src, 1, 1 << 20);
    const int total = 144 * 1024;               // half of a [144, 1024] bf16 activation matrix
    for (int grid : {128}) {
        run(src, 1, 0, total, 16384, 8, grid);
        run(src, 1, 0, total, 36864, 4, grid);
        run(src, 8, 0, total, 18432, 8, grid);
        run(src, 8, 1, total, 18432, 1, grid);
        run(src, 8, 1, total, 9216, 2, grid);
        run(src, 8, 1, total, 4608, 4, grid);
        run(src, 4, 1, total, 18432, 2, grid);
        run(src, 4, 1, total, 9216, 4, grid);
        run(src, 16, 1, total, 9216, 1, grid);
        run(src, 2, 1, total, 18432, 4, grid);
    }
    return 0;
}
