// Probe: cost of an all-to-all split-K exchange through distributed shared memory inside a thread-block cluster
// (what a cluster-resident reduction of fp32 partial tiles would pay), vs the same exchange through L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_dsmem scripts/probe_dsmem.cu && ./probe_dsmem
// Each CTA owns a 72 KB fp32 tile [144 tokens][128 rows]; CTA r of the cluster must end up with the sum over the
// cluster of the token slice r.  DSMEM variant: every CTA stores the slices of its peers straight into their shared
// memory (st.shared::cluster), cluster barrier, local reduce.  Times are per phase, from %globaltimer.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace cg = cooperative_groups;

__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(unsigned addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(unsigned addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// tile rows: ROWS weight rows (128 or 64); tokens 144; S = cluster size = number of K splits
template <int VEC>
__global__ void __launch_bounds__(256, 1) exchange_kernel(long long* stamps, float* out, int S, int rows, int reps) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* recv = reinterpret_cast<float*>(smem);                 // [S][per][rows]
    const int tid = threadIdx.x;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int per = 144 / S;
    const int slice = per * rows;                                  // floats per (source, destination) pair
    const unsigned recv_base = smem_u32(recv);
    float acc = 0.f;
    for (int rep = 0; rep < reps; ++rep) {
        cluster_arrive();
        cluster_wait();
        const long long t0 = gtimer();
        // "drain": every thread produces values for all tokens of its rows and sends each token slice to its owner
        for (int d = 0; d < S; ++d) {
            const unsigned dst = mapa(recv_base, d) + static_cast<unsigned>(rank) * slice * 4;
            if (VEC == 1) {
                for (int i = tid; i < slice; i += 256) st_cluster_f32(dst + i * 4, static_cast<float>(i + rep));
            } else {
                for (int i = tid; i < slice / 4; i += 256)
                    st_cluster_v4(dst + i * 16, make_float4(i + rep, i, rep, 1.f));
            }
        }
        const long long t1 = gtimer();
        cluster_arrive();
        cluster_wait();
        const long long t2 = gtimer();
        // local reduce of the S slices
        for (int i = tid; i < slice; i += 256) {
            float a = 0.f;
            for (int s2 = 0; s2 < S; ++s2) a += recv[s2 * slice + i];
            acc += a;
        }
        const long long t3 = gtimer();
        if (tid == 0 && rep == reps - 1) {
            long long* st = stamps + blockIdx.x * 4;
            st[0] = t1 - t0; st[1] = t2 - t1; st[2] = t3 - t2; st[3] = t3 - t0;
        }
    }
    if (acc == 12345.678f) out[blockIdx.x] = acc;
}

template <int VEC>
static void run(int S, int rows, int grid) {
    long long* stamps;
    float* out;
    cudaMalloc(&stamps, grid * 4 * sizeof(long long));
    cudaMalloc(&out, grid * sizeof(float));
    const int smem = 144 * rows * 4 + 130 * 1024;      // the receive tile + what the GEMM's operands would occupy
    auto kern = exchange_kernel<VEC>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (S > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int reps = 5;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, stamps, out, S, rows, reps);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("S=%2d rows=%3d vec=%d grid=%d smem=%d: %s\n", S, rows, VEC, grid, smem, cudaGetErrorString(e));
        cudaGetLastError();
        return;
    }
    std::vector<long long> h(grid * 4);
    cudaMemcpy(h.data(), stamps, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long med[4], mx[4];
    for (int k = 0; k < 4; ++k) {
        std::vector<long long> v;
        for (int b = 0; b < grid; ++b) v.push_back(h[b * 4 + k]);
        std::sort(v.begin(), v.end());
        med[k] = v[v.size() / 2];
        mx[k] = v.back();
    }
    printf("S=%2d rows=%3d vec=%d grid=%3d: send %lld/%lld ns, barrier %lld/%lld, local reduce %lld/%lld, total %lld/%lld (median/max), "
           "%.1f KB out per CTA\n", S, rows, VEC, grid, med[0], mx[0], med[1], mx[1], med[2], mx[2], med[3], mx[3],
           144.0 * rows * 4 / 1024 * (S - 1) / S);
    cudaFree(stamps);
    cudaFree(out);
}

int main() {
    for (int grid : {128}) {
        run<1>(4, 128, grid);
        run<4>(4, 128, grid);
        run<1>(8, 64, grid);
        run<4>(8, 64, grid);
        run<4>(8, 128, grid);
        run<1>(16, 128, grid);
        run<4>(16, 128, grid);
        run<4>(2, 128, grid);
    }
    return 0;
}
