// bf16 GEMM for sm_100a with the K dimension split over the two CTAs of a cluster: out[M,N] = epilogue(A[M,K] @ W[N,K]^T).
//
// Same contract, epilogues and rounding points as gemm_sm100.cu.  Why a third tiled kernel: Mlp.fc2 (K = 4096, N = 1024,
// reference timm Mlp / model/dit.py:171-198) at M <= 1152 rows (B <= 8 last-frame steps, the dense B = 1 window, the context
// pass) has too few output tiles for the machine, and each tile's mainloop is K-serial: 256 MMAs at the ~61-cycle
// floor of a small tcgen05.mma = 8 us whatever the tile width (in-graph 14.4 us, scripts/sweep_gemm_tiles.py --graph).
// Here two CTAs of a cluster compute the SAME 128 x 128 output tile over one half of K each, then reduce through distributed
// shared memory: CTA r owns output columns [64 r, 64 r + 64); four of its epilogue warps push the OTHER half of its fp32
// partial tile into the peer's staging buffer (st.shared::cluster, 32 KB) and arrive on the peer's mbarrier, the other four
// wait for the peer's half, add it to their own TMEM columns and run the fused epilogue.  The sum is partial(k < K/2) +
// partial(k >= K/2) in both CTAs (fp32 addition commutes), so the result does not depend on which CTA finishes a column.
//
// Roles per CTA (416 threads) as in gemm_sm100.cu: warps 0,7 TMA producers of A, warps 6,8 of W (ahead of the previous
// kernel), warp 1 TMEM allocator + MMA issue (warp-uniform loops, elect.sync), warps 2-5 / 9-12 the reduce + epilogue.
#include "common.cuh"
#include "kernels.h"
#include "epilogue.cuh"

namespace gtav {

static constexpr int SP_BM = 128, SP_BN = 128, SP_BK = 64, SP_KC = 2, SP_STAGES = 3;
static constexpr int SP_THREADS = 416;
static constexpr int SP_A_CHUNK = SP_BM * SP_BK * 2;                 // 16 KB
static constexpr int SP_B_CHUNK = SP_BN * SP_BK * 2;                 // 16 KB
static constexpr int SP_STAGE = SP_KC * (SP_A_CHUNK + SP_B_CHUNK);   // 64 KB
static constexpr int SP_STAGING = SP_BM * (SP_BN / 2) * 4;           // 32 KB: the peer's fp32 partial of this CTA's 64 columns
static constexpr int SP_BAR_OFF = SP_STAGES * SP_STAGE + SP_STAGING;
static constexpr int SP_SMEM = SP_BAR_OFF + (2 * SP_STAGES + 2) * 8 + 16 + 1024;

__device__ __forceinline__ uint32_t sp_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void sp_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t sp_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void sp_st_cluster_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sp_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire: the data behind this barrier was written by the peer CTA
__device__ __forceinline__ void sp_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > (1u << 26)) {
            printf("gtav: split-K staging wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(SP_THREADS, 1)
gemm_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                                                   // [stage][chunk][128 rows][128 B]
    uint8_t* sB = smem + SP_STAGES * SP_KC * SP_A_CHUNK;
    float* staging = reinterpret_cast<float*>(smem + SP_STAGES * SP_STAGE);   // [16 column quads][128 rows] float4
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SP_BAR_OFF);
    uint64_t* empty_bar = full_bar + SP_STAGES;
    uint64_t* acc_bar = empty_bar + SP_STAGES;        // this CTA's partial tile complete
    uint64_t* peer_bar = acc_bar + 1;                 // the peer's half of this CTA's columns has landed in `staging`
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = sp_cluster_ctarank();
    const int tile = blockIdx.x >> 1;
    const int n_tiles_n = p.N / SP_BN;
    const int m_blk = tile / n_tiles_n, n_blk = tile % n_tiles_n;
    const int num_ks = p.K / (2 * SP_KC * SP_BK);                         // stages per CTA (half of K)
    const int k0 = static_cast<int>(rank) * num_ks * SP_KC;               // first 64-wide chunk of this CTA's half

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmA);
    if (warp == 6 && lane == 0) tma_prefetch_desc(&tmB);
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < SP_STAGES; ++s) {
                mbar_init(&full_bar[s], 2 * SP_KC);
                mbar_init(&empty_bar[s], 1);
            }
            mbar_init(acc_bar, 1);
            mbar_init(peer_bar, 4);                   // the peer's four sender warps
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, SP_BN);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    sp_cluster_sync();                                // both CTAs' barriers exist before the peer may arrive on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == 0 || warp == 7) {
        const int j = warp == 0 ? 0 : 1;
        pdl_wait();                                   // A is the previous kernel's output
        int s = 0;
        uint32_t ph = 1;
        for (int ks = 0; ks < num_ks; ++ks) {
            mbar_wait(&empty_bar[s], ph);
            if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[s], SP_A_CHUNK);
                tma_load_2d(sA + (s * SP_KC + j) * SP_A_CHUNK, &tmA, &full_bar[s], (k0 + ks * SP_KC + j) * SP_BK, m_blk * SP_BM);
            }
            __syncwarp();
            if (++s == SP_STAGES) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 6 || warp == 8) {
        const int j = warp == 6 ? 0 : 1;
        int s = 0;
        uint32_t ph = 1;
        for (int ks = 0; ks < num_ks; ++ks) {
            mbar_wait(&empty_bar[s], ph);
            if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[s], SP_B_CHUNK);
                tma_load_2d(sB + (s * SP_KC + j) * SP_B_CHUNK, &tmB, &full_bar[s], (k0 + ks * SP_KC + j) * SP_BK, n_blk * SP_BN);
            }
            __syncwarp();
            if (++s == SP_STAGES) { s = 0; ph ^= 1u; }
        }
        pdl_wait();
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(SP_BM, SP_BN);
        const uint64_t da0 = umma_desc_sw128(smem_u32(sA));
        const uint64_t db0 = umma_desc_sw128(smem_u32(sB));
        int s = 0;
        uint32_t ph = 0;
        for (int ks = 0; ks < num_ks; ++ks) {
            mbar_wait(&full_bar[s], ph);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint64_t da = da0 + static_cast<uint64_t>(s * ((SP_KC * SP_A_CHUNK) >> 4));
                const uint64_t db = db0 + static_cast<uint64_t>(s * ((SP_KC * SP_B_CHUNK) >> 4));
#pragma unroll
                for (int c = 0; c < SP_KC; ++c) {
#pragma unroll
                    for (int k = 0; k < SP_BK / 16; ++k)
                        umma_bf16_ss(tmem_base, da + (c * (SP_A_CHUNK >> 4) + 2 * k), db + (c * (SP_B_CHUNK >> 4) + 2 * k), idesc,
                                     (ks | c | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            __syncwarp();
            if (++s == SP_STAGES) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) umma_commit(acc_bar);
        __syncwarp();
        pdl_wait();
    } else {
        pdl_wait();                                   // bias / gate / residual may come from the previous kernel
        const int q = warp & 3;                       // TMEM lane quadrant this warp may read
        const bool sender = warp >= 9;                // warps 9-12 ship the peer's columns, warps 2-5 finish this CTA's
        const int row_in_tile = q * 32 + lane;
        const int row = m_blk * SP_BM + row_in_tile;
        mbar_wait(acc_bar, 0);
        tcgen05_fence_after();
        if (sender) {
            const int cbase = static_cast<int>(1u - rank) * (SP_BN / 2);                  // the peer's column half of the tile
            // staging layout [16 column quads][128 rows] of float4: consecutive lanes (rows) write consecutive 16-byte words
            const uint32_t remote = sp_mapa(smem_u32(staging + row_in_tile * 4), 1u - rank);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cbase;
            uint32_t va[32], vb[32];
            tmem_ld_32x32(taddr, va);
            tmem_ld_32x32(taddr + 32, vb);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) sp_st_cluster_v4(remote + i * (SP_BM * 16), va[4 * i], va[4 * i + 1], va[4 * i + 2], va[4 * i + 3]);
#pragma unroll
            for (int i = 0; i < 8; ++i) sp_st_cluster_v4(remote + (8 + i) * (SP_BM * 16), vb[4 * i], vb[4 * i + 1], vb[4 * i + 2], vb[4 * i + 3]);
            __syncwarp();                             // every lane's stores before lane 0's release-arrive
            if (lane == 0) sp_arrive_cluster(sp_mapa(smem_u32(peer_bar), 1u - rank));
        } else {
            const int cbase = static_cast<int>(rank) * (SP_BN / 2);
            const bf16* gate_row = nullptr;
            if (EPI == EPI_BIAS_GATE_RES && row < p.M) {
                int f = row / p.rows_per_frame;
                if (p.frame_row != nullptr) f = p.frame_row[f];
                gate_row = p.gate + static_cast<size_t>(f) * p.gate_ld;
            }
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cbase;
            uint32_t va[32], vb[32];
            tmem_ld_32x32(taddr, va);
            tmem_ld_32x32(taddr + 32, vb);
            tmem_ld_wait();
            sp_wait_cluster(peer_bar, 0);
            const float4* mine = reinterpret_cast<const float4*>(staging) + row_in_tile;
            // partial(first half of K) + partial(second half of K): fp32 addition commutes, so both CTAs form the same sum
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 o = mine[i * SP_BM];
                va[4 * i] = __float_as_uint(__fadd_rn(__uint_as_float(va[4 * i]), o.x));
                va[4 * i + 1] = __float_as_uint(__fadd_rn(__uint_as_float(va[4 * i + 1]), o.y));
                va[4 * i + 2] = __float_as_uint(__fadd_rn(__uint_as_float(va[4 * i + 2]), o.z));
                va[4 * i + 3] = __float_as_uint(__fadd_rn(__uint_as_float(va[4 * i + 3]), o.w));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 o = mine[(8 + i) * SP_BM];
                vb[4 * i] = __float_as_uint(__fadd_rn(__uint_as_float(vb[4 * i]), o.x));
                vb[4 * i + 1] = __float_as_uint(__fadd_rn(__uint_as_float(vb[4 * i + 1]), o.y));
                vb[4 * i + 2] = __float_as_uint(__fadd_rn(__uint_as_float(vb[4 * i + 2]), o.z));
                vb[4 * i + 3] = __float_as_uint(__fadd_rn(__uint_as_float(vb[4 * i + 3]), o.w));
            }
            const int col0 = n_blk * SP_BN + cbase;
            if (row < p.M) {
                epilogue_chunk<EPI>(p, row, col0, va, gate_row);
                epilogue_chunk<EPI>(p, row, col0 + 32, vb, gate_row);
            }
        }
    }
    tcgen05_fence_before();
    sp_cluster_sync();                                // no CTA leaves while its peer may still write its staging / signal it
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, SP_BN);
    }
}

// ------------------------------------------------------------------------------------------ host
// K large enough for the K-serial mainloop to dominate, both halves whole pipeline stages, and all 2 x tiles CTAs in one round.
bool gemm_splitk_eligible(int M, int N, int K, int sms) {
    if (K < 2048 || K % (2 * SP_KC * SP_BK) != 0 || N % SP_BN != 0 || M <= 0) return false;
    const int tiles = ((M + SP_BM - 1) / SP_BM) * (N / SP_BN);
    return 2 * tiles <= sms;
}

template <int EPI>
static int launch_splitk(const GemmOp* op, cudaStream_t stream) {
    static bool configured = false;
    auto kern = gemm_splitk_kernel<EPI>;
    if (!configured) {
        GTAV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
        configured = true;
    }
    const int tiles = ((op->p.M + SP_BM - 1) / SP_BM) * (op->p.N / SP_BN);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * tiles);
    cfg.blockDim = dim3(SP_THREADS);
    cfg.dynamicSmemBytes = SP_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    GTAV_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, op->tmA, op->tmB, op->p));
    return 0;
}

int gemm_splitk_run(const GemmOp* op, cudaStream_t stream) {
    switch (op->epi) {
        case EPI_STORE: return launch_splitk<EPI_STORE>(op, stream);
        case EPI_BIAS: return launch_splitk<EPI_BIAS>(op, stream);
        case EPI_BIAS_GELU_TANH: return launch_splitk<EPI_BIAS_GELU_TANH>(op, stream);
        case EPI_BIAS_GELU_ERF: return launch_splitk<EPI_BIAS_GELU_ERF>(op, stream);
        case EPI_BIAS_SILU: return launch_splitk<EPI_BIAS_SILU>(op, stream);
        case EPI_BIAS_GATE_RES: return launch_splitk<EPI_BIAS_GATE_RES>(op, stream);
        case EPI_BIAS_RES: return launch_splitk<EPI_BIAS_RES>(op, stream);
        case EPI_BIAS_RES_SILU: return launch_splitk<EPI_BIAS_RES_SILU>(op, stream);
    }
    set_error("split-K gemm: unknown epilogue %d", op->epi);
    return -1;
}

}  // namespace gtav
