// Persistent last-frame DiT step: the 16 SpatioTemporalDiTBlocks (32 half-blocks) of reference model/dit.py:200-225 on
// the 144 tokens of the frame being denoised, as ONE kernel with one CTA per SM.
//
// Why: at M = 144 rows every op of the step is a few microseconds of work behind a chain of fixed latencies (kernel
// boundary, HBM round trip for the first weight bytes, split-K exchange).  Measured on B200 (profiles/r01), the 7
// kernels of a half-block cost 46 us although their 24 MB of weights stream in 3.7 us.  This kernel keeps the same
// arithmetic (same rounding points, same split order) and removes the fixed costs:
//   * one launch per step; phases are separated by grid barriers (one atomic + one acquire-poll per CTA) instead of
//     kernel boundaries;
//   * weights never wait: each CTA knows its (weight-row block, K split) of every future GEMM, so a TMA producer
//     streams the slab of phase g+2 into a 2-slot shared-memory ring while phase g computes - HBM latency is off the
//     critical path;
//   * LayerNorm + adaLN modulate (dit.py:19-27) is fused into the A-operand fill of the qkv / fc1 GEMMs: the epilogue
//     that produces the residual stream also emits per-row (mean, M2) partials over its 128 columns, the consumer
//     merges the 8 partials (Chan's parallel variance) and normalises while it copies its K slice into the
//     128-byte-swizzled UMMA layout - two kernels and two passes over h per half-block disappear;
//   * attention runs as a phase of the same kernel (spatial: the mma.sync body of attn_seq.cuh; temporal: one warp per
//     (position, head) against the cached context K/V).
// GEMM phases are the weight-streaming decomposition of gemm_skinny.cu: UMMA M side = 128 weight rows, N side = the
// 144 tokens, K split over S CTAs (qkv 24x4, out 8x16, fc1 32x4, fc2 8x16 CTAs), fp32 partials through L2, per
// row-block rendezvous, in-order (deterministic) reduction with the fused bias / GELU-tanh / gate*y+residual epilogue.
//
// Warps: 0-7 workers (A fill, TMEM drain, reduce + epilogue, attention, barriers), 8 = TMA weight producer + MMA issuer.
// All cross-CTA data (h, qkv, att, mlp, partials, stats) is read with ld.global.cg and published with
// fence.acq_rel.gpu + barrier, counters are monotonic within a launch and reset by the last CTA to leave.
#include "attn_seq.cuh"
#include "common.cuh"
#include "kernels.h"

namespace gtav {

static constexpr int MG_WORKERS = 256;
static constexpr int MG_THREADS = MG_WORKERS + 32;
static constexpr int MG_TOK = 144;
static constexpr int MG_D = 1024;
static constexpr int MG_W_CHUNK = 128 * 128;                 // 128 weight rows x 64 bf16
static constexpr int MG_A_CHUNK = MG_TOK * 128;              // 144 tokens x 64 bf16
static constexpr int MG_W_SLOT = 4 * MG_W_CHUNK;             // 64 KB
static constexpr int MG_A_BUF = 4 * MG_A_CHUNK;              // 72 KB
static constexpr int MG_TAIL = 2048;                         // mbarriers, TMEM slot, merged LayerNorm stats [144]
// Per-half constant vectors this CTA needs (its K slice of shift/scale, its 128 columns of bias/gate), staged with
// cp.async one half-block ahead: they are HBM-cold every step (805 MB of weights pass through L2 in between).
static constexpr int MG_V_SHIFT1 = 0, MG_V_SCALE1 = 512, MG_V_SHIFT2 = 1024, MG_V_SCALE2 = 1536, MG_V_OUTB = 2048,
                     MG_V_GATE1 = 2304, MG_V_FC1B = 2560, MG_V_FC2B = 2816, MG_V_GATE2 = 3072, MG_VEC_BYTES = 3328;
static constexpr int MG_SMEM = 2 * MG_W_SLOT + MG_A_BUF + MG_TAIL + 2 * MG_VEC_BYTES + 1024;
static constexpr int MG_MAX_SPLIT = 16;

// the four GEMMs of a half-block
struct MegaKind { int N, K, S, chunks, rbs; };
__host__ __device__ constexpr MegaKind mega_kind(int k) {
    return k == 0 ? MegaKind{3 * MG_D, MG_D, 4, 4, 24}        // to_qkv
         : k == 1 ? MegaKind{MG_D, MG_D, 16, 1, 8}            // to_out
         : k == 2 ? MegaKind{4 * MG_D, MG_D, 4, 4, 32}        // fc1
                  : MegaKind{MG_D, 4 * MG_D, 16, 4, 8};       // fc2
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

// Wait until *ctr >= target (one thread), bounded so that a protocol bug traps instead of hanging the box.
__device__ __forceinline__ void spin_until(const unsigned* ctr, unsigned target, int what) {
    uint32_t spins = 0;
    while (ld_acquire_u32(ctr) < target) {
        if (++spins > (1u << 25)) {
            printf("gtav: step kernel wait %d timed out (block %d, target %u, seen %u)\n", what, blockIdx.x, target, ld_acquire_u32(ctr));
            __trap();
        }
    }
}

// Grid barrier for the workers of every CTA: release (each thread fences its own global writes), arrive, acquire.
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target, int tid) {
    fence_gpu();
    worker_bar();
    if (tid == 0) {
        atomicAdd(ctr, 1u);
        spin_until(ctr, target, 0);
    }
    worker_bar();
}

// Chan et al. merge of 8 (mean, M2) partials over 128 elements each -> (mean, rstd) of the 1024-wide row, eps 1e-6.
__device__ __forceinline__ float2 merge_stats(const float2* st8) {
    float2 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = __ldcg(st8 + i);
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) mean += p[i].x;
    mean *= 0.125f;
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float d = p[i].x - mean;
        m2 += p[i].y + 128.f * d * d;
    }
    return make_float2(mean, rsqrtf(m2 * (1.0f / MG_D) + 1e-6f));
}

// (mean, M2) of the 128 values held 4 per lane by one warp
__device__ __forceinline__ float2 warp_stats128(const float (&y)[4]) {
    const float mean = warp_sum(y[0] + y[1] + y[2] + y[3]) * (1.0f / 128.f);
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float d = y[j] - mean;
        m2 += d * d;
    }
    return make_float2(mean, warp_sum(m2));
}

__device__ __forceinline__ long long mg_timer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// profiling aid: thread 0 of every CTA stamps phase boundaries of half-blocks 2 (spatial) and 3 (temporal)
#define MG_STAMP(s, half, slot)                                                                              \
    do {                                                                                                     \
        if ((s).pp->trace != nullptr && (s).tid == 0 && ((half) == 2 || (half) == 3))                        \
            (s).pp->trace[(static_cast<size_t>((s).cta) * 2 + ((half) - 2)) * 32 + (slot)] = mg_timer();     \
    } while (0)

struct MegaStep {
    const MegaParams* pp;
    uint8_t *sW, *sA;
    uint64_t *bar_w, *bar_a, *bar_acc;
    float2* s_stat;
    uint8_t* s_vec;               // [2][MG_VEC_BYTES]
    uint64_t* bar_at;             // TMA-filled A operand (to_out / fc2)
    uint32_t tmem;
    int cta, tid, warp, lane;
    const bf16* mod_row;          // modulation vectors of this step's conditioning row
    uint32_t acc_uses;            // completed uses of bar_acc (its wait parity)
    unsigned grid_epoch;          // grid barriers passed
};

// ---------------------------------------------------------------------------------------------------- per-half vectors
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Stage the constant vectors of half-block `half` this CTA will touch into s_vec[half & 1] (workers 0..207, one 16-byte
// cp.async each; committed as one group per call by every worker).
__device__ __forceinline__ void stage_vectors(MegaStep& s, int half) {
    const MegaParams& p = *s.pp;
    if (half < p.n_halves && s.tid < MG_VEC_BYTES / 16) {
        const MegaHalfDev& hd = p.halves[half];
        const bf16* mrow = s.mod_row + hd.mod_off;
        const int t = s.tid;
        const int k4 = (s.cta & 3) * 256;            // K slice of the S = 4 GEMMs (qkv, fc1)
        const int n16 = (s.cta >> 4) * 128;          // weight-row block of the S = 16 GEMMs (out, fc2)
        const int n4 = (s.cta >> 2) * 128;           // weight-row block of fc1
        const bf16* src;
        if (t < 32) src = mrow + k4 + t * 8;                                         // shift_msa
        else if (t < 64) src = mrow + MG_D + k4 + (t - 32) * 8;                      // scale_msa
        else if (t < 96) src = mrow + 3 * MG_D + k4 + (t - 64) * 8;                  // shift_mlp
        else if (t < 128) src = mrow + 4 * MG_D + k4 + (t - 96) * 8;                 // scale_mlp
        else if (t < 144) src = hd.out_b + n16 + (t - 128) * 8;
        else if (t < 160) src = mrow + 2 * MG_D + n16 + (t - 144) * 8;               // gate_msa
        else if (t < 176) src = hd.fc1_b + n4 + (t - 160) * 8;
        else if (t < 192) src = hd.fc2_b + n16 + (t - 176) * 8;
        else src = mrow + 5 * MG_D + n16 + (t - 192) * 8;                            // gate_mlp
        cp_async16(s.s_vec + (half & 1) * MG_VEC_BYTES + t * 16, src);
    }
    cp_async_commit();
}

// ---------------------------------------------------------------------------------------------------- A operand fill
// LayerNorm + modulate fused into the copy of this CTA's K slice of h [144, 1024] into the 128-byte-swizzled
// [chunk][144 x 64] layout the UMMA descriptor expects (what TMA SWIZZLE_128B would have produced):
// (x - mean) * rstd * bf16(1 + bf16(scale + 1e-6)) + shift, one bf16 rounding, as ln_rows_kernel.
// One mbarrier arrival per thread and chunk: the MMA issuer starts on chunk c while chunk c+1 is being written.
__device__ __forceinline__ void fill_a_ln(MegaStep& s, int kcol0, const uint8_t* v_shift, const uint8_t* v_scale) {
    const int tid = s.tid;
    const int j = tid & 7;                    // 16-byte column group inside a 64-wide chunk
    const int r0 = tid >> 3;                  // rows r0, r0 + 32, ... (5 per chunk, the last one partial)
    const bf16* src = s.pp->h + kcol0 + j * 8;
    uint4 raw[4][5];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int r = r0 + 32 * i;
            if (r < MG_TOK) raw[c][i] = __ldcg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(r) * MG_D + c * 64));
        }
    }
    // merged (mean, rstd) of every token once per CTA: thread t < 144 merges token t, everyone reads it from smem
    if (tid < MG_TOK) s.s_stat[tid] = merge_stats(s.pp->stats + tid * 8);
    cp_async_wait<1>();                       // this half's vectors (the group issued one half-block ago) have landed
    worker_bar();
    float2 st[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int r = r0 + 32 * i;
        if (r < MG_TOK) st[i] = s.s_stat[r];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float mul[8], add[8];
        const uint4 sh = *reinterpret_cast<const uint4*>(v_shift + c * 128 + j * 16);
        const uint4 sc = *reinterpret_cast<const uint4*>(v_scale + c * 128 + j * 16);
        const uint32_t shw[4] = {sh.x, sh.y, sh.z, sh.w}, scw[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 s2 = unpack_bf16x2(shw[q]), c2 = unpack_bf16x2(scw[q]);
            mul[2 * q] = bf16_round(1.0f + bf16_round(c2.x + 1e-6f));
            mul[2 * q + 1] = bf16_round(1.0f + bf16_round(c2.y + 1e-6f));
            add[2 * q] = s2.x;
            add[2 * q + 1] = s2.y;
        }
        uint8_t* dst = s.sA + c * MG_A_CHUNK;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int r = r0 + 32 * i;
            if (r < MG_TOK) {
                const uint4 v = raw[c][i];
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                uint32_t o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 x = unpack_bf16x2(w[q]);
                    o[q] = pack_bf16x2((x.x - st[i].x) * st[i].y * mul[2 * q] + add[2 * q],
                                       (x.y - st[i].x) * st[i].y * mul[2 * q + 1] + add[2 * q + 1]);
                }
                *reinterpret_cast<uint4*>(dst + r * 128 + ((j ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        fence_proxy_async_smem();             // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&s.bar_a[c]);
    }
}

// A operand that needs no transform (attention output, GELU(fc1)): one TMA box, issued by one thread.  The data was
// written with generic-proxy stores by other CTAs and published through the grid barrier this thread has just
// passed; the proxy fence orders that acquire before the async-proxy (TMA) read.
__device__ __forceinline__ void fill_a_tma(MegaStep& s, const CUtensorMap* tm, int chunk0, int chunks) {
    if (s.tid == 0) {
        asm volatile("fence.proxy.async.global;" ::: "memory");
        mbar_arrive_expect_tx(s.bar_at, chunks * MG_A_CHUNK);
        tma_load_3d(s.sA, tm, s.bar_at, 0, 0, chunk0);
    }
}

// ---------------------------------------------------------------------------------------------------- GEMM phase (workers)
template <int KIND>
__device__ __forceinline__ void gemm_phase_workers(MegaStep& s, int half) {
    constexpr MegaKind kd = mega_kind(KIND);
    const MegaParams& p = *s.pp;
    const int rb = s.cta / kd.S, split = s.cta - rb * kd.S;
    const uint8_t* vec = s.s_vec + (half & 1) * MG_VEC_BYTES;
    constexpr int TS = 1 + KIND * 7;           // trace slots of this phase
    MG_STAMP(s, half, TS + 0);
    // ---- A operand
    if (KIND == 0) fill_a_ln(s, split * 256, vec + MG_V_SHIFT1, vec + MG_V_SCALE1);
    else if (KIND == 1) fill_a_tma(s, &p.tm_att, split * kd.chunks, kd.chunks);
    else if (KIND == 2) fill_a_ln(s, split * 256, vec + MG_V_SHIFT2, vec + MG_V_SCALE2);
    else fill_a_tma(s, &p.tm_mlp, split * kd.chunks, kd.chunks);
    MG_STAMP(s, half, TS + 1);
    // ---- what the epilogue needs besides the partials: residual rows (L2), bias / gate (smem) - requested now, used
    // after the rendezvous
    constexpr int per = MG_TOK / kd.S;                      // tokens this CTA reduces
    constexpr int ITER = (per + 7) / 8;                     // per warp
    const int lo = split * per, hi = lo + per;
    const int n = rb * 128 + 4 * s.lane;
    uint2 rv[ITER];
    if (KIND == 1 || KIND == 3) {
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int tok = lo + s.warp + 8 * it;
            if (tok < hi) rv[it] = __ldcg(reinterpret_cast<const uint2*>(p.h + static_cast<size_t>(tok) * MG_D + n));
        }
    }
    // ---- accumulator -> fp32 partial tile in the workspace: ws[cta][token][128 weight rows]
    mbar_wait(s.bar_acc, s.acc_uses & 1);
    s.acc_uses++;
    tcgen05_fence_after();
    MG_STAMP(s, half, TS + 2);
    {
        const int q = s.warp & 3, hsel = s.warp >> 2;
        const int row = q * 32 + s.lane;
        float* mine = p.ws + static_cast<size_t>(s.cta) * MG_TOK * 128 + row;
        const uint32_t tl = s.tmem + (static_cast<uint32_t>(q * 32) << 16) + hsel * 72;
        uint32_t v[9][8];
#pragma unroll
        for (int c = 0; c < 9; ++c) tmem_ld_32x8(tl + c * 8, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 9; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) __stcg(mine + static_cast<size_t>(hsel * 72 + c * 8 + i) * 128, __uint_as_float(v[c][i]));
    }
    tcgen05_fence_before();
    MG_STAMP(s, half, TS + 3);
    // ---- rendezvous of the S CTAs of this row block
    fence_gpu();
    worker_bar();
    unsigned* rdv = p.sync + 8 + KIND * 32 + rb;
    if (s.tid == 0) {
        atomicAdd(rdv, 1u);
        spin_until(rdv, static_cast<unsigned>(kd.S) * (half + 1), 1 + KIND);
    }
    worker_bar();
    MG_STAMP(s, half, TS + 4);
    // ---- reduce my share of the tokens in split order + fused epilogue; warp per token, lane = 4 weight rows.
    // All partial loads of a warp are issued before the first is consumed.
    const float* part = p.ws + static_cast<size_t>(rb) * kd.S * MG_TOK * 128 + 4 * s.lane;
    float4 acc[ITER];
    if (kd.S <= 4) {
        float4 v[ITER][kd.S];
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int tok = lo + s.warp + 8 * it;
#pragma unroll
            for (int s2 = 0; s2 < kd.S; ++s2)
                if (tok < hi) v[it][s2] = __ldcg(reinterpret_cast<const float4*>(part + (static_cast<size_t>(s2) * MG_TOK + tok) * 128));
        }
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            acc[it] = v[it][0];
#pragma unroll
            for (int s2 = 1; s2 < kd.S; ++s2) { acc[it].x += v[it][s2].x; acc[it].y += v[it][s2].y; acc[it].z += v[it][s2].z; acc[it].w += v[it][s2].w; }
        }
    } else {
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int tok = lo + s.warp + 8 * it;
            if (tok < hi) {
                float4 v[kd.S];
#pragma unroll
                for (int s2 = 0; s2 < kd.S; ++s2)
                    v[s2] = __ldcg(reinterpret_cast<const float4*>(part + (static_cast<size_t>(s2) * MG_TOK + tok) * 128));
                acc[it] = v[0];
#pragma unroll
                for (int s2 = 1; s2 < kd.S; ++s2) { acc[it].x += v[s2].x; acc[it].y += v[s2].y; acc[it].z += v[s2].z; acc[it].w += v[s2].w; }
            }
        }
    }
    float b4[4] = {0.f, 0.f, 0.f, 0.f}, g4[4] = {0.f, 0.f, 0.f, 0.f};
    if (KIND != 0) {
        const uint2 bv = *reinterpret_cast<const uint2*>(vec + (KIND == 1 ? MG_V_OUTB : KIND == 2 ? MG_V_FC1B : MG_V_FC2B) + 8 * s.lane);
        const float2 b0 = unpack_bf16x2(bv.x), b1 = unpack_bf16x2(bv.y);
        b4[0] = b0.x; b4[1] = b0.y; b4[2] = b1.x; b4[3] = b1.y;
    }
    if (KIND == 1 || KIND == 3) {
        const uint2 gv = *reinterpret_cast<const uint2*>(vec + (KIND == 1 ? MG_V_GATE1 : MG_V_GATE2) + 8 * s.lane);
        const float2 g0 = unpack_bf16x2(gv.x), g1 = unpack_bf16x2(gv.y);
        g4[0] = g0.x; g4[1] = g0.y; g4[2] = g1.x; g4[3] = g1.y;
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int tok = lo + s.warp + 8 * it;
        if (tok < hi) {
            float y[4] = {acc[it].x, acc[it].y, acc[it].z, acc[it].w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) y[jj] = bf16_round(y[jj] + b4[jj]);          // the Linear's own bf16 output
            bf16* outp;
            if (KIND == 0) {
                outp = p.qkv + static_cast<size_t>(tok) * (3 * MG_D) + n;
            } else if (KIND == 2) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) y[jj] = gelu_tanh_f(y[jj]);
                outp = p.mlp + static_cast<size_t>(tok) * (4 * MG_D) + n;
            } else {
                const float2 r0 = unpack_bf16x2(rv[it].x), r1 = unpack_bf16x2(rv[it].y);
                const float r4[4] = {r0.x, r0.y, r1.x, r1.y};
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) y[jj] = bf16_round(r4[jj] + bf16_round(g4[jj] * y[jj]));
                outp = p.h + static_cast<size_t>(tok) * MG_D + n;
                const float2 st = warp_stats128(y);                              // LayerNorm partials of the new residual row
                if (s.lane == 0) __stcg(p.stats + tok * 8 + rb, st);
            }
            uint2 o;
            o.x = pack_bf16x2(y[0], y[1]);
            o.y = pack_bf16x2(y[2], y[3]);
            __stcg(reinterpret_cast<uint2*>(outp), o);
        }
    }
    MG_STAMP(s, half, TS + 5);
}

// ---------------------------------------------------------------------------------------------------- attention phases
// Spatial attention (attention.py:99-129): 48 work items (16 heads x 3 blocks of 48 queries), CTAs 0..47.  All 8 worker
// warps stage the head's rotated K, V and the item's rotated Q into the A buffer (16-byte ld.global.cg, every load in
// flight at once), then 3 warps run the mma.sync body of attn_seq.cuh on the staged data.
__device__ __forceinline__ void spatial_attention_phase(MegaStep& s) {
    if (s.cta >= 48) return;
    const MegaParams& p = *s.pp;
    const int head = s.cta / 3, qb = s.cta % 3;
    bf16* sK = reinterpret_cast<bf16*>(s.sA);
    bf16* sV = sK + MG_TOK * SROW;
    bf16* sQ = sV + MG_TOK * SROW;
    constexpr int NV = (2 * MG_TOK + 48) * 8;                 // 16-byte vectors: K, V (144 rows each), Q (48 rows)
    constexpr int IT = (NV + MG_WORKERS - 1) / MG_WORKERS;    // 11
    uint4 raw[IT];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int i = s.tid + it * MG_WORKERS;
        if (i < NV) {
            const int r = i >> 3, c8 = (i & 7) * 8;
            const bf16* src = r < MG_TOK       ? p.qkv + static_cast<size_t>(r) * (3 * MG_D) + MG_D + head * 64 + c8
                              : r < 2 * MG_TOK ? p.qkv + static_cast<size_t>(r - MG_TOK) * (3 * MG_D) + 2 * MG_D + head * 64 + c8
                                               : p.qkv + static_cast<size_t>(r - 2 * MG_TOK + qb * 48) * (3 * MG_D) + head * 64 + c8;
            raw[it] = __ldcg(reinterpret_cast<const uint4*>(src));
        }
    }
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int i = s.tid + it * MG_WORKERS;
        if (i < NV) {
            const int r = i >> 3, c8 = (i & 7) * 8;
            uint32_t w[4] = {raw[it].x, raw[it].y, raw[it].z, raw[it].w};
            if (r < MG_TOK || r >= 2 * MG_TOK) {               // K and Q rows get the axial rotary embedding
                const int tok = r < MG_TOK ? r : r - 2 * MG_TOK + qb * 48;
#pragma unroll
                for (int jx = 0; jx < 4; ++jx) w[jx] = rotate_pair(w[jx], p.rot_s[tok * 32 + (c8 >> 1) + jx]);
            }
            bf16* dst = r < MG_TOK ? sK + r * SROW : r < 2 * MG_TOK ? sV + (r - MG_TOK) * SROW : sQ + (r - 2 * MG_TOK) * SROW;
            *reinterpret_cast<uint4*>(dst + c8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    worker_bar();
    if (s.warp < 3)
        attn_seq_body<144, 144, 3, 32, true, true>(p.qkv, p.att, 16, p.rot_s, sK, sV, qb, head, 0, s.tid, 2, sQ);
}

// Temporal attention of the frame being denoised (attention.py:41-66): one warp per (position, head); query at window
// position TC, keys / values = the TC cached context frames + itself; same arithmetic and order as
// attn_temporal_last_kernel.  A warp's problems are processed together so that their loads overlap.
__device__ __forceinline__ void temporal_attention_phase(MegaStep& s, const bf16* cache) {
    const MegaParams& p = *s.pp;
    const int TC = p.ctx_frames;
    const int lane = s.lane;
    constexpr int NP = 3;                                      // 2304 problems over 128 x 8 warps
    uint32_t kc[NP][7], vc[NP][7], qr[NP], kr[NP], vr[NP];
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        const int prob = s.cta * 8 + s.warp + u * p.grid * 8;
        if (prob < MG_TOK * 16) {
            const int head = prob & 15, pos = prob >> 4;
#pragma unroll
            for (int t = 0; t < 7; ++t) {
                if (t < TC) {
                    const bf16* c = cache + (static_cast<size_t>(t) * MG_TOK + pos) * (2 * MG_D) + head * 64 + 2 * lane;
                    kc[u][t] = *reinterpret_cast<const uint32_t*>(c);
                    vc[u][t] = *reinterpret_cast<const uint32_t*>(c + MG_D);
                }
            }
            const bf16* base = p.qkv + static_cast<size_t>(pos) * (3 * MG_D) + head * 64 + 2 * lane;
            qr[u] = __ldcg(reinterpret_cast<const unsigned int*>(base));
            kr[u] = __ldcg(reinterpret_cast<const unsigned int*>(base + MG_D));
            vr[u] = __ldcg(reinterpret_cast<const unsigned int*>(base + 2 * MG_D));
        }
    }
    const float2 cs = p.rot_t[TC * 32 + lane];
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        const int prob = s.cta * 8 + s.warp + u * p.grid * 8;
        if (prob < MG_TOK * 16) {
            const int head = prob & 15, pos = prob >> 4;
            const float2 qx = unpack_bf16x2(qr[u]), kx = unpack_bf16x2(kr[u]), vx = unpack_bf16x2(vr[u]);
            const float2 q = make_float2(bf16_round(qx.x * cs.x - qx.y * cs.y), bf16_round(qx.y * cs.x + qx.x * cs.y));
            const float2 kn = make_float2(bf16_round(kx.x * cs.x - kx.y * cs.y), bf16_round(kx.y * cs.x + kx.x * cs.y));
            float sc[8];
            float m = -INFINITY;
#pragma unroll
            for (int jx = 0; jx < 8; ++jx) {
                if (jx <= TC) {
                    const float2 kk = (jx == TC || jx == 7) ? kn : unpack_bf16x2(kc[u][jx < 7 ? jx : 0]);
                    sc[jx] = warp_sum(q.x * kk.x + q.y * kk.y) * 0.125f;
                    m = fmaxf(m, sc[jx]);
                }
            }
            float l = 0.f;
#pragma unroll
            for (int jx = 0; jx < 8; ++jx) {
                if (jx <= TC) {
                    sc[jx] = __expf(sc[jx] - m);
                    l += sc[jx];
                }
            }
            const float inv = 1.0f / l;
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int jx = 0; jx < 8; ++jx) {
                if (jx <= TC) {
                    const float2 vv = (jx == TC || jx == 7) ? vx : unpack_bf16x2(vc[u][jx < 7 ? jx : 0]);
                    const float pr = bf16_round(sc[jx] * inv);
                    acc.x += pr * vv.x;
                    acc.y += pr * vv.y;
                }
            }
            __stcg(reinterpret_cast<unsigned int*>(p.att + static_cast<size_t>(pos) * MG_D + head * 64 + 2 * lane), pack_bf16x2(acc.x, acc.y));
        }
    }
}

// ---------------------------------------------------------------------------------------------------- the kernel
__device__ __forceinline__ bool mega_active(int cta, int g) {
    const MegaKind kd = mega_kind(g & 3);
    return cta < kd.rbs * kd.S;
}

__global__ void __launch_bounds__(MG_THREADS, 1) dit_step_mega_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    MegaStep s;
    s.pp = &p;
    s.sW = smem;
    s.sA = smem + 2 * MG_W_SLOT;
    s.bar_w = reinterpret_cast<uint64_t*>(s.sA + MG_A_BUF);
    s.bar_a = s.bar_w + 2;
    s.bar_acc = s.bar_w + 6;
    s.bar_at = s.bar_w + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s.bar_w + 8);
    s.s_stat = reinterpret_cast<float2*>(s.sA + MG_A_BUF + 128);
    s.s_vec = s.sA + MG_A_BUF + MG_TAIL;
    s.tid = threadIdx.x;
    s.warp = threadIdx.x >> 5;
    s.lane = threadIdx.x & 31;
    s.cta = blockIdx.x;
    s.acc_uses = 0;
    s.grid_epoch = 0;
    const int n_phases = 4 * p.n_halves;

    if (s.warp == 8) {
        if (s.lane == 0) {
            tma_prefetch_desc(&p.tm_att);
            tma_prefetch_desc(&p.tm_mlp);
            mbar_init(&s.bar_w[0], 1);
            mbar_init(&s.bar_w[1], 1);
            for (int c = 0; c < 4; ++c) mbar_init(&s.bar_a[c], MG_WORKERS);
            mbar_init(s.bar_acc, 1);
            mbar_init(s.bar_at, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    s.tmem = *tmem_slot;
    s.mod_row = p.mod + static_cast<size_t>(p.last_row[0]) * p.mod_ld;

    if (s.warp == 8) {
        // ======================= weight producer + MMA issuer (lane 0); L2 prefetcher (lanes 1..31) =======================
        if (s.lane == 0) {
            uint32_t w_par = 0, a_par = 0, at_par = 0;          // wait parities: bit s of w_par = ring slot s, bit c of a_par = chunk c
            auto issue_w = [&](int g) {
                const MegaKind kd = mega_kind(g & 3);
                const int rb = s.cta / kd.S, split = s.cta - rb * kd.S;
                const CUtensorMap* tm = &p.halves[g >> 2].tm[g & 3];
                mbar_arrive_expect_tx(&s.bar_w[g & 1], kd.chunks * MG_W_CHUNK);
                tma_load_3d(s.sW + (g & 1) * MG_W_SLOT, tm, &s.bar_w[g & 1], 0, rb * 128, split * kd.chunks);
            };
            for (int g = 0; g < 2 && g < n_phases; ++g)
                if (mega_active(s.cta, g)) issue_w(g);
            constexpr uint32_t idesc = umma_idesc_bf16(128, MG_TOK);
            for (int g = 0; g < n_phases; ++g) {
                if (mega_active(s.cta, g)) {
                    const MegaKind kd = mega_kind(g & 3);
                    const int slot = g & 1;
                    const bool tma_a = (g & 1) != 0;               // to_out and fc2 take their A operand by TMA
                    mbar_wait(&s.bar_w[slot], (w_par >> slot) & 1);
                    w_par ^= 1u << slot;
                    if (tma_a) {
                        mbar_wait(s.bar_at, at_par);
                        at_par ^= 1u;
                        tcgen05_fence_after();
                    }
                    for (int c = 0; c < kd.chunks; ++c) {
                        if (!tma_a) {
                            mbar_wait(&s.bar_a[c], (a_par >> c) & 1);
                            a_par ^= 1u << c;
                            tcgen05_fence_after();
                        }
                        const uint64_t dw = umma_desc_sw128(smem_u32(s.sW + slot * MG_W_SLOT + c * MG_W_CHUNK));
                        const uint64_t da = umma_desc_sw128(smem_u32(s.sA + c * MG_A_CHUNK));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_ss(s.tmem, dw + 2 * k, da + 2 * k, idesc, (c | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(s.bar_acc);
                    mbar_wait(s.bar_acc, s.acc_uses & 1);          // MMAs done: ring slot and A buffer are free again
                    s.acc_uses++;
                }
                if (g + 2 < n_phases && mega_active(s.cta, g + 2)) issue_w(g + 2);
            }
        } else if (p.ctx_frames > 0) {
            // pull the context K/V of every temporal layer through L2 shortly before its half-block needs it: it was
            // written by the context pass and has been evicted by the 805 MB of weights of the previous step
            const size_t layer_bytes = static_cast<size_t>(p.ctx_frames) * MG_TOK * 2 * MG_D * sizeof(bf16);
            const size_t lines = layer_bytes >> 7;
            const int part = s.cta * 31 + (s.lane - 1), parts = p.grid * 31;
            unsigned seen = 0;
            for (int layer = 0; layer < p.n_halves / 2; ++layer) {
                // the grid-barrier counter is the clock: layer L is consumed in half-block 2L+1, fetch it during 2L
                const unsigned want = static_cast<unsigned>(p.grid) * (1u + 5u * 2u * layer);
                bool give_up = false;
                for (uint32_t polls = 0; seen < want; ++polls) {
                    __nanosleep(2000);
                    const unsigned now = *reinterpret_cast<volatile const unsigned*>(p.sync);
                    if (now < seen || polls > 4096) { give_up = true; break; }     // counters reset (kernel is ending) / stuck
                    seen = now;
                }
                if (give_up) break;
                const char* base = reinterpret_cast<const char*>(p.kv_cache + static_cast<size_t>(layer) * p.cache_layer_stride);
                for (size_t l = part; l < lines; l += parts) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (l << 7)));
            }
        }
    } else {
        // ======================= workers =======================
        stage_vectors(s, 0);
        // LayerNorm partials of the incoming residual stream (patch-embed output): warp per row, 8 groups of 128
        for (int row = s.cta * 8 + s.warp; row < MG_TOK; row += p.grid * 8) {
            uint2 raw[8];
#pragma unroll
            for (int gi = 0; gi < 8; ++gi) raw[gi] = __ldcg(reinterpret_cast<const uint2*>(p.h + static_cast<size_t>(row) * MG_D + gi * 128 + 4 * s.lane));
#pragma unroll
            for (int gi = 0; gi < 8; ++gi) {
                const float2 a = unpack_bf16x2(raw[gi].x), b = unpack_bf16x2(raw[gi].y);
                const float y[4] = {a.x, a.y, b.x, b.y};
                const float2 st = warp_stats128(y);
                if (s.lane == 0) __stcg(p.stats + row * 8 + gi, st);
            }
        }
        grid_barrier(p.sync, ++s.grid_epoch * p.grid, s.tid);

        for (int half = 0; half < p.n_halves; ++half) {
            MG_STAMP(s, half, 0);
            stage_vectors(s, half + 1);                    // next half-block's vectors; this half's were issued one half ago
            if (mega_active(s.cta, 0)) {
                gemm_phase_workers<0>(s, half);
            } else {
                cp_async_wait<1>();                        // CTAs without a qkv tile still need this half's vectors later
            }
            grid_barrier(p.sync, ++s.grid_epoch * p.grid, s.tid);
            MG_STAMP(s, half, 7);
            if (half & 1) temporal_attention_phase(s, p.kv_cache + static_cast<size_t>(half >> 1) * p.cache_layer_stride);
            else spatial_attention_phase(s);
            MG_STAMP(s, half, 29);
            grid_barrier(p.sync, ++s.grid_epoch * p.grid, s.tid);
            MG_STAMP(s, half, 30);
            if (mega_active(s.cta, 1)) gemm_phase_workers<1>(s, half);
            grid_barrier(p.sync, ++s.grid_epoch * p.grid, s.tid);
            MG_STAMP(s, half, 14);
            if (mega_active(s.cta, 2)) gemm_phase_workers<2>(s, half);
            grid_barrier(p.sync, ++s.grid_epoch * p.grid, s.tid);
            MG_STAMP(s, half, 21);
            if (mega_active(s.cta, 3)) gemm_phase_workers<3>(s, half);
            grid_barrier(p.sync, ++s.grid_epoch * p.grid, s.tid);
            MG_STAMP(s, half, 28);
        }
        cp_async_wait<0>();
        // leave the counters zeroed for the next launch: the last CTA out resets them (everyone is past every wait)
        if (s.tid == 0) {
            const unsigned old = atomicAdd(p.sync + 1, 1u);
            if (old == static_cast<unsigned>(p.grid) - 1) {
                for (int i = 0; i < 8 + 4 * 32; ++i) p.sync[i] = 0;
                __threadfence();
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (s.warp == 8) tmem_dealloc(s.tmem, 256);
}

// ---------------------------------------------------------------------------------------------------- host
size_t mega_sync_bytes() { return (8 + 4 * 32) * sizeof(unsigned); }
size_t mega_stats_bytes() { return static_cast<size_t>(MG_TOK) * 8 * sizeof(float2); }
size_t mega_ws_bytes() { return static_cast<size_t>(128) * MG_TOK * 128 * sizeof(float); }
int mega_grid() { return 128; }

int mega_make_weight_map(CUtensorMap* out, const bf16* W, int kind) {
    const MegaKind kd = mega_kind(kind);
    return make_tmap_3d(out, W, kd.N, kd.K, kd.K, 128, kd.chunks);
}

int mega_make_act_maps(MegaParams* p) {
    int rc = make_tmap_3d(&p->tm_att, p->att, MG_TOK, MG_D, MG_D, MG_TOK, mega_kind(1).chunks);
    if (rc) return rc;
    return make_tmap_3d(&p->tm_mlp, p->mlp, MG_TOK, 4 * MG_D, 4 * MG_D, MG_TOK, mega_kind(3).chunks);
}

int mega_run(const MegaParams& p, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        GTAV_CUDA_OK(cudaFuncSetAttribute(dit_step_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM));
        configured = true;
    }
    int dev = 0, sms = 0;
    GTAV_CUDA_OK(cudaGetDevice(&dev));
    GTAV_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sms < p.grid) {
        set_error("step kernel: needs %d co-resident CTAs (one per SM), device has %d SMs", p.grid, sms);
        return -1;
    }
    // Plain launch (no programmatic dependent launch: the kernel reads the previous kernel's output at once).  All
    // CTAs are co-resident because grid <= SM count and one CTA fits per SM; waits are bounded (trap, not hang).
    dit_step_mega_kernel<<<dim3(p.grid), dim3(MG_THREADS), MG_SMEM, stream>>>(p);
    GTAV_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gtav
