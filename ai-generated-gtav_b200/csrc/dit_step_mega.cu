// Persistent last-frame DiT step: the 16 SpatioTemporalDiTBlocks (32 half-blocks) of reference model/dit.py:200-225 on
// the 144 tokens of the frame being denoised (B = 1), as ONE kernel with one CTA per SM.
//
// At M = 144 rows every op of a step is a few microseconds of work behind fixed latencies.  Measured on B200
// (profiles/r01): the PDL-chained kernels spend ~46 us per half-block although its 24 MB of weights stream in 3.7 us; a
// first persistent version that kept the split-K decomposition was no faster, because the fp32 partial exchange
// (through L2: ~4 us, through DSMEM: 2.5-4 us, scripts/probe_dsmem.cu) and the activation ingest dominate.  What the
// probes also showed (scripts/probe_mcast.cu): plain 1-D bulk copies (cp.async.bulk) feed an SM at ~250-330 GB/s, 3-10x
// what tensor-map boxes or register copies reached - IF the source is already laid out as the shared-memory image.
// Hence this design:
//   * every activation that is a GEMM operand lives in global memory PRE-TILED: [K/64 chunks][144 rows][64] bf16 with
//     the 128-byte swizzle of an UMMA K-major operand, written that way by the producing epilogue; a consumer streams
//     the whole [144, K] matrix through a 4-stage shared-memory ring with one 18 KB bulk copy per chunk;
//   * with activation ingest cheap there is no reason to split K: a CTA owns 32 (to_qkv, fc1, fc2) or 16 (to_out)
//     weight rows x the whole K = 1024 (a 64 KB slab), tokens are the UMMA M side (tile 1 = tokens 0..127, tile 2 =
//     tokens 16..143 of which rows 112..127 = tokens 128..143 are kept), the accumulators [144 x 32] fp32 sit in 64
//     TMEM columns and the epilogue needs no exchange.  Only fc2 (K = 4096) splits K four ways; its partials are
//     summed by the row phase below, in split order (deterministic);
//   * weights never wait: the slab of GEMM g+2 is requested (TMA, 3-D box) as soon as GEMM g has released its ring
//     slot - HBM latency is off the critical path;
//   * residual update + LayerNorm + adaLN modulate (dit.py:19-27,207-223) are a row phase: CTA r owns token row r,
//     sums what to_out / fc2 left for that row, applies bias, gate and residual with the reference's bf16 rounding
//     points, then normalises (two-pass fp32 statistics, as ln_rows_kernel) and writes the next GEMM's operand pre-tiled;
//   * the rotary embedding moves into the to_qkv epilogue (same fp32 rotate + one bf16 rounding as
//     apply_rotary_emb), which writes q, k, v per head in the row-padded layout the attention body wants, so spatial
//     attention stages K, V, Q with three bulk copies; temporal attention is one warp per (position, head).
// STATUS: opt-in (GTAV_MEGA=1), parity-green, measured 52 us (spatial) / 61 us (temporal) per half-block against 46 us
// for the PDL-chained kernels (profiles/r01/step_kernel_trace_v3.txt): a GEMM phase costs ~7.8 us (128 small MMAs at
// >= 61 cycles each, scripts/probe_umma_rate.cu, plus single-thread issue overhead) and each of the 7 grid barriers
// ~2 us (128 same-address atomics + store drain).  Kept as the starting point for a tokens-as-N variant (64 MMAs).
// Phases are separated by grid barriers (7 per half-block); counters are monotonic within a launch and reset by the
// last CTA to leave.  Warps 0-7 = workers (warp 0 lane 0 also issues the activation copies), warp 8 lane 0 = weight
// producer + MMA issuer.  Cross-CTA data is read with ld.global.cg or bulk copies after fence.proxy.async.
#include "attn_seq.cuh"
#include "common.cuh"
#include "kernels.h"

namespace gtav {

static constexpr int MG_WORKERS = 256;
static constexpr int MG_THREADS = MG_WORKERS + 32;
static constexpr int MG_TOK = 144;
static constexpr int MG_D = 1024;
static constexpr int MG_CHUNK = MG_TOK * 128;                 // one [144 x 64] bf16 chunk of a pre-tiled activation
static constexpr int MG_STAGES = 4;
static constexpr int MG_KCH = 16;                             // 64-wide chunks per GEMM (K = 1024 per CTA)
static constexpr int MG_W_SLOT = 64 * 1024;
static constexpr int MG_TAIL = 1024;                          // mbarriers, TMEM slot, block-reduction scratch
static constexpr int MG_SMEM = 2 * MG_W_SLOT + MG_STAGES * MG_CHUNK + MG_TAIL + 1024;

// GEMM kinds of a half-block: 0 to_qkv, 1 to_out, 2 fc1, 3 fc2
__host__ __device__ constexpr int mega_rows(int kind) { return kind == 1 ? 16 : 32; }          // weight rows per CTA
__host__ __device__ constexpr int mega_ctas(int kind) { return kind == 0 ? 96 : kind == 1 ? 64 : 128; }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void spin_until(const unsigned* ctr, unsigned target) {
    uint32_t spins = 0;
    while (ld_acquire_u32(ctr) < target) {
        if (++spins > (1u << 25)) {          // a protocol bug traps instead of hanging the box
            printf("gtav: step kernel grid barrier timed out (block %d, target %u, seen %u)\n", blockIdx.x, target, ld_acquire_u32(ctr));
            __trap();
        }
    }
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
    uint64_t *bar_w, *bar_full, *bar_empty, *bar_acc, *bar_att;
    float* s_red;                 // [16] block-reduction scratch
    uint32_t tmem;
    int cta, tid, warp, lane;
    const bf16* mod_row;          // modulation vectors of this step's conditioning row
    uint32_t acc_uses, att_uses;  // completed uses of bar_acc / bar_att (their wait parities)
    unsigned grid_epoch;          // grid barriers passed
};

// Grid barrier for the workers of every CTA.  Release: the CTA barrier orders every worker's global writes before
// thread 0's gpu-scope fence (cumulativity), which orders them before its arrival; acquire: thread 0's acquire load,
// then the CTA barrier.
__device__ __forceinline__ void grid_barrier(MegaStep& s) {
    worker_bar();
    if (s.tid == 0) {
        fence_gpu();
        atomicAdd(s.pp->sync, 1u);
        spin_until(s.pp->sync, ++s.grid_epoch * static_cast<unsigned>(s.pp->grid));
    }
    worker_bar();
}

__device__ __forceinline__ float block_sum(MegaStep& s, float v, int slot) {
    v = warp_sum(v);
    if (s.lane == 0) s.s_red[slot * 8 + s.warp] = v;
    worker_bar();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s.s_red[slot * 8 + i];
    return t;
}

// ---------------------------------------------------------------------------------------------------- row phase
// CTA r owns token row r (CTAs 0..15 also row r + 128).  SRC 0: h is taken as it is (patch-embed output); SRC 1: h +=
// gate_msa * (to_out result + bias) from ws_out [64 tiles][144][16]; SRC 2: h += gate_mlp * (fc2 result + bias) from
// the 4 K-split partials ws_fc2 [32 blocks x 4 splits][144][32].  Then, if LN, the modulated LayerNorm of the new row
// goes to hn_t (pre-tiled).  Rounding points as the reference's autocast graph: Linear output -> bf16, gate * y ->
// bf16, residual sum -> bf16; LN statistics two-pass in fp32 on the bf16 row, one rounding of the modulated result.
template <int SRC, bool LN>
__device__ __forceinline__ void row_phase(MegaStep& s, const bf16* bias, int gate_off, int shift_off, int scale_off) {
    const MegaParams& p = *s.pp;
    const int col = s.tid * 4;
    uint2 bv = make_uint2(0, 0), gv = make_uint2(0, 0), shv = make_uint2(0, 0), scv = make_uint2(0, 0);
    if (SRC != 0) {
        bv = *reinterpret_cast<const uint2*>(bias + col);
        gv = *reinterpret_cast<const uint2*>(s.mod_row + gate_off + col);
    }
    if (LN) {
        shv = *reinterpret_cast<const uint2*>(s.mod_row + shift_off + col);
        scv = *reinterpret_cast<const uint2*>(s.mod_row + scale_off + col);
    }
    for (int row = s.cta; row < MG_TOK; row += p.grid) {
        bf16* hrow = p.h + static_cast<size_t>(row) * MG_D + col;
        const uint2 hv = __ldcg(reinterpret_cast<const uint2*>(hrow));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (SRC == 1) {
            acc = __ldcg(reinterpret_cast<const float4*>(p.ws_out + (static_cast<size_t>(col >> 4) * MG_TOK + row) * 16 + (col & 15)));
        } else if (SRC == 2) {
            float4 v[4];
#pragma unroll
            for (int sp = 0; sp < 4; ++sp)
                v[sp] = __ldcg(reinterpret_cast<const float4*>(p.ws_fc2 + (static_cast<size_t>((col >> 5) * 4 + sp) * MG_TOK + row) * 32 + (col & 31)));
            acc = v[0];
#pragma unroll
            for (int sp = 1; sp < 4; ++sp) { acc.x += v[sp].x; acc.y += v[sp].y; acc.z += v[sp].z; acc.w += v[sp].w; }
        }
        const float2 h0 = unpack_bf16x2(hv.x), h1 = unpack_bf16x2(hv.y);
        float x[4] = {h0.x, h0.y, h1.x, h1.y};
        if (SRC != 0) {
            const float2 b0 = unpack_bf16x2(bv.x), b1 = unpack_bf16x2(bv.y), g0 = unpack_bf16x2(gv.x), g1 = unpack_bf16x2(gv.y);
            const float a4[4] = {acc.x, acc.y, acc.z, acc.w}, b4[4] = {b0.x, b0.y, b1.x, b1.y}, g4[4] = {g0.x, g0.y, g1.x, g1.y};
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = bf16_round(x[j] + bf16_round(g4[j] * bf16_round(a4[j] + b4[j])));
            uint2 o;
            o.x = pack_bf16x2(x[0], x[1]);
            o.y = pack_bf16x2(x[2], x[3]);
            __stcg(reinterpret_cast<uint2*>(hrow), o);
        }
        if (LN) {
            const float mean = block_sum(s, x[0] + x[1] + x[2] + x[3], 0) * (1.0f / MG_D);
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) sq += (x[j] - mean) * (x[j] - mean);
            const float rstd = rsqrtf(block_sum(s, sq, 1) * (1.0f / MG_D) + 1e-6f);
            const float2 s0 = unpack_bf16x2(shv.x), s1 = unpack_bf16x2(shv.y), c0 = unpack_bf16x2(scv.x), c1 = unpack_bf16x2(scv.y);
            const float sh4[4] = {s0.x, s0.y, s1.x, s1.y}, sc4[4] = {c0.x, c0.y, c1.x, c1.y};
            float y[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] = (x[j] - mean) * rstd * bf16_round(1.0f + bf16_round(sc4[j] + 1e-6f)) + sh4[j];
            uint2 o;
            o.x = pack_bf16x2(y[0], y[1]);
            o.y = pack_bf16x2(y[2], y[3]);
            uint8_t* dst = p.hn_t + static_cast<size_t>(col >> 6) * MG_CHUNK + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4) + (col & 7) * 2;
            __stcg(reinterpret_cast<uint2*>(dst), o);
        }
    }
}

// ---------------------------------------------------------------------------------------------------- GEMM phase (workers)
// D[144 tokens x n weight rows] = act[144, 1024] @ W_slab[n, 1024]^T, accumulators in TMEM columns [0, n) (tokens
// 0..127) and [32, 32 + n) (tokens 16..143).  Worker warp 0 lane 0 streams the 16 activation chunks; the epilogue reads
// token rows 0..127 with warps 0..3 and tokens 128..143 with the upper half of warp 7 (TMEM lanes 112..127 of tile 2).
template <int KIND>
__device__ __forceinline__ void gemm_phase_workers(MegaStep& s, int half, const MegaHalfDev& hd) {
    const MegaParams& p = *s.pp;
    constexpr int NR = mega_rows(KIND);
    // ---- activation stream: lane 0 of worker warp w feeds ring stage w (a thread gets a bulk copy accepted only every
    // ~0.4 us, so one issuing thread would serialise the ring - measured: 16 copies, 7.8 us)
    if (s.warp < MG_STAGES) {
        if (s.lane == 0) {
            const uint8_t* src = KIND == 1 ? p.att_t : KIND == 3 ? p.mlp_t + static_cast<size_t>(s.cta & 3) * MG_KCH * MG_CHUNK : p.hn_t;
            fence_proxy_async_global();      // generic-proxy writes of other CTAs (acquired at the grid barrier) -> async-proxy reads
            const int st = s.warp;
#pragma unroll 1
            for (int u = 0; u < MG_KCH / MG_STAGES; ++u) {
                const int c = u * MG_STAGES + st;
                mbar_wait(&s.bar_empty[st], (u & 1) ^ 1);
                mbar_arrive_expect_tx(&s.bar_full[st], MG_CHUNK);
                bulk_load(s.sA + st * MG_CHUNK, src + static_cast<size_t>(c) * MG_CHUNK, MG_CHUNK, &s.bar_full[st]);
            }
        }
        __syncwarp();
    }
    // ---- epilogue operands that do not depend on the accumulator
    const bool tile1 = s.warp < 4, tile2 = s.warp == 7 && s.lane >= 16;
    const int tok = tile1 ? s.warp * 32 + s.lane : 128 + ((s.lane - 16) & 15);
    const int f0 = NR * s.cta;                                   // first output feature of this CTA (kinds 0..2)
    float2 rot[16];
    uint4 braw[4];
    if (KIND == 0 && (tile1 || tile2) && f0 < 2 * MG_D) {
        const float2* r = (half & 1) ? p.rot_t + p.ctx_frames * 32 + ((f0 >> 5) & 1) * 16 : p.rot_s + tok * 32 + ((f0 >> 5) & 1) * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) rot[i] = r[i];
    }
    if (KIND == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) braw[i] = *reinterpret_cast<const uint4*>(hd.fc1_b + f0 + 8 * i);
    }
    // ---- accumulator
    mbar_wait(s.bar_acc, s.acc_uses & 1);
    s.acc_uses++;
    tcgen05_fence_after();
    if (tile1 || s.warp == 7) {
        uint32_t acc[32];
        const uint32_t taddr = tile1 ? s.tmem + (static_cast<uint32_t>(s.warp * 32) << 16) : s.tmem + (96u << 16) + 32u;
        if (NR == 32) {
            tmem_ld_32x32(taddr, acc);
        } else {
            uint32_t a16[16];
            tmem_ld_32x16(taddr, a16);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = a16[i];
        }
        tmem_ld_wait();
        if (tile1 || tile2) {
            if (KIND == 0) {
                // to_qkv has no bias; rotary on q and k (fp32 rotate, one bf16 rounding), row-padded per-head layout
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) y[i] = bf16_round(__uint_as_float(acc[i]));
                uint32_t o[16];
                const int which = f0 >> 10, head = (f0 >> 6) & 15, hh = (f0 >> 5) & 1;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (which < 2) o[i] = pack_bf16x2(y[2 * i] * rot[i].x - y[2 * i + 1] * rot[i].y, y[2 * i + 1] * rot[i].x + y[2 * i] * rot[i].y);
                    else o[i] = pack_bf16x2(y[2 * i], y[2 * i + 1]);
                }
                bf16* dst = p.qkv_h + (static_cast<size_t>(which * 16 + head) * MG_TOK + tok) * SROW + hh * 32;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    __stcg(reinterpret_cast<uint4*>(dst) + q, make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
            } else if (KIND == 2) {
                // fc1: + bias -> bf16 -> GELU(tanh) -> bf16, written as chunk f0/64 of the pre-tiled fc2 operand
                uint8_t* dst = p.mlp_t + static_cast<size_t>(f0 >> 6) * MG_CHUNK + tok * 128;
                const int j0 = ((f0 >> 5) & 1) * 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t bw[4] = {braw[q].x, braw[q].y, braw[q].z, braw[q].w};
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 b2 = unpack_bf16x2(bw[j]);
                        o[j] = pack_bf16x2(gelu_tanh_f(bf16_round(__uint_as_float(acc[8 * q + 2 * j]) + b2.x)),
                                           gelu_tanh_f(bf16_round(__uint_as_float(acc[8 * q + 2 * j + 1]) + b2.y)));
                    }
                    __stcg(reinterpret_cast<uint4*>(dst + (((j0 + q) ^ (tok & 7)) << 4)), make_uint4(o[0], o[1], o[2], o[3]));
                }
            } else {
                // to_out / fc2: raw fp32 sums for the row phase (bias, gate, residual and - fc2 - the split sum happen there)
                float* dst = KIND == 1 ? p.ws_out + (static_cast<size_t>(s.cta) * MG_TOK + tok) * 16
                                       : p.ws_fc2 + (static_cast<size_t>(s.cta) * MG_TOK + tok) * 32;
#pragma unroll
                for (int q = 0; q < NR / 4; ++q)
                    __stcg(reinterpret_cast<float4*>(dst) + q, make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]),
                                                                           __uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3])));
            }
        }
    }
    tcgen05_fence_before();
}

// ---------------------------------------------------------------------------------------------------- attention phases
// Spatial attention (attention.py:99-129): 48 work items (16 heads x 3 blocks of 48 queries), CTAs 0..47.  K, V (whole
// head) and the item's Q rows arrive already rotated and row-padded (to_qkv epilogue): three bulk copies into the
// activation ring, then 3 warps run the mma.sync body of attn_seq.cuh and write the to_out operand pre-tiled.
__device__ __forceinline__ void spatial_attention_phase(MegaStep& s) {
    const MegaParams& p = *s.pp;
    if (s.cta >= 48) return;
    const int head = s.cta / 3, qb = s.cta % 3;
    constexpr uint32_t KV_BYTES = MG_TOK * SROW * 2, Q_BYTES = 48 * SROW * 2;
    bf16* sK = reinterpret_cast<bf16*>(s.sA);
    bf16* sV = sK + MG_TOK * SROW;
    bf16* sQ = sV + MG_TOK * SROW;
    if (s.tid == 0) {
        fence_proxy_async_global();
        mbar_arrive_expect_tx(s.bar_att, 2 * KV_BYTES + Q_BYTES);
        bulk_load(sK, p.qkv_h + static_cast<size_t>(16 + head) * MG_TOK * SROW, KV_BYTES, s.bar_att);
        bulk_load(sV, p.qkv_h + static_cast<size_t>(32 + head) * MG_TOK * SROW, KV_BYTES, s.bar_att);
        bulk_load(sQ, p.qkv_h + (static_cast<size_t>(head) * MG_TOK + qb * 48) * SROW, Q_BYTES, s.bar_att);
    }
    if (s.warp < 3) {
        mbar_wait(s.bar_att, s.att_uses & 1);
        attn_seq_body<144, 144, 3, 32, true, true, true>(nullptr, reinterpret_cast<bf16*>(p.att_t), 16, nullptr, sK, sV, qb, head, 0,
                                                         s.tid, 2, sQ);
    }
    s.att_uses++;
}

// Temporal attention of the frame being denoised (attention.py:41-66): one warp per (position, head); query at window
// position TC, keys / values = the TC cached context frames + itself (q, k already rotated by the to_qkv epilogue);
// same arithmetic and order as attn_temporal_last_kernel.  A warp's problems are processed together (loads overlap).
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
            const bf16* base = p.qkv_h + (static_cast<size_t>(head) * MG_TOK + pos) * SROW + 2 * lane;
            qr[u] = __ldcg(reinterpret_cast<const unsigned int*>(base));
            kr[u] = __ldcg(reinterpret_cast<const unsigned int*>(base + static_cast<size_t>(16) * MG_TOK * SROW));
            vr[u] = __ldcg(reinterpret_cast<const unsigned int*>(base + static_cast<size_t>(32) * MG_TOK * SROW));
        }
    }
#pragma unroll
    for (int u = 0; u < NP; ++u) {
        const int prob = s.cta * 8 + s.warp + u * p.grid * 8;
        if (prob < MG_TOK * 16) {
            const int head = prob & 15, pos = prob >> 4;
            const float2 q = unpack_bf16x2(qr[u]), kn = unpack_bf16x2(kr[u]), vx = unpack_bf16x2(vr[u]);
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
            // column head*64 + 2*lane of the pre-tiled to_out operand: chunk = head, 16-byte group = lane / 4
            uint8_t* dst = p.att_t + static_cast<size_t>(head) * MG_CHUNK + pos * 128 + (((lane >> 2) ^ (pos & 7)) << 4) + (lane & 3) * 4;
            __stcg(reinterpret_cast<unsigned int*>(dst), pack_bf16x2(acc.x, acc.y));
        }
    }
}

// Pull what the next half-block will read cold into L2 (the 805 MB of weights that stream through every step evict
// everything else): its six modulation vectors + three bias vectors, and - before a temporal half - the context K/V
// of its layer.  Called by CTAs that have nothing else to do in the attention phase.
__device__ __forceinline__ void prefetch_next(MegaStep& s, int next_half) {
    const MegaParams& p = *s.pp;
    if (next_half >= p.n_halves) return;
    const MegaHalfDev& hd = p.halves[next_half];
    if (s.cta == p.grid - 1) {
        const int t = s.tid;
        if (t < 96) l2_prefetch_line(reinterpret_cast<const char*>(s.mod_row + hd.mod_off) + t * 128);            // 6 x 2 KB
        else if (t < 112) l2_prefetch_line(reinterpret_cast<const char*>(hd.out_b) + (t - 96) * 128);
        else if (t < 176) l2_prefetch_line(reinterpret_cast<const char*>(hd.fc1_b) + (t - 112) * 128);
        else if (t < 192) l2_prefetch_line(reinterpret_cast<const char*>(hd.fc2_b) + (t - 176) * 128);
    } else if ((next_half & 1) && p.ctx_frames > 0 && s.cta >= 48) {
        const size_t lines = (static_cast<size_t>(p.ctx_frames) * MG_TOK * 2 * MG_D * sizeof(bf16)) >> 7;
        const char* base = reinterpret_cast<const char*>(p.kv_cache + static_cast<size_t>(next_half >> 1) * p.cache_layer_stride);
        const int parts = (p.grid - 1 - 48) * MG_WORKERS;
        for (size_t l = static_cast<size_t>(s.cta - 48) * MG_WORKERS + s.tid; l < lines; l += parts) l2_prefetch_line(base + (l << 7));
    }
}

// ---------------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(MG_THREADS, 1) dit_step_mega_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    MegaStep s;
    s.pp = &p;
    s.sW = smem;
    s.sA = smem + 2 * MG_W_SLOT;
    uint8_t* tail = s.sA + MG_STAGES * MG_CHUNK;
    s.bar_w = reinterpret_cast<uint64_t*>(tail);
    s.bar_full = s.bar_w + 2;
    s.bar_empty = s.bar_w + 6;
    s.bar_acc = s.bar_w + 10;
    s.bar_att = s.bar_w + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s.bar_w + 12);
    s.s_red = reinterpret_cast<float*>(tail + 128);
    s.tid = threadIdx.x;
    s.warp = threadIdx.x >> 5;
    s.lane = threadIdx.x & 31;
    s.cta = blockIdx.x;
    s.acc_uses = 0;
    s.att_uses = 0;
    s.grid_epoch = 0;
    const int n_gemms = 4 * p.n_halves;

    if (s.warp == 8) {
        if (s.lane == 0) {
            mbar_init(&s.bar_w[0], 1);
            mbar_init(&s.bar_w[1], 1);
            for (int i = 0; i < MG_STAGES; ++i) {
                mbar_init(&s.bar_full[i], 1);
                mbar_init(&s.bar_empty[i], 1);
            }
            mbar_init(s.bar_acc, 1);
            mbar_init(s.bar_att, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 64);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    s.tmem = *tmem_slot;
    s.mod_row = p.mod + static_cast<size_t>(p.last_row[0]) * p.mod_ld;

    if (s.warp == 8) {
        // ======================= weight producer + MMA issuer (one thread) =======================
        if (s.lane == 0) {
            uint32_t w_par = 0;                                  // bit i = wait parity of ring slot i
            auto active = [&](int g) { return s.cta < mega_ctas(g & 3); };
            auto issue_w = [&](int g) {
                const int kind = g & 3, nr = mega_rows(kind);
                const int row0 = kind == 3 ? 32 * (s.cta >> 2) : nr * s.cta;
                const int chunk0 = kind == 3 ? MG_KCH * (s.cta & 3) : 0;
                mbar_arrive_expect_tx(&s.bar_w[g & 1], nr * 128 * MG_KCH);
                tma_load_3d(s.sW + (g & 1) * MG_W_SLOT, &p.halves[g >> 2].tm[kind], &s.bar_w[g & 1], 0, row0, chunk0);
            };
            for (int g = 0; g < 2 && g < n_gemms; ++g)
                if (active(g)) issue_w(g);
            for (int g = 0; g < n_gemms; ++g) {
                if (active(g)) {
                    const int kind = g & 3, nr = mega_rows(kind), slot = g & 1;
                    const uint32_t idesc = umma_idesc_bf16(128, nr);
                    mbar_wait(&s.bar_w[slot], (w_par >> slot) & 1);
                    w_par ^= 1u << slot;
#pragma unroll 1
                    for (int c = 0; c < MG_KCH; ++c) {
                        const int st = c & (MG_STAGES - 1), u = c / MG_STAGES;
                        mbar_wait(&s.bar_full[st], u & 1);
                        tcgen05_fence_after();
                        const uint32_t a_addr = smem_u32(s.sA + st * MG_CHUNK);
                        const uint64_t da1 = umma_desc_sw128(a_addr), da2 = umma_desc_sw128(a_addr + 16 * 128);
                        const uint64_t db = umma_desc_sw128(smem_u32(s.sW + slot * MG_W_SLOT + c * nr * 128));
                        // cost model (scripts/probe_umma_rate.cu): every tcgen05.mma costs >= ~61 cycles however small N is
                        // (75 at N = 144, 130 at N = 256 = full rate), so this phase is bound by its 128 MMAs (~4 us)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_bf16_ss(s.tmem, da1 + 2 * k, db + 2 * k, idesc, (c | k) != 0 ? 1u : 0u);
                            umma_bf16_ss(s.tmem + 32, da2 + 2 * k, db + 2 * k, idesc, (c | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(&s.bar_empty[st]);           // ring stage reusable once these MMAs have read it
                    }
                    umma_commit(s.bar_acc);
                    mbar_wait(s.bar_acc, s.acc_uses & 1);        // all MMAs done: the weight slot is free again
                    s.acc_uses++;
                }
                if (g + 2 < n_gemms && active(g + 2)) issue_w(g + 2);
            }
        }
    } else {
        // ======================= workers =======================
        const MegaHalfDev* H = p.halves;
        if (s.cta == p.grid - 1 && s.tid < 96) l2_prefetch_line(reinterpret_cast<const char*>(s.mod_row + H[0].mod_off) + s.tid * 128);
        // LayerNorm + modulate of the incoming residual stream (patch-embed output) with half-block 0's shift/scale_msa
        row_phase<0, true>(s, nullptr, 0, H[0].mod_off, H[0].mod_off + MG_D);
        grid_barrier(s);

        for (int half = 0; half < p.n_halves; ++half) {
            const MegaHalfDev& hd = H[half];
            MG_STAMP(s, half, 0);
            if (s.cta < mega_ctas(0)) gemm_phase_workers<0>(s, half, hd);
            MG_STAMP(s, half, 1);
            grid_barrier(s);
            MG_STAMP(s, half, 2);
            if (half & 1) temporal_attention_phase(s, p.kv_cache + static_cast<size_t>(half >> 1) * p.cache_layer_stride);
            else spatial_attention_phase(s);
            prefetch_next(s, half + 1);
            MG_STAMP(s, half, 3);
            grid_barrier(s);
            MG_STAMP(s, half, 4);
            if (s.cta < mega_ctas(1)) gemm_phase_workers<1>(s, half, hd);
            MG_STAMP(s, half, 5);
            grid_barrier(s);
            MG_STAMP(s, half, 6);
            row_phase<1, true>(s, hd.out_b, hd.mod_off + 2 * MG_D, hd.mod_off + 3 * MG_D, hd.mod_off + 4 * MG_D);
            MG_STAMP(s, half, 7);
            grid_barrier(s);
            MG_STAMP(s, half, 8);
            gemm_phase_workers<2>(s, half, hd);
            MG_STAMP(s, half, 9);
            grid_barrier(s);
            MG_STAMP(s, half, 10);
            gemm_phase_workers<3>(s, half, hd);
            MG_STAMP(s, half, 11);
            grid_barrier(s);
            MG_STAMP(s, half, 12);
            if (half + 1 < p.n_halves) {
                row_phase<2, true>(s, hd.fc2_b, hd.mod_off + 5 * MG_D, H[half + 1].mod_off, H[half + 1].mod_off + MG_D);
                MG_STAMP(s, half, 13);
                grid_barrier(s);
                MG_STAMP(s, half, 14);
            } else {
                row_phase<2, false>(s, hd.fc2_b, hd.mod_off + 5 * MG_D, 0, 0);
            }
        }
        // leave the counters zeroed for the next launch: the last CTA out resets them (everyone is past every wait)
        if (s.tid == 0) {
            const unsigned old = atomicAdd(p.sync + 1, 1u);
            if (old == static_cast<unsigned>(p.grid) - 1) {
                p.sync[0] = 0;
                p.sync[1] = 0;
                __threadfence();
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (s.warp == 8) tmem_dealloc(s.tmem, 64);
}

// ---------------------------------------------------------------------------------------------------- host
size_t mega_sync_bytes() { return 64; }
int mega_grid() { return 128; }
size_t mega_buffer_bytes(int which) {
    switch (which) {
        case 0: return static_cast<size_t>(MG_KCH) * MG_CHUNK;                   // hn_t
        case 1: return static_cast<size_t>(MG_KCH) * MG_CHUNK;                   // att_t
        case 2: return static_cast<size_t>(4 * MG_KCH) * MG_CHUNK;               // mlp_t
        case 3: return static_cast<size_t>(3 * 16) * MG_TOK * SROW * 2;          // qkv_h
        case 4: return static_cast<size_t>(64) * MG_TOK * 16 * sizeof(float);    // ws_out
        default: return static_cast<size_t>(128) * MG_TOK * 32 * sizeof(float);  // ws_fc2
    }
}

int mega_make_weight_map(CUtensorMap* out, const bf16* W, int kind) {
    const int N = kind == 0 ? 3 * MG_D : kind == 2 ? 4 * MG_D : MG_D, K = kind == 3 ? 4 * MG_D : MG_D;
    return make_tmap_3d(out, W, N, K, K, mega_rows(kind), MG_KCH);
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
