// Temporal attention of ONE (rollout, position, head) problem for the frame being denoised: the query is the last
// frame's, keys/values are the TC cached context frames (rotated K and V, reference model/attention.py:52-66)
// followed by the frame's own.  One warp, lane = one rotary pair (2 of the 64 head dims).  Shared by
// attn_temporal_last_kernel (attn_temporal.cu) and the attention fused into the weight-streaming to_qkv GEMM's
// reduce (gemm_skinny.cu), so both produce the same bits from the same q / k / v.
#pragma once
#include "common.cuh"

namespace gtav {

// This lane's pairs of the cached rotated K and V of context frames 0 .. tc-1 (packed bf16x2), TMAX >= tc.
// cache: kv_cache + (first context row of this (b, pos)) * 2D + head*64 + 2*lane, consecutive frames frame_stride
// elements apart.  Separate from the core so that callers can issue the loads early.
template <int TMAX>
__device__ __forceinline__ void temporal_cache_load(uint32_t (&kc)[TMAX], uint32_t (&vc)[TMAX], int tc, const bf16* cache,
                                                    size_t frame_stride, int D) {
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
        kc[t] = 0u;
        vc[t] = 0u;
        if (t < tc) {
            const bf16* c = cache + static_cast<size_t>(t) * frame_stride;
            kc[t] = *reinterpret_cast<const uint32_t*>(c);
            vc[t] = *reinterpret_cast<const uint32_t*>(c + D);
        }
    }
}

// qx, kx, vx: this lane's un-rotated q / k pair and v pair of the last frame (bf16 values as floats); kc / vc: the
// first TC entries are the cached context pairs (temporal_cache_load); cs: (cos, sin) of window position TC for this
// lane's pair.  Returns the packed bf16 output pair.
template <int TC, int TMAX>
__device__ __forceinline__ uint32_t temporal_last_core(float2 qx, float2 kx, float2 vx, const uint32_t (&kc)[TMAX],
                                                       const uint32_t (&vc)[TMAX], float2 cs) {
    static_assert(TC <= TMAX, "context frames exceed the cache registers");
    float2 k[TC + 1], v[TC + 1];
#pragma unroll
    for (int t = 0; t < TC; ++t) {
        k[t] = unpack_bf16x2(kc[t]);
        v[t] = unpack_bf16x2(vc[t]);
    }
    // rotate in fp32, round once to bf16 (apply_rotary_emb casts back to the input dtype)
    const float2 q = make_float2(bf16_round(qx.x * cs.x - qx.y * cs.y), bf16_round(qx.y * cs.x + qx.x * cs.y));
    k[TC] = make_float2(bf16_round(kx.x * cs.x - kx.y * cs.y), bf16_round(kx.y * cs.x + kx.x * cs.y));
    v[TC] = vx;

    float s[TC + 1];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j <= TC; ++j) {
        s[j] = warp_sum(q.x * k[j].x + q.y * k[j].y) * 0.125f;
        m = fmaxf(m, s[j]);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j <= TC; ++j) {
        s[j] = __expf(s[j] - m);
        l += s[j];
    }
    const float inv = 1.0f / l;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j <= TC; ++j) {
        const float p = bf16_round(s[j] * inv);              // probabilities enter P@V as bf16
        acc.x += p * v[j].x;
        acc.y += p * v[j].y;
    }
    return pack_bf16x2(acc.x, acc.y);
}

}  // namespace gtav
