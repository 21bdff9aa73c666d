// One LayerNorm row (D = 256 * CHUNKS features) handled by one warp: fp32 two-pass statistics in registers, then either
// the adaLN modulate of the DiT (reference model/dit.py:19-27) or the VAE's affine transform, one rounding to bf16.
// Shared by ln_rows_kernel (norm_mod.cu) and the LayerNorm fused into the weight-streaming GEMM's reduce
// (gemm_skinny.cu), so both produce the same bits from the same row.
#pragma once
#include "common.cuh"

namespace gtav {

// Explicit round-to-nearest intrinsics throughout: the same source inlined into different kernels must produce the
// same bits (no compiler-chosen FMA contraction), and LN(x) * (1 + scale) + shift is evaluated unfused like the
// reference's separate fp32 tensor ops (model/dit.py:26-27).
// xu: this lane's 8-element slices of the row (chunk c covers features c*256 + lane*8 .. +7).
// modulate: shu / scu = the same slices of the shift / scale vectors (bf16).
template <int CHUNKS>
__device__ __forceinline__ void ln_row_stats(const uint4 (&xu)[CHUNKS], float (&v)[CHUNKS][8], float& mean, float& rstd) {
    constexpr int D = CHUNKS * 256;
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const uint32_t uw[4] = {xu[c].x, xu[c].y, xu[c].z, xu[c].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = unpack_bf16x2(uw[j]);
            v[c][2 * j] = f.x;
            v[c][2 * j + 1] = f.y;
            sum = __fadd_rn(sum, __fadd_rn(f.x, f.y));
        }
    }
    mean = __fmul_rn(warp_sum(sum), 1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = __fsub_rn(v[c][j], mean);
            sq = __fmaf_rn(d, d, sq);
        }
    rstd = rsqrtf(__fadd_rn(__fmul_rn(warp_sum(sq), 1.0f / D), 1e-6f));
}

// y = LN(x) * bf16(1 + bf16(scale + 1e-6)) + shift for one 8-element slice -> packed bf16
__device__ __forceinline__ uint4 ln_modulate_slice(const float (&v)[8], float mean, float rstd, uint4 sh, uint4 sc) {
    const uint32_t shw[4] = {sh.x, sh.y, sh.z, sh.w}, scw[4] = {sc.x, sc.y, sc.z, sc.w};
    float y[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 s2 = unpack_bf16x2(shw[j]), c2 = unpack_bf16x2(scw[j]);
        // scale + 1e-6 and 1 + scale are bf16 tensor ops in the reference (model/dit.py:26-27)
        const float m0 = bf16_round(__fadd_rn(1.0f, bf16_round(__fadd_rn(c2.x, 1e-6f))));
        const float m1 = bf16_round(__fadd_rn(1.0f, bf16_round(__fadd_rn(c2.y, 1e-6f))));
        y[2 * j] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[2 * j], mean), rstd), m0), s2.x);
        y[2 * j + 1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[2 * j + 1], mean), rstd), m1), s2.y);
    }
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]);
    o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]);
    o.w = pack_bf16x2(y[6], y[7]);
    return o;
}

}  // namespace gtav
