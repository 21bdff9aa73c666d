// One LayerNorm row (D = 256 * CHUNKS features): fp32 two-pass statistics, then either the adaLN modulate of the DiT
// (reference model/dit.py:19-27) or the VAE's affine transform, one rounding to bf16.
// Shared by ln_rows_kernel (norm_mod.cu, one warp per row) and the LayerNorm fused into the weight-streaming GEMM's
// reduce (gemm_skinny.cu, EIGHT warps per row, 128 features each), so both produce the same bits from the same row.
// That fixes the summation order of the statistics, which is defined on the 8-warp layout:
//   * quad q = features 4q .. 4q + 3:          (x0 + x1) + (x2 + x3)        [squares: fma chain over (x - mean)^2]
//   * segment w = features 128 w .. 128 w + 127: xor-butterfly over its 32 quads with offsets 16, 8, 4, 2, 1
//   * row: ((S0 + S1) + (S2 + S3)) + ((S4 + S5) + (S6 + S7))   (pairwise over the 2 * CHUNKS segments)
// The one-warp form below holds, per 256-feature chunk, quads 2m and 2m + 1 of segment 2c + lane / 16 in lane m = lane % 16
// and walks the same butterfly (offsets 16 .. 2 of the quad index = lane offsets 8 .. 1, offset 1 = its own two quads).
#pragma once
#include "common.cuh"

namespace gtav {

// Explicit round-to-nearest intrinsics throughout: the same source inlined into different kernels must produce the
// same bits (no compiler-chosen FMA contraction), and LN(x) * (1 + scale) + shift is evaluated unfused like the
// reference's separate fp32 tensor ops (model/dit.py:26-27).
__device__ __forceinline__ float ln_quad_sum(float a, float b, float c, float d) { return __fadd_rn(__fadd_rn(a, b), __fadd_rn(c, d)); }
__device__ __forceinline__ float ln_quad_sq(float a, float b, float c, float d, float mean) {
    const float d0 = __fsub_rn(a, mean), d1 = __fsub_rn(b, mean), d2 = __fsub_rn(c, mean), d3 = __fsub_rn(d, mean);
    return __fmaf_rn(d3, d3, __fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0))));
}
// the 8 segment sums of a 1024-wide row -> row sum
__device__ __forceinline__ float ln_combine8(float s0, float s1, float s2, float s3, float s4, float s5, float s6, float s7) {
    return __fadd_rn(__fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3)), __fadd_rn(__fadd_rn(s4, s5), __fadd_rn(s6, s7)));
}
__device__ __forceinline__ float ln_rstd(float sq_sum, int D) { return rsqrtf(__fadd_rn(__fmul_rn(sq_sum, 1.0f / D), 1e-6f)); }

// One warp, whole row.  a0 / a1: this lane's two quad values of every chunk; returns the row total in every lane.
template <int CHUNKS>
__device__ __forceinline__ float ln_row_total(float (&a0)[CHUNKS], float (&a1)[CHUNKS], int lane) {
    static_assert(CHUNKS == 4, "the pairwise segment order is written out for 8 segments (D = 1024)");
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            a0[c] = __fadd_rn(a0[c], __shfl_xor_sync(0xffffffffu, a0[c], off));
            a1[c] = __fadd_rn(a1[c], __shfl_xor_sync(0xffffffffu, a1[c], off));
        }
    }
    float pair[CHUNKS];
    const bool hi = lane >= 16;                               // this lane's half holds segment 2c + 1, the other half 2c
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const float own = __fadd_rn(a0[c], a1[c]);            // segment 2c + hi
        const float oth = __shfl_xor_sync(0xffffffffu, own, 16);
        pair[c] = hi ? __fadd_rn(oth, own) : __fadd_rn(own, oth);   // S(2c) + S(2c + 1)
    }
    return __fadd_rn(__fadd_rn(pair[0], pair[1]), __fadd_rn(pair[2], pair[3]));
}

// xu: this lane's 8-element slices of the row (chunk c covers features c*256 + lane*8 .. +7).
template <int CHUNKS>
__device__ __forceinline__ void ln_row_stats(const uint4 (&xu)[CHUNKS], float (&v)[CHUNKS][8], float& mean, float& rstd) {
    constexpr int D = CHUNKS * 256;
    const int lane = threadIdx.x & 31;
    float a0[CHUNKS], a1[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const uint32_t uw[4] = {xu[c].x, xu[c].y, xu[c].z, xu[c].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = unpack_bf16x2(uw[j]);
            v[c][2 * j] = f.x;
            v[c][2 * j + 1] = f.y;
        }
        a0[c] = ln_quad_sum(v[c][0], v[c][1], v[c][2], v[c][3]);
        a1[c] = ln_quad_sum(v[c][4], v[c][5], v[c][6], v[c][7]);
    }
    mean = __fmul_rn(ln_row_total<CHUNKS>(a0, a1, lane), 1.0f / D);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        a0[c] = ln_quad_sq(v[c][0], v[c][1], v[c][2], v[c][3], mean);
        a1[c] = ln_quad_sq(v[c][4], v[c][5], v[c][6], v[c][7], mean);
    }
    rstd = ln_rstd(ln_row_total<CHUNKS>(a0, a1, lane), D);
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

// the same for one quad (4 features): sh / sc = packed bf16 shift / scale of these features
__device__ __forceinline__ uint2 ln_modulate_quad(const float (&v)[4], float mean, float rstd, uint2 sh, uint2 sc) {
    const uint32_t shw[2] = {sh.x, sh.y}, scw[2] = {sc.x, sc.y};
    float y[4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float2 s2 = unpack_bf16x2(shw[j]), c2 = unpack_bf16x2(scw[j]);
        const float m0 = bf16_round(__fadd_rn(1.0f, bf16_round(__fadd_rn(c2.x, 1e-6f))));
        const float m1 = bf16_round(__fadd_rn(1.0f, bf16_round(__fadd_rn(c2.y, 1e-6f))));
        y[2 * j] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[2 * j], mean), rstd), m0), s2.x);
        y[2 * j + 1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[2 * j + 1], mean), rstd), m1), s2.y);
    }
    uint2 o;
    o.x = pack_bf16x2(y[0], y[1]);
    o.y = pack_bf16x2(y[2], y[3]);
    return o;
}

}  // namespace gtav
