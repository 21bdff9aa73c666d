// Fused GEMM epilogues shared by the tiled and the frame GEMM kernels: bias, GELU / SiLU, adaLN gate, residual,
// with a bf16 rounding at every point where the reference's autocast graph materialises a bf16 tensor
// (reference model/dit.py:207-223, model/vae.py:147-156, timm Mlp).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace gtav {

// GROUPS * 8 consecutive columns of one output row, handled as groups of 8 (16-byte vectors).  acc holds the fp32
// accumulators (bit patterns, as tcgen05.ld returns them).
template <int EPI, int GROUPS = 4>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int col0, const uint32_t (&acc)[GROUPS * 8],
                                               const bf16* gate_row) {
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
        const int col = col0 + g * 8;
        if (col >= p.N) break;                               // N is a multiple of 8
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(acc[g * 8 + j]);
        if (EPI != EPI_STORE) {
            uint4 bv = *reinterpret_cast<const uint4*>(p.bias + col);
            const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 b2 = unpack_bf16x2(bw[j]);
                y[2 * j] += b2.x;
                y[2 * j + 1] += b2.y;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = bf16_round(y[j]);  // the Linear's own bf16 output
        if (EPI == EPI_BIAS_GELU_TANH) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = gelu_tanh_f(y[j]);
        } else if (EPI == EPI_BIAS_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = gelu_erf_f(y[j]);
        } else if (EPI == EPI_BIAS_SILU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = silu_f(y[j]);
        } else if (EPI == EPI_BIAS_GATE_RES || EPI == EPI_BIAS_RES || EPI == EPI_BIAS_RES_SILU) {
            float r[8];
            if (p.res != nullptr) {
                uint4 rv = *reinterpret_cast<const uint4*>(p.res + static_cast<size_t>(row) * p.ldr + col);
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 r2 = unpack_bf16x2(rw[j]);
                    r[2 * j] = r2.x;
                    r[2 * j + 1] = r2.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = 0.f;
            }
            if (EPI == EPI_BIAS_GATE_RES) {
                uint4 gv = *reinterpret_cast<const uint4*>(gate_row + col);
                const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 g2 = unpack_bf16x2(gw[j]);
                    y[2 * j] = bf16_round(g2.x * y[2 * j]);
                    y[2 * j + 1] = bf16_round(g2.y * y[2 * j + 1]);
                }
            }
            if (EPI == EPI_BIAS_RES_SILU) {
                if (p.res != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = bf16_round(r[j] + y[j]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = silu_f(y[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = r[j] + y[j];
            }
        }
        uint4 o;
        o.x = pack_bf16x2(y[0], y[1]);
        o.y = pack_bf16x2(y[2], y[3]);
        o.z = pack_bf16x2(y[4], y[5]);
        o.w = pack_bf16x2(y[6], y[7]);
        *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(row) * p.ldo + col) = o;
    }
}

}  // namespace gtav
