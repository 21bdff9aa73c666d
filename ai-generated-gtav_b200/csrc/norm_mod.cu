// Row kernels: LayerNorm fused with adaLN modulate (DiT) or with the affine transform (VAE).
//
// Replaces `modulate(norm(x), shift, scale)` (reference model/dit.py:19-27 with the LayerNorms of
// 133,163,170,181,189) and the affine LayerNorms of the VAE blocks (model/vae.py:155-156,313,331).
// HBM/L2-bound: one warp per 1024-wide row, 16-byte loads/stores, fp32 statistics (two-pass in
// registers), output rounded once to bf16 - the rounding the following Linear's autocast applies.
#include "common.cuh"
#include "kernels.h"
#include "ln_row.cuh"

namespace gtav {

static constexpr int LN_WARPS = 8;

// D = 32 lanes * 8 elements * CHUNKS
template <int CHUNKS, bool AFFINE>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_rows_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int M, const bf16* __restrict__ mod, int mod_ld,
               int shift_off, int scale_off, const int* __restrict__ frame_row, int rows_per_frame,
               const float* __restrict__ w, const float* __restrict__ b) {
    constexpr int D = CHUNKS * 256;
    // shift | scale of the CTA's frame, staged once per CTA with cp.async right after the dependency wait: as plain
    // loads ptxas sinks them to their uses after the statistics (three exposed round trips instead of one)
    __shared__ __align__(16) uint8_t smod[2 * D * 2];
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const bool staged = !AFFINE && CHUNKS == 4 && rows_per_frame % LN_WARPS == 0;   // all rows of the CTA share a frame
    if (staged) {
        int f = (blockIdx.x * LN_WARPS) / rows_per_frame;
        if (frame_row != nullptr) f = frame_row[f];
        const int i = threadIdx.x & 127;
        const bf16* src = mod + static_cast<size_t>(f) * mod_ld + (threadIdx.x < 128 ? shift_off : scale_off) + i * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smod + threadIdx.x * 16)), "l"(src) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const bool active = row < M;
    const bf16* xr = x + static_cast<size_t>(active ? row : 0) * D;
    uint4 xu[CHUNKS], shu[CHUNKS], scu[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) xu[c] = *reinterpret_cast<const uint4*>(xr + c * 256 + lane * 8);
    if (!AFFINE && !staged) {
        int f = (active ? row : 0) / rows_per_frame;
        if (frame_row != nullptr) f = frame_row[f];
        const bf16* mrow = mod + static_cast<size_t>(f) * mod_ld;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            shu[c] = *reinterpret_cast<const uint4*>(mrow + shift_off + c * 256 + lane * 8);
            scu[c] = *reinterpret_cast<const uint4*>(mrow + scale_off + c * 256 + lane * 8);
        }
    }
    float v[CHUNKS][8];
    float mean, rstd;
    ln_row_stats<CHUNKS>(xu, v, mean, rstd);
    if (staged) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            shu[c] = *reinterpret_cast<const uint4*>(smod + (c * 256 + lane * 8) * 2);
            scu[c] = *reinterpret_cast<const uint4*>(smod + D * 2 + (c * 256 + lane * 8) * 2);
        }
    }
    if (!active) return;

#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = c * 256 + lane * 8;
        uint4 o;
        if (AFFINE) {
            float y[8];
            const float4 w0 = *reinterpret_cast<const float4*>(w + col), w1 = *reinterpret_cast<const float4*>(w + col + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(b + col), b1 = *reinterpret_cast<const float4*>(b + col + 4);
            const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = (v[c][j] - mean) * rstd * ww[j] + bb[j];
            o.x = pack_bf16x2(y[0], y[1]);
            o.y = pack_bf16x2(y[2], y[3]);
            o.z = pack_bf16x2(y[4], y[5]);
            o.w = pack_bf16x2(y[6], y[7]);
        } else {
            o = ln_modulate_slice(v[c], mean, rstd, shu[c], scu[c]);
        }
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * D + col) = o;
    }
}

int launch_ln_modulate(const bf16* x, bf16* out, int M, int D, const bf16* mod, int mod_ld, int shift_off,
                       int scale_off, const int* frame_row, int rows_per_frame, cudaStream_t s) {
    if (D != 1024 || (mod_ld % 8) || (shift_off % 8) || (scale_off % 8) || rows_per_frame <= 0) {
        set_error("ln_modulate: unsupported D=%d mod_ld=%d offsets=%d,%d", D, mod_ld, shift_off, scale_off);
        return -1;
    }
    if (M <= 0) return 0;
    GTAV_CUDA_OK(launch_k(ln_rows_kernel<4, false>, dim3((M + LN_WARPS - 1) / LN_WARPS), dim3(LN_WARPS * 32), 0, s, x, out, M, mod,
                          mod_ld, shift_off, scale_off, frame_row, rows_per_frame, static_cast<const float*>(nullptr),
                          static_cast<const float*>(nullptr)));
    return 0;
}

int launch_ln_affine(const bf16* x, bf16* out, int M, int D, const float* w, const float* b, cudaStream_t s) {
    if (D != 1024) {
        set_error("ln_affine: unsupported D=%d", D);
        return -1;
    }
    if (M <= 0) return 0;
    GTAV_CUDA_OK(launch_k(ln_rows_kernel<4, true>, dim3((M + LN_WARPS - 1) / LN_WARPS), dim3(LN_WARPS * 32), 0, s, x, out, M,
                          static_cast<const bf16*>(nullptr), 0, 0, 0, static_cast<const int*>(nullptr), 1, w, b));
    return 0;
}

}  // namespace gtav
