// Small HBM-bound kernels around the GEMMs: conditioning inputs, (un)patchify, latent casts, the
// pixel epilogue and the fused DDIM update.  Each cites the reference lines it replaces.
#include "common.cuh"
#include "kernels.h"

namespace gtav {

// ------------------------------------------------------------------------------------------
// Sinusoidal timestep features + action embedding (reference model/dit.py:95-118 and the
// `external_cond` Linear(25 -> D) of 263-267, 363-364).  One block per conditioning row.
// cosf/sinf are the full-range-reduction versions (arguments reach 999 rad); no fast-math.
// ------------------------------------------------------------------------------------------
__global__ void cond_prep_kernel(const int64_t* __restrict__ t, const float* __restrict__ actions, int act_dim,
                                 const float* __restrict__ freqs, const bf16* __restrict__ Wa,
                                 const bf16* __restrict__ ba, bf16* __restrict__ temb, bf16* __restrict__ aemb, int D) {
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x;
    const float tv = static_cast<float>(t[r]);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        const float a = tv * freqs[i];
        temb[r * 256 + i] = __float2bfloat16_rn(cosf(a));
        temb[r * 256 + 128 + i] = __float2bfloat16_rn(sinf(a));
    }
    if (actions != nullptr) {
        __shared__ float sa[64];
        for (int k = threadIdx.x; k < act_dim; k += blockDim.x) sa[k] = bf16_round(actions[r * act_dim + k]);
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += blockDim.x) {
            float acc = 0.f;
            for (int k = 0; k < act_dim; ++k) acc += sa[k] * __bfloat162float(Wa[c * act_dim + k]);
            aemb[static_cast<size_t>(r) * D + c] = __float2bfloat16_rn(acc + __bfloat162float(ba[c]));
        }
    }
}

int launch_cond_prep(const int64_t* t, const float* actions, int act_dim, int R, const float* freqs, const bf16* Wa,
                     const bf16* ba, bf16* temb, bf16* aemb, int D, cudaStream_t s) {
    if (R <= 0) return 0;
    if (act_dim > 64) {
        set_error("cond_prep: action dimension %d > 64", act_dim);
        return -1;
    }
    GTAV_CUDA_OK(launch_k(cond_prep_kernel, dim3(R), dim3(256), 0, s, t, actions, act_dim, freqs, Wa, ba, temb, aemb, D));
    return 0;
}

// ------------------------------------------------------------------------------------------
// Patchify: the stride-p convolution's im2col (reference model/dit.py:60-72 for both the DiT's
// Conv2d(16->D, k=s=2) and the VAE's Conv2d(3->D, k=s=20)); k = c*p*p + ph*p + pw, rounded to bf16
// (autocast casts the conv input), columns [C*p*p, ldo) zero-filled.
// ------------------------------------------------------------------------------------------
template <typename Tin>
__global__ void patchify_kernel(const Tin* __restrict__ x, bf16* __restrict__ out, int ldo, int F, int C, int H, int W,
                                int p, long total, int frames_per_group, long group_stride) {
    pdl_trigger();
    pdl_wait();
    const int gw = W / p, gh = H / p, pp = p * p, K = C * pp;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % ldo);
        const long row = i / ldo;
        float v = 0.f;
        if (k < K) {
            const int wx = static_cast<int>(row % gw), hy = static_cast<int>((row / gw) % gh);
            const long f = row / (static_cast<long>(gw) * gh);
            const int c = k / pp, ph = (k % pp) / p, pw = k % p;
            // frame f = (group g, j): groups of frames_per_group contiguous frames, group_stride elements apart
            const long fbase = (f / frames_per_group) * group_stride + (f % frames_per_group) * static_cast<long>(C) * H * W;
            v = static_cast<float>(x[fbase + (static_cast<long>(c) * H + hy * p + ph) * W + wx * p + pw]);
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

int launch_patchify(const void* x, int x_is_bf16, bf16* out, int ldo, int F, int C, int H, int W, int p,
                    int frames_per_group, long group_stride, cudaStream_t s) {
    if (H % p || W % p || C * p * p > ldo) {
        set_error("patchify: bad geometry C=%d H=%d W=%d p=%d ldo=%d", C, H, W, p, ldo);
        return -1;
    }
    const long total = static_cast<long>(F) * (H / p) * (W / p) * ldo;
    if (total <= 0) return 0;
    const int grid = static_cast<int>(min(static_cast<long>(148 * 16), (total + 255) / 256));
    if (frames_per_group <= 0) { frames_per_group = F > 0 ? F : 1; group_stride = 0; }
    if (x_is_bf16)
        GTAV_CUDA_OK(launch_k(patchify_kernel<bf16>, dim3(grid), dim3(256), 0, s, static_cast<const bf16*>(x), out, ldo, F, C, H, W, p,
                              total, frames_per_group, group_stride));
    else
        GTAV_CUDA_OK(launch_k(patchify_kernel<float>, dim3(grid), dim3(256), 0, s, static_cast<const float*>(x), out, ldo, F, C, H, W, p,
                              total, frames_per_group, group_stride));
    return 0;
}

// ------------------------------------------------------------------------------------------
// DiT un-patchify (reference model/dit.py:328-341, "nhwpqc->nchpwq")
// ------------------------------------------------------------------------------------------
__global__ void dit_unpatchify_kernel(const bf16* __restrict__ y, bf16* __restrict__ out, int C, int gh, int gw, int p,
                                      long total) {
    pdl_trigger();
    pdl_wait();
    const int H = gh * p, W = gw * p, ldy = p * p * C;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int X = static_cast<int>(i % W), Y = static_cast<int>((i / W) % H);
        const int c = static_cast<int>((i / (static_cast<long>(W) * H)) % C);
        const long f = i / (static_cast<long>(W) * H * C);
        const long row = (f * gh + Y / p) * gw + X / p;
        out[i] = y[row * ldy + (Y % p) * (p * C) + (X % p) * C + c];
    }
}

int launch_dit_unpatchify(const bf16* y, bf16* out, int F, int C, int gh, int gw, int p, cudaStream_t s) {
    const long total = static_cast<long>(F) * C * gh * p * gw * p;
    if (total <= 0) return 0;
    const int grid = static_cast<int>(min(static_cast<long>(148 * 16), (total + 255) / 256));
    GTAV_CUDA_OK(launch_k(dit_unpatchify_kernel, dim3(grid), dim3(256), 0, s, y, out, C, gh, gw, p, total));
    return 0;
}

// ------------------------------------------------------------------------------------------
// VAE un-patchify (reference model/vae.py:279-304) and, optionally fused, the pixel epilogue of
// generate.py:241-244: (y+1)/2 in bf16, *255 in bf16, clamp to [0,255], truncate to uint8, HWC.
// ------------------------------------------------------------------------------------------
template <bool TO_U8>
__global__ void vae_unpatchify_kernel(const bf16* __restrict__ y, void* __restrict__ out, int sh, int sw, int p,
                                      long total_pixels) {
    pdl_trigger();
    pdl_wait();
    const int H = sh * p, W = sw * p, pp = p * p, ldy = 3 * pp;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total_pixels;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int X = static_cast<int>(i % W), Y = static_cast<int>((i / W) % H);
        const long f = i / (static_cast<long>(W) * H);
        const bf16* src = y + ((f * sh + Y / p) * sw + X / p) * ldy + (Y % p) * p + (X % p);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const bf16 v = src[c * pp];
            if (TO_U8) {
                float a = bf16_round(__bfloat162float(v) + 1.0f);
                a = bf16_round(a * 0.5f);
                a = bf16_round(a * 255.0f);
                a = fminf(fmaxf(a, 0.0f), 255.0f);
                static_cast<uint8_t*>(out)[i * 3 + c] = static_cast<uint8_t>(a);
            } else {
                static_cast<bf16*>(out)[((f * 3 + c) * H + Y) * W + X] = v;
            }
        }
    }
}

int launch_vae_unpatchify(const bf16* y, void* out, int to_u8, int F, int sh, int sw, int p, cudaStream_t s) {
    const long total = static_cast<long>(F) * sh * p * sw * p;
    if (total <= 0) return 0;
    const int grid = static_cast<int>(min(static_cast<long>(148 * 32), (total + 255) / 256));
    if (to_u8) GTAV_CUDA_OK(launch_k(vae_unpatchify_kernel<true>, dim3(grid), dim3(256), 0, s, y, out, sh, sw, p, total));
    else GTAV_CUDA_OK(launch_k(vae_unpatchify_kernel<false>, dim3(grid), dim3(256), 0, s, y, out, sh, sw, p, total));
    return 0;
}

// ------------------------------------------------------------------------------------------
// latent casts around the VAE bottleneck (generate.py:56 `mean * scaling_factor`, :241 `x / 0.0784`)
// ------------------------------------------------------------------------------------------
__global__ void cast_pad_kernel(const float* __restrict__ z, bf16* __restrict__ out, long rows, int C, int ldo,
                                float divisor) {
    pdl_trigger();
    pdl_wait();
    const long total = rows * ldo;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % ldo);
        const float v = k < C ? __fdiv_rn(z[(i / ldo) * C + k], divisor) : 0.f;
        out[i] = __float2bfloat16_rn(v);
    }
}

int launch_cast_pad(const float* z, bf16* out, int rows, int C, int ldo, float divisor, cudaStream_t s) {
    const long total = static_cast<long>(rows) * ldo;
    if (total <= 0) return 0;
    const int grid = static_cast<int>(min(static_cast<long>(148 * 16), (total + 255) / 256));
    GTAV_CUDA_OK(launch_k(cast_pad_kernel, dim3(grid), dim3(256), 0, s, z, out, static_cast<long>(rows), C, ldo, divisor));
    return 0;
}

__global__ void take_mean_kernel(const bf16* __restrict__ moments, int ldm, float* __restrict__ out, long rows, int C,
                                 float scale, int round_bf16) {
    pdl_trigger();
    pdl_wait();
    const long total = rows * C;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        float v = __bfloat162float(moments[(i / C) * ldm + (i % C)]) * scale;
        out[i] = round_bf16 ? bf16_round(v) : v;
    }
}

int launch_take_mean(const bf16* moments, int ldm, float* out, int rows, int C, float scale, int round_bf16,
                     cudaStream_t s) {
    const long total = static_cast<long>(rows) * C;
    if (total <= 0) return 0;
    const int grid = static_cast<int>(min(static_cast<long>(148 * 16), (total + 255) / 256));
    GTAV_CUDA_OK(launch_k(take_mean_kernel, dim3(grid), dim3(256), 0, s, moments, ldm, out, static_cast<long>(rows), C, scale, round_bf16));
    return 0;
}

// ------------------------------------------------------------------------------------------
// Fused v-prediction DDIM update (reference train_dit.py:110-123).  fp32, same operation order as
// the reference's tensor expression; _rn intrinsics keep nvcc from contracting into FMAs so the
// result is bit-identical to the separate torch kernels.
// ------------------------------------------------------------------------------------------
__global__ void ddim_kernel(const float* __restrict__ x, long x_stride, const bf16* __restrict__ v, long v_stride,
                            float* __restrict__ out, long out_stride, int n, const float* __restrict__ abar_t,
                            const float* __restrict__ abar_next, const int* __restrict__ final_flag) {
    pdl_trigger();
    pdl_wait();
    const int f = blockIdx.y;
    const float a = abar_t[f], an = abar_next[f];
    const float c_x = sqrtf(a), c_v = sqrtf(__fsub_rn(1.0f, a));
    const float inv_a = __fdiv_rn(1.0f, a);
    const float c_ix = sqrtf(inv_a), c_den = sqrtf(__fsub_rn(inv_a, 1.0f));
    const float c_n0 = sqrtf(an), c_ne = sqrtf(__fsub_rn(1.0f, an));
    const bool fin = (*final_flag) != 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float xv = x[f * x_stride + i], vv = __bfloat162float(v[f * v_stride + i]);
        const float x0 = __fsub_rn(__fmul_rn(c_x, xv), __fmul_rn(c_v, vv));
        float r = x0;
        if (!fin) {
            const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(c_ix, xv), x0), c_den);
            r = __fadd_rn(__fmul_rn(c_n0, x0), __fmul_rn(c_ne, eps));
        }
        out[f * out_stride + i] = r;
    }
}

int launch_ddim(const float* x, long x_stride, const bf16* v, long v_stride, float* out, long out_stride, int F, int n,
                const float* abar_t, const float* abar_next, const int* final_flag, cudaStream_t s) {
    if (F <= 0 || n <= 0) return 0;
    GTAV_CUDA_OK(launch_k(ddim_kernel, dim3((n + 255) / 256, F), dim3(256), 0, s, x, x_stride, v, v_stride, out, out_stride, n, abar_t,
                          abar_next, final_flag));
    return 0;
}

}  // namespace gtav

namespace gtav {

// ------------------------------------------------------------------------------------------
// Per-step sampler bookkeeping (replaces the host-side torch.full / indexing of train_dit.py:64-99,
// 110,116-117).  Conditioning-table row layout: context frame j of rollout b -> b*(T-1)+j;
// last frame of rollout b at step k -> B*(T-1) + b*(steps+1) + k.
// ------------------------------------------------------------------------------------------
__global__ void step_prep_kernel(int* counter, const int* __restrict__ levels, const float* __restrict__ abar, int B,
                                 int T, int steps, int* __restrict__ frame_row, int* __restrict__ last_row,
                                 float* __restrict__ abar_t, float* __restrict__ abar_next, int* __restrict__ final_flag) {
    pdl_trigger();
    pdl_wait();
    const int k = *counter;
    __syncthreads();
    for (int f = threadIdx.x; f < B * T; f += blockDim.x) {
        const int b = f / T, j = f % T;
        frame_row[f] = (j < T - 1) ? b * (T - 1) + j : B * (T - 1) + b * (steps + 1) + k;
    }
    const float at = abar[levels[k]], an = abar[levels[k > 0 ? k - 1 : 0]];
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        abar_t[b] = at;
        abar_next[b] = an;
        last_row[b] = B * (T - 1) + b * (steps + 1) + k;       // the same row, for the last-frame-only backbone
    }
    if (threadIdx.x == 0) {
        *final_flag = k <= 0 ? 1 : 0;
        *counter = k - 1;
    }
}

int launch_step_prep(int* counter, const int* levels, const float* abar, int B, int T, int steps, int* frame_row,
                     int* last_row, float* abar_t, float* abar_next, int* final_flag, cudaStream_t s) {
    GTAV_CUDA_OK(launch_k(step_prep_kernel, dim3(1), dim3(256), 0, s, counter, levels, abar, B, T, steps, frame_row, last_row, abar_t,
                          abar_next, final_flag));
    return 0;
}

__global__ void set_int_kernel(int* dst, int value) {
    pdl_trigger();
    pdl_wait(); *dst = value; }

int launch_set_int(int* dst, int value, cudaStream_t s) {
    GTAV_CUDA_OK(launch_k(set_int_kernel, dim3(1), dim3(1), 0, s, dst, value));
    return 0;
}

__global__ void noise_clamp_kernel(const float* __restrict__ noise, float* __restrict__ x, long x_stride, int n,
                                   float amax) {
    pdl_trigger();
    pdl_wait();
    const int f = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        x[f * x_stride + i] = fminf(fmaxf(noise[static_cast<long>(f) * n + i], -amax), amax);
}

int launch_noise_clamp(const float* noise, float* x, long x_stride, int F, int n, float amax, cudaStream_t s) {
    if (F <= 0 || n <= 0) return 0;
    GTAV_CUDA_OK(launch_k(noise_clamp_kernel, dim3((n + 255) / 256, F), dim3(256), 0, s, noise, x, x_stride, n, amax));
    return 0;
}

}  // namespace gtav
