// Non-causal multi-head attention over short sequences with the rotary embedding fused on load.
//
// Replaces SpatialAxialAttention's rearrange + get_axial_freqs + apply_rotary_emb + SDPA
// (reference model/attention.py:111-129, S = 144 tokens, rotary on all 64 head dims) and the VAE
// Attention (model/vae.py:83-107, S = 576, rotary on head dims 0..31 only).  Head dim is 64.
//
// This is 1.2 % (DiT) / 8.5 % (VAE) of the step's FLOPs, so round 1 keeps it on warp-level
// mma.sync m16n8k16 tiles (flash-style: scores stay in registers, online softmax across key
// blocks, P rounded to bf16 before P@V like the reference's fused SDPA backends).  q/k are rotated
// in fp32 and rounded to bf16 once, as apply_rotary_emb does (rotary_embedding_torch.py:46-73).
#include "attn_seq.cuh"

namespace gtav {

// grid: (SEQ / (WARPS*16), heads, groups).  qkv rows of one group are consecutive.
template <int SEQ, int KB, int WARPS, int ROT_PAIRS>
__global__ void __launch_bounds__(WARPS * 32)
attn_seq_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int heads, const float2* __restrict__ rot) {
    __shared__ __align__(16) bf16 sK[KB * SROW];
    __shared__ __align__(16) bf16 sV[KB * SROW];
    pdl_trigger();
    pdl_wait();
    attn_seq_body<SEQ, KB, WARPS, ROT_PAIRS, false>(qkv, out, heads, rot, sK, sV, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, 0);
}

int launch_attention_seq(const bf16* qkv, bf16* out, int groups, int seq, int heads, const float2* rot, int rot_pairs,
                         cudaStream_t s) {
    if (groups <= 0) return 0;
    if (seq == 144 && rot_pairs == 32) {
        // all 144 keys of a head staged in one pass (every global load of the block in flight at once); the queries
        // are split over 3 CTAs of 3 warps so that 48 SMs share the loads at B = 1
        GTAV_CUDA_OK(launch_k(attn_seq_kernel<144, 144, 3, 32>, dim3(3, heads, groups), dim3(3 * 32), 0, s, qkv, out, heads, rot));
    } else if (seq == 576 && rot_pairs == 16) {
        GTAV_CUDA_OK(launch_k(attn_seq_kernel<576, 64, 4, 16>, dim3(9, heads, groups), dim3(4 * 32), 0, s, qkv, out, heads, rot));
    } else {
        set_error("attention: unsupported (seq=%d, rot_pairs=%d); built for (144,32) and (576,16)", seq, rot_pairs);
        return -1;
    }
    return 0;
}

}  // namespace gtav
