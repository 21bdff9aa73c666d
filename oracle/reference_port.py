"""CPU restatement ("port") of the reference inference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (ai-generated-gtav_b200/) never does and has no CPU fallback.

Everything is a pure function of a state_dict (see oracle/weights.py), written as plain fp32 torch
on whatever device the tensors live on (CPU in practice).  Each function cites the reference
code it restates.  `Rounding` lets the same code reproduce the places where the reference running
under `torch.autocast("cuda", bfloat16)` rounds to bf16, so the CUDA product can be compared both
with the exact-fp32 oracle (stated tolerance) and with the bf16-rounding-aware oracle (tight).

Pinned against outputs of the unmodified reference modules (oracle/make_golden.py ->
tests/golden/*.safetensors); see tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from .weights import DiTConfig, VAEConfig

SCALING_FACTOR = 0.07843137255  # reference generate.py:51,241


# --------------------------------------------------------------------------------------------
# rounding model
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Rounding:
    """bf16=False: exact fp32 everywhere (what the reference computes on CPU).
    bf16=True : round to bf16 wherever CUDA autocast would (SURVEY.md §3.3 numerics note).
    sdpa=True : attention through F.scaled_dot_product_attention like the reference (model/attention.py:62,127) instead
                of the explicit softmax - used with tensors on a CUDA device under torch.autocast(bf16) to time the
                reference's own eager graph (cuBLAS + SDPA) on the GPU (bench.py: gpu_eager_baseline)."""
    bf16: bool = False
    sdpa: bool = False

    def r(self, x: torch.Tensor) -> torch.Tensor:
        return x.to(torch.bfloat16).to(torch.float32) if self.bf16 else x


FP32 = Rounding(False)
BF16 = Rounding(True)
EAGER = Rounding(False, True)


def _linear(rd: Rounding, x, w, b=None):
    """nn.Linear under autocast: operands rounded to bf16, fp32 accumulate, result rounded."""
    y = rd.r(x) @ rd.r(w).t()
    if b is not None:
        y = y + rd.r(b)
    return rd.r(y)


def _layer_norm(x, w=None, b=None, eps=1e-6):
    """nn.LayerNorm runs in fp32 under autocast (autocast fp32 list); eps 1e-6 everywhere on this
    path (reference model/dit.py:133,163; model/vae.py:174)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


# --------------------------------------------------------------------------------------------
# schedule + sampler (reference utils.py:30-48, train_dit.py:30-125, generate.py:192-220)
# --------------------------------------------------------------------------------------------
def sigmoid_beta_schedule(timesteps: int, start=-3.0, end=3.0, tau=1.0, clamp_min=1e-4) -> torch.Tensor:
    """float64 betas of the sigmoid schedule rescaled into [clamp_min, 1] (reference utils.py:30-48)."""
    u = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
    lo = torch.sigmoid(torch.tensor(start / tau))
    hi = torch.sigmoid(torch.tensor(end / tau))
    abar = (hi - torch.sigmoid((u * (end - start) + start) / tau)) / (hi - lo)
    abar = abar / abar[0]
    abar = abar * (1 - clamp_min) + clamp_min
    return (1 - abar[1:] / abar[:-1]).clamp(0, 0.999)


def alphas_cumprod_table(max_noise_level: int = 1000) -> torch.Tensor:
    """fp32 cumprod(1 - betas) exactly as reference generate.py:195-197 builds it."""
    betas = sigmoid_beta_schedule(max_noise_level).float()
    return torch.cumprod(1.0 - betas, dim=0)


def noise_levels(noise_steps: int, max_noise_level: int = 1000) -> list[int]:
    """Integer timesteps visited by the sampler: linspace(0, 999, steps+1) truncated by
    `torch.full(..., dtype=long)` (reference generate.py:194, train_dit.py:72-87)."""
    return [int(v) for v in torch.linspace(0, max_noise_level - 1, noise_steps + 1).tolist()]


def ddim_update(x, v, abar_t, abar_next, final: bool):
    """v-prediction DDIM algebra of reference train_dit.py:110-123 (all fp32, broadcast coefficients)."""
    x0 = abar_t.sqrt() * x - (1 - abar_t).sqrt() * v
    eps = ((1 / abar_t).sqrt() * x - x0) / (1 / abar_t - 1).sqrt()
    if final:
        return x0
    return abar_next.sqrt() * x0 + (1 - abar_next).sqrt() * eps


def denoise_step(sd, cfg: DiTConfig, x_noisy, actions, noise_idx, stabilization_level, levels,
                 abar, start_frame=0, rd: Rounding = FP32):
    """One sampler step (reference train_dit.py:30-125): context frames at t=stabilization_level,
    last frame at levels[noise_idx]; returns (x_pred, v_pred) for the window."""
    B, F_all = x_noisy.shape[:2]
    t = torch.full((B, F_all), stabilization_level, dtype=torch.long, device=x_noisy.device)
    t_next = t.clone()
    t[:, -1] = levels[noise_idx]
    t_next[:, -1] = levels[max(0, noise_idx - 1)]
    xw, tw, tnw = x_noisy[:, start_frame:], t[:, start_frame:], t_next[:, start_frame:]
    aw = None if actions is None else actions[:, start_frame:start_frame + xw.shape[1]]
    v = dit_forward(sd, cfg, xw, tw, aw, rd)
    a_t = abar[tw].view(B, -1, 1, 1, 1)
    a_n = abar[tnw].view(B, -1, 1, 1, 1).clone()
    a_n[:, :-1] = 1.0
    return ddim_update(xw.float(), v, a_t, a_n, noise_idx <= 0), v


def rollout(dit_sd, dcfg: DiTConfig, prompt_latents, actions, total_frames, noise_steps, noise_fn,
            stabilization_level=15, noise_abs_max=20.0, rd: Rounding = FP32, on_step=None, abar=None):
    """Autoregressive loop of reference generate.py:192-220.  `noise_fn(i)` returns the [B,1,C,H,W]
    Gaussian draw for frame i (the caller owns the RNG so product and oracle see the same noise).
    abar: the cumulative-alpha table; default generate.py's (clamp_min 1e-4) - DiffusionTrainer.predict runs the same
    loop on its clamp_min 1e-6 table (train_dit.py:297-307, 400-450)."""
    abar = alphas_cumprod_table() if abar is None else abar
    levels = noise_levels(noise_steps)
    x = prompt_latents.float()
    n_prompt = x.shape[1]
    for i in range(n_prompt, total_frames):
        chunk = noise_fn(i).clamp(-noise_abs_max, noise_abs_max)
        x = torch.cat([x, chunk], dim=1)
        start = max(0, i + 1 - dcfg.max_frames)
        for k in reversed(range(noise_steps + 1)):
            xp, v = denoise_step(dit_sd, dcfg, x, actions, k, stabilization_level, levels, abar, start, rd)
            x[:, -1:] = xp[:, -1:]
            if on_step is not None:
                on_step(i, k, x, v)
    return x


# --------------------------------------------------------------------------------------------
# rotary tables (reference model/rotary_embedding_torch.py:39-73, 186-209, 290-345)
# --------------------------------------------------------------------------------------------
def _interleave2(a):  # "... n -> ... (n r)", r=2
    return a.repeat_interleave(2, dim=-1)


def axial_angles(base: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """[h, w, 4*len(base)] angle table of `get_axial_freqs(h, w)` for freqs_for="pixel": positions
    linspace(-1,1,n) per axis, each angle repeated twice, row angles first then column angles."""
    ah = _interleave2(torch.linspace(-1, 1, h, device=base.device)[:, None] * base[None])
    aw = _interleave2(torch.linspace(-1, 1, w, device=base.device)[:, None] * base[None])
    return torch.cat([ah[:, None].expand(h, w, -1), aw[None].expand(h, w, -1)], dim=-1)


def temporal_angles(base: torch.Tensor, T: int) -> torch.Tensor:
    """[T, 2*len(base)] table of `rotate_queries_or_keys`: window-relative positions arange(T)."""
    return _interleave2(torch.arange(T, dtype=torch.float32, device=base.device)[:, None] * base[None])


def apply_rotary(rd: Rounding, ang: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """Rotate adjacent pairs of the first ang.shape[-1] features: out[2p] = x[2p]cos - x[2p+1]sin,
    out[2p+1] = x[2p+1]cos + x[2p]sin; fp32 math, one rounding (rotary_embedding_torch.py:39-73)."""
    n = ang.shape[-1]
    head, tail = x[..., :n], x[..., n:]
    pair = head.reshape(*head.shape[:-1], n // 2, 2)
    swapped = torch.stack((-pair[..., 1], pair[..., 0]), dim=-1).reshape(head.shape)
    return torch.cat([rd.r(head * ang.cos() + swapped * ang.sin()), tail], dim=-1)


def _attention(rd: Rounding, q, k, v, causal: bool):
    """softmax(q k^T / sqrt(d)) v; fp32 softmax, probabilities rounded before the second product
    when emulating the bf16 SDPA (reference calls F.scaled_dot_product_attention,
    model/attention.py:62,127; model/vae.py:101)."""
    if rd.sdpa:
        return F.scaled_dot_product_attention(q, k, v, is_causal=causal)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(q.shape[-1]))
    if causal:
        n = s.shape[-1]
        s = s.masked_fill(torch.ones(n, n, dtype=torch.bool, device=s.device).triu(1), float("-inf"))
    p = torch.softmax(s, dim=-1)
    return rd.r(rd.r(p) @ v)


# --------------------------------------------------------------------------------------------
# DiT (reference model/dit.py, model/attention.py)
# --------------------------------------------------------------------------------------------
def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: float = 10000.0):
    """[cos(t f), sin(t f)] with f = exp(-ln(max_period) * arange(half)/half) (model/dit.py:95-118)."""
    half = dim // 2
    f = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t[:, None].float() * f[None]
    return torch.cat([a.cos(), a.sin()], dim=-1)


def _modulate(rd: Rounding, xn, shift, scale):
    """model/dit.py:19-27: scale+1e-6 and 1+scale are evaluated in the modulation dtype (bf16 under
    autocast), the product with the fp32 LayerNorm output and the sum stay fp32."""
    s1 = rd.r(1 + rd.r(scale + 1e-6))
    return xn * s1 + shift


def _half_block(rd, sd, cfg, pre, x, c_act, axis):
    """One adaLN-zero half (attention + MLP) of SpatioTemporalDiTBlock (model/dit.py:200-225).
    x: [B,T,S,D] residual stream; c_act: SiLU(c) [B,T,D]; axis "s" or "t"."""
    B, T, S, D = x.shape
    H, dh = cfg.num_heads, cfg.head_dim
    mod = _linear(rd, c_act, sd[f"{pre}_adaLN_modulation.1.weight"], sd[f"{pre}_adaLN_modulation.1.bias"])
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = [m[:, :, None, :] for m in mod.chunk(6, dim=-1)]

    h = _modulate(rd, _layer_norm(x), sh_a, sc_a)
    qkv = _linear(rd, h, sd[f"{pre}_attn.to_qkv.weight"])
    q, k, v = [z.reshape(B, T, S, H, dh) for z in qkv.chunk(3, dim=-1)]
    if axis == "s":  # per frame over the S = grid_h*grid_w tokens (model/attention.py:99-136)
        ang = axial_angles(sd["spatial_rotary_emb.freqs"], cfg.grid_h, cfg.grid_w).reshape(S, dh)
        q, k, v = [z.permute(0, 1, 3, 2, 4) for z in (q, k, v)]          # B T H S d
        o = _attention(rd, apply_rotary(rd, ang, q), apply_rotary(rd, ang, k), v, causal=False)
        o = o.permute(0, 1, 3, 2, 4)
    else:            # per spatial position over the T frames, causal (model/attention.py:41-71)
        ang = temporal_angles(sd["temporal_rotary_emb.freqs"], T)
        q, k, v = [z.permute(0, 2, 3, 1, 4) for z in (q, k, v)]          # B S H T d
        o = _attention(rd, apply_rotary(rd, ang, q), apply_rotary(rd, ang, k), v, causal=True)
        o = o.permute(0, 3, 1, 2, 4)
    o = o.reshape(B, T, S, D)
    y = _linear(rd, o, sd[f"{pre}_attn.to_out.weight"], sd[f"{pre}_attn.to_out.bias"])
    x = rd.r(x + rd.r(g_a * y))

    h = _modulate(rd, _layer_norm(x), sh_m, sc_m)
    h = _linear(rd, h, sd[f"{pre}_mlp.fc1.weight"], sd[f"{pre}_mlp.fc1.bias"])
    h = rd.r(F.gelu(h, approximate="tanh"))
    y = _linear(rd, h, sd[f"{pre}_mlp.fc2.weight"], sd[f"{pre}_mlp.fc2.bias"])
    return rd.r(x + rd.r(g_m * y))


def dit_patch_embed(rd, sd, cfg, x):
    """Conv2d(k=s=patch) as a per-patch linear map, output [B,T,S,D] (model/dit.py:38-76, 353-358)."""
    B, T, C, Hh, Ww = x.shape
    p = cfg.patch_size
    patches = x.reshape(B, T, C, Hh // p, p, Ww // p, p).permute(0, 1, 3, 5, 2, 4, 6)
    patches = patches.reshape(B, T, cfg.tokens, C * p * p)               # k = c*p*p + ph*p + pw
    w = sd["x_embedder.proj.weight"].reshape(cfg.hidden_size, -1)
    return _linear(rd, patches.float(), w, sd["x_embedder.proj.bias"])


def dit_conditioning(rd, sd, cfg, t, external_cond):
    """c = t_embedder(t) (+ external_cond(actions)) (model/dit.py:86-92, 120-123, 359-364) -> [B,T,D]."""
    B, T = t.shape
    e = timestep_embedding(t.reshape(-1))
    h = _linear(rd, e, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    h = rd.r(F.silu(h))
    c = _linear(rd, h, sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"]).reshape(B, T, -1)
    if external_cond is not None:
        c = rd.r(c + _linear(rd, external_cond.float(), sd["external_cond.weight"], sd["external_cond.bias"]))
    return c


def dit_unpatchify(cfg, y):
    """[B,T,S,p*p*C] -> [B,T,C,H,W]; feature index = ph*(p*C) + pw*C + c (model/dit.py:328-341)."""
    B, T = y.shape[:2]
    p, C = cfg.patch_size, cfg.in_channels
    y = y.reshape(B, T, cfg.grid_h, cfg.grid_w, p, p, C).permute(0, 1, 6, 2, 4, 3, 5)
    return y.reshape(B, T, C, cfg.grid_h * p, cfg.grid_w * p)


def dit_forward(sd, cfg: DiTConfig, x, t, external_cond=None, rd: Rounding = FP32):
    """DiT.forward(x[B,T,C,H,W], t[B,T] long, external_cond[B,T,25]|None) -> v[B,T,C,H,W]
    (reference model/dit.py:343-376)."""
    h = dit_patch_embed(rd, sd, cfg, x)
    c = dit_conditioning(rd, sd, cfg, t, external_cond)
    c_act = rd.r(F.silu(c))
    for n in range(cfg.depth):
        h = _half_block(rd, sd, cfg, f"blocks.{n}.s", h, c_act, "s")
        h = _half_block(rd, sd, cfg, f"blocks.{n}.t", h, c_act, "t")
    mod = _linear(rd, c_act, sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"])
    shift, scale = [m[:, :, None, :] for m in mod.chunk(2, dim=-1)]
    y = _linear(rd, _modulate(rd, _layer_norm(h), shift, scale),
                sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    return dit_unpatchify(cfg, y)


# --------------------------------------------------------------------------------------------
# VAE (reference model/vae.py)
# --------------------------------------------------------------------------------------------
def vae_rotary_angles(cfg: VAEConfig) -> torch.Tensor:
    """`RotaryEmbedding(dim=head_dim//4, "pixel", max_freq=seq_h*seq_w).get_axial_freqs(seq_h, seq_w)`
    flattened to [seq_len, head_dim//2] (model/vae.py:71-76): only the first half of each head rotates."""
    dim = cfg.head_dim // 4
    base = torch.linspace(1.0, cfg.seq_len / 2, dim // 2) * math.pi
    return axial_angles(base, cfg.seq_h, cfg.seq_w).reshape(cfg.seq_len, -1)


def _vae_block(rd, sd, cfg, pre, x, ang):
    """Pre-LN attention block with plain residuals (model/vae.py:78-112, 154-157)."""
    N, S, D = x.shape
    H, dh = cfg.heads, cfg.head_dim
    h = _layer_norm(x, sd[f"{pre}.norm1.weight"], sd[f"{pre}.norm1.bias"])
    qkv = _linear(rd, h, sd[f"{pre}.attn.qkv.weight"], sd[f"{pre}.attn.qkv.bias"])
    q, k, v = qkv.reshape(N, S, 3, H, dh).permute(2, 0, 3, 1, 4)
    o = _attention(rd, apply_rotary(rd, ang, q), apply_rotary(rd, ang, k), v, causal=False)
    o = o.transpose(1, 2).reshape(N, S, D)
    x = rd.r(x + _linear(rd, o, sd[f"{pre}.attn.proj.weight"], sd[f"{pre}.attn.proj.bias"]))
    h = _layer_norm(x, sd[f"{pre}.norm2.weight"], sd[f"{pre}.norm2.bias"])
    h = _linear(rd, h, sd[f"{pre}.mlp.fc1.weight"], sd[f"{pre}.mlp.fc1.bias"])
    h = rd.r(F.gelu(h))
    return rd.r(x + _linear(rd, h, sd[f"{pre}.mlp.fc2.weight"], sd[f"{pre}.mlp.fc2.bias"]))


def vae_encode_mean(sd, cfg: VAEConfig, img, rd: Rounding = FP32):
    """`vae.encode(img).mean`: img [N,3,H,W] in [-1,1] -> [N, seq_len, latent_dim]
    (model/vae.py:306-322; the posterior's logvar/std are unused by generate.py:56)."""
    return vae_encode_moments(sd, cfg, img, rd)[..., : cfg.latent_dim]


def vae_encode_moments(sd, cfg: VAEConfig, img, rd: Rounding = FP32):
    """quant_conv's output [N, seq_len, 2*latent_dim] = (mean | logvar) of DiagonalGaussianDistribution
    (model/vae.py:19-45, 306-322), logvar unclamped."""
    N = img.shape[0]
    p = cfg.patch_size
    patches = img.reshape(N, 3, cfg.seq_h, p, cfg.seq_w, p).permute(0, 2, 4, 1, 3, 5)
    patches = patches.reshape(N, cfg.seq_len, cfg.patch_dim)            # k = c*p*p + ph*p + pw
    x = _linear(rd, patches.float(), sd["patch_embed.proj.weight"].reshape(cfg.dim, -1), sd["patch_embed.proj.bias"])
    ang = vae_rotary_angles(cfg)
    for n in range(cfg.enc_depth):
        x = _vae_block(rd, sd, cfg, f"encoder.{n}", x, ang)
    x = _layer_norm(x, sd["enc_norm.weight"], sd["enc_norm.bias"])
    return _linear(rd, x, sd["quant_conv.weight"], sd["quant_conv.bias"])


def vae_decode(sd, cfg: VAEConfig, z, rd: Rounding = FP32):
    """`vae.decode(z)`: z [N, seq_len, latent_dim] -> [N,3,H,W] (model/vae.py:324-338, 279-304);
    predictor feature index = c*p*p + ph*p + pw."""
    N = z.shape[0]
    x = _linear(rd, z.float(), sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    ang = vae_rotary_angles(cfg)
    for n in range(cfg.dec_depth):
        x = _vae_block(rd, sd, cfg, f"decoder.{n}", x, ang)
    x = _layer_norm(x, sd["dec_norm.weight"], sd["dec_norm.bias"])
    y = _linear(rd, x, sd["predictor.weight"], sd["predictor.bias"])
    p = cfg.patch_size
    y = y.reshape(N, cfg.seq_h, cfg.seq_w, 3, p, p).permute(0, 3, 1, 4, 2, 5)
    return y.reshape(N, 3, cfg.input_height, cfg.input_width)


def encode_prompt(sd, cfg: VAEConfig, video, rd: Rounding = FP32):
    """`vae_encode` of reference generate.py:50-66: video [B,n,3,H,W] in [0,1] ->
    latents [B,n,latent_dim,seq_h,seq_w] scaled by SCALING_FACTOR."""
    B, n = video.shape[:2]
    m = rd.r(vae_encode_mean(sd, cfg, video.reshape(B * n, *video.shape[2:]) * 2 - 1, rd) * SCALING_FACTOR)
    return m.reshape(B, n, cfg.seq_h, cfg.seq_w, cfg.latent_dim).permute(0, 1, 4, 2, 3).contiguous()


def decode_to_uint8(sd, cfg: VAEConfig, latents, rd: Rounding = FP32):
    """Decode + pixel epilogue of reference generate.py:238-244: latents [B,F,C,h,w] ->
    uint8 [B,F,H,W,3]; (y+1)/2, *255, clamp, truncate (`.byte()`)."""
    B, Fr = latents.shape[:2]
    z = latents.permute(0, 1, 3, 4, 2).reshape(B * Fr, cfg.seq_len, cfg.latent_dim)
    y = vae_decode(sd, cfg, z / SCALING_FACTOR, rd)
    y = rd.r(rd.r(y + 1) / 2)
    y = rd.r(y * 255).clamp(0, 255).to(torch.uint8)
    return y.reshape(B, Fr, 3, cfg.input_height, cfg.input_width).permute(0, 1, 3, 4, 2).contiguous()
