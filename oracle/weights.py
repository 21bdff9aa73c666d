"""Deterministic synthetic weights for the DiT / VAE hot path.  TEST INFRASTRUCTURE.

The reference ships no checkpoints offline and its default init makes every DiT block an
exact identity (reference model/dit.py:316-320 zero-inits both adaLN linears), so parity needs a
weight set that (a) exercises attention/MLP, (b) can be rebuilt bit-identically on the GPU box
without /root/reference.  Each tensor is drawn from its own torch CPU generator seeded with
crc32(key) ^ seed, so a tensor's values depend only on (key, shape, seed) and never on
construction order.

Key names / shapes follow the reference state_dict contract:
  DiT: reference model/dit.py:228-326 (334 entries at depth 16, incl. the aliased rotary `freqs`)
  VAE: reference model/vae.py:160-236 (228 entries at enc 6 / dec 12)
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class DiTConfig:
    """Shape of the spatio-temporal DiT (reference model/dit.py:233-244, 379-389)."""
    input_h: int = 18
    input_w: int = 32
    patch_size: int = 2
    in_channels: int = 16
    hidden_size: int = 1024
    depth: int = 16
    num_heads: int = 16
    mlp_ratio: float = 4.0
    external_cond_dim: int = 25
    max_frames: int = 5

    @property
    def grid_h(self): return self.input_h // self.patch_size
    @property
    def grid_w(self): return self.input_w // self.patch_size
    @property
    def tokens(self): return self.grid_h * self.grid_w
    @property
    def head_dim(self): return self.hidden_size // self.num_heads
    @property
    def mlp_hidden(self): return int(self.hidden_size * self.mlp_ratio)


@dataclass(frozen=True)
class VAEConfig:
    """Shape of the ViT-L-20 shallow-encoder VAE (reference model/vae.py:363-380)."""
    latent_dim: int = 16
    input_height: int = 360
    input_width: int = 640
    patch_size: int = 20
    dim: int = 1024
    enc_depth: int = 6
    dec_depth: int = 12
    heads: int = 16
    mlp_ratio: float = 4.0

    @property
    def seq_h(self): return self.input_height // self.patch_size
    @property
    def seq_w(self): return self.input_width // self.patch_size
    @property
    def seq_len(self): return self.seq_h * self.seq_w
    @property
    def patch_dim(self): return 3 * self.patch_size ** 2
    @property
    def head_dim(self): return self.dim // self.heads
    @property
    def mlp_hidden(self): return int(self.dim * self.mlp_ratio)


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(key, shape, std, seed):
    return torch.empty(shape, dtype=torch.float32).normal_(0.0, std, generator=_gen(key, seed))


def _uniform(key, shape, bound, seed):
    return torch.empty(shape, dtype=torch.float32).uniform_(-bound, bound, generator=_gen(key, seed))


def spatial_rotary_base(cfg: DiTConfig) -> torch.Tensor:
    """`RotaryEmbedding(dim=head_dim//2, freqs_for="pixel", max_freq=256).freqs`
    (reference model/dit.py:259-261, model/rotary_embedding_torch.py:124-125)."""
    dim = cfg.head_dim // 2
    return torch.linspace(1.0, 256 / 2, dim // 2) * math.pi


def temporal_rotary_base(cfg: DiTConfig) -> torch.Tensor:
    """`RotaryEmbedding(dim=head_dim).freqs`, "lang" flavour, theta 1e4
    (reference model/dit.py:262, model/rotary_embedding_torch.py:120-123)."""
    dim = cfg.head_dim
    return 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))


def make_dit_state(cfg: DiTConfig = DiTConfig(), seed: int = 0, degenerate: bool = False,
                   include_rotary: bool = True) -> dict[str, torch.Tensor]:
    """fp32 state_dict for the DiT.  `degenerate=True` zeroes the per-block adaLN linears like the
    reference's default init (every block an identity); otherwise they are N(0, 0.02) so that
    attention and MLP contribute ("Init B" of SURVEY.md §8(c), with our own generator)."""
    D, P, C = cfg.hidden_size, cfg.patch_size, cfg.in_channels
    sd: dict[str, torch.Tensor] = {}
    sd["x_embedder.proj.weight"] = _normal("x_embedder.proj.weight", (D, C, P, P), 0.02, seed)
    sd["x_embedder.proj.bias"] = _normal("x_embedder.proj.bias", (D,), 0.02, seed)
    sd["t_embedder.mlp.0.weight"] = _normal("t_embedder.mlp.0.weight", (D, 256), 0.02, seed)
    sd["t_embedder.mlp.0.bias"] = _normal("t_embedder.mlp.0.bias", (D,), 0.02, seed)
    sd["t_embedder.mlp.2.weight"] = _normal("t_embedder.mlp.2.weight", (D, D), 0.02, seed)
    sd["t_embedder.mlp.2.bias"] = _normal("t_embedder.mlp.2.bias", (D,), 0.02, seed)
    if cfg.external_cond_dim > 0:
        sd["external_cond.weight"] = _normal("external_cond.weight", (D, cfg.external_cond_dim), 0.05, seed)
        sd["external_cond.bias"] = _normal("external_cond.bias", (D,), 0.02, seed)
    if include_rotary:
        sd["spatial_rotary_emb.freqs"] = spatial_rotary_base(cfg)
        sd["temporal_rotary_emb.freqs"] = temporal_rotary_base(cfg)
    Hm = cfg.mlp_hidden
    for n in range(cfg.depth):
        for h in ("s", "t"):
            p = f"blocks.{n}.{h}"
            sd[f"{p}_attn.to_qkv.weight"] = _normal(f"{p}_attn.to_qkv.weight", (3 * D, D), 0.02, seed)
            sd[f"{p}_attn.to_out.weight"] = _normal(f"{p}_attn.to_out.weight", (D, D), 0.02, seed)
            sd[f"{p}_attn.to_out.bias"] = _normal(f"{p}_attn.to_out.bias", (D,), 0.02, seed)
            if include_rotary:
                sd[f"{p}_attn.rotary_emb.freqs"] = (
                    sd["spatial_rotary_emb.freqs"] if h == "s" else sd["temporal_rotary_emb.freqs"])
            sd[f"{p}_mlp.fc1.weight"] = _normal(f"{p}_mlp.fc1.weight", (Hm, D), 0.02, seed)
            sd[f"{p}_mlp.fc1.bias"] = _normal(f"{p}_mlp.fc1.bias", (Hm,), 0.02, seed)
            sd[f"{p}_mlp.fc2.weight"] = _normal(f"{p}_mlp.fc2.weight", (D, Hm), 0.02, seed)
            sd[f"{p}_mlp.fc2.bias"] = _normal(f"{p}_mlp.fc2.bias", (D,), 0.02, seed)
            k = f"{p}_adaLN_modulation.1"
            if degenerate:
                sd[f"{k}.weight"] = torch.zeros(6 * D, D)
                sd[f"{k}.bias"] = torch.zeros(6 * D)
            else:
                sd[f"{k}.weight"] = _normal(f"{k}.weight", (6 * D, D), 0.02, seed)
                sd[f"{k}.bias"] = _normal(f"{k}.bias", (6 * D,), 0.02, seed)
    sd["final_layer.linear.weight"] = _normal("final_layer.linear.weight", (P * P * C, D), 0.02, seed)
    sd["final_layer.linear.bias"] = _normal("final_layer.linear.bias", (P * P * C,), 0.02, seed)
    sd["final_layer.adaLN_modulation.1.weight"] = _normal("final_layer.adaLN_modulation.1.weight", (2 * D, D), 0.01, seed)
    sd["final_layer.adaLN_modulation.1.bias"] = _normal("final_layer.adaLN_modulation.1.bias", (2 * D,), 0.02, seed)
    return sd


def make_vae_state(cfg: VAEConfig = VAEConfig(), seed: int = 0) -> dict[str, torch.Tensor]:
    """fp32 state_dict for the VAE: xavier-uniform-scaled linears like the reference
    (model/vae.py:243-260) but with non-trivial biases / LayerNorm affines so those paths are tested."""
    D, L, Pd, Hm = cfg.dim, cfg.latent_dim, cfg.patch_dim, cfg.mlp_hidden
    sd: dict[str, torch.Tensor] = {}

    def xavier(key, out_f, in_f, shape=None):
        b = math.sqrt(6.0 / (in_f + out_f))
        return _uniform(key, shape or (out_f, in_f), b, seed)

    p = cfg.patch_size
    sd["patch_embed.proj.weight"] = xavier("patch_embed.proj.weight", D, Pd, (D, 3, p, p))
    sd["patch_embed.proj.bias"] = _normal("patch_embed.proj.bias", (D,), 0.02, seed)
    for side, depth in (("encoder", cfg.enc_depth), ("decoder", cfg.dec_depth)):
        for n in range(depth):
            q = f"{side}.{n}"
            for nm in ("norm1", "norm2"):
                sd[f"{q}.{nm}.weight"] = 1.0 + _normal(f"{q}.{nm}.weight", (D,), 0.1, seed)
                sd[f"{q}.{nm}.bias"] = _normal(f"{q}.{nm}.bias", (D,), 0.05, seed)
            sd[f"{q}.attn.qkv.weight"] = xavier(f"{q}.attn.qkv.weight", 3 * D, D)
            sd[f"{q}.attn.qkv.bias"] = _normal(f"{q}.attn.qkv.bias", (3 * D,), 0.02, seed)
            sd[f"{q}.attn.proj.weight"] = xavier(f"{q}.attn.proj.weight", D, D)
            sd[f"{q}.attn.proj.bias"] = _normal(f"{q}.attn.proj.bias", (D,), 0.02, seed)
            sd[f"{q}.mlp.fc1.weight"] = xavier(f"{q}.mlp.fc1.weight", Hm, D)
            sd[f"{q}.mlp.fc1.bias"] = _normal(f"{q}.mlp.fc1.bias", (Hm,), 0.02, seed)
            sd[f"{q}.mlp.fc2.weight"] = xavier(f"{q}.mlp.fc2.weight", D, Hm)
            sd[f"{q}.mlp.fc2.bias"] = _normal(f"{q}.mlp.fc2.bias", (D,), 0.02, seed)
    for nm in ("enc_norm", "dec_norm"):
        sd[f"{nm}.weight"] = 1.0 + _normal(f"{nm}.weight", (D,), 0.1, seed)
        sd[f"{nm}.bias"] = _normal(f"{nm}.bias", (D,), 0.05, seed)
    sd["quant_conv.weight"] = xavier("quant_conv.weight", 2 * L, D)
    sd["quant_conv.bias"] = _normal("quant_conv.bias", (2 * L,), 0.02, seed)
    sd["post_quant_conv.weight"] = xavier("post_quant_conv.weight", D, L)
    sd["post_quant_conv.bias"] = _normal("post_quant_conv.bias", (D,), 0.02, seed)
    sd["predictor.weight"] = xavier("predictor.weight", Pd, D)
    sd["predictor.bias"] = _normal("predictor.bias", (Pd,), 0.02, seed)
    return sd


def dummy_prompt(n_frames: int = 5, height: int = 360, width: int = 640) -> torch.Tensor:
    """The offline input fixture of BASELINE config 1: solid-colour frames ramping blue -> red,
    values in [0,1], shape [n,3,H,W] (what reference dummy_dataset.py:16-28 yields as "video")."""
    w = torch.linspace(0, 1, n_frames).view(n_frames, 1)
    blue = torch.tensor([0.0, 0.0, 1.0]).view(1, 3)
    red = torch.tensor([1.0, 0.0, 0.0]).view(1, 3)
    col = (1 - w) * blue + w * red
    return col.view(n_frames, 3, 1, 1).expand(n_frames, 3, height, width).contiguous()


def w_key_actions(batch: int, total_frames: int, dim: int = 25) -> torch.Tensor:
    """Constant "W" (index 3) one-hot action rows (reference generate.py:158-159, 172-181)."""
    a = torch.zeros(batch, total_frames, dim)
    a[:, :, 3] = 1.0
    return a
