"""Drop-in for the reference's model/vae.py (ViT VAE, reference model/vae.py:160-384): same factory,
state_dict keys, `.encode(x).mean`, `.decode(z)`, `.patch_size`; arithmetic in libgtav_b200.so.

Observable differences: CUDA-only, inference-only, bf16 results as under autocast; the
posterior exposes mean / logvar (the reference also materialises std/var, unused on this path).
"""
from __future__ import annotations

import ctypes as C
import functools
import math

import torch
from torch import nn

try:
    from .. import _native as N
except ImportError:
    import _native as N


class _Holder(nn.Module):
    pass


def _linear(i, o):
    return nn.utils.skip_init(nn.Linear, i, o)


class DiagonalGaussianDistribution:
    """The posterior of reference model/vae.py:19-45 for dim=2 moments [N, seq, 2*latent] (already split): `mean` is what
    generate.py:56 consumes; `logvar` is clamped to [-30, 20], `std` / `var` derived from it, `sample()` draws
    mean + std * randn, `mode()` is the mean; `deterministic` zeroes std / var."""

    def __init__(self, mean: torch.Tensor, logvar: torch.Tensor | None = None, deterministic: bool = False):
        self.mean = mean
        self.deterministic = deterministic or logvar is None
        self.logvar = torch.zeros_like(mean) if logvar is None else torch.clamp(logvar, -30.0, 20.0)
        if self.deterministic:
            self.std = self.var = torch.zeros_like(mean)
        else:
            self.std = torch.exp(0.5 * self.logvar)
            self.var = torch.exp(self.logvar)
        self.parameters = torch.cat([self.mean, self.logvar], dim=-1)
        self.dims = [1, 2]

    def mode(self):
        return self.mean

    def sample(self):
        return self.mean + self.std * torch.randn(self.mean.shape, device=self.mean.device, dtype=self.mean.dtype)


class AutoencoderKL(nn.Module):
    def __init__(self, latent_dim, input_height=270, input_width=480, patch_size=24, enc_dim=768, enc_depth=6,
                 enc_heads=12, dec_dim=768, dec_depth=6, dec_heads=12, mlp_ratio=4.0,
                 norm_layer=functools.partial(nn.LayerNorm, eps=1e-6), use_variational=True, **kwargs):
        super().__init__()
        if enc_dim != dec_dim or enc_heads != dec_heads or not use_variational:
            raise RuntimeError("gtav_b200 VAE kernels cover the symmetric-width variational ViT (vit-l-20-shallow-encoder)")
        self.input_height, self.input_width = input_height, input_width
        self.patch_size = patch_size
        self.seq_h, self.seq_w = input_height // patch_size, input_width // patch_size
        self.seq_len = self.seq_h * self.seq_w
        self.patch_dim = 3 * patch_size ** 2
        self.latent_dim, self.enc_dim, self.dec_dim = latent_dim, enc_dim, dec_dim
        self.heads, self.enc_depth, self.dec_depth = enc_heads, enc_depth, dec_depth
        self.use_variational = use_variational
        D, Hm = enc_dim, int(enc_dim * mlp_ratio)

        self.patch_embed = _Holder()
        self.patch_embed.proj = nn.utils.skip_init(nn.Conv2d, 3, D, kernel_size=patch_size, stride=patch_size)

        def block():
            b = _Holder()
            b.norm1, b.norm2 = norm_layer(D), norm_layer(D)
            b.attn = _Holder()
            b.attn.qkv, b.attn.proj = _linear(D, 3 * D), _linear(D, D)
            b.mlp = _Holder()
            b.mlp.fc1, b.mlp.fc2 = _linear(D, Hm), _linear(Hm, D)
            return b

        self.encoder = nn.ModuleList([block() for _ in range(enc_depth)])
        self.enc_norm = norm_layer(D)
        self.quant_conv = _linear(D, 2 * latent_dim)
        self.post_quant_conv = _linear(latent_dim, D)
        self.decoder = nn.ModuleList([block() for _ in range(dec_depth)])
        self.dec_norm = norm_layer(D)
        self.predictor = _linear(D, self.patch_dim)
        self.initialize_weights()
        self._engine = None
        self._packed_sig = None
        self._plans = {}
        self._dirty = False
        self._pack_generation = 0
        self.register_load_state_dict_post_hook(lambda module, _keys: module._mark_dirty())

    @torch.no_grad()
    def initialize_weights(self):
        """Xavier-uniform linears, zero biases, unit LayerNorms (model/vae.py:239-260)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                m.bias.zero_()
            elif isinstance(m, nn.LayerNorm):
                m.bias.zero_()
                m.weight.fill_(1.0)
        w = self.patch_embed.proj.weight
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        self.patch_embed.proj.bias.zero_()

    # ------------------------------------------------------------------ packing
    def _signature(self):
        return N.param_signature(self)

    def _mark_dirty(self, *_):
        self._dirty = True

    def _apply(self, fn, *a, **k):          # .to() / .cuda() / .half(): new storage -> repack on the next call
        self._dirty = True
        return super()._apply(fn, *a, **k)

    def repack(self):
        """Force a re-pack of the bf16 weights on the next call (after editing parameters in place in a way the version
        counters cannot show, e.g. on inference tensors)."""
        self._dirty = True

    def _release(self):
        lib = N.load()
        for plan, _ws in self._plans.values():
            lib.gtav_vae_plan_destroy(plan)
        self._plans = {}
        if self._engine is not None:
            lib.gtav_vae_destroy(self._engine[0])
            self._engine = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    @torch.no_grad()
    def _pack(self):
        sig = self._signature()
        if self._engine is not None and sig == self._packed_sig and not self._dirty:
            return
        self._release()
        self._dirty = False
        self._pack_generation += 1
        lib = N.load()
        dev = self.predictor.weight.device
        if dev.type != "cuda":
            raise RuntimeError("gtav_b200.AutoencoderKL: parameters must live on a CUDA device; no CPU fallback")
        bf = lambda p: p.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        f32 = lambda p: p.detach().to(device=dev, dtype=torch.float32).contiguous()
        keep = []

        def k(t):
            keep.append(t)
            return t.data_ptr()

        def pack_blocks(blocks):
            arr = (N.VaeBlock * max(len(blocks), 1))()
            for i, b in enumerate(blocks):
                a = arr[i]
                a.norm1_w, a.norm1_b = k(f32(b.norm1.weight)), k(f32(b.norm1.bias))
                a.norm2_w, a.norm2_b = k(f32(b.norm2.weight)), k(f32(b.norm2.bias))
                a.qkv_w, a.qkv_b = k(bf(b.attn.qkv.weight)), k(bf(b.attn.qkv.bias))
                a.proj_w, a.proj_b = k(bf(b.attn.proj.weight)), k(bf(b.attn.proj.bias))
                a.fc1_w, a.fc1_b = k(bf(b.mlp.fc1.weight)), k(bf(b.mlp.fc1.bias))
                a.fc2_w, a.fc2_b = k(bf(b.mlp.fc2.weight)), k(bf(b.mlp.fc2.bias))
            keep.append(arr)
            return C.cast(arr, C.POINTER(N.VaeBlock))

        D = self.enc_dim
        w = N.VaeWeights()
        w.patch_w, w.patch_b = k(bf(self.patch_embed.proj.weight.reshape(D, -1))), k(bf(self.patch_embed.proj.bias))
        w.enc_norm_w, w.enc_norm_b = k(f32(self.enc_norm.weight)), k(f32(self.enc_norm.bias))
        w.dec_norm_w, w.dec_norm_b = k(f32(self.dec_norm.weight)), k(f32(self.dec_norm.bias))
        w.quant_w, w.quant_b = k(bf(self.quant_conv.weight)), k(bf(self.quant_conv.bias))
        post = torch.zeros((D, 64), dtype=torch.bfloat16, device=dev)      # K padded 16 -> 64 for one TMA k-block
        post[:, : self.latent_dim] = bf(self.post_quant_conv.weight)
        w.post_w, w.post_b = k(post), k(bf(self.post_quant_conv.bias))
        w.pred_w, w.pred_b = k(bf(self.predictor.weight)), k(bf(self.predictor.bias))
        # rotary table exactly as model/vae.py:71-76 builds its buffer: base linspace(1, seq/2, 8)*pi,
        # positions linspace(-1,1,n); 16 pairs = 8 row + 8 column; head dims 32..63 are not rotated
        dim = (D // self.heads) // 4
        base = (torch.linspace(1.0, self.seq_len / 2, dim // 2) * math.pi).to(dev)
        ah = torch.linspace(-1, 1, steps=self.seq_h, device=dev)[:, None] * base[None]
        aw = torch.linspace(-1, 1, steps=self.seq_w, device=dev)[:, None] * base[None]
        ang = torch.cat([ah[:, None].expand(self.seq_h, self.seq_w, -1), aw[None].expand(self.seq_h, self.seq_w, -1)],
                        dim=-1).reshape(self.seq_len, -1)
        assert ang.shape[1] == 16
        w.rot = k(torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous())
        w.enc, w.dec = pack_blocks(self.encoder), pack_blocks(self.decoder)
        cfg = N.VaeConfig(D, self.heads, self.enc_depth, self.dec_depth, self.latent_dim, self.patch_size, self.seq_h,
                          self.seq_w)
        handle = N.vp()
        N.check(lib.gtav_vae_create(C.byref(cfg), C.byref(w), C.byref(handle)), "gtav_vae_create")
        self._engine = (handle, keep)
        self._packed_sig = sig

    def _plan(self, n_frames):
        if n_frames not in self._plans:
            lib = N.load()
            handle = self._engine[0]
            nbytes = lib.gtav_vae_workspace_bytes(handle, n_frames)
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.predictor.weight.device)
            base = (ws.data_ptr() + 1023) & ~1023
            plan = N.vp()
            N.check(lib.gtav_vae_plan_create(handle, n_frames, base, nbytes, C.byref(plan)), "gtav_vae_plan_create")
            if len(self._plans) >= 4:       # plans own sizeable workspaces; keep a few shapes
                old = next(iter(self._plans))
                lib.gtav_vae_plan_destroy(self._plans.pop(old)[0])
            self._plans[n_frames] = (plan, ws)
        return self._plans[n_frames][0]

    # ------------------------------------------------------------------ API
    @torch.no_grad()
    def encode_mean(self, x, scale: float = 1.0, round_bf16: bool = False) -> torch.Tensor:
        """[N,3,H,W] in [-1,1] -> fp32 [N, seq_len, latent] holding bf16(mean) * scale (re-rounded to bf16
        when round_bf16, which is what `vae.encode(x).mean * s` yields under autocast, generate.py:56)."""
        N.require_cuda(x, "x")
        if x.shape[1:] != (3, self.input_height, self.input_width):
            raise AssertionError(f"Input image size {tuple(x.shape[1:])} doesn't match model "
                                 f"(3, {self.input_height}, {self.input_width}).")
        self._pack()
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        n = x.shape[0]
        out = torch.empty((n, self.seq_len, self.latent_dim), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            N.check(N.load().gtav_vae_encode(self._plan(n), x.data_ptr(), int(x.dtype == torch.bfloat16), out.data_ptr(),
                                             float(scale), int(round_bf16), N.current_stream()), "gtav_vae_encode")
        return out

    @torch.no_grad()
    def encode(self, x):
        """`vae.encode(x)` (model/vae.py:306-322): the full posterior, both halves of quant_conv's bf16 output."""
        N.require_cuda(x, "x")
        if x.shape[1:] != (3, self.input_height, self.input_width):
            raise AssertionError(f"Input image size {tuple(x.shape[1:])} doesn't match model "
                                 f"(3, {self.input_height}, {self.input_width}).")
        self._pack()
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        n = x.shape[0]
        mean = torch.empty((n, self.seq_len, self.latent_dim), dtype=torch.float32, device=x.device)
        logvar = torch.empty_like(mean)
        with torch.cuda.device(x.device):
            N.check(N.load().gtav_vae_encode_moments(self._plan(n), x.data_ptr(), int(x.dtype == torch.bfloat16), mean.data_ptr(),
                                                     logvar.data_ptr(), N.current_stream()), "gtav_vae_encode_moments")
        return DiagonalGaussianDistribution(mean.to(torch.bfloat16), logvar.to(torch.bfloat16))

    @torch.no_grad()
    def decode(self, z, divisor: float = 1.0, to_uint8: bool = False):
        """z [N, seq_len, latent] -> bf16 [N,3,H,W]; with to_uint8 the pixel epilogue of generate.py:241-244
        is fused and the result is uint8 [N,H,W,3]."""
        N.require_cuda(z, "z")
        if z.shape[1:] != (self.seq_len, self.latent_dim):
            raise AssertionError(f"latent shape {tuple(z.shape[1:])} != ({self.seq_len}, {self.latent_dim})")
        self._pack()
        z = z.to(torch.float32).contiguous()
        n = z.shape[0]
        if to_uint8:
            out = torch.empty((n, self.input_height, self.input_width, 3), dtype=torch.uint8, device=z.device)
        else:
            out = torch.empty((n, 3, self.input_height, self.input_width), dtype=torch.bfloat16, device=z.device)
        with torch.cuda.device(z.device):
            N.check(N.load().gtav_vae_decode(self._plan(n), z.data_ptr(), float(divisor), out.data_ptr(), int(to_uint8),
                                             N.current_stream()), "gtav_vae_decode")
        return out

    def autoencode(self, input, sample_posterior=True):
        posterior = self.encode(input)
        z = posterior.sample() if sample_posterior else posterior.mode()
        return self.decode(z), posterior, z

    def forward(self, inputs, labels=None, split="train"):
        return self.autoencode(inputs)

    def get_last_layer(self):
        return self.predictor.weight


def ViT_L_20_Shallow_Encoder(**kwargs):
    latent_dim = kwargs.pop("latent_dim", 16)
    return AutoencoderKL(latent_dim=latent_dim, patch_size=20, enc_dim=1024, enc_depth=6, enc_heads=16, dec_dim=1024,
                         dec_depth=12, dec_heads=16, input_height=360, input_width=640, **kwargs)


VAE_models = {"vit-l-20-shallow-encoder": ViT_L_20_Shallow_Encoder}
