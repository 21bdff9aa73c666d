"""Drop-in for the one function generate.py imports from the reference's train_dit.py:
`denoise_step` (reference train_dit.py:30-125).  The training side (DiffusionTrainer) is out of scope.

Same keyword signature and return values; the DiT call goes through whatever `dit_model` is passed
(normally gtav_b200's DiT) and the v-prediction DDIM algebra runs in the fused CUDA kernel
(`gtav_ddim_update`).  For the graph-captured whole-rollout path see sampler.py.
"""
from __future__ import annotations

import torch

try:
    from . import _native as N
except ImportError:
    import _native as N


@torch.inference_mode()
def denoise_step(dit_model, x_noisy, actions, noise_idx, stabilization_level, noise_range, alphas_cumprod,
                 start_frame=0, dtype=torch.bfloat16):
    """One DDIM step on the sliding window x_noisy[:, start_frame:].

    Context frames sit at t = stabilization_level, the last frame at noise_range[noise_idx] (truncated to
    an integer, as torch.full(dtype=long) does in the reference); returns (x_pred, v_pred) for the window,
    x_pred = x0 when noise_idx <= 0.
    """
    N.require_cuda(x_noisy, "x_noisy")
    if dtype != torch.bfloat16:
        raise RuntimeError("gtav_b200 computes in bf16 only")
    B, F_all = x_noisy.shape[:2]
    dev = x_noisy.device
    t_cur = int(noise_range[noise_idx])
    t_nxt = int(noise_range[max(0, noise_idx - 1)])
    t = torch.full((B, F_all), int(stabilization_level), dtype=torch.long, device=dev)
    t_next = t.clone()
    t[:, -1] = t_cur
    t_next[:, -1] = t_nxt
    x_curr = x_noisy[:, start_frame:].to(torch.float32).contiguous()
    t, t_next = t[:, start_frame:].contiguous(), t_next[:, start_frame:].contiguous()
    T = x_curr.shape[1]
    if actions is not None:
        actions = actions[:, start_frame:start_frame + T]

    v_pred = dit_model(x_curr, t, actions)

    abar = alphas_cumprod.to(device=dev, dtype=torch.float32).reshape(-1)
    a_t = abar[t.reshape(-1)].contiguous()
    a_n = abar[t_next].clone()
    a_n[:, :-1] = 1.0
    a_n = a_n.reshape(-1).contiguous()
    final = torch.full((1,), int(noise_idx <= 0), dtype=torch.int32, device=dev)
    v = v_pred.to(torch.bfloat16).contiguous()
    x_pred = torch.empty_like(x_curr)
    n = x_curr[0, 0].numel()
    with torch.cuda.device(dev):
        N.check(N.load().gtav_ddim_update(x_curr.data_ptr(), v.data_ptr(), x_pred.data_ptr(), B * T, n, a_t.data_ptr(),
                                          a_n.data_ptr(), final.data_ptr(), N.current_stream()), "gtav_ddim_update")
    return x_pred, v_pred


# ------------------------------------------------------------------------------------------------------
# Inference side of the reference's DiffusionTrainer (train_dit.py:163-552): encode_frames / decode_frames /
# predict / predict_noise call the same denoise_step + VAE API as generate.py, under @torch.inference_mode, so they
# run on the same kernels.  Training (backward, AdamW, DDP, checkpoints, wandb: train_dit.py:553-1094) is out of scope
# and raises.
# ------------------------------------------------------------------------------------------------------
import dataclasses
from typing import Optional

try:
    from .sampler import SCALING_FACTOR, Sampler
    from .utils import sigmoid_beta_schedule
except ImportError:
    from sampler import SCALING_FACTOR, Sampler
    from utils import sigmoid_beta_schedule


@dataclasses.dataclass
class TrainingConfig:
    """The fields of the reference's TrainingConfig (train_dit.py:128-157) that the inference methods read, with the
    reference's defaults; the optimiser / logging / dataset fields are accepted and ignored."""
    ddim_noise_steps: int = 16
    ddim_noise_steps_inference: int = 16
    ctx_max_noise_idx: int = 3
    noise_abs_max: float = 20.0
    n_prompt_frames: int = 1
    use_action_conditioning: bool = True
    model_name: str = "dit"
    seed: int = 42

    @classmethod
    def from_dict(cls, d: dict) -> "TrainingConfig":
        names = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in names})

    @classmethod
    def from_yaml(cls, yaml_path: str) -> "TrainingConfig":
        import yaml
        with open(yaml_path, "r") as f:
            return cls.from_dict(yaml.safe_load(f) or {})


class DiffusionTrainer:
    """Inference-only stand-in: `DiffusionTrainer(config, dit, vae)` instead of the reference's
    `DiffusionTrainer(config)` (which builds the models, Accelerator and data loaders itself, train_dit.py:164-266)."""

    def __init__(self, config: TrainingConfig, dit, vae, device: Optional[torch.device] = None, dtype=torch.bfloat16):
        if dtype != torch.bfloat16:
            raise RuntimeError("gtav_b200 computes in bf16 only")
        self.config, self.dit, self.vae, self.dtype = config, dit, vae, dtype
        self.device = device if device is not None else next(dit.parameters()).device
        self.register_buffers()
        self._samplers = {}

    def register_buffers(self):
        """train_dit.py:288-327: sigmoid-beta schedule (clamp_min 1e-6), cumulative alphas as [T,1,1,1], integer DDIM
        level tables for training and inference, stabilization_level = noise_range[1]."""
        self.max_noise_level = 1000
        self.ctx_max_noise_idx = self.config.ctx_max_noise_idx
        self.betas = sigmoid_beta_schedule(self.max_noise_level, clamp_min=0.000001).to(device=self.device, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0).view(-1, 1, 1, 1)
        self.betas_inference = self.betas.clone()
        self.alphas_inference = 1.0 - self.betas_inference
        self.alphas_cumprod_inference = torch.cumprod(self.alphas_inference, dim=0).view(-1, 1, 1, 1)
        self.noise_range = torch.linspace(0, self.max_noise_level - 1, self.config.ddim_noise_steps + 1).long().to(self.device)
        self.noise_range_inference = torch.linspace(0, self.max_noise_level - 1,
                                                    self.config.ddim_noise_steps_inference + 1).long().to(self.device)
        self.stabilization_level = self.noise_range[1]

    def _sampler(self) -> Sampler:
        key = (self.config.ddim_noise_steps_inference, int(self.stabilization_level), float(self.config.noise_abs_max))
        if key not in self._samplers:
            self._samplers[key] = Sampler(self.dit, self.vae, noise_steps=key[0], stabilization_level=key[1], noise_abs_max=key[2],
                                          max_noise_level=self.max_noise_level, alphas_cumprod=self.alphas_cumprod_inference)
        return self._samplers[key]

    @torch.inference_mode()
    def encode_frames(self, frames, dtype=torch.bfloat16):
        """frames [b, t, 3, H, W] in [0, 1] -> latents [b, t, C, H/p, W/p] (train_dit.py:329-352)."""
        return self._sampler().encode_prompt(frames)

    @torch.inference_mode()
    def decode_frames(self, frames, num_frames, dtype=torch.bfloat16):
        """latents [b, t, C, h, w] -> uint8 pixels [b, t, H, W, 3] (train_dit.py:354-371)."""
        if frames.shape[1] != num_frames:
            raise RuntimeError(f"decode_frames: {frames.shape[1]} latent frames, num_frames={num_frames}")
        return self._sampler().decode_frames(frames)

    def _prompt_actions(self, prompt, num_frames):
        """train_dit.py:381-398: first batch entry only; missing actions are padded with the "W" key (index 3)."""
        if not self.config.use_action_conditioning:
            return None
        actions = prompt["actions"][:1].to(self.device)
        if num_frames is not None and actions.shape[1] < num_frames:
            pad = torch.zeros((actions.shape[0], num_frames - actions.shape[1], actions.shape[2]), device=actions.device)
            pad[:, :, 3] = 1
            actions = torch.cat([actions, pad], dim=1)
        return actions

    @torch.inference_mode()
    def predict(self, test_loader, epoch=0, global_step=0, num_frames=32, generator=None, video_path=None, stepwise=False,
                noise=None):
        """Generate `num_frames` frames from the first n_prompt_frames of the loader's first batch (train_dit.py:373-469).
        Returns (pixels uint8 [1, num_frames, H, W, 3], latents); writes an mp4 only when video_path is given (the
        reference always writes debug_visualizations/test_*.mp4).  stepwise=True runs the reference's literal
        per-step loop through denoise_step instead of the graph-captured Sampler (same results).  noise: optional
        [1, num_frames - n_prompt, C, h, w] N(0,1) draws used instead of torch.randn (the reference draws them from the
        global RNG, train_dit.py:419-421)."""
        self.dit.eval()
        prompt = next(iter(test_loader))
        frames = prompt["video"][:1, : self.config.n_prompt_frames].to(self.device)
        actions = self._prompt_actions(prompt, num_frames)
        smp = self._sampler()
        x = self.encode_frames(frames, dtype=self.dtype)
        if not stepwise:
            x = smp.sample_latents(x, actions, num_frames, generator=generator, noise=noise)
        else:
            n_prompt = x.shape[1]
            x = x.float()
            for i in range(n_prompt, num_frames):
                if noise is not None:
                    new_frame = noise[:, i - n_prompt: i - n_prompt + 1].to(device=self.device, dtype=torch.float32)
                else:
                    new_frame = torch.randn((x.shape[0], 1, *x.shape[2:]), device=self.device, generator=generator)
                new_frame = torch.clamp(new_frame, -self.config.noise_abs_max, self.config.noise_abs_max)
                x = torch.cat([x, new_frame], dim=1)
                start_frame = max(0, i + 1 - self.dit.max_frames)
                for noise_idx in reversed(range(0, self.config.ddim_noise_steps_inference + 1)):
                    x_pred, _ = denoise_step(dit_model=self.dit, x_noisy=x, actions=actions, noise_idx=noise_idx,
                                             stabilization_level=self.stabilization_level, noise_range=self.noise_range_inference,
                                             alphas_cumprod=self.alphas_cumprod_inference, start_frame=start_frame, dtype=self.dtype)
                    x[:, -1:] = x_pred[:, -1:]
        pixels = self.decode_frames(x, num_frames, dtype=self.dtype)
        if video_path is not None:
            try:
                from .generate import write_video
            except ImportError:
                from generate import write_video
            write_video(video_path, pixels[0].cpu(), fps=10)
        return pixels, x

    @torch.inference_mode()
    def predict_noise(self, test_loader, epoch=0, global_step=0, generator=None, noise=None):
        """Noise the context frames of a clip to stabilization_level - 1, replace the last frame by clamped noise and
        denoise it (train_dit.py:471-552, without the matplotlib panel).  Returns (x_noisy after denoising, clean
        latents, v_pred of the final step)."""
        self.dit.eval()
        prompt = next(iter(test_loader))
        frames = prompt["video"][:1].to(self.device)
        num_frames = frames.shape[1]
        actions = self._prompt_actions(prompt, None)
        latents = self.encode_frames(frames).float()
        B = latents.shape[0]
        x_noisy = latents.clone()
        if noise is not None:
            noise = noise.to(device=self.device, dtype=torch.float32)
            ctx_noise = noise[:, :-1]
        else:
            ctx_noise = torch.randn(x_noisy[:, :-1].shape, device=self.device, generator=generator)
        ctx_noise = torch.clamp(ctx_noise, -self.config.noise_abs_max, self.config.noise_abs_max)
        t_ctx = torch.full((B, num_frames - 1), int(self.stabilization_level) - 1, dtype=torch.long, device=self.device)
        alpha_ctx = self.alphas_cumprod[t_ctx]
        x_noisy[:, :-1] = alpha_ctx.sqrt() * x_noisy[:, :-1] + (1 - alpha_ctx).sqrt() * ctx_noise
        new_frame = noise[:, -1:] if noise is not None else torch.randn((B, 1, *x_noisy.shape[2:]), device=self.device, generator=generator)
        x_noisy[:, -1:] = torch.clamp(new_frame, -self.config.noise_abs_max, self.config.noise_abs_max)
        start_frame = max(0, num_frames - self.dit.max_frames)
        v_pred = None
        for noise_idx in reversed(range(0, self.config.ddim_noise_steps_inference + 1)):
            x_pred, v_pred = denoise_step(dit_model=self.dit, x_noisy=x_noisy, actions=actions, noise_idx=noise_idx,
                                          stabilization_level=self.stabilization_level, noise_range=self.noise_range_inference,
                                          alphas_cumprod=self.alphas_cumprod_inference, start_frame=start_frame, dtype=self.dtype)
            x_noisy[:, -1:] = x_pred[:, -1:]
        return x_noisy, latents, v_pred

    def train(self, *a, **k):
        raise NotImplementedError("gtav_b200 covers the inference path only: training (train_dit.py:553-1094) is out of scope")

    validate = save_checkpoint = load_checkpoint = _shared_step = train
