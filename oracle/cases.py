"""Shared definitions of the golden / parity cases (inputs are re-derived from seeds).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import torch

# t values mirror what the sampler feeds the model: context frames at 15, last frame on the
# truncated linspace grid {0, 9, 19, ...}.
CASES_DIT = {
    # depth 2 keeps the CPU suite fast; depth 16 is the real DiT-S/2 (607.9 M parameters).
    "d2_b2_t3": dict(depth=2, degenerate=False, B=2, T=3, t=[15, 15, 999, 15, 15, 509], actions=False, seed=11),
    "d2_b1_t5_act": dict(depth=2, degenerate=False, B=1, T=5, t=[15, 15, 15, 15, 749], actions=True, seed=12),
    "d2_b1_t1": dict(depth=2, degenerate=False, B=1, T=1, t=[0], actions=True, seed=13),
    "d16_b1_t5_act": dict(depth=16, degenerate=False, B=1, T=5, t=[15, 15, 15, 15, 989], actions=True, seed=14),
    "d16_degenerate": dict(depth=16, degenerate=True, B=1, T=2, t=[15, 99], actions=False, seed=15),
}

CASES_DENOISE = {
    "mid_step": dict(depth=2, B=1, frames=6, start_frame=1, noise_steps=10, noise_idx=4, actions=True, seed=21),
    "final_step": dict(depth=2, B=2, frames=3, start_frame=0, noise_steps=10, noise_idx=0, actions=False, seed=22),
    "first_step": dict(depth=2, B=1, frames=5, start_frame=0, noise_steps=100, noise_idx=100, actions=True, seed=23),
}

CASES_VAE = {
    "e1_d1": dict(enc_depth=1, dec_depth=1, N=2, seed=31),
    "e6_d12": dict(enc_depth=6, dec_depth=12, N=1, seed=32),
}

ROLLOUT = dict(depth=2, enc_depth=1, dec_depth=1, n_prompt=4, total_frames=7, noise_steps=4, actions=True, seed=41)


def seeded_randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def seeded_rand(shape, seed):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed))


def subsample_image(img: torch.Tensor, stride: int = 8) -> torch.Tensor:
    """[N,3,H,W] -> every `stride`-th pixel, fp32 (keeps the decode goldens small)."""
    return img[:, :, ::stride, ::stride].float()


# ---- full-size cases (oracle/make_golden_full.py): the BASELINE configs themselves, depth-16 DiT + VAE 6/12 ----------
# C1 = BASELINE config 1 exactly: B=1, 360x640, dummy blue->red prompt (4 frames), 8 frames, 10 DDIM steps (11 DiT
# evaluations per frame, generate.py:206), minted for the non-degenerate weights ("initB") and for zero adaLN linears
# ("initA": every block an identity, like the reference's default init).
C1 = dict(depth=16, enc_depth=6, dec_depth=12, n_prompt=4, total_frames=8, noise_steps=10, actions=False, seed=51)
# One denoise_step of BASELINE config 3's shape: B=8 action-conditioned rollouts, full 5-frame window (M = 5760 rows).
C3_STEP = dict(depth=16, B=8, frames=5, start_frame=0, noise_steps=100, noise_idx=57, actions=True, seed=52)
# DiffusionTrainer.predict / predict_noise of the reference (train_dit.py:373-552) on the small models of ROLLOUT.
TRAINER = dict(depth=2, enc_depth=1, dec_depth=1, ddim_noise_steps=16, ddim_noise_steps_inference=4, n_prompt_frames=2,
               num_frames=6, noise_abs_max=20.0, seed_predict=61, seed_predict_noise=62)
