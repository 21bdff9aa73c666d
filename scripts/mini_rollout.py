"""A short rollout (random-init weights) for profiling: python scripts/mini_rollout.py [B] [total_frames] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_models, synthetic_prompt  # noqa: E402
from gtav_b200.sampler import Sampler  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
total = int(sys.argv[2]) if len(sys.argv) > 2 else 7
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda")
dit, vae = build_models(dev)
s = Sampler(dit, vae, noise_steps=steps)
acts = torch.zeros(B, total, 25, device=dev)
acts[:, :, 3] = 1.0
frames, lat = s.generate(synthetic_prompt(B, 4).to(dev), acts, total, generator=torch.Generator(device=dev).manual_seed(0))
torch.cuda.synchronize()
print("ok", tuple(frames.shape), float(lat.abs().mean()))
