"""ms per DiT step of the shipped path (frame cache) and of the dense window at B rollouts: two generated frames, device-timed.
python scripts/bench_step_b8.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_models  # noqa: E402
from gtav_b200.sampler import Sampler  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda")
dit, vae = build_models(dev)
lat = torch.randn(B, 4, 16, 18, 32, device=dev)
acts = torch.zeros(B, 8, 25, device=dev)
acts[:, :, 3] = 1
for cache in (True, False):
    s = Sampler(dit, None, noise_steps=100, frame_cache=cache)
    g = torch.Generator(device=dev).manual_seed(0)
    s.sample_latents(lat, acts, 5, generator=g)
    s.sample_latents(lat, acts, 5, generator=g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.sample_latents(lat, acts, 6, generator=g)
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B} {'frame cache' if cache else 'dense window'}: {e0.elapsed_time(e1) / (2 * 101):.4f} ms per DiT step", flush=True)
    s.close()
