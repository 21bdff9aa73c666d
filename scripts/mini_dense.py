"""A few dense DiT forwards (the reference's schedule: whole 5-frame window) at B rollouts, for profiling:
python scripts/mini_dense.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_models  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda")
dit, _ = build_models(dev)
x = torch.randn(B, 5, 16, 18, 32, device=dev)
t = torch.tensor([[15, 15, 15, 15, 499]], device=dev).expand(B, 5).contiguous()
a = torch.zeros(B, 5, 25, device=dev)
a[:, :, 3] = 1
for _ in range(3):
    v = dit(x, t, a)
torch.cuda.synchronize()
print("ok", float(v.float().abs().mean()))
