"""Phase time stamps of the persistent last-frame step kernel (GTAV_MEGA_TRACE): median / max over CTAs of the time
spent between consecutive stamps of half-blocks 2 (spatial) and 3 (temporal).  Profiling aid, not a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtav_b200.model.dit import DiT_models

dev = torch.device("cuda")
torch.manual_seed(0)
dit = DiT_models["DiT-S/2"]().to(dev).eval()
with torch.no_grad():
    for b_ in dit.blocks:
        for h_ in ("s", "t"):
            lin = getattr(b_, f"{h_}_adaLN_modulation")[-1]
            lin.weight.normal_(std=0.02)
            lin.bias.normal_(std=0.02)
x = torch.randn((1, 5, 16, 18, 32), device=dev)
t = torch.tensor([[15, 15, 15, 15, 500]], device=dev)
dit.forward_last_frame(x, t)
torch.cuda.synchronize()
G = 128
trace = torch.zeros((G, 2, 32), dtype=torch.int64, device=dev)
os.environ["GTAV_MEGA_TRACE"] = str(trace.data_ptr())
for _ in range(3):
    dit.forward_last_frame(x, t)
torch.cuda.synchronize()
del os.environ["GTAV_MEGA_TRACE"]
tr = trace.cpu()
names = {0: "half start"}
for k, nm in enumerate(("qkv", "out", "fc1", "fc2")):
    for j, st in enumerate(("start", "A filled", "acc done", "partials stored", "rendezvous", "reduced")):
        names[1 + 7 * k + j] = f"{nm}: {st}"
names.update({7: "barrier after qkv", 29: "attention done", 30: "barrier after attention", 14: "barrier after out", 21: "barrier after fc1",
              28: "barrier after fc2"})
order = [0, 1, 2, 3, 4, 5, 6, 7, 29, 30, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28]
for hsel, label in ((0, "half 2 (spatial)"), (1, "half 3 (temporal)")):
    t0 = tr[:, hsel, 0].min()
    print(f"# {label}: ns since the first CTA entered the half; min / median / max over the CTAs that stamped")
    for sl in order:
        col = tr[:, hsel, sl]
        col = col[col > 0] - t0
        if col.numel() == 0:
            continue
        print(f"  {names[sl]:<26} n={col.numel():3d}  {int(col.min()):7d} {int(col.median()):7d} {int(col.max()):7d}")
