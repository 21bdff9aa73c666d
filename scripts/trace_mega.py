"""Phase time stamps of the persistent last-frame step kernel (GTAV_MEGA_TRACE): median / max over CTAs of the time
spent between consecutive stamps of half-blocks 2 (spatial) and 3 (temporal).  Profiling aid, not a benchmark."""
import os
import sys

import torch

os.environ["GTAV_MEGA"] = "1"
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
names = ["half start", "qkv gemm done", "barrier", "attention done", "barrier", "out gemm done", "barrier", "row phase (h += gate*out, LN2) done",
         "barrier", "fc1 gemm done", "barrier", "fc2 gemm done", "barrier", "row phase (h += gate*fc2, LN1) done", "barrier"]
order = list(range(15))
for hsel, label in ((0, "half 2 (spatial)"), (1, "half 3 (temporal)")):
    t0 = tr[:, hsel, 0].min()
    print(f"# {label}: ns since the first CTA entered the half; min / median / max over the CTAs that stamped")
    for sl in order:
        col = tr[:, hsel, sl]
        col = col[col > 0] - t0
        if col.numel() == 0:
            continue
        print(f"  {names[sl]:<44} n={col.numel():3d}  {int(col.min()):7d} {int(col.median()):7d} {int(col.max()):7d}")
