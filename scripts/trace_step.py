"""Phase trace of the weight-streaming GEMMs INSIDE the real last-frame DiT step (16 blocks, B=1, T=5), replayed
from a CUDA graph with programmatic dependent launch as in the sampler: for each of a few consecutive launches the
globaltimer stamps of the kernel's phases (entry, set-up, W landed, A landed, accumulator done, partials written,
rendezvous passed, reduced) relative to the first CTA entry of the first launch shown - min / median / max over CTAs.

    python scripts/trace_step.py [first_launch [count]]      (GTAV_FUSE=0 for the unfused chain)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtav_b200._native as N
from gtav_b200.model.dit import DiT_models

lib = N.load()
dev = torch.device("cuda")
first = int(sys.argv[1]) if len(sys.argv) > 1 else 8
count = int(sys.argv[2]) if len(sys.argv) > 2 else 9
torch.manual_seed(0)
x = torch.randn((1, 5, 16, 18, 32), device=dev)
t = torch.tensor([[15, 15, 15, 15, 500]], device=dev)
rows = torch.arange(5, dtype=torch.int32, device=dev)
out = torch.empty((1, 1, 16, 18, 32), dtype=torch.bfloat16, device=dev)
dit = DiT_models["DiT-S/2"]().to(dev).eval()
with torch.no_grad():
    for b_ in dit.blocks:
        for h_ in ("s", "t"):
            lin = getattr(b_, f"{h_}_adaLN_modulation")[-1]
            lin.weight.normal_(std=0.02)
            lin.bias.normal_(std=0.02)
dit.forward_last_frame(x, t)
plan = dit._plan(1, 5)
trace = torch.zeros((128, 160, 8), dtype=torch.int64, device=dev)
os.environ["GTAV_ENGINE_TRACE"] = str(trace.data_ptr())


def step():
    N.check(lib.gtav_dit_last_frame(plan, x.data_ptr(), 0, rows[4:].data_ptr(), out.data_ptr(), N.current_stream()), "last_frame")


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    trace.zero_()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
tr = trace.cpu()
names = ["entry", "setup", "W_landed", "A_landed", "acc_done", "partials", "rendezv", "reduced"]
kinds = ["qkv", "out", "fc1", "fc2"]
t0 = None
print("# launch kind ctas | " + " ".join(f"{n:>22}" for n in names) + "   (ns since the first entry shown: min/median/max)")
for k in range(first, min(first + count, 128)):
    a = tr[k]
    used = a[:, 0] > 0
    if not used.any():
        continue
    a = a[used]
    if t0 is None:
        t0 = int(a[:, 0].min())
    cells = []
    for j in range(8):
        col = a[:, j]
        col = col[col > 0]
        if len(col) == 0:
            cells.append(f"{'-':>22}")
            continue
        col = col - t0
        cells.append(f"{int(col.min()):>6}/{int(col.median()):>6}/{int(col.max()):>6}".rjust(22))
    print(f"# {k:>4} {kinds[k % 4]:>4} {int(used.sum()):>4} | " + " ".join(cells))
