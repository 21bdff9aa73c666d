"""Attention kernels in isolation: tcgen05 (attn_tc.cu) vs mma.sync (attn_mma.cu), CUDA-event timing with rotating
buffers larger than L2 between iterations.   python scripts/bench_attn.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtav_b200._native as N  # noqa: E402

lib = N.load()
H, d = 16, 64


def run(seq, pairs, groups, impl, reps=20):
    os.environ["GTAV_ATTN"] = impl
    nbuf = max(2, int(300e6 // (groups * seq * 3 * H * d * 2)) + 1)
    nbuf = min(nbuf, 64)
    qkvs = [torch.randn((groups * seq, 3 * H * d), device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    out = torch.empty((groups * seq, H * d), dtype=torch.bfloat16, device="cuda")
    ang = torch.rand((seq, pairs), device="cuda") * 20 - 10
    rot = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()
    s = N.current_stream()
    for i in range(3):
        N.check(lib.gtav_attention_seq(qkvs[i % nbuf].data_ptr(), out.data_ptr(), groups, seq, H, rot.data_ptr(), pairs, s), "attn")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        N.check(lib.gtav_attention_seq(qkvs[i % nbuf].data_ptr(), out.data_ptr(), groups, seq, H, rot.data_ptr(), pairs, s), "attn")
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    fl = 4.0 * seq * seq * d * H * groups
    return us, fl / us / 1e6


for seq, pairs, groups in ((576, 16, 1), (576, 16, 4), (576, 16, 32), (576, 16, 64), (144, 32, 1), (144, 32, 5), (144, 32, 40), (144, 32, 64), (144, 32, 96), (144, 32, 160), (144, 32, 256), (144, 32, 320)):
    row = []
    for impl in ("mma", "tc"):
        us, tf = run(seq, pairs, groups, impl)
        row.append(f"{impl}: {us:8.1f} us {tf:7.1f} TFLOP/s")
    print(f"S={seq} groups={groups:4d}  " + "   ".join(row), flush=True)
