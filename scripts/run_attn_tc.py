"""A few launches of the tcgen05 attention kernel (for ncu): python scripts/run_attn_tc.py [seq] [groups]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtav_b200._native as N  # noqa: E402

lib = N.load()
seq = int(sys.argv[1]) if len(sys.argv) > 1 else 576
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 32
pairs = 16 if seq == 576 else 32
H, d = 16, 64
os.environ["GTAV_ATTN"] = "tc"
qkv = torch.randn((groups * seq, 3 * H * d), device="cuda").to(torch.bfloat16)
out = torch.empty((groups * seq, H * d), dtype=torch.bfloat16, device="cuda")
ang = torch.rand((seq, pairs), device="cuda") * 20 - 10
rot = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()
for _ in range(4):
    N.check(lib.gtav_attention_seq(qkv.data_ptr(), out.data_ptr(), groups, seq, H, rot.data_ptr(), pairs, N.current_stream()), "attn")
torch.cuda.synchronize()
