"""Phase time stamps of the tcgen05 attention kernel (GTAV_ATTN_TRACE): per role (first thread of softmax group A / B, MMA thread, loader
thread 0) of CTA 0, ns since that CTA's entry.   python scripts/trace_attn.py [seq] [groups]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtav_b200._native as N  # noqa: E402

lib = N.load()
seq = int(sys.argv[1]) if len(sys.argv) > 1 else 576
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pairs = 16 if seq == 576 else 32
H, d = 16, 64
os.environ["GTAV_ATTN"] = "tc"
qkv = torch.randn((groups * seq, 3 * H * d), device="cuda").to(torch.bfloat16)
out = torch.empty((groups * seq, H * d), dtype=torch.bfloat16, device="cuda")
ang = torch.rand((seq, pairs), device="cuda") * 20 - 10
rot = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()
s = N.current_stream()
for _ in range(3):
    N.check(lib.gtav_attention_seq(qkv.data_ptr(), out.data_ptr(), groups, seq, H, rot.data_ptr(), pairs, s), "attn")
torch.cuda.synchronize()
trace = torch.zeros((groups * H, 4, 64), dtype=torch.int64, device="cuda")
os.environ["GTAV_ATTN_TRACE"] = str(trace.data_ptr() + (8 if len(sys.argv) > 3 else 0))
N.check(lib.gtav_attention_seq(qkv.data_ptr(), out.data_ptr(), groups, seq, H, rot.data_ptr(), pairs, s), "attn")
torch.cuda.synchronize()
del os.environ["GTAV_ATTN_TRACE"]
tr = trace.cpu()
t0 = int(tr[:, 0, 0].min())
for cta in (0,):
    for role, name in enumerate(("softmaxA", "mma", "loader", "softmaxB")):
        v = [int(x) - t0 for x in tr[cta, role].tolist() if x > 0]
        print(f"cta {cta} {name:8s}", " ".join(f"{x}" for x in v))
