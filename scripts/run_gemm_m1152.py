"""The three big GEMM shapes of a B = 8 last-frame step (M = 1152) a few times each, L2 flushed in between (for ncu)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtav_b200._native as N  # noqa: E402

lib = N.load()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1152
for Nn, K, epi in ((3072, 1024, 0), (4096, 1024, 2), (1024, 4096, 1)):
    A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    W = (torch.randn((Nn, K), device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    out = torch.empty((M, Nn), device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(Nn, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        flush.zero_()
        N.check(lib.gtav_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), Nn, M, Nn, K, epi, bias.data_ptr(),
                                   None, 0, None, 0, None, 1, 0, N.current_stream()), "gemm")
    torch.cuda.synchronize()
