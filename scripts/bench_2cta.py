"""Single-CTA vs CTA-pair (cta_group::2) tiled GEMM vs cuBLAS at large row counts (CUDA events, L2 flushed)."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtav_b200._native as N

lib = N.load()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=12, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for M, Nn, K, epi in ((1152, 3072, 1024, 0), (1152, 4096, 1024, 2), (1152, 1024, 4096, 1), (5760, 3072, 1024, 0), (5760, 4096, 1024, 2),
                      (5760, 1024, 4096, 1), (9216, 3072, 1024, 0), (9216, 4096, 1024, 2), (9216, 1024, 4096, 1), (9216, 1024, 1024, 1),
                      (18432, 4096, 1024, 3), (8192, 8192, 8192, 0)):
    A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    W = (torch.randn((Nn, K), device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    out = torch.empty((M, Nn), device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(Nn, device="cuda", dtype=torch.bfloat16)
    s = N.current_stream()

    def ours():
        N.check(lib.gtav_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), Nn, M, Nn, K, epi, bias.data_ptr(),
                                   None, 0, None, 0, None, 1, 0, s), "gemm")
    res = {}
    for mode in ("0", "1"):
        os.environ["GTAV_GEMM_2CTA"] = mode
        res[mode] = timeit(ours)
    tc = timeit(lambda: torch.matmul(A, W.t()))
    fl = 2.0 * M * Nn * K / 1e9
    print(json.dumps(dict(M=M, N=Nn, K=K, epi=epi, one_cta_tflops=round(fl / res["0"], 1), cta_pair_tflops=round(fl / res["1"], 1),
                          cublas_tflops=round(fl / tc, 1), one_cta_us=round(res["0"] * 1e3, 1), cta_pair_us=round(res["1"] * 1e3, 1))), flush=True)
