"""Tile-width sweep of the tiled GEMM at small row counts (M = 720 dense B=1 window, 1152 B=8 last-frame step, 576 context
pass): CUDA events, L2 flushed; bn = 0 is the library's own choice.   python scripts/sweep_gemm_tiles.py [M ...]"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtav_b200._native as N  # noqa: E402

lib = N.load()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warm=3):
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
    return ts[len(ts) // 2] * 1e3


Ms = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1152, 720]
for M in ([] if "--graph" in sys.argv else Ms):
    for name, Nn, K, epi in (("qkv", 3072, 1024, 0), ("out", 1024, 1024, 1), ("fc1", 4096, 1024, 2), ("fc2", 1024, 4096, 1)):
        A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
        W = (torch.randn((Nn, K), device="cuda") / math.sqrt(K)).to(torch.bfloat16)
        out = torch.empty((M, Nn), device="cuda", dtype=torch.bfloat16)
        bias = torch.zeros(Nn, device="cuda", dtype=torch.bfloat16)
        row = dict(M=M, gemm=name)
        for bn in (0, 64, 128, 256):
            def f():
                N.check(lib.gtav_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), Nn, M, Nn, K, epi, bias.data_ptr(),
                                           None, 0, None, 0, None, 1, bn, N.current_stream()), "gemm")
            row[f"bn{bn}_us"] = round(timeit(f), 1)
        row["cublas_us"] = round(timeit(lambda: torch.matmul(A, W.t())), 1)
        print(json.dumps(row), flush=True)


def graph_sweep(Ms):
    """The same sweep timed inside CUDA graphs (32 launches per graph over 32 different weight matrices, PDL between them):
    no event granularity, no launch cost - what a GEMM costs inside a step.   python scripts/sweep_gemm_tiles.py --graph [M ...]"""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from bench_graph import NW, b, graph_time, rnd
    for M in Ms:
        A1, A4 = rnd(M, 1024), rnd(M, 4096)
        for name, A, Ws, Nn, K, epi in (("qkv", A1, b.w_qkv, 3072, 1024, 0), ("out", A1, b.w_out, 1024, 1024, 1),
                                        ("fc1", A1, b.w_fc1, 4096, 1024, 2), ("fc2", A4, b.w_fc2, 1024, 4096, 1)):
            out = torch.empty((M, Nn), device="cuda", dtype=torch.bfloat16)
            row = dict(M=M, gemm=name)
            for bn in (0, 64, 128, 256):
                def mk(i, bn=bn):
                    W = Ws[i % NW]

                    def f():
                        N.check(lib.gtav_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), Nn, M, Nn, K, epi, b.bias.data_ptr(),
                                                   None, 0, None, 0, None, 1, bn, N.current_stream()), "gemm")
                    return f
                row[f"bn{bn}_us"] = round(graph_time([mk(i) for i in range(32)]), 2)
            row["tflops_best"] = round(2.0 * M * Nn * K / min(row[f"bn{x}_us"] for x in (0, 64, 128, 256)) / 1e6, 1)
            print(json.dumps(row), flush=True)


if "--graph" in sys.argv:
    graph_sweep(Ms)
