"""Micro-benchmarks of the stand-alone kernels at the hot-path shapes (CUDA events, L2 flushed
between timed launches).  Prints one JSON object per line; cuBLAS (torch.matmul) beside each GEMM."""
import json
import math
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import gtav_b200._native as N

lib = N.load()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20, warm=3):
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


def gemm_case(M, Nn, K, bn=0, epi=0):
    A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    W = (torch.randn((Nn, K), device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    out = torch.empty((M, Nn), device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(Nn, device="cuda", dtype=torch.bfloat16)
    s = N.current_stream()

    def ours():
        N.check(lib.gtav_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), Nn, M, Nn, K, epi, bias.data_ptr(),
                                   None, 0, None, 0, None, 1, bn, s), "gemm")
    t = timeit(ours)
    tc = timeit(lambda: torch.matmul(A, W.t()))
    fl = 2.0 * M * Nn * K
    print(json.dumps(dict(kernel="gemm", M=M, N=Nn, K=K, bn=bn, ms=round(t, 4), tflops=round(fl / t / 1e9, 1),
                          cublas_ms=round(tc, 4), cublas_tflops=round(fl / tc / 1e9, 1))), flush=True)


if __name__ == "__main__":
    for B in (1, 8):
        M = 720 * B
        for (Nn, K) in ((3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096)):
            for bn in ((128, 256) if Nn % 256 == 0 else (128,)):
                gemm_case(M, Nn, K, bn)
    gemm_case(5, 198656, 1024, 256, epi=1)
    gemm_case(46080, 4096, 1024, 256)
    gemm_case(8192, 8192, 8192, 256)
