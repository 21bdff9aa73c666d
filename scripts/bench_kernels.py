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


def skinny_case(Nn, K, epi=0):
    M = 144
    A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    W = (torch.randn((Nn, K), device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    out = torch.empty((M, Nn), device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(Nn, device="cuda", dtype=torch.bfloat16)
    ws = torch.empty(lib.gtav_gemm_skinny_workspace_bytes(M), dtype=torch.uint8, device="cuda")
    counters = torch.zeros(512, dtype=torch.int32, device="cuda")
    s = N.current_stream()
    for splits in (0, 1, 2, 4, 8, 16):
        def ours():
            N.check(lib.gtav_gemm_skinny_bf16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), Nn, M, Nn, K, epi, bias.data_ptr(),
                                              None, 0, None, 0, None, 144, splits, ws.data_ptr(), counters.data_ptr(), s), "skinny")
        try:
            t = timeit(ours)
        except RuntimeError:
            continue
        by = 2.0 * (Nn * K + M * K + M * Nn)
        print(json.dumps(dict(kernel="gemm_skinny", M=M, N=Nn, K=K, splits=splits, us=round(t * 1e3, 2),
                              gbs=round(by / t / 1e6, 1))), flush=True)

    def tiled():
        N.check(lib.gtav_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), Nn, M, Nn, K, epi, bias.data_ptr(),
                                   None, 0, None, 0, None, 1, 0, s), "gemm")
    t = timeit(tiled)
    print(json.dumps(dict(kernel="gemm_tiled", M=M, N=Nn, K=K, us=round(t * 1e3, 2))), flush=True)
    tc = timeit(lambda: torch.matmul(A, W.t()))
    print(json.dumps(dict(kernel="cublas", M=M, N=Nn, K=K, us=round(tc * 1e3, 2))), flush=True)


def light_cases():
    """The non-GEMM kernels of a last-frame step at B=1 (144 rows)."""
    D, P, H = 1024, 144, 16
    s = N.current_stream()
    x = torch.randn((P, D), device="cuda").to(torch.bfloat16)
    out = torch.empty_like(x)
    mod = torch.randn((4, 6 * D), device="cuda").to(torch.bfloat16)
    t = timeit(lambda: N.check(lib.gtav_ln_modulate(x.data_ptr(), out.data_ptr(), P, D, mod.data_ptr(), 6 * D, 0, D, None, P, s), "ln"))
    print(json.dumps(dict(kernel="ln_modulate", rows=P, us=round(t * 1e3, 2))), flush=True)
    qkv = torch.randn((P, 3 * D), device="cuda").to(torch.bfloat16)
    rot = torch.randn((P, 32, 2), device="cuda")
    t = timeit(lambda: N.check(lib.gtav_attention_seq(qkv.data_ptr(), out.data_ptr(), 1, P, H, rot.data_ptr(), 32, s), "attn"))
    print(json.dumps(dict(kernel="attention_seq", groups=1, us=round(t * 1e3, 2))), flush=True)
    cache = torch.randn((4 * P, 2 * D), device="cuda").to(torch.bfloat16)
    rott = torch.randn((5, 32, 2), device="cuda")
    t = timeit(lambda: N.check(lib.gtav_attention_temporal_last(qkv.data_ptr(), out.data_ptr(), 1, 4, P, H, rott.data_ptr(),
                                                                cache.data_ptr(), s), "attn_t"))
    print(json.dumps(dict(kernel="attention_temporal_last", ctx=4, us=round(t * 1e3, 2))), flush=True)
    e = torch.empty(1, device="cuda")
    t = timeit(lambda: e.zero_())
    print(json.dumps(dict(kernel="(torch zero_ of 4 bytes: event + launch floor)", us=round(t * 1e3, 2))), flush=True)


if __name__ == "__main__":
    if "--light" in sys.argv:
        light_cases()
        for (Nn, K) in ((3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096)):
            skinny_case(Nn, K)
        sys.exit(0)
    for B in (1, 8):
        M = 720 * B
        for (Nn, K) in ((3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096)):
            for bn in ((128, 256) if Nn % 256 == 0 else (128,)):
                gemm_case(M, Nn, K, bn)
    gemm_case(5, 198656, 1024, 256, epi=1)
    gemm_case(46080, 4096, 1024, 256)
    gemm_case(8192, 8192, 8192, 256)
