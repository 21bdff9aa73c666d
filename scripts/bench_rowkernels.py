"""In-graph time of the row-wise kernels of a large-batch last-frame step (LayerNorm + modulate, last-frame temporal attention,
spatial attention) at B rollouts, against the bytes they move.   python scripts/bench_rowkernels.py [B ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_graph import D, H, N, P, dev, graph_time, lib, rnd, st  # noqa: E402

for B in [int(a) for a in sys.argv[1:]] or [8, 64]:
    M = B * P
    h, hn = rnd(M, D), rnd(M, D)
    qkv, att = rnd(M, 3 * D), rnd(M, D)
    mod = rnd(B + 4, 6 * D, scale=0.1)
    rows = torch.arange(B, dtype=torch.int32, device=dev)
    rot = torch.randn((P, 32, 2), device=dev)
    rott = torch.randn((5, 32, 2), device=dev)
    cache = rnd(B * 4 * P, 2 * D)

    def ln():
        N.check(lib.gtav_ln_modulate(h.data_ptr(), hn.data_ptr(), M, D, mod.data_ptr(), 6 * D, 0, D, rows.data_ptr(), P, st()), "ln")

    def attn_s():
        N.check(lib.gtav_attention_seq(qkv.data_ptr(), att.data_ptr(), B, P, H, rot.data_ptr(), 32, st()), "attn")

    def attn_t():
        N.check(lib.gtav_attention_temporal_last(qkv.data_ptr(), att.data_ptr(), B, 4, P, H, rott.data_ptr(), cache.data_ptr(), st()), "attn_t")

    for name, f, mb in (("ln_modulate", ln, M * D * 2 * 2 / 1e6), ("attention_temporal_last ctx=4", attn_t, (M * 4 * D * 2 + B * 4 * P * 2 * D * 2) / 1e6),
                        ("attention_seq S=144", attn_s, M * 4 * D * 2 / 1e6)):
        us = graph_time([f] * 32)
        print(json.dumps(dict(kernel=name, B=B, rows=M, us=round(us, 2), mbytes=round(mb, 1), gbs=round(mb / us * 1e3, 0))), flush=True)
