"""BASELINE config 4: ViT-L-20 VAE encode + decode throughput sweep, 360x640 frames, batch 1..256, bf16, one B200.
CUDA events on the current stream, 3 warm-ups; weights (0.46 GB bf16) exceed L2, so every pass streams them from HBM.
One JSON line per batch size: frames/s and TFLOP/s (96.58 GFLOP encode, 191.69 GFLOP decode per frame)
against the measured sustained bf16 peak.

    python scripts/bench_vae.py [--max 256]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtav_b200.model.vae import VAE_models  # noqa: E402

ENC_GF, DEC_GF = 96.58, 191.69


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max", type=int, default=256)
    args = ap.parse_args()
    peak = 1404.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["bf16_tflops_sustained"]
    torch.manual_seed(0)
    vae = VAE_models["vit-l-20-shallow-encoder"]().cuda().eval()
    n = 1
    while n <= args.max:
        img = torch.rand((n, 3, 360, 640), device="cuda", generator=torch.Generator("cuda").manual_seed(0)) * 2 - 1
        z = vae.encode_mean(img)
        reps = max(2, min(20, 64 // n))
        ms_enc = timed(lambda: vae.encode_mean(img), reps)
        ms_dec = timed(lambda: vae.decode(z, to_uint8=True), reps)
        tf_enc, tf_dec = ENC_GF * n / ms_enc, DEC_GF * n / ms_dec
        print(json.dumps(dict(config="c4 VAE sweep", frames=n, encode_ms=round(ms_enc, 3), decode_ms=round(ms_dec, 3),
                              encode_frames_per_s=round(1e3 * n / ms_enc, 1), decode_frames_per_s=round(1e3 * n / ms_dec, 1),
                              roundtrip_frames_per_s=round(1e3 * n / (ms_enc + ms_dec), 1),
                              encode_tflops=round(tf_enc, 1), decode_tflops=round(tf_dec, 1),
                              frac_of_sustained_peak=round((ENC_GF + DEC_GF) * n / (ms_enc + ms_dec) / peak, 4))), flush=True)
        del img, z
        n *= 2


if __name__ == "__main__":
    main()
