#!/usr/bin/env python
"""Headline benchmark: generated frames/sec end-to-end (100-step DiT + VAE decode), BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c1]

A *step* is one complete rollout batch of the README inference shape (BASELINE config 2: B=1 rollout,
32 frames, 4 prompt frames, 100 noise steps => 28 generated frames x 101 DiT evaluations, plus VAE
encode of the prompt and decode of all 32 frames), run through gtav_b200's Sampler (C ABI -> sm_100a
kernels).  One JSON line is printed by rank 0:
  value  : generated frames/s with the prompt already resident in HBM (device-timed, max over ranks)
  e2e    : same metric through the public API with HOST buffers: pinned prompt video H2D and uint8 frames
           D2H inside the timed region
  roofline: the dominant kernel of the default (frame-cache) algorithm at B = 1, the weight-streaming tcgen05 GEMM,
           against the measured HBM bandwidth: algorithmic bytes of its 128 launches per step / their CUDA-event time
           (weights HBM-cold, as in the real step); at B > 1 and in the `dense` leg the tiled / CTA-pair GEMM against
           the measured sustained bf16 tensor throughput
  cpu_baseline: the CPU port of the reference (oracle/) timed on this box's host cores on a bounded sample
With --impl reference only the CPU arm runs (the Python reference cannot travel to the GPU box; the
oracle port is its restatement) and prints the same line with "impl": "reference".
N > 1 (torchrun): every rank runs its own rollouts (weak scaling, no collective on the data path).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "generated frames/sec end-to-end (100-step DiT + VAE decode)"
UNIT = "frames/s"
WORKLOADS = {
    # name: (B rollouts per GPU, total frames, prompt frames, noise steps, actions)
    "c2": dict(B=1, total=32, n_prompt=4, steps=100, actions=False,
               desc="README shape: unconditioned DiT-S/2, B=1, 32 frames (4 prompt), 100 noise steps, 360x640"),
    "c3": dict(B=8, total=32, n_prompt=4, steps=100, actions=True,
               desc="action-conditioned (W key), B=8 rollouts, 32 frames (4 prompt), 100 noise steps"),
    "c1": dict(B=1, total=8, n_prompt=4, steps=10, actions=False,
               desc="generate.py CPU case: B=1, 8 frames (4 prompt), 10 noise steps"),
    # 64 rollouts in total, sharded over the ranks (strong scaling): B = 64 / world per GPU
    "c5": dict(B=64, total=32, n_prompt=4, steps=100, actions=True, shard=True,
               desc="64 independent action-conditioned rollouts batch-sharded over the GPUs, 32 frames (4 prompt), 100 noise steps"),
}
DIT_STEP_GFLOP = 589.09          # per sample, 5-frame window
DIT_GEMM_GFLOP = 579.8           # the four per-half GEMMs only
VAE_ENC_GFLOP, VAE_DEC_GFLOP = 96.58, 191.69


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], hbm=d["hbm_gbs"], src="measured")
    return dict(tf=1400.0, tf_burst=1590.0, hbm=6650.0, src="fallback")


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, reasons, smax = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # "under load": upper half of the samples (the loop is busy the whole time; idle tails drop out)
        med = sm[len(sm) * 3 // 4] if sm else None
        return dict(sm_mhz=med, sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------ product arm
def build_models(device):
    from gtav_b200.model.dit import DiT_models
    from gtav_b200.model.vae import VAE_models
    torch.manual_seed(0)
    dit = DiT_models["DiT-S/2"]()
    vae = VAE_models["vit-l-20-shallow-encoder"]()
    with torch.no_grad():      # random-init of the architecture, adaLN made non-zero so no block is an identity
        g = torch.Generator().manual_seed(1)
        for b in dit.blocks:
            for h in ("s", "t"):
                lin = getattr(b, f"{h}_adaLN_modulation")[-1]
                lin.weight.normal_(std=0.02, generator=g)
                lin.bias.normal_(std=0.02, generator=g)
        dit.final_layer.linear.weight.normal_(std=0.02, generator=g)
    return dit.to(device).eval(), vae.to(device).eval()


def synthetic_prompt(B, n):
    """Blue->red solid-colour ramp like the reference's dummy dataset, [B, n, 3, 360, 640] fp32 in [0,1]."""
    w = torch.linspace(0, 1, 5)[:n].view(n, 1)
    col = (1 - w) * torch.tensor([0.0, 0.0, 1.0]) + w * torch.tensor([1.0, 0.0, 0.0])
    return col.view(1, n, 3, 1, 1).expand(B, n, 3, 360, 640).contiguous()


def launches_per_rollout(wl, algorithm, depth=16, enc=6, dec=12):
    """Kernels of OURS launched per rollout batch (mirrors dit_engine.cu / vae_engine.cu / sampler.cu)."""
    backbone = 2 + 2 * depth * 7 + 3                           # patchify, patch GEMM, 7 per half, final LN/GEMM/unpatchify
    last = backbone
    if algorithm == "cached" and wl["B"] == 1 and os.environ.get("GTAV_FUSE", "1") != "0" and os.environ.get("GTAV_SKINNY", "1") != "0":
        # one rollout: LayerNorm + modulate and the temporal attention run inside the weight-streaming GEMMs' reduce
        # (dit_engine.cu: fuse_ln / fuse_tattn) - first LN, then 5 kernels per spatial half and 4 per temporal half
        last = 2 + 1 + depth * (5 + 4) + 2
    step = 1 + (last if algorithm == "cached" else backbone) + 1   # step_prep, backbone (window or last frame), ddim
    gen = wl["total"] - wl["n_prompt"]
    context = (2 + 2 * depth * 7) if algorithm == "cached" else 0
    per_frame = 1 + 1 + 4 + context + (wl["steps"] + 1) * step  # clamp, set_int, conditioning(4), context pass, steps
    vae_enc = 2 + enc * 7 + 3
    vae_dec = 2 + dec * 7 + 3
    return gen * per_frame + vae_enc + vae_dec * ((wl["B"] * wl["total"] + 31) // 32)


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/r*/roofline_traffic.json)."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*", "roofline_traffic.json")), reverse=True):
        try:
            d = json.load(open(path)).get(kernel)
        except Exception:
            d = None
        if d:
            return d["traffic_bytes_per_launch"]
    return None


def _hot_weights(dit):
    """Packed bf16 weights in _pack order: per half qkv_w, out_w, out_b, fc1_w, fc1_b, fc2_w, fc2_b."""
    dit._pack()
    keep = dit._engine[1]
    return [keep[i * 7:(i + 1) * 7] for i in range(2 * dit.depth)]


def _time_passes(one_pass, reps=10):
    """GPU time of one_pass: captured into a CUDA graph (the launches keep their programmatic-dependent-launch
    edges, as in the sampler) and replayed, so host launch cost is not in the number; CUDA events on the
    replay stream."""
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        one_pass()                                   # first-use configuration outside capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            one_pass()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gemm_roofline(dit, B, pk):
    """Tensor roofline of the tiled tcgen05 GEMM (the dense-window step's dominant kernel): the four hot GEMM
    shapes of a DiT step with CUDA events through the C ABI, cycling through the 16 blocks' own weight matrices so
    weights come from HBM (1.2 GB >> 126 MB L2) as in the real step."""
    import gtav_b200._native as N
    lib = N.load()
    M = B * 5 * 144
    dev = torch.device("cuda")
    D = 1024
    bufs = dict(a1=torch.randn((M, D), device=dev).to(torch.bfloat16), a4=torch.randn((M, 4 * D), device=dev).to(torch.bfloat16),
                o1=torch.empty((M, D), device=dev, dtype=torch.bfloat16), o3=torch.empty((M, 3 * D), device=dev, dtype=torch.bfloat16),
                o4=torch.empty((M, 4 * D), device=dev, dtype=torch.bfloat16))
    bias = torch.zeros(4 * D, device=dev, dtype=torch.bfloat16)
    halves = _hot_weights(dit)

    def run(a, w, out, n, k, epi):
        N.check(lib.gtav_gemm_bf16(a.data_ptr(), k, w.data_ptr(), k, out.data_ptr(), n, M, n, k, epi, bias.data_ptr(), None, 0,
                                   None, 0, None, 1, 0, N.current_stream()), "gemm")

    def one_pass():
        for h in halves:
            run(bufs["a1"], h[0], bufs["o3"], 3 * D, D, N.EPI_STORE)
            run(bufs["a1"], h[1], bufs["o1"], D, D, N.EPI_BIAS)
            run(bufs["a1"], h[3], bufs["o4"], 4 * D, D, N.EPI_BIAS_GELU_TANH)
            run(bufs["a4"], h[5], bufs["o1"], D, 4 * D, N.EPI_BIAS)
    ms = _time_passes(one_pass)                         # GEMM time of one DiT step (128 launches)
    tf = DIT_GEMM_GFLOP * B / ms                        # GFLOP / ms = TFLOP/s
    return dict(bound="tensor", achieved=round(tf, 1), peak=pk["tf"], unit="TFLOP/s", frac=round(tf / pk["tf"], 4),
                traffic=measured_traffic("gemm_bf16_kernel"), kernel="gemm_bf16_kernel (tcgen05, 128 launches per dense DiT step)",
                gemm_ms_per_dit_step=round(ms, 4), peak_source=f"{pk['src']} sustained bf16")


def skinny_roofline(dit, B, pk):
    """HBM roofline of the weight-streaming GEMM (the last-frame step's dominant kernel): algorithmic bytes per
    launch = the weight matrix (read once) + the token operand and result, / the average launch time, over the 128
    launches of one step on the blocks' own weights (805 MB per pass >> L2, so every launch streams from HBM)."""
    import gtav_b200._native as N
    lib = N.load()
    M = B * 144
    if B != 1:
        return None
    dev = torch.device("cuda")
    D = 1024
    a1 = torch.randn((M, D), device=dev).to(torch.bfloat16)
    a4 = torch.randn((M, 4 * D), device=dev).to(torch.bfloat16)
    o1 = torch.empty((M, D), device=dev, dtype=torch.bfloat16)
    o3 = torch.empty((M, 3 * D), device=dev, dtype=torch.bfloat16)
    o4 = torch.empty((M, 4 * D), device=dev, dtype=torch.bfloat16)
    bias = torch.zeros(4 * D, device=dev, dtype=torch.bfloat16)
    ws = torch.empty(lib.gtav_gemm_skinny_workspace_bytes(M), dtype=torch.uint8, device=dev)
    counters = torch.zeros(512, dtype=torch.int32, device=dev)
    halves = _hot_weights(dit)
    shapes = [(0, a1, o3, 3 * D, D, N.EPI_STORE), (1, a1, o1, D, D, N.EPI_BIAS), (3, a1, o4, 4 * D, D, N.EPI_BIAS_GELU_TANH),
              (5, a4, o1, D, 4 * D, N.EPI_BIAS)]
    def one_pass():
        for h in halves:
            for wi, a, out, n, k, epi in shapes:
                N.check(lib.gtav_gemm_skinny_bf16(a.data_ptr(), k, h[wi].data_ptr(), k, out.data_ptr(), n, M, n, k, epi, bias.data_ptr(),
                                                  None, 0, None, 0, None, 144, 0, ws.data_ptr(), counters.data_ptr(), N.current_stream()),
                        "gemm_skinny")
    ms = _time_passes(one_pass)
    per_half = sum(2 * (n * k + M * k + M * n) for _, _, _, n, k, _ in shapes)          # bf16 bytes: W + A + out
    gb = len(halves) * per_half / 1e9
    gbs = gb / (ms / 1e3)
    return dict(bound="hbm", achieved=round(gbs, 1), peak=pk["hbm"], unit="GB/s", frac=round(gbs / pk["hbm"], 4),
                traffic=measured_traffic("gemm_skinny_kernel"),
                kernel="gemm_skinny_kernel (tcgen05 weight-streaming GEMM, 128 launches per last-frame DiT step)",
                algorithmic_mb_per_launch=round(gb * 1e3 / (4 * len(halves)), 3), us_per_launch=round(ms * 1e3 / (4 * len(halves)), 3),
                gemm_ms_per_last_frame_step=round(ms, 4), peak_source=f"{pk['src']} HBM copy bandwidth",
                note="latency-bound, not bandwidth-bound: a launch is a chain of L2 round trips (CTA entry, operand slab 1.5 us, "
                     "MMA 0.8, split-K partials out 2.1, rendezvous 0.7-1.3, reduce + fused LayerNorm / temporal attention 1.8-2.7; "
                     "profiles/r01/skinny_in_step_trace_v8.txt); its DRAM traffic equals the algorithmic bytes "
                     "(profiles/r01/roofline_traffic.json)")


def cpu_baseline(wl, max_seconds=25.0):
    """The reference's CPU fp32 path (oracle port of it) on this box's cores, bounded sample, extrapolated."""
    from oracle import reference_port as rp
    from oracle.weights import DiTConfig, VAEConfig, make_dit_state, make_vae_state, w_key_actions
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dcfg, vcfg = DiTConfig(), VAEConfig()
    dsd, vsd = make_dit_state(dcfg, seed=0), make_vae_state(vcfg, seed=0)
    B = 1
    x = torch.randn(B, 5, 16, 18, 32)
    t = torch.tensor([[15, 15, 15, 15, 999]])
    a = w_key_actions(B, 5) if wl["actions"] else None
    with torch.inference_mode():
        rp.dit_forward(dsd, dcfg, x[:, :2], t[:, :2], None if a is None else a[:, :2])      # warm-up
        n_dit, t0 = 0, time.perf_counter()
        while n_dit < 3 or (time.perf_counter() - t0 < max_seconds * 0.7 and n_dit < 12):
            rp.dit_forward(dsd, dcfg, x, t, a)
            n_dit += 1
        t_dit = (time.perf_counter() - t0) / n_dit
        img = torch.rand(1, 3, 360, 640) * 2 - 1
        t0 = time.perf_counter(); z = rp.vae_encode_mean(vsd, vcfg, img); t_enc = time.perf_counter() - t0
        t0 = time.perf_counter(); rp.vae_decode(vsd, vcfg, z); t_dec = time.perf_counter() - t0
    gen = wl["total"] - wl["n_prompt"]
    per_rollout = gen * (wl["steps"] + 1) * t_dit + wl["n_prompt"] * t_enc + wl["total"] * t_dec
    return dict(value=round(gen / per_rollout, 6), unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=(f"{n_dit} DiT window steps (B=1,T=5) at {t_dit:.3f} s + 1 VAE encode ({t_enc:.3f} s) + 1 decode "
                        f"({t_dec:.3f} s) of oracle/reference_port.py fp32, extrapolated to the {gen}x{wl['steps'] + 1}-step rollout"),
                s_per_dit_step=round(t_dit, 4))


def run_reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    t0 = time.perf_counter()
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(wl, max_seconds=max(6.0, 120.0 / (args.warmup + args.steps)))
        if i >= args.warmup:
            vals.append(cb)
    v = sum(c["value"] for c in vals) / len(vals)
    cb = dict(vals[-1], value=round(v, 6))
    line = dict(impl="reference", metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=round(1000.0 * (wl["total"] - wl["n_prompt"]) / v, 1),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=wl["desc"], l2="n/a (CPU)"), cpu_baseline=cb,
                e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                wall_s=round(time.perf_counter() - t0, 1))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--algorithm", default="cached", choices=["cached", "dense"],
                    help="cached: context pass per frame + last-frame-only steps (default, same results); "
                         "dense: every step recomputes the whole window like the reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-algorithm comparison leg")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    strong = bool(wl.get("shard"))
    if strong:                                   # fixed total work: this rank's share of the rollouts (shard.py: r % world)
        from gtav_b200.shard import shard_rollouts
        wl["B"] = len(shard_rollouts(wl["B"], rank, world))

    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gtav_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from gtav_b200.sampler import Sampler
    pk = peaks()
    dit, vae = build_models(dev)
    cached = args.algorithm == "cached"
    sampler = Sampler(dit, vae, noise_steps=wl["steps"], frame_cache=cached)
    B, total, n_prompt = wl["B"], wl["total"], wl["n_prompt"]
    gen = total - n_prompt
    prompt_host = synthetic_prompt(B, n_prompt).pin_memory()
    prompt_dev = prompt_host.to(dev)
    actions = None
    if wl["actions"]:
        actions = torch.zeros(B, total, 25, device=dev)
        actions[:, :, 3] = 1.0
    gen_rng = torch.Generator(device=dev).manual_seed(1000 + rank)
    out_host = torch.empty((B, total, 360, 640, 3), dtype=torch.uint8).pin_memory()

    def rollout_resident():
        frames, _ = sampler.generate(prompt_dev, actions, total, generator=gen_rng)
        return frames

    def rollout_e2e():
        p = prompt_host.to(dev, non_blocking=True)
        frames, _ = sampler.generate(p, actions, total, generator=gen_rng)
        out_host.copy_(frames, non_blocking=True)
        return frames

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(args.warmup):
        rollout_resident()
    with ClockSampler(local) as cs:
        ms_total = timed(rollout_resident, args.steps)
    clocks = cs.summary()
    rollout_e2e()
    ms_e2e = timed(rollout_e2e, args.steps)

    # ms per DiT step: two generated frames (context pass, if cached, + steps+1 steps each), device-timed
    def time_frames(smp, n_frames=2):
        lat = smp.encode_prompt(prompt_dev)
        smp.sample_latents(lat, actions, n_prompt + 1, generator=gen_rng)          # capture / warm
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        smp.sample_latents(lat, actions, n_prompt + n_frames, generator=gen_rng)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (n_frames * (wl["steps"] + 1))
    ms_dit_step = time_frames(sampler)

    if rank == 0:
        frames_total = world * B * gen * args.steps
        value = frames_total / (ms_total / 1000.0)
        e2e_v = frames_total / (ms_e2e / 1000.0)
        T = 5
        step_dense_gf = DIT_STEP_GFLOP * B
        if cached:
            # executed FLOPs per generated frame: one (T-1)-frame context pass + (steps+1) last-frame steps
            last_gf = (116.411 + 1.359 + 0.009437 * T) * B
            ctx_gf = (116.411 + 1.359) * (T - 1) * B + 0.009437 * (T - 1) ** 2 * B
            exec_gf_per_step = (ctx_gf + (wl["steps"] + 1) * last_gf) / (wl["steps"] + 1)
            roof = skinny_roofline(dit, B, pk) or gemm_roofline(dit, B, pk)
            algo = ("frame cache: per generated frame one context pass over the T-1 context frames stores every temporal "
                    "layer's K/V, then each of the steps+1 DDIM steps recomputes only the frame being denoised (144*B rows) "
                    "against it - same arithmetic per row as the dense window (bit-identical with the tiled GEMM, "
                    "tests/test_model_gpu.py), 4.8x fewer executed FLOPs")
        else:
            exec_gf_per_step = step_dense_gf
            roof = gemm_roofline(dit, B, pk)
            algo = "dense (every step recomputes the whole 5-frame window, like the reference)"
        line = dict(metric=METRIC, value=round(value, 3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=round(ms_total / args.steps, 2), higher_is_better=True, scaling="strong" if strong else "weak",
                    vs_baseline=None,
                    dtype="bf16", data="synthetic",
                    config=dict(workload=wl["desc"], rollouts_per_gpu=B, frames=total, prompt_frames=n_prompt,
                                noise_steps=wl["steps"], dit_evals_per_rollout=gen * (wl["steps"] + 1),
                                weights="random-init DiT-S/2 (607.9M) + ViT-L-20 VAE (229.2M), adaLN non-zero",
                                l2="inputs larger than L2: 0.8-1.2 GB of bf16 weights streamed per DiT step (L2 126 MB)",
                                algorithm=algo),
                    ms_per_dit_step=round(ms_dit_step, 4),
                    step_roofline=dict(bound="tensor", executed_tflops=round(exec_gf_per_step / ms_dit_step, 1),
                                       reference_equivalent_tflops=round(step_dense_gf / ms_dit_step, 1), peak=pk["tf"],
                                       unit="TFLOP/s", frac_executed=round(exec_gf_per_step / ms_dit_step / pk["tf"], 4),
                                       note="executed = FLOPs this algorithm runs per DDIM step / measured ms per step; "
                                            "reference-equivalent = the dense window's 589.09 GFLOP x B / the same time"),
                    roofline=roof,
                    e2e=dict(value=round(e2e_v, 3), unit=UNIT, h2d_bytes_per_step=prompt_host.numel() * 4,
                             d2h_bytes_per_step=out_host.numel()),
                    gpu_launches=launches_per_rollout(wl, args.algorithm) * args.steps, clocks=clocks)
        if cached and not args.no_dense:
            # the dense algorithm on the same box, for the record: ms per dense step and the tiled GEMM's tensor roofline
            dense = Sampler(dit, vae, noise_steps=wl["steps"], frame_cache=False)
            ms_dense = time_frames(dense, n_frames=1)
            dense.close()
            line["dense"] = dict(ms_per_dit_step=round(ms_dense, 4), frames_per_s=round(B * 1000.0 / (ms_dense * (wl["steps"] + 1)), 3),
                                 step_tflops=round(step_dense_gf / ms_dense, 1), frac=round(step_dense_gf / ms_dense / pk["tf"], 4),
                                 gemm_roofline=gemm_roofline(dit, B, pk),
                                 note="every step recomputes the whole window; DiT steps only (no VAE), 1 generated frame timed")
        if not args.no_cpu_baseline and world == 1:             # rank 0 at N = 1 only
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
