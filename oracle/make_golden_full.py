"""Mint the FULL-SIZE golden vectors: the BASELINE configs themselves, run through the UNMODIFIED reference (CPU, fp32).

Run in the build container only (needs /root/reference; ~10 min on 8 cores):   python -m oracle.make_golden_full
Outputs (committed): tests/golden/c1_rollout.safetensors, c3_step.safetensors, trainer.safetensors, full_meta.json.
TEST INFRASTRUCTURE.

  * C1  = BASELINE config 1 exactly (reference generate.py:186-244 driven literally): depth-16 DiT + VAE 6/12, B=1,
          dummy blue->red prompt, 8 frames, 10 DDIM steps, fixed noise; latents of every frame and the decoded uint8
          frames (every 8th pixel) for the non-degenerate weights and for zero adaLN linears.  The same rollout is then
          repeated with oracle/reference_port.py in its bf16-rounding mode: its distance to the fp32 reference (latent
          max-abs per frame, PSNR of the decoded frames) is what bf16 arithmetic costs on this config and calibrates the
          bound of the GPU test (PSNR(product) >= PSNR(bf16 model of the reference) - 3 dB).
  * C3 step = one reference denoise_step (train_dit.py:30-125) at BASELINE config 3's shape (B=8, actions, T=5).
  * trainer = DiffusionTrainer.predict / predict_noise (train_dit.py:373-552) of the reference, instantiated without
          its constructor (which needs accelerate + datasets) and with write_video / visualize_step intercepted.
The wall times of the reference's C1 run on this container's cores are recorded in full_meta.json.
"""
from __future__ import annotations

import json
import math
import os
import sys
import tempfile
import time
import types
import warnings

import torch
from safetensors.torch import save_file

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims  # noqa: E402
from oracle import reference_port as rp  # noqa: E402
from oracle.cases import C1, C3_STEP, TRAINER, seeded_randn  # noqa: E402
from oracle.make_golden import OUT, build_ref_dit, build_ref_vae  # noqa: E402
from oracle.weights import (DiTConfig, VAEConfig, dummy_prompt, make_dit_state, make_vae_state,  # noqa: E402
                            w_key_actions)

SCALE = 0.07843137255


def psnr_u8(a, b):
    mse = float(((a.float() - b.float()) ** 2).mean())
    return 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


def c1_noise(c):
    g = torch.Generator().manual_seed(c["seed"])
    return [torch.randn((1, 1, 16, 18, 32), generator=g) for _ in range(c["total_frames"] - c["n_prompt"])]


@torch.inference_mode()
def mint_c1(ref, meta):
    c = C1
    dcfg = DiTConfig(depth=c["depth"])
    vcfg = VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"])
    vsd = make_vae_state(vcfg, seed=0)
    vae = build_ref_vae(ref, vcfg, vsd)
    betas = ref.sigmoid_beta_schedule(1000)
    abar = torch.cumprod(1.0 - betas.float(), dim=0)
    n_prompt, total, steps = c["n_prompt"], c["total_frames"], c["noise_steps"]
    noise_range = torch.linspace(0, 999, steps + 1)
    video = dummy_prompt(5)[None]
    out = {}
    for tag, degenerate in (("initB", False), ("initA", True)):
        dsd = make_dit_state(dcfg, seed=0, degenerate=degenerate)
        model = build_ref_dit(ref, dcfg, dsd)
        model.max_frames = 5
        noise = c1_noise(c)
        t0 = time.perf_counter()
        frames = video[:, :n_prompt].reshape(n_prompt, 3, 360, 640)
        lat = vae.encode(frames * 2 - 1).mean * SCALE
        x = lat.reshape(1, n_prompt, 18, 32, 16).permute(0, 1, 4, 2, 3).contiguous()
        prompt_lat = x.clone()
        t_enc = time.perf_counter() - t0
        t0 = time.perf_counter()
        for i in range(n_prompt, total):
            chunk = noise[i - n_prompt].clamp(-20, 20)
            x = torch.cat([x, chunk], dim=1)
            start = max(0, i + 1 - model.max_frames)
            for k in reversed(range(steps + 1)):
                xp, _ = ref.denoise_step(dit_model=model, x_noisy=x, actions=None, noise_idx=k, stabilization_level=15,
                                         noise_range=noise_range, alphas_cumprod=abar.reshape(-1, 1, 1, 1),
                                         start_frame=start, dtype=torch.bfloat16)
                x[:, -1:] = xp[:, -1:]
        t_dit = time.perf_counter() - t0
        t0 = time.perf_counter()
        z = x.permute(0, 1, 3, 4, 2).reshape(total, 576, 16)
        img = (vae.decode(z / SCALE) + 1) / 2
        u8 = torch.clamp(img * 255, 0, 255).byte().reshape(1, total, 3, 360, 640).permute(0, 1, 3, 4, 2)
        t_dec = time.perf_counter() - t0
        out[f"{tag}.prompt_latents"] = prompt_lat.contiguous()
        out[f"{tag}.latents"] = x.contiguous()
        out[f"{tag}.frames_u8_sub"] = u8[:, :, ::8, ::8].contiguous()
        n_steps = (total - n_prompt) * (steps + 1)
        m = dict(cpu_threads=torch.get_num_threads(), encode_s=round(t_enc, 2), dit_loop_s=round(t_dit, 2),
                 decode_s=round(t_dec, 2), total_s=round(t_enc + t_dit + t_dec, 2), dit_steps=n_steps,
                 s_per_dit_step=round(t_dit / n_steps, 4),
                 generated_frames_per_s=round((total - n_prompt) / (t_enc + t_dit + t_dec), 5),
                 latents_std=float(x[:, n_prompt:].std()), latents_absmax=float(x[:, n_prompt:].abs().max()))
        print(tag, "reference fp32:", m, flush=True)
        # ---- the same rollout with the bf16-rounding model of the reference's autocast graph (oracle port)
        it = iter(noise)
        xb = rp.rollout(dsd, dcfg, rp.encode_prompt(vsd, vcfg, video[:, :n_prompt], rp.BF16), None, total, steps,
                        lambda i: next(it), rd=rp.BF16)
        u8b = rp.decode_to_uint8(vsd, vcfg, xb, rp.BF16)
        err = (xb - x).abs()
        m["bf16_model"] = dict(latent_max_abs_per_frame=[round(float(err[:, f].max()), 5) for f in range(total)],
                               latent_mean_abs=round(float(err[:, n_prompt:].mean()), 6),
                               psnr_db=round(psnr_u8(u8b[:, :, ::8, ::8], u8[:, :, ::8, ::8]), 2),
                               psnr_generated_db=round(psnr_u8(u8b[:, n_prompt:, ::8, ::8], u8[:, n_prompt:, ::8, ::8]), 2))
        print(tag, "bf16 model vs fp32 reference:", m["bf16_model"], flush=True)
        meta[f"c1_{tag}"] = m
        del model
    save_file(out, os.path.join(OUT, "c1_rollout.safetensors"))


@torch.inference_mode()
def mint_c3_step(ref, meta):
    c = C3_STEP
    dcfg = DiTConfig(depth=c["depth"])
    model = build_ref_dit(ref, dcfg, make_dit_state(dcfg, seed=0))
    betas = ref.sigmoid_beta_schedule(1000)
    abar = torch.cumprod(1.0 - betas.float(), dim=0)
    x = seeded_randn((c["B"], c["frames"], 16, 18, 32), c["seed"])
    a = w_key_actions(c["B"], c["frames"])
    for b in range(c["B"]):                     # rollouts differ in their actions too (a second key on some frames)
        a[b, b % c["frames"]:, 5 + b] = 1.0
    t0 = time.perf_counter()
    xp, v = ref.denoise_step(dit_model=model, x_noisy=x, actions=a, noise_idx=c["noise_idx"], stabilization_level=15,
                             noise_range=torch.linspace(0, 999, c["noise_steps"] + 1),
                             alphas_cumprod=abar.reshape(-1, 1, 1, 1), start_frame=c["start_frame"], dtype=torch.bfloat16)
    meta["c3_step"] = dict(seconds=round(time.perf_counter() - t0, 2), v_std=float(v.std()), v_absmax=float(v.abs().max()))
    print("c3 step:", meta["c3_step"], flush=True)
    save_file({"actions": a.contiguous(), "x_pred": xp.float().contiguous(), "v_pred": v.float().contiguous()},
              os.path.join(OUT, "c3_step.safetensors"))


@torch.inference_mode()
def mint_trainer(ref, meta):
    """DiffusionTrainer.predict / predict_noise of the reference (train_dit.py:373-552).  The constructor builds an
    Accelerator, datasets and optimiser (train_dit.py:174-266), none of which exist offline, so the object is created
    with __new__ and given exactly the attributes the two methods read; register_buffers is the reference's own."""
    c = TRAINER
    rtrain = ref.train_module
    dcfg = DiTConfig(depth=c["depth"])
    vcfg = VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"])
    tr = object.__new__(rtrain.DiffusionTrainer)
    tr.config = types.SimpleNamespace(ddim_noise_steps=c["ddim_noise_steps"], ddim_noise_steps_inference=c["ddim_noise_steps_inference"],
                                      ctx_max_noise_idx=3, noise_abs_max=c["noise_abs_max"], n_prompt_frames=c["n_prompt_frames"],
                                      use_action_conditioning=True, model_name="dit")
    tr.accelerator = types.SimpleNamespace(device=torch.device("cpu"), is_local_main_process=False, process_index=0)
    tr.dtype = torch.bfloat16
    tr.dit = build_ref_dit(ref, dcfg, make_dit_state(dcfg, seed=0))
    tr.vae = build_ref_vae(ref, vcfg, make_vae_state(vcfg, seed=0))
    tr.register_buffers()
    video = dummy_prompt(5)[None]
    loader = [dict(video=video, actions=w_key_actions(1, 5))]
    cap = {}
    orig_decode = tr.decode_frames

    def decode_spy(frames, num_frames, dtype=torch.bfloat16):
        cap["latents"] = frames.clone()
        return orig_decode(frames, num_frames, dtype=dtype)

    tr.decode_frames = decode_spy
    rtrain.write_video = lambda path, pixels, fps=10: cap.__setitem__("pixels", pixels.clone())
    rtrain.visualize_step = lambda **kw: cap.update({f"viz_{k}": (v.clone() if torch.is_tensor(v) else v) for k, v in kw.items()})
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            torch.manual_seed(c["seed_predict"])
            tr.predict(loader, 0, 0, num_frames=c["num_frames"])
            out = {"predict.latents": cap["latents"].float().contiguous(),
                   "predict.frames_u8_sub": cap["pixels"][None][:, :, ::8, ::8].contiguous()}
            torch.manual_seed(c["seed_predict_noise"])
            tr.predict_noise(loader, 0, 0)
        finally:
            os.chdir(cwd)
    out["predict_noise.latents"] = cap["viz_x_curr"].float().contiguous()
    out["predict_noise.x_noisy_in"] = cap["viz_x_noisy"].float().contiguous()
    out["predict_noise.noise"] = cap["viz_noise"].float().contiguous()
    out["predict_noise.x_pred"] = cap["viz_pred"].float().contiguous()
    out["predict_noise.v_pred"] = cap["viz_v"].float().contiguous()
    out["stabilization_level"] = torch.tensor([int(tr.stabilization_level)])
    out["noise_range_inference"] = tr.noise_range_inference.clone()
    save_file(out, os.path.join(OUT, "trainer.safetensors"))
    meta["trainer"] = dict(stabilization_level=int(tr.stabilization_level), levels=tr.noise_range_inference.tolist(),
                           predict_latents_std=float(out["predict.latents"].std()))
    print("trainer:", meta["trainer"], flush=True)


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shims.load()
    only = set(sys.argv[1:])
    path = os.path.join(OUT, "full_meta.json")
    meta = json.load(open(path)) if os.path.exists(path) else {}
    meta["host"] = dict(cpu_count=os.cpu_count(), torch=torch.__version__)
    if not only or "trainer" in only:
        mint_trainer(ref, meta)
    if not only or "c3" in only:
        mint_c3_step(ref, meta)
    if not only or "c1" in only:
        mint_c1(ref, meta)
    json.dump(meta, open(path, "w"), indent=1, sort_keys=True)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
