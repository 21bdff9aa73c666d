"""Mint golden vectors by executing the UNMODIFIED reference (CPU, fp32) on deterministic weights.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
Outputs: tests/golden/*.safetensors (small; committed).  TEST INFRASTRUCTURE.

The reference has no tests, seeds or golden tensors of its own, so these files are
what pins oracle/reference_port.py: weights come from oracle/weights.py (rebuildable anywhere),
inputs from seeded CPU generators, outputs from the reference's own nn.Modules / denoise_step.
"""
from __future__ import annotations

import os
import sys
import warnings

import torch
from safetensors.torch import save_file

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims  # noqa: E402
from oracle.weights import (DiTConfig, VAEConfig, make_dit_state, make_vae_state, dummy_prompt,  # noqa: E402
                            w_key_actions)
from oracle.cases import (CASES_DIT, CASES_DENOISE, CASES_VAE, ROLLOUT, seeded_randn, seeded_rand,  # noqa: E402
                          subsample_image)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_ref_dit(ref, cfg: DiTConfig, sd):
    m = ref.DiT(input_h=cfg.input_h, input_w=cfg.input_w, patch_size=cfg.patch_size, in_channels=cfg.in_channels,
                hidden_size=cfg.hidden_size, depth=cfg.depth, num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio,
                external_cond_dim=cfg.external_cond_dim, max_frames=cfg.max_frames)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m.eval()


def build_ref_vae(ref, cfg: VAEConfig, sd):
    m = ref.AutoencoderKL(latent_dim=cfg.latent_dim, patch_size=cfg.patch_size, enc_dim=cfg.dim, enc_depth=cfg.enc_depth,
                          enc_heads=cfg.heads, dec_dim=cfg.dim, dec_depth=cfg.dec_depth, dec_heads=cfg.heads,
                          input_height=cfg.input_height, input_width=cfg.input_width)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m.eval()


def mint_trainer_schedule(ref):
    """DiffusionTrainer.register_buffers (reference train_dit.py:288-327): the same schedule with clamp_min 1e-6, the integer
    DDIM level table of 16 steps and the stabilization level it yields."""
    betas = ref.sigmoid_beta_schedule(1000, clamp_min=0.000001)
    abar = torch.cumprod(1.0 - betas.to(torch.float32), dim=0)
    levels = torch.linspace(0, 999, 17).long()
    save_file({"betas_f64": betas.contiguous(), "alphas_cumprod_f32": abar.contiguous(), "noise_range_16": levels.contiguous()},
              os.path.join(OUT, "schedule_trainer.safetensors"))


@torch.inference_mode()
def mint_vae_posterior(ref):
    """The whole DiagonalGaussianDistribution of `vae.encode(img)` (reference model/vae.py:19-45, 306-322) for the small
    VAE case: raw moments (mean | logvar), clamped logvar and std."""
    c = CASES_VAE["e1_d1"]
    cfg = VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"])
    vae = build_ref_vae(ref, cfg, make_vae_state(cfg, seed=0))
    img = seeded_rand((c["N"], 3, 360, 640), c["seed"]) * 2 - 1
    post = vae.encode(img)
    save_file({"moments": post.parameters.float().contiguous(), "logvar": post.logvar.float().contiguous(),
               "std": post.std.float().contiguous(), "mode": post.mode().float().contiguous()},
              os.path.join(OUT, "vae_posterior.safetensors"))


@torch.inference_mode()
def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shims.load()

    # ---- schedule KATs (reference utils.py:30-48 + generate.py:195-197) -------------------
    betas = ref.sigmoid_beta_schedule(1000)
    abar = torch.cumprod(1.0 - betas.float(), dim=0)
    save_file({"betas_f64": betas.contiguous(), "alphas_cumprod_f32": abar.contiguous()},
              os.path.join(OUT, "schedule.safetensors"))

    mint_trainer_schedule(ref)

    # ---- DiT.forward -----------------------------------------------------------------------
    out = {}
    models = {}
    for name, c in CASES_DIT.items():
        cfg = DiTConfig(depth=c["depth"])
        key = (c["depth"], c["degenerate"])
        if key not in models:
            models[key] = build_ref_dit(ref, cfg, make_dit_state(cfg, seed=0, degenerate=c["degenerate"]))
        x = seeded_randn((c["B"], c["T"], 16, 18, 32), c["seed"])
        t = torch.tensor(c["t"], dtype=torch.long).reshape(c["B"], c["T"])
        a = w_key_actions(c["B"], c["T"]) if c["actions"] else None
        v = models[key](x, t, a)
        out[f"{name}.v"] = v.float().contiguous()
        out[f"{name}.x_sum"] = x.double().sum().reshape(1)
        print(name, "v std", float(v.std()), "absmax", float(v.abs().max()))
    save_file(out, os.path.join(OUT, "dit_forward.safetensors"))

    # ---- denoise_step (reference train_dit.py:30-125) --------------------------------------
    out = {}
    for name, c in CASES_DENOISE.items():
        cfg = DiTConfig(depth=c["depth"])
        model = models[(c["depth"], False)]
        x = seeded_randn((c["B"], c["frames"], 16, 18, 32), c["seed"])
        a = w_key_actions(c["B"], c["frames"]) if c["actions"] else None
        noise_range = torch.linspace(0, 999, c["noise_steps"] + 1)
        xp, v = ref.denoise_step(dit_model=model, x_noisy=x, actions=a, noise_idx=c["noise_idx"],
                                 stabilization_level=15, noise_range=noise_range,
                                 alphas_cumprod=abar.reshape(-1, 1, 1, 1), start_frame=c["start_frame"],
                                 dtype=torch.bfloat16)
        out[f"{name}.x_pred"] = xp.float().contiguous()
        out[f"{name}.v_pred"] = v.float().contiguous()
    save_file(out, os.path.join(OUT, "denoise_step.safetensors"))

    # ---- VAE encode / decode ---------------------------------------------------------------
    out = {}
    vaes = {}
    for name, c in CASES_VAE.items():
        cfg = VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"])
        vae = vaes.setdefault((c["enc_depth"], c["dec_depth"]), build_ref_vae(ref, cfg, make_vae_state(cfg, seed=0)))
        img = seeded_rand((c["N"], 3, 360, 640), c["seed"]) * 2 - 1
        mean = vae.encode(img).mean
        z = seeded_randn((c["N"], 576, 16), c["seed"] + 1)
        dec = vae.decode(z)
        out[f"{name}.mean"] = mean.float().contiguous()
        out[f"{name}.dec_sub"] = subsample_image(dec).contiguous()
        out[f"{name}.dec_sum"] = dec.double().sum().reshape(1)
        out[f"{name}.dec_abs_sum"] = dec.double().abs().sum().reshape(1)
        print(name, "mean std", float(mean.std()), "dec std", float(dec.std()))
    save_file(out, os.path.join(OUT, "vae.safetensors"))
    mint_vae_posterior(ref)

    # ---- a short autoregressive rollout driven exactly like reference generate.py:186-244 ----
    c = ROLLOUT
    dcfg = DiTConfig(depth=c["depth"])
    vcfg = VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"])
    model = models.get((c["depth"], False)) or build_ref_dit(ref, dcfg, make_dit_state(dcfg, seed=0))
    vae = vaes.get((c["enc_depth"], c["dec_depth"])) or build_ref_vae(ref, vcfg, make_vae_state(vcfg, seed=0))
    model.max_frames = 5
    video = dummy_prompt(5)[None]                                          # [1,5,3,360,640]
    n_prompt, total, steps = c["n_prompt"], c["total_frames"], c["noise_steps"]
    actions = w_key_actions(1, total) if c["actions"] else None
    frames = video[:, :n_prompt].reshape(n_prompt, 3, 360, 640)
    lat = vae.encode(frames * 2 - 1).mean * 0.07843137255
    x = lat.reshape(1, n_prompt, 18, 32, 16).permute(0, 1, 4, 2, 3).contiguous()
    prompt_lat = x.clone()
    noise_range = torch.linspace(0, 999, steps + 1)
    g = torch.Generator().manual_seed(c["seed"])
    for i in range(n_prompt, total):
        chunk = torch.randn((1, 1, 16, 18, 32), generator=g).clamp(-20, 20)
        x = torch.cat([x, chunk], dim=1)
        start = max(0, i + 1 - model.max_frames)
        for k in reversed(range(steps + 1)):
            xp, _ = ref.denoise_step(dit_model=model, x_noisy=x, actions=actions, noise_idx=k,
                                     stabilization_level=15, noise_range=noise_range,
                                     alphas_cumprod=abar.reshape(-1, 1, 1, 1), start_frame=start,
                                     dtype=torch.bfloat16)
            x[:, -1:] = xp[:, -1:]
    z = x.permute(0, 1, 3, 4, 2).reshape(total, 576, 16)
    img = (vae.decode(z / 0.07843137255) + 1) / 2
    u8 = torch.clamp(img * 255, 0, 255).byte().reshape(1, total, 3, 360, 640).permute(0, 1, 3, 4, 2)
    save_file({"prompt_latents": prompt_lat.contiguous(), "latents": x.contiguous(),
               "frames_u8_sub": u8[:, :, ::8, ::8].contiguous()},
              os.path.join(OUT, "rollout.safetensors"))
    print("rollout latents std", float(x.std()), "u8 mean", float(u8.float().mean()))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
