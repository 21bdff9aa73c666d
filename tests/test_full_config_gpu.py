"""Parity at the BASELINE configs themselves (full depth, full step count), against golden vectors minted by running the
UNMODIFIED reference on the CPU in fp32 (oracle/make_golden_full.py -> tests/golden/{c1_rollout,c3_step,trainer}.safetensors):

  * config 1 exactly - generate.py's loop, depth-16 DiT + VAE 6/12, dummy prompt, 8 frames, 10 DDIM steps, fixed noise -
    for the non-degenerate weights and for zero adaLN linears (every block an identity, like the default init);
  * one denoise_step at config 3's shape (B = 8 action-conditioned rollouts, M = 5760 rows) and the same step through the
    frame-cache split (M = 1152 last-frame rows);
  * DiffusionTrainer.predict / predict_noise of the reference (train_dit.py:373-552).

Stated tolerances (floating point, bf16 compute vs the fp32 reference).  They are calibrated on the bf16 model of the
reference's own autocast graph (oracle/reference_port.py, Rounding BF16; numbers in tests/golden/full_meta.json):
on config 1 that model sits at max-abs 0.021 per generated frame / 44.9 dB PSNR from the fp32 reference, so
    latents  : max-abs <= 0.08 per generated frame (4 x the bf16 model's own distance), mean-abs <= 0.012
    frames   : PSNR(product uint8, reference fp32 uint8) >= 44.9 - 3 = 41.9 dB   
"""
import json
import math
import os

import pytest
import torch

from oracle.cases import C1, C3_STEP, TRAINER, seeded_randn
from oracle.weights import DiTConfig, VAEConfig, dummy_prompt, make_dit_state, make_vae_state, w_key_actions

pytestmark = pytest.mark.gpu
META = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "full_meta.json")))


def psnr(a, b):
    mse = float(((a.float() - b.float()) ** 2).mean())
    return 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


@pytest.fixture(scope="module")
def vae_full():
    from gtav_b200.model.vae import VAE_models
    vae = VAE_models["vit-l-20-shallow-encoder"]()
    vae.load_state_dict(make_vae_state(VAEConfig(), seed=0), strict=True)
    return vae.cuda().eval()


def dit_full(degenerate=False):
    from gtav_b200.model.dit import DiT_models
    dit = DiT_models["DiT-S/2"]()
    dit.load_state_dict(make_dit_state(DiTConfig(), seed=0, degenerate=degenerate), strict=True)
    return dit.cuda().eval()


@pytest.mark.parametrize("tag", ["initB", "initA"])
def test_config1_rollout_matches_reference(golden, vae_full, tag):
    """generate.py:186-244 at BASELINE config 1: encode the dummy prompt, 4 generated frames x 11 DiT steps through the
    graph-captured frame-cache Sampler (the shipped path), decode to uint8 - vs the reference's fp32 run."""
    from gtav_b200.sampler import Sampler
    c = C1
    g = golden("c1_rollout")
    dit = dit_full(degenerate=(tag == "initA"))
    s = Sampler(dit, vae_full, noise_steps=c["noise_steps"])
    gen = torch.Generator().manual_seed(c["seed"])
    noise = torch.stack([torch.randn((1, 1, 16, 18, 32), generator=gen)[:, 0] for _ in range(c["total_frames"] - c["n_prompt"])], dim=1)
    video = dummy_prompt(5)[None, : c["n_prompt"]].cuda()
    frames, lat = s.generate(video, None, c["total_frames"], noise=noise.cuda())
    ref_lat, ref_u8 = g[f"{tag}.latents"], g[f"{tag}.frames_u8_sub"]
    err = (lat.float().cpu() - ref_lat).abs()
    per_frame = [float(err[:, f].max()) for f in range(c["total_frames"])]
    p_all = psnr(frames[:, :, ::8, ::8].cpu(), ref_u8)
    p_gen = psnr(frames[:, c["n_prompt"]:, ::8, ::8].cpu(), ref_u8[:, c["n_prompt"]:])
    cal = META[f"c1_{tag}"]["bf16_model"]
    print(f"C1 {tag}: latent max-abs per frame {[round(v, 4) for v in per_frame]} (bf16 model of the reference: "
          f"{cal['latent_max_abs_per_frame']}), mean-abs {float(err[:, c['n_prompt']:].mean()):.5f}; PSNR {p_all:.2f} dB "
          f"(generated frames {p_gen:.2f}; bf16 model {cal['psnr_db']} / {cal['psnr_generated_db']})")
    assert max(per_frame[: c["n_prompt"]]) < 2e-2                  # VAE-encoded prompt latents
    assert max(per_frame[c["n_prompt"]:]) < 8e-2 and float(err[:, c["n_prompt"]:].mean()) < 1.2e-2
    assert p_all >= cal["psnr_db"] - 3.0 and p_gen >= cal["psnr_generated_db"] - 3.0
    s.close()


def test_config3_step_matches_reference(golden):
    """One DDIM step of 8 action-conditioned rollouts over the full window (CTA-pair / tiled GEMMs at M = 5760) and the
    same step's last frame through the context pass + last-frame split (M = 1152 rows), vs reference denoise_step."""
    from gtav_b200.train_dit import denoise_step
    from oracle import reference_port as rp
    c = C3_STEP
    g = golden("c3_step")
    dit = dit_full()
    x = seeded_randn((c["B"], c["frames"], 16, 18, 32), c["seed"]).cuda()
    a = g["actions"].cuda()
    abar = rp.alphas_cumprod_table()
    xp, v = denoise_step(dit_model=dit, x_noisy=x, actions=a, noise_idx=c["noise_idx"], stabilization_level=15,
                         noise_range=torch.linspace(0, 999, c["noise_steps"] + 1),
                         alphas_cumprod=abar.reshape(-1, 1, 1, 1).cuda(), start_frame=c["start_frame"])
    ev = (v.float().cpu() - g["v_pred"]).abs()
    t_last = rp.noise_levels(c["noise_steps"])[c["noise_idx"]]
    ex = (xp[:, -1].cpu() - g["x_pred"][:, -1]).abs()
    print(f"C3 step: v max-abs {float(ev.max()):.4f} mean-abs {float(ev.mean()):.5f}; x_pred(last) max-abs {float(ex.max()):.4f}")
    assert float(ev.max()) < 6e-2 and float(ev.mean()) < 1.2e-2
    assert float(ex.max()) < 6e-2 * math.sqrt(1 - float(abar[t_last])) + 1e-3
    t = torch.full((c["B"], c["frames"]), 15, dtype=torch.long)
    t[:, -1] = t_last
    v_last = dit.forward_last_frame(x, t.cuda(), a)
    el = (v_last.float().cpu() - g["v_pred"][:, -1:]).abs()
    print(f"C3 step, frame-cache split (M = {144 * c['B']} rows): v max-abs {float(el.max()):.4f} mean-abs {float(el.mean()):.5f}")
    assert float(el.max()) < 6e-2 and float(el.mean()) < 1.2e-2


@pytest.fixture(scope="module")
def trainer():
    from gtav_b200.model.dit import DiT
    from gtav_b200.model.vae import AutoencoderKL
    from gtav_b200.train_dit import DiffusionTrainer, TrainingConfig
    c = TRAINER
    dit = DiT(depth=c["depth"])
    dit.load_state_dict(make_dit_state(DiTConfig(depth=c["depth"]), seed=0), strict=True)
    vae = AutoencoderKL(latent_dim=16, patch_size=20, enc_dim=1024, enc_depth=c["enc_depth"], enc_heads=16, dec_dim=1024,
                        dec_depth=c["dec_depth"], dec_heads=16, input_height=360, input_width=640)
    vae.load_state_dict(make_vae_state(VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"]), seed=0), strict=True)
    cfg = TrainingConfig(ddim_noise_steps=c["ddim_noise_steps"], ddim_noise_steps_inference=c["ddim_noise_steps_inference"],
                         n_prompt_frames=c["n_prompt_frames"], noise_abs_max=c["noise_abs_max"], use_action_conditioning=True)
    return DiffusionTrainer(cfg, dit.cuda().eval(), vae.cuda().eval())


@pytest.mark.parametrize("stepwise", [False, True])
def test_trainer_predict_matches_reference(golden, trainer, stepwise):
    """DiffusionTrainer.predict (train_dit.py:373-469) - graph-captured Sampler and the literal per-step loop - vs the
    latents / frames the reference's own trainer produced from the same global-RNG noise draws."""
    c = TRAINER
    g = golden("trainer")
    assert int(trainer.stabilization_level) == int(g["stabilization_level"])
    assert trainer.noise_range_inference.tolist() == g["noise_range_inference"].tolist()
    gen = torch.Generator().manual_seed(c["seed_predict"])                     # == torch.manual_seed + global torch.randn draws
    noise = torch.stack([torch.randn((1, 1, 16, 18, 32), generator=gen)[:, 0] for _ in range(c["num_frames"] - c["n_prompt_frames"])], dim=1)
    loader = [dict(video=dummy_prompt(5)[None].cuda(), actions=w_key_actions(1, 5).cuda())]
    pix, lat = trainer.predict(loader, num_frames=c["num_frames"], noise=noise.cuda(), stepwise=stepwise)
    err = (lat.float().cpu() - g["predict.latents"]).abs()
    p = psnr(pix[:, :, ::8, ::8].cpu(), g["predict.frames_u8_sub"])
    print(f"trainer.predict(stepwise={stepwise}): latents max-abs {float(err.max()):.4f} mean-abs {float(err.mean()):.5f}, PSNR {p:.1f} dB")
    assert float(err.max()) < 8e-2 and float(err.mean()) < 1.2e-2
    assert p >= 38.0


def test_trainer_predict_noise_matches_reference(golden, trainer):
    """DiffusionTrainer.predict_noise (train_dit.py:471-552): context frames noised to stabilization_level - 1, last frame
    replaced by clamped noise and denoised; vs what the reference handed to its visualisation at the final step."""
    g = golden("trainer")
    loader = [dict(video=dummy_prompt(5)[None].cuda(), actions=w_key_actions(1, 5).cuda())]
    x_noisy, latents, v = trainer.predict_noise(loader, noise=g["predict_noise.noise"].cuda())
    e_lat = float((latents.float().cpu() - g["predict_noise.latents"]).abs().max())
    e_ctx = float((x_noisy[:, :-1].float().cpu() - g["predict_noise.x_noisy_in"][:, :-1]).abs().max())
    e_x = float((x_noisy[:, -1].float().cpu() - g["predict_noise.x_pred"][:, -1]).abs().max())
    e_v = (v.float().cpu() - g["predict_noise.v_pred"]).abs()
    print(f"trainer.predict_noise: latents {e_lat:.4f}, noised context {e_ctx:.4f}, final frame {e_x:.4f}, v max-abs {float(e_v.max()):.4f}")
    assert e_lat < 2e-2 and e_ctx < 2e-2
    assert e_x < 8e-2 and float(e_v.max()) < 8e-2 and float(e_v.mean()) < 1.2e-2
