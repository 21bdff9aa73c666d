"""The CPU port (oracle/reference_port.py) against outputs of the unmodified reference
(tests/golden/, minted by oracle/make_golden.py).  fp32 vs fp32, so the bound is accumulation noise."""
import pytest
import torch

from oracle import reference_port as rp
from oracle.cases import CASES_DIT, CASES_DENOISE, CASES_VAE, ROLLOUT, seeded_rand, seeded_randn, subsample_image
from oracle.weights import DiTConfig, VAEConfig, dummy_prompt, make_dit_state, make_vae_state, w_key_actions

torch.set_num_threads(8)
_state = {}


def dit_state(depth, degenerate=False):
    key = ("dit", depth, degenerate)
    if key not in _state:
        _state[key] = make_dit_state(DiTConfig(depth=depth), seed=0, degenerate=degenerate)
    return _state[key]


def vae_state(enc, dec):
    key = ("vae", enc, dec)
    if key not in _state:
        _state[key] = make_vae_state(VAEConfig(enc_depth=enc, dec_depth=dec), seed=0)
    return _state[key]


def test_schedule_known_answers(golden):
    g = golden("schedule")
    betas = rp.sigmoid_beta_schedule(1000)
    assert betas.dtype == torch.float64
    assert torch.equal(betas, g["betas_f64"])
    abar = rp.alphas_cumprod_table()
    assert torch.equal(abar, g["alphas_cumprod_f32"])
    # KATs recorded in SURVEY.md §8(a3)
    assert abs(float(betas[0]) - 3.0024917e-4) < 1e-10
    assert abs(float(abar[15]) - 0.99499559) < 1e-7
    assert abs(float(abar[999]) - 1.0000775e-4) < 1e-9


def test_noise_levels_truncate():
    assert rp.noise_levels(100)[:4] == [0, 9, 19, 29] and rp.noise_levels(100)[-1] == 999
    assert rp.noise_levels(10) == [0, 99, 199, 299, 399, 499, 599, 699, 799, 899, 999]


@pytest.mark.parametrize("name", list(CASES_DIT))
def test_dit_forward_matches_reference(golden, name):
    c = CASES_DIT[name]
    cfg = DiTConfig(depth=c["depth"])
    x = seeded_randn((c["B"], c["T"], 16, 18, 32), c["seed"])
    g = golden("dit_forward")
    assert float(x.double().sum()) == float(g[f"{name}.x_sum"])      # RNG stream unchanged
    t = torch.tensor(c["t"]).reshape(c["B"], c["T"])
    a = w_key_actions(c["B"], c["T"]) if c["actions"] else None
    v = rp.dit_forward(dit_state(c["depth"], c["degenerate"]), cfg, x, t, a)
    ref = g[f"{name}.v"]
    assert v.shape == ref.shape
    assert float((v - ref).abs().max()) < 2e-4, float((v - ref).abs().max())


@pytest.mark.parametrize("name", list(CASES_DENOISE))
def test_denoise_step_matches_reference(golden, name):
    c = CASES_DENOISE[name]
    cfg = DiTConfig(depth=c["depth"])
    x = seeded_randn((c["B"], c["frames"], 16, 18, 32), c["seed"])
    a = w_key_actions(c["B"], c["frames"]) if c["actions"] else None
    xp, v = rp.denoise_step(dit_state(c["depth"]), cfg, x, a, c["noise_idx"], 15,
                            rp.noise_levels(c["noise_steps"]), rp.alphas_cumprod_table(), c["start_frame"])
    g = golden("denoise_step")
    assert float((v - g[f"{name}.v_pred"]).abs().max()) < 2e-4
    # x_pred divides by sqrt(1/abar - 1): tiny at t=15 for context frames, so compare relatively
    ref = g[f"{name}.x_pred"]
    assert float(((xp - ref).abs() / (1 + ref.abs())).max()) < 2e-3
    assert float((xp[:, -1] - ref[:, -1]).abs().max()) < 5e-4


@pytest.mark.parametrize("name", list(CASES_VAE))
def test_vae_matches_reference(golden, name):
    c = CASES_VAE[name]
    cfg = VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"])
    sd = vae_state(c["enc_depth"], c["dec_depth"])
    g = golden("vae")
    img = seeded_rand((c["N"], 3, 360, 640), c["seed"]) * 2 - 1
    mean = rp.vae_encode_mean(sd, cfg, img)
    assert float((mean - g[f"{name}.mean"]).abs().max()) < 5e-4
    dec = rp.vae_decode(sd, cfg, seeded_randn((c["N"], 576, 16), c["seed"] + 1))
    assert float((subsample_image(dec) - g[f"{name}.dec_sub"]).abs().max()) < 5e-4
    assert abs(float(dec.double().sum()) - float(g[f"{name}.dec_sum"])) < 1e-3 * float(g[f"{name}.dec_abs_sum"])


def test_rollout_matches_reference(golden):
    c = ROLLOUT
    dcfg = DiTConfig(depth=c["depth"])
    vcfg = VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"])
    g = golden("rollout")
    vsd = vae_state(c["enc_depth"], c["dec_depth"])
    lat = rp.encode_prompt(vsd, vcfg, dummy_prompt(5)[None, : c["n_prompt"]])
    assert float((lat - g["prompt_latents"]).abs().max()) < 5e-4
    gen = torch.Generator().manual_seed(c["seed"])
    x = rp.rollout(dit_state(c["depth"]), dcfg, g["prompt_latents"], w_key_actions(1, c["total_frames"]),
                   c["total_frames"], c["noise_steps"], lambda i: torch.randn((1, 1, 16, 18, 32), generator=gen))
    assert float((x - g["latents"]).abs().max()) < 2e-3
    u8 = rp.decode_to_uint8(vsd, vcfg, g["latents"])
    diff = (u8[:, :, ::8, ::8].int() - g["frames_u8_sub"].int()).abs()
    assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) < 0.01   # truncation flips only


def test_bf16_rounding_model_is_close_to_fp32():
    """The autocast-rounding emulation stays within the calibrated bf16-vs-fp32 gap."""
    c = CASES_DIT["d2_b1_t5_act"]
    cfg = DiTConfig(depth=2)
    x = seeded_randn((1, 5, 16, 18, 32), c["seed"])
    t = torch.tensor(c["t"]).reshape(1, 5)
    a = w_key_actions(1, 5)
    v32 = rp.dit_forward(dit_state(2), cfg, x, t, a, rp.FP32)
    v16 = rp.dit_forward(dit_state(2), cfg, x, t, a, rp.BF16)
    err = (v32 - v16).abs()
    assert 0 < float(err.max()) < 6e-2 and float(err.mean()) < 1.2e-2


def test_trainer_schedule_known_answers(golden):
    """DiffusionTrainer.register_buffers of the reference (train_dit.py:288-327): clamp_min 1e-6 schedule, integer level
    table, stabilization level - the oracle port, the product's utils and the product's inference-side trainer (its
    buffers are built on the host, so this runs without a GPU) against the golden minted from the unmodified reference."""
    g = golden("schedule_trainer")
    betas = rp.sigmoid_beta_schedule(1000, clamp_min=0.000001)
    assert torch.equal(betas, g["betas_f64"])
    from gtav_b200.utils import sigmoid_beta_schedule
    assert torch.equal(sigmoid_beta_schedule(1000, clamp_min=0.000001), g["betas_f64"])
    from gtav_b200.train_dit import DiffusionTrainer, TrainingConfig
    tr = DiffusionTrainer(TrainingConfig(), dit=None, vae=None, device=torch.device("cpu"))
    assert torch.equal(tr.alphas_cumprod.reshape(-1), g["alphas_cumprod_f32"])
    assert torch.equal(tr.alphas_cumprod_inference.reshape(-1), g["alphas_cumprod_f32"])
    assert torch.equal(tr.noise_range, g["noise_range_16"]) and torch.equal(tr.noise_range_inference, g["noise_range_16"])
    assert int(tr.stabilization_level) == int(g["noise_range_16"][1]) == 62


def test_vae_posterior_matches_reference(golden):
    """Both halves of quant_conv's output (mean | logvar), the clamp and std of DiagonalGaussianDistribution
    (reference model/vae.py:19-45) - the oracle's vae_encode_moments against the unmodified reference."""
    c = CASES_VAE["e1_d1"]
    cfg = VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"])
    g = golden("vae_posterior")
    img = seeded_rand((c["N"], 3, 360, 640), c["seed"]) * 2 - 1
    mom = rp.vae_encode_moments(vae_state(c["enc_depth"], c["dec_depth"]), cfg, img)
    assert mom.shape == (c["N"], 576, 32)
    assert float((mom - g["moments"]).abs().max()) < 5e-4
    logvar = torch.clamp(mom[..., 16:], -30.0, 20.0)
    assert float((logvar - g["logvar"]).abs().max()) < 5e-4
    assert float((torch.exp(0.5 * logvar) - g["std"]).abs().max()) < 5e-4 * float(g["std"].max())
    assert torch.equal(g["mode"], g["moments"][..., :16])


# ---- full-size goldens (oracle/make_golden_full.py): the port against the unmodified reference at BASELINE shapes ----
def test_config1_first_generated_frame_matches_reference(golden):
    """BASELINE config 1, depth-16 DiT: the first generated frame (11 window steps from the golden prompt latents) of the
    port equals the reference's; the remaining frames repeat the same code on other inputs (GPU suite runs them all)."""
    from oracle.cases import C1
    c = C1
    g = golden("c1_rollout")
    gen = torch.Generator().manual_seed(c["seed"])
    noise = torch.randn((1, 1, 16, 18, 32), generator=gen)
    cfg = DiTConfig(depth=c["depth"])
    x = rp.rollout(dit_state(c["depth"]), cfg, g["initB.prompt_latents"], None, c["n_prompt"] + 1, c["noise_steps"], lambda i: noise)
    err = float((x - g["initB.latents"][:, : c["n_prompt"] + 1]).abs().max())
    assert err < 5e-4, err


def test_config3_step_matches_reference(golden):
    from oracle.cases import C3_STEP
    c = C3_STEP
    g = golden("c3_step")
    x = seeded_randn((c["B"], c["frames"], 16, 18, 32), c["seed"])
    xp, v = rp.denoise_step(dit_state(c["depth"]), DiTConfig(depth=c["depth"]), x, g["actions"], c["noise_idx"], 15,
                            rp.noise_levels(c["noise_steps"]), rp.alphas_cumprod_table(), c["start_frame"])
    assert float((v - g["v_pred"]).abs().max()) < 3e-4
    assert float((xp - g["x_pred"]).abs().max()) < 3e-4


def test_trainer_predict_matches_reference(golden):
    """The reference's DiffusionTrainer.predict (train_dit.py:373-469) is generate.py's loop on the trainer's schedule
    (clamp_min 1e-6, stabilization_level = noise_range[1]): the port's rollout with that table reproduces its latents."""
    from oracle.cases import TRAINER
    c = TRAINER
    g = golden("trainer")
    dcfg = DiTConfig(depth=c["depth"])
    vcfg = VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"])
    abar = torch.cumprod(1.0 - rp.sigmoid_beta_schedule(1000, clamp_min=1e-6).float(), dim=0)
    stab = int(torch.linspace(0, 999, c["ddim_noise_steps"] + 1).long()[1])
    assert stab == int(g["stabilization_level"])
    gen = torch.Generator().manual_seed(c["seed_predict"])
    lat = rp.encode_prompt(vae_state(c["enc_depth"], c["dec_depth"]), vcfg, dummy_prompt(5)[None, : c["n_prompt_frames"]])
    x = rp.rollout(dit_state(c["depth"]), dcfg, lat, w_key_actions(1, c["num_frames"]), c["num_frames"],
                   c["ddim_noise_steps_inference"], lambda i: torch.randn((1, 1, 16, 18, 32), generator=gen),
                   stabilization_level=stab, abar=abar)
    assert float((x - g["predict.latents"]).abs().max()) < 5e-4
