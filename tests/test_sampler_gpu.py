"""Rollout-level parity on the GPU: the graph-captured Sampler against (a) the golden rollout of the
unmodified reference (CPU fp32), (b) the same loop written with the drop-in denoise_step (same kernels,
per-step conditioning) and (c) itself without graph capture.

Autoregressive feedback amplifies bf16 noise, so the latent tolerance vs the fp32 reference is looser than
the per-step one; the decoded frames must stay within a stated PSNR of the reference frames."""
import math

import pytest
import torch

from oracle import reference_port as rp
from oracle.cases import ROLLOUT
from oracle.weights import DiTConfig, VAEConfig, dummy_prompt, make_dit_state, make_vae_state, w_key_actions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models():
    from gtav_b200.model.dit import DiT
    from gtav_b200.model.vae import AutoencoderKL
    c = ROLLOUT
    dit = DiT(depth=c["depth"])
    dit.load_state_dict(make_dit_state(DiTConfig(depth=c["depth"]), seed=0), strict=True)
    vae = AutoencoderKL(latent_dim=16, patch_size=20, enc_dim=1024, enc_depth=c["enc_depth"], enc_heads=16, dec_dim=1024,
                        dec_depth=c["dec_depth"], dec_heads=16, input_height=360, input_width=640)
    vae.load_state_dict(make_vae_state(VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"]), seed=0), strict=True)
    return dit.cuda().eval(), vae.cuda().eval()


def rollout_noise():
    c = ROLLOUT
    g = torch.Generator().manual_seed(c["seed"])
    return torch.stack([torch.randn((1, 1, 16, 18, 32), generator=g)[:, 0] for _ in range(c["total_frames"] - c["n_prompt"])], dim=1)


def psnr(a, b):
    mse = float(((a.float() - b.float()) ** 2).mean())
    return 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


def test_rollout_matches_reference_golden(golden, models):
    from gtav_b200.sampler import Sampler
    c = ROLLOUT
    dit, vae = models
    g = golden("rollout")
    s = Sampler(dit, vae, noise_steps=c["noise_steps"])
    video = dummy_prompt(5)[None, : c["n_prompt"]].cuda()
    lat = s.encode_prompt(video)
    assert float((lat.cpu() - g["prompt_latents"]).abs().max()) < 5e-2
    acts = w_key_actions(1, c["total_frames"]).cuda()
    x = s.sample_latents(g["prompt_latents"].cuda(), acts, c["total_frames"], noise=rollout_noise().cuda())
    err = (x.cpu() - g["latents"]).abs()
    print(f"rollout latents: max-abs {float(err.max()):.4f} mean-abs {float(err.mean()):.5f}")
    assert float(err.max()) < 1.5e-1 and float(err.mean()) < 1.5e-2
    frames = s.decode_frames(x)
    assert frames.shape == (1, c["total_frames"], 360, 640, 3) and frames.dtype == torch.uint8
    p = psnr(frames[:, :, ::8, ::8].cpu(), g["frames_u8_sub"])
    print(f"decoded frames PSNR vs reference fp32 frames: {p:.1f} dB")
    assert p > 35.0


def test_sampler_equals_stepwise_dropin_loop(models):
    """Hoisted conditioning table + device-side step bookkeeping + graph replay == the reference-shaped loop
    over the drop-in denoise_step (same kernels), bit for bit."""
    from gtav_b200.sampler import Sampler
    from gtav_b200.train_dit import denoise_step
    c = ROLLOUT
    dit, _ = models
    steps, total, n_prompt = 3, 7, 2
    prompt = torch.randn((2, n_prompt, 16, 18, 32), generator=torch.Generator().manual_seed(5)).cuda()
    noise = torch.randn((2, total - n_prompt, 16, 18, 32), generator=torch.Generator().manual_seed(6)).cuda()
    acts = torch.zeros(2, total, 25, device="cuda")
    acts[0, :, 3] = 1.0
    acts[1, :, 7] = 1.0
    acts[1, 3:, 2] = 1.0                                      # actions that change along the rollout
    outs = []
    for use_graph in (True, False):
        s = Sampler(dit, None, noise_steps=steps, use_graph=use_graph, frame_cache=False)
        outs.append(s.sample_latents(prompt, acts, total, noise=noise))
        s.close()
    assert torch.equal(outs[0], outs[1])
    abar = rp.alphas_cumprod_table().cuda().reshape(-1, 1, 1, 1)
    noise_range = torch.linspace(0, 999, steps + 1)
    x = prompt.float()
    for i in range(n_prompt, total):
        x = torch.cat([x, noise[:, i - n_prompt: i - n_prompt + 1].clamp(-20, 20)], dim=1)
        start = max(0, i + 1 - dit.max_frames)
        for k in reversed(range(steps + 1)):
            xp, _ = denoise_step(dit_model=dit, x_noisy=x, actions=acts, noise_idx=k, stabilization_level=15,
                                 noise_range=noise_range, alphas_cumprod=abar, start_frame=start)
            x[:, -1:] = xp[:, -1:]
    assert torch.equal(outs[0], x), float((outs[0] - x).abs().max())


def test_frame_cache_equals_dense_sampling(models):
    """Context pass + last-frame-only steps reproduce the dense sampler (every step recomputing the whole window):
    same arithmetic per row, so the latents agree to fp32-summation-order noise of the split-K GEMM."""
    from gtav_b200.sampler import Sampler
    dit, _ = models
    steps, total, n_prompt = 4, 8, 2
    prompt = torch.randn((2, n_prompt, 16, 18, 32), generator=torch.Generator().manual_seed(15)).cuda()
    noise = torch.randn((2, total - n_prompt, 16, 18, 32), generator=torch.Generator().manual_seed(16)).cuda()
    acts = torch.zeros(2, total, 25, device="cuda")
    acts[:, :, 3] = 1.0
    acts[1, 4:, 9] = 1.0
    res = {}
    for name, kw in (("dense", dict(frame_cache=False)), ("cache", dict(frame_cache=True)),
                     ("cache_nograph", dict(frame_cache=True, use_graph=False))):
        s = Sampler(dit, None, noise_steps=steps, **kw)
        res[name] = s.sample_latents(prompt, acts, total, noise=noise)
        s.close()
    assert torch.equal(res["cache"], res["cache_nograph"])
    err = (res["cache"] - res["dense"]).abs()
    print(f"frame cache vs dense sampling: max-abs {float(err.max()):.5f} mean-abs {float(err.mean()):.6f} "
          f"equal={torch.equal(res['cache'], res['dense'])}")
    assert float(err.max()) < 3e-2 and float(err.mean()) < 2e-3


def test_sampler_without_actions_and_growing_window(models):
    """n_prompt = 1 (the --start_frame path): the window grows 2,3,4,5,5 and needs one plan per length."""
    from gtav_b200.sampler import Sampler
    dit, _ = models
    s = Sampler(dit, None, noise_steps=2)
    prompt = torch.randn((1, 1, 16, 18, 32), generator=torch.Generator().manual_seed(9)).cuda()
    noise = torch.randn((1, 6, 16, 18, 32), generator=torch.Generator().manual_seed(10))
    x = s.sample_latents(prompt, None, 7, noise=noise.cuda())
    sd = make_dit_state(DiTConfig(depth=ROLLOUT["depth"]), seed=0)
    it = iter(range(6))
    ref = rp.rollout(sd, DiTConfig(depth=ROLLOUT["depth"]), prompt.cpu(), None, 7, 2, lambda i: noise[:, next(it)][:, None], rd=rp.BF16)
    err = (x.cpu() - ref).abs()
    print(f"growing window vs bf16 oracle: max-abs {float(err.max()):.4f} mean-abs {float(err.mean()):.5f}")
    assert float(err.max()) < 1e-1 and float(err.mean()) < 1e-2


def test_frame_stream_equals_batch_generate(models):
    """Interactive frame-at-a-time generation (per-frame action, single-frame decode) == Sampler.generate on the same
    noise draws and the same action sequence: latents bit for bit, decoded frames equal up to the rounding difference of the two
    attention kernels the VAE uses at different frame counts (<= 4 LSB, mean < 0.5 LSB: bf16 carries 8 bits at pixel scale)."""
    from gtav_b200.sampler import Sampler
    dit, vae = models
    steps, total, n_prompt, B = 3, 8, 2, 2
    video = dummy_prompt(5)[None, :n_prompt].expand(B, -1, -1, -1, -1).contiguous().cuda()
    noise = torch.randn((B, total - n_prompt, 16, 18, 32), generator=torch.Generator().manual_seed(16)).cuda()
    acts = torch.zeros(B, total, 25, device="cuda")
    acts[0, :, 3] = 1.0
    acts[1, :, 5] = 1.0
    acts[:, 4:, 9] = 1.0
    s = Sampler(dit, vae, noise_steps=steps)
    frames, lat = s.generate(video, acts, total, noise=noise)
    st = s.stream(video, prompt_actions=acts[:, :n_prompt])
    for i in range(n_prompt, total):
        f, z = st.next(action=acts[:, i], noise=noise[:, i - n_prompt])
        assert torch.equal(z, lat[:, i]), f"frame {i}: streamed latent differs from the batch rollout"
        assert f.shape == (B, 360, 640, 3) and f.dtype == torch.uint8
        # same latent, but a 2-frame decode runs the mma.sync attention and the 16-frame decode the tcgen05 one (attn_mma.cu:
        # selection by size): the two round differently inside bf16, so pixels may differ by an LSB or two
        d = (f.int() - frames[:, i].int()).abs()
        if i == n_prompt:
            print(f"streamed vs batch-decoded pixels: max {int(d.max())} LSB, mean {float(d.float().mean()):.4f} LSB")
        assert int(d.max()) <= 4 and float(d.float().mean()) < 0.5, f"frame {i}: streamed pixels differ (max {int(d.max())}, mean {float(d.float().mean()):.4f})"
    assert st.frames_generated == total - n_prompt
    with pytest.raises(RuntimeError, match="action"):
        st.next()                                                   # opened with actions: every frame needs one
    # unconditioned stream, noise from the generator
    st2 = s.stream(video, generator=torch.Generator(device="cuda").manual_seed(3))
    f, z = st2.next()
    assert torch.isfinite(z).all() and f.shape == (B, 360, 640, 3)


def test_trainer_inference_methods(models):
    """DiffusionTrainer's inference side (train_dit.py:329-552 of the reference) on the drop-in modules:
    encode/decode against the CPU oracle, predict == its own literal stepwise loop, predict_noise runs the window."""
    from gtav_b200.train_dit import DiffusionTrainer, TrainingConfig
    dit, vae = models
    c = ROLLOUT
    cfg = TrainingConfig(ddim_noise_steps=16, ddim_noise_steps_inference=3, n_prompt_frames=2, use_action_conditioning=True)
    tr = DiffusionTrainer(cfg, dit, vae)
    assert int(tr.stabilization_level) == 62 and tr.noise_range_inference.tolist() == [0, 333, 666, 999]
    assert tuple(tr.alphas_cumprod.shape) == (1000, 1, 1, 1)
    video = dummy_prompt(5)[None].cuda()                                        # [1, 5, 3, 360, 640]
    loader = [dict(video=video, actions=w_key_actions(1, 5).cuda())]
    lat = tr.encode_frames(video[:, :2])
    vsd = make_vae_state(VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"]), seed=0)
    vcfg = VAEConfig(enc_depth=c["enc_depth"], dec_depth=c["dec_depth"])
    ref = rp.vae_encode_mean(vsd, vcfg, video[0, :2].cpu() * 2 - 1) * 0.07843137255
    ref = ref.reshape(1, 2, 18, 32, 16).permute(0, 1, 4, 2, 3)
    assert float((lat.float().cpu() - ref).abs().max()) < 1e-2
    pix = tr.decode_frames(lat, 2)
    assert pix.shape == (1, 2, 360, 640, 3) and pix.dtype == torch.uint8
    with pytest.raises(RuntimeError, match="num_frames"):
        tr.decode_frames(lat, 3)
    g1 = torch.Generator(device="cuda").manual_seed(11)
    g2 = torch.Generator(device="cuda").manual_seed(11)
    pa, xa = tr.predict(loader, num_frames=6, generator=g1)
    pb, xb = tr.predict(loader, num_frames=6, generator=g2, stepwise=True)
    assert pa.shape == (1, 6, 360, 640, 3)
    err = float((xa.float() - xb.float()).abs().max())
    print(f"trainer.predict: Sampler vs literal stepwise loop max-abs {err:.5f}")
    assert err <= 5e-2                                                          # frame cache + skinny GEMM vs dense window
    x_noisy, latents, v = tr.predict_noise(loader, generator=torch.Generator(device="cuda").manual_seed(12))
    assert x_noisy.shape == latents.shape == (1, 5, 16, 18, 32) and torch.isfinite(x_noisy).all() and v.shape == (1, 5, 16, 18, 32)
    with pytest.raises(NotImplementedError):
        tr.train()
