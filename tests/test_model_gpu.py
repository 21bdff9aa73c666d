"""Model-level parity on the GPU: gtav_b200's DiT / VAE / denoise_step (CUDA kernels through the C ABI)
against (a) the golden outputs of the unmodified reference (tests/golden, CPU fp32) and (b) the CPU
oracle on the same seeded inputs, both exact-fp32 and with the autocast rounding points emulated.

Stated tolerances (floating point, bf16 compute as BASELINE.json's north_star allows), for a v-prediction
of std 0.63 / abs-max 2.9 on the non-degenerate weights:
    vs fp32 reference : max-abs <= 6e-2, mean-abs <= 1.2e-2   (measured reference-bf16 vs fp32 gap: 0.038 / 0.0070)
    vs bf16-rounding oracle: max-abs <= 4e-2, mean-abs <= 4e-3 (only accumulation-order / SDPA-internals noise left)
"""
import pytest
import torch

from oracle import reference_port as rp
from oracle.cases import CASES_DENOISE, CASES_DIT, CASES_VAE, seeded_rand, seeded_randn, subsample_image
from oracle.weights import DiTConfig, VAEConfig, make_dit_state, make_vae_state, w_key_actions

pytestmark = pytest.mark.gpu

TOL_FP32 = (6e-2, 1.2e-2)
TOL_BF16 = (4e-2, 4e-3)
_cache = {}


def dit_pair(depth, degenerate=False):
    key = ("dit", depth, degenerate)
    if key not in _cache:
        from gtav_b200.model.dit import DiT
        sd = make_dit_state(DiTConfig(depth=depth), seed=0, degenerate=degenerate)
        m = DiT(depth=depth)
        m.load_state_dict(sd, strict=True)
        _cache[key] = (sd, m.cuda().eval())
    return _cache[key]


def vae_pair(enc, dec):
    key = ("vae", enc, dec)
    if key not in _cache:
        from gtav_b200.model.vae import AutoencoderKL
        cfg = VAEConfig(enc_depth=enc, dec_depth=dec)
        sd = make_vae_state(cfg, seed=0)
        m = AutoencoderKL(latent_dim=16, patch_size=20, enc_dim=1024, enc_depth=enc, enc_heads=16, dec_dim=1024,
                          dec_depth=dec, dec_heads=16, input_height=360, input_width=640)
        m.load_state_dict(sd, strict=True)
        _cache[key] = (sd, m.cuda().eval())
    return _cache[key]


def check(out, ref, tol, what):
    err = (out.float().cpu() - ref).abs()
    mx, mean = float(err.max()), float(err.mean())
    print(f"{what}: max-abs {mx:.4f} mean-abs {mean:.5f} (ref std {float(ref.std()):.3f})")
    assert mx <= tol[0] and mean <= tol[1], (what, mx, mean)


@pytest.mark.parametrize("name", list(CASES_DIT))
def test_dit_forward(golden, name):
    c = CASES_DIT[name]
    sd, model = dit_pair(c["depth"], c["degenerate"])
    cfg = DiTConfig(depth=c["depth"])
    x = seeded_randn((c["B"], c["T"], 16, 18, 32), c["seed"])
    t = torch.tensor(c["t"]).reshape(c["B"], c["T"])
    a = w_key_actions(c["B"], c["T"]) if c["actions"] else None
    v = model(x.cuda(), t.cuda(), None if a is None else a.cuda())
    assert v.dtype == torch.bfloat16 and v.shape == x.shape
    check(v, golden("dit_forward")[f"{name}.v"], TOL_FP32, f"{name} vs reference fp32 golden")
    if c["depth"] <= 2:
        check(v, rp.dit_forward(sd, cfg, x, t, a, rp.BF16), TOL_BF16, f"{name} vs bf16-rounding oracle")


def test_dit_forward_bf16_input_and_replan():
    """bf16 latents (the dtype generate.py's first window has) and a second shape through the same module."""
    sd, model = dit_pair(2)
    cfg = DiTConfig(depth=2)
    for B, T, seed in ((1, 4, 71), (3, 2, 72), (1, 4, 73)):
        x = seeded_randn((B, T, 16, 18, 32), seed).to(torch.bfloat16)
        t = torch.randint(0, 1000, (B, T), generator=torch.Generator().manual_seed(seed))
        v = model(x.cuda(), t.cuda())
        check(v, rp.dit_forward(sd, cfg, x.float(), t, None, rp.BF16), TOL_BF16, f"bf16 input B={B} T={T}")


LAST_FRAME_MODES = {
    # name: (GTAV_SKINNY, max-abs bound vs the dense window)
    "tiled": ("0", "0", 0.0),       # same full-K tiled kernels, same arithmetic per row: bit-identical
    "tiled_splitk": ("0", "1", 2e-2),  # fc2 of the small pass on the split-K pair kernel: (k < K/2) + (k >= K/2) summation
    "skinny": ("1", "1", 2e-2),     # weight-streaming GEMM: fp32 summation order inside the K splits differs
}


@pytest.mark.parametrize("mode", list(LAST_FRAME_MODES))
def test_dit_last_frame_split_equals_dense(monkeypatch, mode):
    """Context pass + last-frame-only pass (the sampler's frame cache) == the last frame of the dense forward."""
    from gtav_b200.model.dit import DiT
    skinny, splitk, bound = LAST_FRAME_MODES[mode]
    monkeypatch.setenv("GTAV_SKINNY", skinny)
    if splitk == "0":
        monkeypatch.setenv("GTAV_GEMM_SPLITK", "0")
    sd = make_dit_state(DiTConfig(depth=2), seed=0)
    model = DiT(depth=2)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    for B, T, seed, actions in ((1, 5, 81, True), (2, 3, 82, False), (1, 2, 83, True), (4, 5, 84, True), (1, 1, 85, False)):
        x = seeded_randn((B, T, 16, 18, 32), seed).cuda()
        t = torch.randint(0, 1000, (B, T), generator=torch.Generator().manual_seed(seed)).cuda()
        a = w_key_actions(B, T).cuda() if actions else None
        dense = model(x, t, a)[:, -1:]
        split = model.forward_last_frame(x, t, a)
        again = model.forward_last_frame(x, t, a)
        err = float((dense.float() - split.float()).abs().max())
        print(f"{mode} B={B} T={T}: last-frame split vs dense max-abs {err:.5f} equal={torch.equal(dense, split)}")
        assert err <= bound, err
        assert torch.equal(split, again), "last-frame pass is not deterministic"


def test_fused_reduce_equals_separate_kernels(monkeypatch):
    """LayerNorm + modulate and the last-frame temporal attention folded into the weight-streaming GEMMs' reduce
    (GTAV_FUSE, default on) give the same bits as the stand-alone kernels (GTAV_FUSE=0): same row code, same order."""
    from gtav_b200.model.dit import DiT
    sd = make_dit_state(DiTConfig(depth=3), seed=0)
    outs = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("GTAV_FUSE", fuse)
        monkeypatch.setenv("GTAV_SKINNY", "1")
        model = DiT(depth=3)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().eval()
        res = []
        for B, T, seed, actions in ((1, 5, 91, True), (1, 3, 92, False), (1, 1, 93, False), (2, 4, 94, True), (1, 5, 95, True)):
            x = seeded_randn((B, T, 16, 18, 32), seed).cuda()
            t = torch.randint(0, 1000, (B, T), generator=torch.Generator().manual_seed(seed)).cuda()
            a = w_key_actions(B, T).cuda() if actions else None
            res.append(model.forward_last_frame(x, t, a).clone())
        outs[fuse] = res
    for i, (a, b) in enumerate(zip(outs["1"], outs["0"])):
        assert torch.isfinite(a.float()).all()
        assert torch.equal(a, b), f"case {i}: fused reduce differs from separate kernels, max-abs {float((a.float() - b.float()).abs().max())}"

@pytest.mark.gpu
def test_tagged_exchange_matches_counter_rendezvous(monkeypatch):
    """The weight-streaming GEMMs exchange their split-K partial sums tagged with the launch's parity in the last mantissa
    bit (default; gemm_skinny.cu, SkTag) instead of meeting on a counter (GTAV_SK_TAG=0).  Same sums in the same order up to
    that cleared bit of each fp32 partial - flips of single bf16 roundings that the following layers spread like any change
    of summation order (cf. test_dit_last_frame_split_equals_dense: 2e-2 at depth 2), so the bound is a few bf16 ulps.
    Repeated passes on one plan give the same bits (every pass leaves each workspace at parity 0), for 1, 2 and 3 rollouts
    and with or without the fused reduces."""
    from gtav_b200.model.dit import DiT
    sd = make_dit_state(DiTConfig(depth=3), seed=0)
    cases = ((1, 5, 191, True), (2, 4, 192, True), (3, 3, 193, False), (1, 1, 194, False))
    for fuse in ("1", "0"):
        outs = {}
        for tag in ("1", "0"):
            monkeypatch.setenv("GTAV_FUSE", fuse)
            monkeypatch.setenv("GTAV_SK_TAG", tag)
            monkeypatch.setenv("GTAV_SKINNY", "1")
            model = DiT(depth=3)
            model.load_state_dict(sd, strict=True)
            model = model.cuda().eval()
            res = []
            for B, T, seed, actions in cases:
                x = seeded_randn((B, T, 16, 18, 32), seed).cuda()
                t = torch.randint(0, 1000, (B, T), generator=torch.Generator().manual_seed(seed)).cuda()
                a = w_key_actions(B, T).cuda() if actions else None
                first = model.forward_last_frame(x, t, a).clone()
                for _ in range(2):                                   # the same plan again: parity state must be back at its start
                    assert torch.equal(model.forward_last_frame(x, t, a), first)
                res.append(first)
            outs[tag] = res
        for i, (a, b) in enumerate(zip(outs["1"], outs["0"])):
            assert torch.isfinite(a.float()).all()
            d = (a.float() - b.float()).abs()
            print(f"tagged vs counter exchange (fuse={fuse}, case {i}): max-abs {float(d.max()):.5f} mean-abs {float(d.mean()):.6f} "
                  f"differing {float((d > 0).float().mean()):.4f}")
            assert float(d.max()) <= 4e-2 and float(d.mean()) <= 3e-3



def test_last_frame_pass_full_depth_vs_reference_golden(golden):
    """The frame-cache split (context pass + weight-streaming last-frame pass) on the real 16-block DiT (B = 1, T = 5): last
    frame of the v-prediction against the unmodified reference's fp32 output, same tolerance as the dense forward."""
    from gtav_b200.model.dit import DiT
    c = CASES_DIT["d16_b1_t5_act"]
    sd = make_dit_state(DiTConfig(depth=16), seed=0)
    model = DiT(depth=16)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    x = seeded_randn((1, 5, 16, 18, 32), c["seed"])
    t = torch.tensor(c["t"]).reshape(1, 5)
    a = w_key_actions(1, 5)
    v_last = model.forward_last_frame(x.cuda(), t.cuda(), a.cuda())
    check(v_last, golden("dit_forward")["d16_b1_t5_act.v"][:, -1:], TOL_FP32, "last-frame pass, depth 16, vs reference fp32 golden")


def test_dit_rejects_cpu_and_bad_shapes():
    _, model = dit_pair(2)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(1, 1, 16, 18, 32), torch.zeros(1, 1, dtype=torch.long))
    with pytest.raises(AssertionError):
        model(torch.zeros(1, 1, 16, 20, 32, device="cuda"), torch.zeros(1, 1, dtype=torch.long, device="cuda"))
    with pytest.raises(RuntimeError, match="max_frames"):
        model(torch.zeros(1, 6, 16, 18, 32, device="cuda"), torch.zeros(1, 6, dtype=torch.long, device="cuda"))


@pytest.mark.parametrize("name", list(CASES_DENOISE))
def test_denoise_step(golden, name):
    from gtav_b200.train_dit import denoise_step
    c = CASES_DENOISE[name]
    sd, model = dit_pair(c["depth"])
    x = seeded_randn((c["B"], c["frames"], 16, 18, 32), c["seed"])
    a = w_key_actions(c["B"], c["frames"]) if c["actions"] else None
    abar = rp.alphas_cumprod_table()
    xp, v = denoise_step(dit_model=model, x_noisy=x.cuda(), actions=None if a is None else a.cuda(),
                         noise_idx=c["noise_idx"], stabilization_level=15,
                         noise_range=torch.linspace(0, 999, c["noise_steps"] + 1),
                         alphas_cumprod=abar.reshape(-1, 1, 1, 1).cuda(), start_frame=c["start_frame"],
                         dtype=torch.bfloat16)
    g = golden("denoise_step")
    check(v, g[f"{name}.v_pred"], TOL_FP32, f"{name} v_pred")
    # x_pred of the last frame: error of v scaled by the DDIM coefficients (<= 1 in magnitude here)
    check(xp[:, -1], g[f"{name}.x_pred"][:, -1], TOL_FP32, f"{name} x_pred[last]")
    # exact algebra check: recompute the update from OUR v with the oracle's formula
    T = xp.shape[1]
    t = torch.full((c["B"], T), 15)
    tn = t.clone()
    lv = rp.noise_levels(c["noise_steps"])
    t[:, -1], tn[:, -1] = lv[c["noise_idx"]], lv[max(0, c["noise_idx"] - 1)]
    a_t = abar[t].view(c["B"], T, 1, 1, 1)
    a_n = abar[tn].view(c["B"], T, 1, 1, 1).clone()
    a_n[:, :-1] = 1.0
    ref = rp.ddim_update(x[:, c["start_frame"]:], v.float().cpu(), a_t, a_n, c["noise_idx"] <= 0)
    assert float(((xp.cpu() - ref).abs() / (1 + ref.abs())).max()) < 1e-5


@pytest.mark.parametrize("name", list(CASES_VAE))
def test_vae(golden, name):
    c = CASES_VAE[name]
    sd, vae = vae_pair(c["enc_depth"], c["dec_depth"])
    g = golden("vae")
    img = seeded_rand((c["N"], 3, 360, 640), c["seed"]) * 2 - 1
    mean = vae.encode(img.cuda()).mean
    assert mean.dtype == torch.bfloat16 and mean.shape == (c["N"], 576, 16)
    # latents have std 1.4; bf16 through up to 6 blocks
    check(mean, g[f"{name}.mean"], (8e-2, 1e-2), f"{name} encode mean vs reference fp32")      # measured 0.051
    z = seeded_randn((c["N"], 576, 16), c["seed"] + 1)
    dec = vae.decode(z.cuda())
    assert dec.dtype == torch.bfloat16 and dec.shape == (c["N"], 3, 360, 640)
    check(subsample_image(dec), g[f"{name}.dec_sub"], (6e-2, 1e-2), f"{name} decode vs reference fp32")   # measured 0.037
    if c["enc_depth"] == 1:
        cfg = VAEConfig(enc_depth=1, dec_depth=1)
        check(mean, rp.vae_encode_mean(sd, cfg, img, rp.BF16), (6e-2, 6e-3), f"{name} encode vs bf16 oracle")
        check(dec, rp.vae_decode(sd, cfg, z, rp.BF16), (6e-2, 6e-3), f"{name} decode vs bf16 oracle")


def test_vae_decode_uint8_matches_pixel_epilogue():
    """Fused pixel epilogue == (decode + 1)/2 * 255 clamp truncate applied to our own bf16 decode (exact)."""
    _, vae = vae_pair(1, 1)
    z = seeded_randn((2, 576, 16), 91).cuda() * 0.078
    dec = vae.decode(z, divisor=0.07843137255)
    u8 = vae.decode(z, divisor=0.07843137255, to_uint8=True)
    ref = torch.clamp(((dec + 1) / 2) * 255, 0, 255).byte().permute(0, 2, 3, 1)
    assert u8.shape == (2, 360, 640, 3) and torch.equal(u8, ref)


def test_vae_posterior_moments(golden):
    """`vae.encode(x)` returns the whole DiagonalGaussianDistribution of the reference (model/vae.py:19-45): mean AND
    logvar (clamped), std / var, sample(), mode(); both halves against the CPU oracle's quant_conv output and against the
    golden minted from the unmodified reference (fp32)."""
    sd, vae = vae_pair(1, 1)
    cfg = VAEConfig(enc_depth=1, dec_depth=1)
    c = CASES_VAE["e1_d1"]
    gp = golden("vae_posterior")
    pg = vae.encode((seeded_rand((c["N"], 3, 360, 640), c["seed"]) * 2 - 1).cuda())
    check(pg.mean, gp["moments"][..., :16], (8e-2, 1e-2), "posterior mean vs reference fp32")
    check(pg.logvar, gp["logvar"], (8e-2, 1e-2), "posterior logvar vs reference fp32")
    img = seeded_rand((2, 3, 360, 640), 97) * 2 - 1
    post = vae.encode(img.cuda())
    mom = rp.vae_encode_moments(sd, cfg, img, rp.BF16)
    check(post.mean, mom[..., :16], (6e-2, 6e-3), "posterior mean vs bf16 oracle")
    check(post.logvar, torch.clamp(mom[..., 16:], -30.0, 20.0), (6e-2, 6e-3), "posterior logvar vs bf16 oracle")
    assert torch.equal(post.mean, vae.encode_mean(img.cuda()).to(torch.bfloat16))       # same kernels as the mean-only path
    assert torch.equal(post.mode(), post.mean) and post.parameters.shape == (2, 576, 32)
    assert torch.allclose(post.std.float(), torch.exp(0.5 * post.logvar.float()), rtol=1e-2)
    assert torch.allclose(post.var.float(), torch.exp(post.logvar.float()), rtol=2e-2)
    torch.manual_seed(0)
    s1 = post.sample()
    assert s1.shape == post.mean.shape and not torch.equal(s1, post.mean)
    z = (s1.float() - post.mean.float()) / post.std.float()
    assert abs(float(z.mean())) < 0.05 and abs(float(z.std()) - 1.0) < 0.05          # mean + std * N(0, 1)
    rec, post2, zz = vae.autoencode(img.cuda(), sample_posterior=False)
    assert rec.shape == (2, 3, 360, 640) and torch.equal(zz, post2.mean)


def test_models_built_and_run_under_inference_mode():
    """The reference's generate.py builds, loads and moves both models under @torch.inference_mode (load_models / main,
    generate.py:28,69): parameters are then inference tensors (no version counter).  Same results as modules built the
    normal way; a later load_state_dict re-packs the weights, and a Sampler created before it rebuilds its graphs."""
    from gtav_b200.model.dit import DiT
    from gtav_b200.model.vae import AutoencoderKL
    from gtav_b200.sampler import Sampler
    sd = make_dit_state(DiTConfig(depth=2), seed=0)
    vsd = make_vae_state(VAEConfig(enc_depth=1, dec_depth=1), seed=0)
    x = seeded_randn((1, 3, 16, 18, 32), 61).cuda()
    t = torch.tensor([[15, 15, 700]]).cuda()
    z = seeded_randn((1, 576, 16), 62).cuda()
    _, ref_dit = dit_pair(2)
    _, ref_vae = vae_pair(1, 1)
    v_ref, d_ref = ref_dit(x, t), ref_vae.decode(z)
    with torch.inference_mode():
        dit = DiT(depth=2)
        dit.load_state_dict(sd, strict=True)
        vae = AutoencoderKL(latent_dim=16, patch_size=20, enc_dim=1024, enc_depth=1, enc_heads=16, dec_dim=1024, dec_depth=1,
                            dec_heads=16, input_height=360, input_width=640)
        vae.load_state_dict(vsd, strict=True)
        dit, vae = dit.cuda().eval(), vae.cuda().eval()
        assert all(p.is_inference() for p in dit.parameters())
        assert torch.equal(dit(x, t), v_ref) and torch.equal(vae.decode(z), d_ref)
        # a Sampler that has run, then new weights: the packed copies, plans and captured graphs are rebuilt, not reused
        s = Sampler(dit, None, noise_steps=2)
        prompt = x[:, :2].float()
        noise = seeded_randn((1, 1, 16, 18, 32), 63).cuda()
        a = s.sample_latents(prompt, None, 3, noise=noise)
        sd2 = make_dit_state(DiTConfig(depth=2), seed=5)
        dit.load_state_dict(sd2, strict=True)
        b = s.sample_latents(prompt, None, 3, noise=noise)
        fresh = DiT(depth=2)
        fresh.load_state_dict(sd2, strict=True)
        fresh = fresh.cuda().eval()
        c = Sampler(fresh, None, noise_steps=2).sample_latents(prompt, None, 3, noise=noise)
        assert not torch.equal(a, b), "the sampler kept replaying graphs built on the old weights"
        assert torch.equal(b, c)
