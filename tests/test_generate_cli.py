"""The generate.py drop-in CLI (reference generate.py:69-247): flag contract on CPU, a small end-to-end run on the GPU."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# flag -> (default, type) exactly as the reference declares them (generate.py:71-118)
REFERENCE_FLAGS = {
    "total_frames": 32, "dit_model_path": "checkpoints/oasis500m.pt", "vae_model_path": "checkpoints/vit-l-20.safetensors",
    "noise_steps": 100, "use_actions": False, "output_path": "video1.mp4", "start_frame": None,
}


def test_cli_keeps_the_reference_flags_and_defaults():
    from gtav_b200.generate import build_parser
    args = build_parser().parse_args([])
    for k, v in REFERENCE_FLAGS.items():
        assert getattr(args, k) == v, k
    a = build_parser().parse_args(["--total-frames", "8", "--noise_steps", "10", "--use_actions", "--start_frame", "x.jpg",
                                   "--output_path", "o.mp4", "--dit_model_path", "d", "--vae_model_path", "v"])
    assert (a.total_frames, a.noise_steps, a.use_actions, a.start_frame) == (8, 10, True, "x.jpg")
    # additive flags default to the reference's behaviour
    assert (a.rollouts, a.seed, a.random_init, a.stepwise) == (1, None, False, False)


def test_dummy_clip_is_the_dummy_dataset_fixture():
    from gtav_b200.generate import dummy_clip
    from oracle.weights import dummy_prompt
    assert torch.equal(dummy_clip(5), dummy_prompt(5))


def test_write_video_formats(tmp_path):
    from gtav_b200.generate import write_video
    frames = (torch.arange(4 * 36 * 64 * 3) % 251).to(torch.uint8).reshape(4, 36, 64, 3)
    p = str(tmp_path / "v.npy")
    write_video(p, frames)
    assert np.array_equal(np.load(p), frames.numpy())
    p = str(tmp_path / "v.pt")
    write_video(p, frames)
    assert torch.equal(torch.load(p), frames)
    p = str(tmp_path / "v.mp4")
    write_video(p, frames)
    assert os.path.getsize(p) > 0


def test_missing_checkpoint_keys_raise(tmp_path):
    """The reference prints and continues on a key mismatch (generate.py:32-38); the drop-in raises."""
    from safetensors.torch import save_file
    from gtav_b200 import generate as G
    p = str(tmp_path / "bad.safetensors")
    save_file({"not_a_key": torch.zeros(1)}, p)
    with pytest.raises(RuntimeError, match="Missing keys"):
        G.load_models(p, p, torch.device("cpu"))


@pytest.mark.gpu
def test_cli_end_to_end_sampler_equals_stepwise(tmp_path):
    """Random-init weights, 6 frames x 3 noise steps: the graph-captured sampler and the literal per-step loop
    (reference generate.py:200-220 through the drop-in denoise_step) must produce the same frames for one seed."""
    from gtav_b200.generate import main
    outs = {}
    for mode in ("sampler", "stepwise"):
        out = str(tmp_path / f"{mode}.npy")
        tj = str(tmp_path / f"{mode}.json")
        argv = ["--random_init", "--total-frames", "6", "--noise_steps", "3", "--use_actions", "--seed", "11",
                "--output_path", out, "--timing_json", tj] + (["--stepwise"] if mode == "stepwise" else [])
        torch.manual_seed(0)            # same random-init weights in both runs
        assert main(argv) == 0
        outs[mode] = np.load(out)
        rep = json.load(open(tj))
        assert rep["total_frames"] == 6 and rep["prompt_frames"] == 4 and rep["generated_frames_per_s"] > 0
    a, b = outs["sampler"].astype(np.int32), outs["stepwise"].astype(np.int32)
    assert a.shape == (6, 360, 640, 3)
    # default init leaves every DiT block an identity, so both paths see the same arithmetic except for the
    # weight-streaming GEMM's summation order on last-frame steps: allow 1 LSB on a tiny fraction of pixels
    diff = np.abs(a - b)
    assert diff.max() <= 2 and (diff > 0).mean() < 0.02, (diff.max(), (diff > 0).mean())


@pytest.mark.gpu
def test_cli_start_frame_path(tmp_path):
    """--start_frame: one prompt frame, window grows 2,3,4,5 (generate.py:135,150-161)."""
    import cv2
    from gtav_b200.generate import main
    img = (np.random.RandomState(0).rand(90, 160, 3) * 255).astype(np.uint8)
    src = str(tmp_path / "start.png")
    cv2.imwrite(src, img)
    out = str(tmp_path / "o.npy")
    torch.manual_seed(0)
    assert main(["--random_init", "--total-frames", "4", "--noise_steps", "2", "--start_frame", src, "--output_path", out,
                 "--seed", "3"]) == 0
    assert np.load(out).shape == (4, 360, 640, 3)
