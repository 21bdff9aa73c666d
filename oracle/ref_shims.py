"""Import the UNMODIFIED reference modules from /root/reference in this container.  TEST INFRASTRUCTURE.

/root/reference does not exist on the GPU box, so nothing on the gpu-test / smoke / bench path
imports this file; it is used only by oracle/make_golden.py (run here, outputs committed under
tests/golden/) and by the optional CPU test that re-validates the port against the live reference.

timm, diffusers, accelerate, webdataset, matplotlib are not installed and have no wheel offline;
the reference imports them at module scope (model/dit.py:14-15, model/embeddings.py:11,
train_dit.py:9-25), so minimal stand-ins are registered in sys.modules first.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GTAV_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "dit.py"))


def _module(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__spec__ = None
    sys.modules[name] = m
    return m


def install():
    """Register the stand-ins and put the reference on sys.path.  Idempotent."""
    if getattr(install, "_done", False):
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch
    from torch import nn

    class Mlp(nn.Module):
        """Stand-in for timm.models.vision_transformer.Mlp: fc1 -> act -> fc2 (dropout p=0)."""

        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0, **_):
            super().__init__()
            self.fc1 = nn.Linear(in_features, hidden_features or in_features)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    if "timm" not in sys.modules:
        _module("timm"); _module("timm.models"); _module("timm.layers")
        _module("timm.models.vision_transformer", Mlp=Mlp)
        _module("timm.layers.helpers", to_2tuple=to_2tuple)
    if "diffusers" not in sys.modules:
        _module("diffusers"); _module("diffusers.models")
        _module("diffusers.models.embeddings", TimestepEmbedding=type("TimestepEmbedding", (nn.Module,), {}))

    # train_dit.py extras.  transformers.optimization must be imported BEFORE a fake `accelerate`
    # exists (its availability probe chokes on a module whose __spec__ is None).
    try:
        import transformers.optimization  # noqa: F401
    except Exception:
        _module("transformers"); _module("transformers.optimization",
                                         get_cosine_with_min_lr_schedule_with_warmup=None)
    if "accelerate" not in sys.modules:
        _module("accelerate", Accelerator=object, DistributedDataParallelKwargs=object)
        _module("accelerate.logging", get_logger=lambda *a, **k: None)
        _module("accelerate.utils", set_seed=lambda *a, **k: None)
    for name in ("webdataset", "matplotlib", "matplotlib.pyplot", "wandb", "huggingface_hub", "datasets"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _module(name, load_dataset=None)
    import torchvision.io
    if not hasattr(torchvision.io, "write_video"):
        torchvision.io.write_video = lambda *a, **k: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    install._done = True


def load():
    """Returns a namespace with the reference's DiT / VAE classes, denoise_step and schedule."""
    install()
    saved = {k: sys.modules.pop(k) for k in ("model", "model.dit", "model.vae", "utils", "train_dit")
             if k in sys.modules and not getattr(sys.modules[k], "__file__", "").startswith(REFERENCE_ROOT)}
    try:
        import model.dit as rdit
        import model.vae as rvae
        import utils as rutils
        import train_dit as rtrain
    finally:
        # leave the reference modules importable only through the namespace we return
        for k in ("model", "model.dit", "model.vae", "model.attention", "model.embeddings",
                  "model.rotary_embedding_torch", "utils", "train_dit", "dummy_dataset", "hf_dataset",
                  "web_dataset"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    return types.SimpleNamespace(DiT=rdit.DiT, AutoencoderKL=rvae.AutoencoderKL,
                                 denoise_step=rtrain.denoise_step,
                                 sigmoid_beta_schedule=rutils.sigmoid_beta_schedule,
                                 train_module=rtrain)
