"""Drop-in for the one function generate.py imports from the reference's train_dit.py:
`denoise_step` (reference train_dit.py:30-125).  The training side (DiffusionTrainer) is out of scope.

Same keyword signature and return values; the DiT call goes through whatever `dit_model` is passed
(normally gtav_b200's DiT) and the v-prediction DDIM algebra runs in the fused CUDA kernel
(`gtav_ddim_update`).  For the graph-captured whole-rollout path see sampler.py.
"""
from __future__ import annotations

import torch

try:
    from . import _native as N
except ImportError:
    import _native as N


@torch.inference_mode()
def denoise_step(dit_model, x_noisy, actions, noise_idx, stabilization_level, noise_range, alphas_cumprod,
                 start_frame=0, dtype=torch.bfloat16):
    """One DDIM step on the sliding window x_noisy[:, start_frame:].

    Context frames sit at t = stabilization_level, the last frame at noise_range[noise_idx] (truncated to
    an integer, as torch.full(dtype=long) does in the reference); returns (x_pred, v_pred) for the window,
    x_pred = x0 when noise_idx <= 0.
    """
    N.require_cuda(x_noisy, "x_noisy")
    if dtype != torch.bfloat16:
        raise RuntimeError("gtav_b200 computes in bf16 only")
    B, F_all = x_noisy.shape[:2]
    dev = x_noisy.device
    t_cur = int(noise_range[noise_idx])
    t_nxt = int(noise_range[max(0, noise_idx - 1)])
    t = torch.full((B, F_all), int(stabilization_level), dtype=torch.long, device=dev)
    t_next = t.clone()
    t[:, -1] = t_cur
    t_next[:, -1] = t_nxt
    x_curr = x_noisy[:, start_frame:].to(torch.float32).contiguous()
    t, t_next = t[:, start_frame:].contiguous(), t_next[:, start_frame:].contiguous()
    T = x_curr.shape[1]
    if actions is not None:
        actions = actions[:, start_frame:start_frame + T]

    v_pred = dit_model(x_curr, t, actions)

    abar = alphas_cumprod.to(device=dev, dtype=torch.float32).reshape(-1)
    a_t = abar[t.reshape(-1)].contiguous()
    a_n = abar[t_next].clone()
    a_n[:, :-1] = 1.0
    a_n = a_n.reshape(-1).contiguous()
    final = torch.full((1,), int(noise_idx <= 0), dtype=torch.int32, device=dev)
    v = v_pred.to(torch.bfloat16).contiguous()
    x_pred = torch.empty_like(x_curr)
    n = x_curr[0, 0].numel()
    with torch.cuda.device(dev):
        N.check(N.load().gtav_ddim_update(x_curr.data_ptr(), v.data_ptr(), x_pred.data_ptr(), B * T, n, a_t.data_ptr(),
                                          a_n.data_ptr(), final.data_ptr(), N.current_stream()), "gtav_ddim_update")
    return x_pred, v_pred
