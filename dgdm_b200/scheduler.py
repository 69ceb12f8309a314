"""Host mirror of the noise schedule the reference takes from diffusers==0.11.1 ``DDIMScheduler``
(constructed at generator/train.py:83; used at generator/diffusion.py:103,571,575-576).

Only the schedule (a 15-entry table) and the per-step scalar coefficients live on the host; the update
itself is the fused CUDA kernel K4 (``dgdm_ddim_guided_update``).  Same attribute names as diffusers so
code written against the reference reads the same: ``config.num_train_timesteps``, ``alphas_cumprod``,
``timesteps``, ``set_timesteps``, ``step(...).prev_sample``.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib


class DDIMScheduler:
    def __init__(self, num_train_timesteps: int = 15, beta_schedule: str = "squaredcos_cap_v2",
                 clip_sample: bool = True, prediction_type: str = "epsilon"):
        if beta_schedule != "squaredcos_cap_v2":
            raise ValueError("only the reference's 'squaredcos_cap_v2' schedule is supported (train.py:83)")
        if prediction_type != "epsilon":
            raise ValueError("only prediction_type='epsilon' is supported (train.py:83)")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_schedule=beta_schedule,
                                      clip_sample=clip_sample, prediction_type=prediction_type)
        T = num_train_timesteps
        bar = lambda s: math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
        betas = torch.tensor([min(1 - bar((i + 1) / T) / bar(i / T), 0.999) for i in range(T)], dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)          # fp32, CPU
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, T)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int) -> None:
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def coefficients(self, t: int) -> Tuple[float, float, float, float]:
        """(sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev)) evaluated in fp32 like the reference's
        0-dim tensor arithmetic; a_prev := 1 below t = 0 (set_alpha_to_one)."""
        t = int(t)
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return (float((1 - a_t) ** 0.5), float(a_t ** 0.5), float(a_p ** 0.5), float((1 - a_p) ** 0.5))

    def guided_step(self, eps: torch.Tensor, t: int, sample: torch.Tensor, grad: Optional[torch.Tensor],
                    scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """generator/diffusion.py:575-576 in one kernel: eps_hat = eps - sqrt(1-a_t)*grad*scale; DDIM(eta=0)."""
        c = self.coefficients(t)
        out = torch.empty_like(sample) if out is None else out
        l = _lib.lib()
        # launch in the context (and on the current stream) of the device that owns the tensors, whatever the
        # process-wide current device is
        with torch.cuda.device(sample.device):
            _lib.check(l.dgdm_ddim_guided_update(_lib.ptr(out), _lib.ptr(sample), _lib.ptr(eps), _lib.ptr(grad),
                                                 sample.numel(), c[0], c[1], c[2], c[3], float(scale),
                                                 int(bool(self.config.clip_sample)), _lib.stream_ptr(sample.device)),
                       "dgdm_ddim_guided_update")
        return out

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor):
        """diffusers-style ``scheduler.step(eps, t, x).prev_sample``."""
        return SimpleNamespace(prev_sample=self.guided_step(model_output.contiguous(), int(timestep),
                                                            sample.contiguous(), None, 0.0))
