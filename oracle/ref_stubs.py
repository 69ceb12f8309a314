"""Third-party stubs that let the reference's own sampler modules import without the packages this image lacks
(matplotlib, pytorch_lightning, diffusers, the MuJoCo evaluators).  TEST INFRASTRUCTURE (see dgdm_oracle.py): used by
tests/golden/make_golden.py (authoring container) and by oracle/ref_arm.py (bench.py --impl reference).

The only arithmetic here is ``StubDDIM``: diffusers==0.11.1 ``DDIMScheduler(num_train_timesteps,
'squaredcos_cap_v2', clip_sample=True, prediction_type='epsilon')`` with eta = 0, restated from that release's
published algorithm (requirements.txt:1; not vendored, not installable offline => "parity unpinned" against
diffusers itself; call sites generator/train.py:83, generator/diffusion.py:103,575-576).
"""
import math
import sys
import types

import numpy as np
import torch
import torch.nn as nn


class StubDDIM:
    class _Cfg:
        pass

    class _Out:
        def __init__(self, prev_sample):
            self.prev_sample = prev_sample

    def __init__(self, num_train_timesteps, beta_schedule="squaredcos_cap_v2", clip_sample=True,
                 prediction_type="epsilon"):
        assert beta_schedule == "squaredcos_cap_v2" and prediction_type == "epsilon"
        self.config = self._Cfg()
        self.config.num_train_timesteps = num_train_timesteps
        self.config.clip_sample = clip_sample
        bar = lambda s: math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
        T = num_train_timesteps
        betas = torch.tensor([min(1 - bar((i + 1) / T) / bar(i / T), 0.999) for i in range(T)], dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, T)[::-1].copy().astype(np.int64))

    def set_timesteps(self, n):
        self.num_inference_steps = n
        ratio = self.config.num_train_timesteps // n
        self.timesteps = torch.from_numpy((np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64))

    def step(self, model_output, timestep, sample):
        prev = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            x0 = torch.clamp(x0, -1, 1)
        std = 0.0 * ((1 - a_p) / (1 - a_t) * (1 - a_t / a_p)) ** 0.5
        direction = (1 - a_p - std ** 2) ** 0.5 * model_output
        return self._Out(a_p ** 0.5 * x0 + direction)


def install_stubs(ref_root, sim_test_batch=None, sim_test_batch_3d=None):
    """Make ``generator.diffusion`` & co. importable from ``ref_root``.  The two simulator entry points default to
    stubs that raise; bench.py's reference arm passes recorders instead (oracle/ref_arm.py)."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("matplotlib").pyplot = mod("matplotlib.pyplot")

    class LightningModule(nn.Module):
        @property
        def device(self):
            return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    mod("pytorch_lightning", LightningModule=LightningModule)

    class _EMA:
        def __init__(self, *a, **k):
            pass

    d = mod("diffusers", UNet2DModel=type("UNet2DModel", (), {}))
    d.schedulers = mod("diffusers.schedulers")
    mod("diffusers.schedulers.scheduling_ddim", DDIMScheduler=StubDDIM, DDIMSchedulerOutput=StubDDIM._Out)
    mod("diffusers.schedulers.scheduling_ddpm", DDPMScheduler=type("DDPMScheduler", (), {}),
        DDPMSchedulerOutput=type("DDPMSchedulerOutput", (), {}))
    mod("diffusers.training_utils", EMAModel=_EMA)

    def _no_sim(*a, **k):
        raise RuntimeError("MuJoCo evaluator is out of scope")

    # the real `dynamics` package must stay importable, only these two modules are stubbed
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import dynamics  # noqa: F401  (namespace package under ref_root)
    mod("dynamics.sim_test_mj", sim_test_batch=sim_test_batch or _no_sim)
    mod("dynamics.sim_test_mj_3d", sim_test_batch_3d=sim_test_batch_3d or _no_sim)
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            mod("wandb")
