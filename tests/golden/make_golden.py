#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REAL reference code.

Runs only in the authoring container, where /root/reference exists (it does not exist on the GPU
box; nothing in tests/ reads it at run time).  The reference has no tests or golden vectors of its
own, so these fixtures are the pin: the reference's real modules
    generator/diffusion.py      (Diffusion.cond_fn, deltas_to_objective, get_convergence_centers)
    generator/diffusion_utils.py (ConditionalUnet1D)
    dynamics/profile_forward_2d.py, profile_forward_3d.py, models/pointnet2*.py
are imported unmodified, with only third-party packages that are not installed here stubbed
(matplotlib, pytorch_lightning, diffusers, the MuJoCo evaluators).  The DDIM scheduler stub is a
restatement of diffusers==0.11.1 (not installable offline) -- that arithmetic is "parity unpinned".

Weights and inputs come from dgdm_b200.synthetic (numpy RandomState => reproducible anywhere), and
are loaded into the reference modules with strict=True, which also validates the name/shape tables.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)


# ---------------------------------------------------------------------------------------------
# third-party stubs
# ---------------------------------------------------------------------------------------------
class _StubDDIM:
    """diffusers==0.11.1 DDIMScheduler(num_train_timesteps, 'squaredcos_cap_v2', clip_sample=True,
    prediction_type='epsilon'), eta=0 -- restated, see module docstring."""

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


def install_stubs():
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
    mod("diffusers.schedulers.scheduling_ddim", DDIMScheduler=_StubDDIM, DDIMSchedulerOutput=_StubDDIM._Out)
    mod("diffusers.schedulers.scheduling_ddpm", DDPMScheduler=type("DDPMScheduler", (), {}),
        DDPMSchedulerOutput=type("DDPMSchedulerOutput", (), {}))
    mod("diffusers.training_utils", EMAModel=_EMA)

    def _no_sim(*a, **k):
        raise RuntimeError("MuJoCo evaluator is out of scope")

    # the real `dynamics` package must stay importable, only these two modules are stubbed
    sys.path.insert(0, REF)
    import dynamics  # noqa: F401  (namespace package under /root/reference)
    mod("dynamics.sim_test_mj", sim_test_batch=_no_sim)
    mod("dynamics.sim_test_mj_3d", sim_test_batch_3d=_no_sim)
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            mod("wandb")


# ---------------------------------------------------------------------------------------------
def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    install_stubs()
    import warnings
    warnings.filterwarnings("ignore")
    from generator.diffusion import Diffusion
    from generator.diffusion_utils import ConditionalUnet1D
    from dynamics.profile_forward_2d import ProfileForward2DModel
    from dynamics.profile_forward_3d import ProfileForward3DModel
    from dynamics.models import pointnet2_utils
    from dgdm_b200 import synthetic as syn

    T, NINF = 15, 5

    def build(mode, B, grid, npos, objs, sub_bs=16, seed=0):
        P = 14 if mode == "point" else 42
        unet = ConditionalUnet1D(input_dim=1, global_cond_dim=0, down_dims=[128, 256], diffusion_step_embed_dim=32)
        unet.load_state_dict(syn.unet1d_state_dict(seed), strict=True)
        if mode == "point":
            net = nn.DataParallel(ProfileForward2DModel(output_ch=3, params_ch=P, object_ch=200))
            net.load_state_dict(syn.dynamics2d_state_dict(seed), strict=True)
        else:
            net = nn.DataParallel(ProfileForward3DModel(output_ch=3, params_ch=P))
            net.load_state_dict(syn.dynamics3d_state_dict(seed), strict=True)
        for p in net.parameters():
            p.requires_grad = False
        sched = _StubDDIM(num_train_timesteps=T)
        dm = Diffusion(noise_pred_net=unet, noise_scheduler=sched, num_inference_steps=NINF, mode=mode, input_dim=1,
                       num_points=P, class_cond=True, classifier_model=net, grid_size=grid, num_pos=npos,
                       object_vertices=objs, object_ids=list(range(len(objs))), sub_batch_size=sub_bs, seed=seed)
        dm.eval()     # Lightning's validate() does this => BatchNorm uses running stats
        return dm

    # DataParallel on a CPU-only box just calls .module (no device_ids) -- same arithmetic.

    # ------------------------------------------------------------------ 2D
    B, grid, npos = 4, 6, 2
    objs2 = syn.objects_2d(2)
    dm = build("point", B, grid, npos, objs2)
    noise = syn.initial_noise(B, 14, seed=0)
    out = {"noise": noise.numpy(), "objects": objs2.numpy(), "grid_size": grid, "num_pos": npos,
           "alphas_cumprod": dm.noise_scheduler.alphas_cumprod.numpy(),
           "timesteps": dm.noise_scheduler.timesteps.numpy()}
    objectives = ["rotate", "rotate_clockwise", "clockwise_up", "shift_left", "counterclockwise_right"]
    for oi in range(2):
        for t in (12, 3):
            ts = t * torch.ones((B,), dtype=torch.int64)
            for name in objectives:
                g = dm.cond_fn(noise, ts, opt_obj=name, object_vertices=objs2[oi], ori_range=[-1.0, 1.0])
                out[f"grad_o{oi}_t{t}_{name}"] = g.detach().numpy()
        # narrower orientation range as well
        g = dm.cond_fn(noise, 6 * torch.ones((B,), dtype=torch.int64), opt_obj="rotate", object_vertices=objs2[oi],
                       ori_range=[-0.5, 0.25])
        out[f"grad_o{oi}_t6_rotate_narrow"] = g.detach().numpy()
    # raw dynamics-net logits on explicit rows (forward signature of ProfileForward2DModel)
    rs = np.random.RandomState(7)
    n = 37
    xc = torch.from_numpy(rs.randn(n, 14).astype(np.float32))
    xo = torch.from_numpy(rs.uniform(-1, 1, (n, 1)).astype(np.float32))
    xp = torch.from_numpy(rs.uniform(-1, 1, (n, 2)).astype(np.float32))
    tt = torch.from_numpy((rs.randint(0, 15, n) / 15.0).astype(np.float32))
    ov = objs2[rs.randint(0, 2, n)].reshape(n, -1)
    with torch.no_grad():
        out["fwd_logits"] = dm.classifier_model(xc, xo, xp, tt, object_vertices=ov).numpy()
    out.update(fwd_ctrl=xc.numpy(), fwd_ori=xo.numpy(), fwd_pos=xp.numpy(), fwd_t=tt.numpy(), fwd_obj=ov.numpy())
    # UNet
    with torch.no_grad():
        for t in (12, 0):
            out[f"unet_t{t}"] = dm.noise_pred_net(noise, t * torch.ones((B,), dtype=torch.int64)).numpy()

    # guided_sample loop body (diffusion.py:570-576) with the real cond_fn / UNet
    def loop(dm, noise, objs, name, scale, multi):
        tr = {}
        if multi:
            sample = noise.clone().detach()
            for i, t in enumerate(dm.noise_scheduler.timesteps):
                ts = t * torch.ones((noise.shape[0],), dtype=torch.int64)
                with torch.no_grad():
                    eps = dm.noise_pred_net(sample, ts)
                grad = 0.0
                for ov_ in objs:
                    grad += dm.cond_fn(sample, ts, opt_obj=name, object_vertices=ov_, ori_range=[-1.0, 1.0])
                grad /= len(objs)
                with torch.no_grad():
                    eh = eps - (1 - dm.noise_scheduler.alphas_cumprod[t]).sqrt() * grad * scale
                    sample = dm.noise_scheduler.step(eh, t, sample).prev_sample
                tr[f"eps_s{i}"], tr[f"grad_s{i}"], tr[f"sample_s{i}"] = eps.numpy(), grad.numpy(), sample.numpy()
            return tr
        for oi, ov_ in enumerate(objs):
            sample = noise.clone().detach()
            for i, t in enumerate(dm.noise_scheduler.timesteps):
                ts = t * torch.ones((noise.shape[0],), dtype=torch.int64)
                with torch.no_grad():
                    eps = dm.noise_pred_net(sample, ts)
                grad = dm.cond_fn(sample, ts, opt_obj=name, object_vertices=ov_, ori_range=[-1.0, 1.0])
                with torch.no_grad():
                    eh = eps - (1 - dm.noise_scheduler.alphas_cumprod[t]).sqrt() * grad * scale
                    sample = dm.noise_scheduler.step(eh, t, sample).prev_sample
                tr[f"eps_o{oi}_s{i}"], tr[f"grad_o{oi}_s{i}"], tr[f"sample_o{oi}_s{i}"] = \
                    eps.numpy(), grad.numpy(), sample.numpy()
        return tr

    for name in ("rotate_clockwise", "rotate"):
        for k, v in loop(dm, noise, objs2, name, 0.001, False).items():
            out[f"loop_{name}_{k}"] = v
    for k, v in loop(dm, noise, objs2, "shift_up", 0.001, True).items():
        out[f"multi_shift_up_{k}"] = v

    # convergence objective (f-2): centres from the unguided sample + one cond_fn call
    with torch.no_grad():
        ung = noise.clone()
        for t in dm.noise_scheduler.timesteps:
            ts = t * torch.ones((B,), dtype=torch.int64)
            ung = dm.noise_scheduler.step(dm.noise_pred_net(ung, ts), t, ung).prev_sample
    out["unguided"] = ung.numpy()
    for oi in range(2):
        c = dm.get_convergence_centers(ung, objs2[oi], B, ori_range=[-1.0, 1.0])
        out[f"conv_centers_o{oi}"] = c.numpy()
        g = dm.cond_fn(noise, 9 * torch.ones((B,), dtype=torch.int64), opt_obj="convergence",
                       object_vertices=objs2[oi], ori_range=[-1.0, 1.0], convergence_centers=c)
        out[f"grad_o{oi}_t9_convergence"] = g.detach().numpy()
    # profile pass used for scoring (get_convergence_centers input convention, diffusion.py:509-516)
    with torch.no_grad():
        final = torch.from_numpy(out["loop_rotate_clockwise_sample_o0_s4"])
        ori = torch.linspace(-1.0, 1.0, grid)
        ori = torch.concat([o.repeat(B).reshape((-1, 1)) for o in ori], dim=0)
        pos = torch.zeros((B * grid, 2))
        pts = torch.concat([final for _ in range(grid)], dim=0).reshape((B * grid, -1))
        ova = torch.stack([objs2[0].reshape(-1) for _ in range(B * grid)], dim=0)
        out["profile_logits_o0"] = dm.classifier_model(pts, ori, pos, torch.zeros(B * grid), object_vertices=ova).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_2d.npz"), **out)
    print("golden_2d.npz:", len(out), "arrays")

    # ------------------------------------------------------------------ 3D
    B, grid, npos, sub_bs = 3, 3, 2, 16
    objs3 = syn.objects_3d(2)
    starts = syn.fps_starts(2)
    dm3 = build("point_3d", B, grid, npos, objs3, sub_bs=sub_bs)
    noise3 = syn.initial_noise(B, 42, seed=0)

    class PinnedRandint:
        """Replaces pointnet2_utils.torch.randint: FPS start = host-supplied index, shared by every row of the
        object; calls alternate sa1, sa2 within one forward (pointnet2.py:28-29)."""
        def __init__(self):
            self.obj, self.calls = 0, 0

        def __call__(self, lo, hi, shape, dtype=torch.long):
            lvl = self.calls % 2
            self.calls += 1
            return torch.full(shape, int(starts[self.obj, lvl]), dtype=dtype)

    pinned = PinnedRandint()
    real_torch = pointnet2_utils.torch

    class _TorchProxy:
        def __getattr__(self, k):
            return pinned if k == "randint" else getattr(real_torch, k)

    pointnet2_utils.torch = _TorchProxy()
    out3 = {"noise": noise3.numpy(), "objects": objs3.numpy(), "fps_starts": starts.numpy(), "grid_size": grid,
            "num_pos": npos, "sub_bs": sub_bs}
    with torch.no_grad():
        for oi in range(2):
            pinned.obj, pinned.calls = oi, 0
            code, _ = dm3.classifier_model.module.object_encoder(objs3[oi].t()[None])
            out3[f"code_o{oi}"] = code.numpy()
        for t in (12, 0):
            out3[f"unet_t{t}"] = dm3.noise_pred_net(noise3, t * torch.ones((B,), dtype=torch.int64)).numpy()
    for oi in range(2):
        for t, name in ((12, "rotate_clockwise"), (6, "rotate"), (0, "counterclockwise_down")):
            pinned.obj, pinned.calls = oi, 0
            ts = t * torch.ones((B,), dtype=torch.int64)
            g = dm3.cond_fn(noise3, ts, opt_obj=name, object_vertices=objs3[oi], ori_range=[-1.0, 1.0])
            out3[f"grad_o{oi}_t{t}_{name}"] = g.detach().numpy()
    # raw logits on explicit rows with per-row clouds (ProfileForward3DModel.forward signature)
    rs = np.random.RandomState(11)
    n = 5
    xc = torch.from_numpy(rs.randn(n, 3, 42).astype(np.float32))
    xo = torch.from_numpy(rs.uniform(-1, 1, (n, 1)).astype(np.float32))
    xp = torch.from_numpy(rs.uniform(-1, 1, (n, 2)).astype(np.float32))
    tt = torch.from_numpy((rs.randint(0, 15, n) / 15.0).astype(np.float32))
    pinned.obj, pinned.calls = 1, 0
    with torch.no_grad():
        out3["fwd_logits_o1"] = dm3.classifier_model(xc, xo, xp, tt,
                                                     object_vertices=objs3[1].t()[None].repeat(n, 1, 1)).numpy()
    out3.update(fwd_ctrl=xc.numpy(), fwd_ori=xo.numpy(), fwd_pos=xp.numpy(), fwd_t=tt.numpy())

    # short guided loop, object 0, 3D scale
    tr = {}
    sample = noise3.clone()
    for i, t in enumerate(dm3.noise_scheduler.timesteps):
        ts = t * torch.ones((B,), dtype=torch.int64)
        with torch.no_grad():
            eps = dm3.noise_pred_net(sample, ts)
        pinned.obj, pinned.calls = 0, 0
        grad = dm3.cond_fn(sample, ts, opt_obj="rotate_clockwise", object_vertices=objs3[0], ori_range=[-1.0, 1.0])
        with torch.no_grad():
            eh = eps - (1 - dm3.noise_scheduler.alphas_cumprod[t]).sqrt() * grad * 0.5
            sample = dm3.noise_scheduler.step(eh, t, sample).prev_sample
        out3[f"loop_eps_s{i}"], out3[f"loop_grad_s{i}"], out3[f"loop_sample_s{i}"] = \
            eps.numpy(), grad.numpy(), sample.numpy()
    pointnet2_utils.torch = real_torch
    np.savez_compressed(os.path.join(HERE, "golden_3d.npz"), **out3)
    print("golden_3d.npz:", len(out3), "arrays")


if __name__ == "__main__":
    main()
