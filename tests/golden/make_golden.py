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

    python tests/golden/make_golden.py            # rewrites tests/golden/golden_{2d,3d}.npz
    python tests/golden/make_golden.py bigG       # golden_bigG.npz: cond_fn at BASELINE pose-grid sizes (G = 900 / 1125)
    python tests/golden/make_golden.py metrics    # golden_metrics.json: metric2objective / get_best_ids_* tables
"""
import json
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
# third-party stubs: shared with bench.py's reference arm (oracle/ref_stubs.py)
# ---------------------------------------------------------------------------------------------
sys.path.insert(0, os.path.join(REPO, "oracle"))
import ref_stubs  # noqa: E402

_StubDDIM = ref_stubs.StubDDIM


def install_stubs():
    ref_stubs.install_stubs(REF)


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


# ---------------------------------------------------------------------------------------------
def _build(mode, grid, npos, objs, sub_bs=16, seed=0):
    """The reference's Diffusion module around the reference's real networks, synthetic weights (strict load)."""
    from generator.diffusion import Diffusion
    from generator.diffusion_utils import ConditionalUnet1D
    from dynamics.profile_forward_2d import ProfileForward2DModel
    from dynamics.profile_forward_3d import ProfileForward3DModel
    from dgdm_b200 import synthetic as syn
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
    dm = Diffusion(noise_pred_net=unet, noise_scheduler=_StubDDIM(num_train_timesteps=15), num_inference_steps=5,
                   mode=mode, input_dim=1, num_points=P, class_cond=True, classifier_model=net, grid_size=grid,
                   num_pos=npos, object_vertices=objs, object_ids=list(range(len(objs))), sub_batch_size=sub_bs,
                   seed=seed)
    dm.eval()
    return dm


def main_big_g():
    """cond_fn of the real reference at the BASELINE pose grids (2D: 36 x 5 x 5 = 900 rows per candidate, 3D: 45 x 5 x 5 =
    1125, stock sub-batch 512), a few candidates: the fixtures the unscaled 1e-3 / 2e-2 bounds are checked against."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    install_stubs()
    import warnings
    warnings.filterwarnings("ignore")
    from dynamics.models import pointnet2_utils
    from dgdm_b200 import synthetic as syn
    out = {}
    B2 = 16
    objs2 = syn.objects_2d(2)
    dm = _build("point", 36, 5, objs2)
    n2 = syn.initial_noise(B2, 14, seed=3)
    out["noise_2d"], out["objects_2d"] = n2.numpy(), objs2.numpy()
    for t, name in ((6, "rotate_clockwise"), (12, "rotate"), (0, "counterclockwise_left")):
        g = dm.cond_fn(n2, t * torch.ones((B2,), dtype=torch.int64), opt_obj=name, object_vertices=objs2[1],
                       ori_range=[-1.0, 1.0])
        out[f"grad_2d_t{t}_{name}"] = g.detach().numpy()
    B3 = 8
    objs3, starts = syn.objects_3d(2), syn.fps_starts(2)
    dm3 = _build("point_3d", 45, 5, objs3, sub_bs=512)
    n3 = syn.initial_noise(B3, 42, seed=4)
    out["noise_3d"], out["objects_3d"], out["fps_starts"] = n3.numpy(), objs3.numpy(), starts.numpy()
    calls = {"n": 0}
    real_torch = pointnet2_utils.torch

    def pinned(lo, hi, shape, dtype=torch.long):
        lvl = calls["n"] % 2
        calls["n"] += 1
        return torch.full(shape, int(starts[1, lvl]), dtype=dtype)

    class _TorchProxy:
        def __getattr__(self, k):
            return pinned if k == "randint" else getattr(real_torch, k)

    pointnet2_utils.torch = _TorchProxy()
    for t, name in ((6, "rotate_clockwise"), (3, "rotate")):
        calls["n"] = 0
        g = dm3.cond_fn(n3, t * torch.ones((B3,), dtype=torch.int64), opt_obj=name, object_vertices=objs3[1],
                        ori_range=[-1.0, 1.0])
        out[f"grad_3d_t{t}_{name}"] = g.detach().numpy()
    pointnet2_utils.torch = real_torch
    np.savez_compressed(os.path.join(HERE, "golden_bigG.npz"), **out)
    print("golden_bigG.npz:", len(out), "arrays")


OBJECTIVES_15 = ["rotate", "rotate_clockwise", "rotate_counterclockwise", "shift_up", "shift_down", "shift_left",
                 "shift_right", "clockwise_up", "clockwise_down", "clockwise_left", "clockwise_right",
                 "counterclockwise_up", "counterclockwise_down", "counterclockwise_left", "counterclockwise_right"]


def main_metrics():
    """The reference's own metric tables on a seeded set of simulator-style metric dicts (dynamics/metrics.py:67-233
    metric2objective; generator/diffusion.py:354-428 get_average_best_ids / get_best_ids_all_metrics / get_best_ids),
    for all 15 non-convergence objectives.  Candidates 2 and 5 are exact copies of candidates 0 and 1, so every
    arg-min / arg-max has a tie and the first-occurrence rule is pinned too."""
    install_stubs()
    from generator.diffusion import Diffusion
    from dynamics.metrics import metric2objective
    rs = np.random.RandomState(42)
    n_cand, n_rot = 7, 36
    metrics = []
    for c in range(n_cand):
        m = {"profile": rs.randint(0, 3, n_rot), "profile_x": rs.randint(0, 3, n_rot), "profile_y": rs.randint(0, 3, n_rot),
             "delta_theta": rs.randn(n_rot) * 3.0, "delta_pos": rs.randn(n_rot, 2) * 0.004,
             "final_delta_theta": rs.randn(n_rot) * 20.0, "final_pos": rs.randn(n_rot, 2) * 0.02}
        metrics.append(m)
    for dst, src in ((2, 0), (5, 1)):
        metrics[dst] = {k: v.copy() for k, v in metrics[src].items()}
    py = lambda v: v.item() if hasattr(v, "item") else v
    out = {"metrics": [{k: v.tolist() for k, v in m.items()} for m in metrics], "objectives": {}}
    for name in OBJECTIVES_15:
        objs = [metric2objective(m, name) for m in metrics]
        best_all = Diffusion.get_best_ids_all_metrics(None, objs, opt_obj=name)
        best_avg = Diffusion.get_average_best_ids(None, objs, opt_obj=name)
        out["objectives"][name] = {"per_candidate": [{k: py(v) for k, v in o.items()} for o in objs],
                                   "dtypes": {k: type(v).__name__ for k, v in objs[0].items()},
                                   "best_ids_all_metrics": {k: int(v) for k, v in best_all.items()},
                                   "average_best_id": int(best_avg)}

    # get_best_ids: per-object blocks of num_grippers candidates, ids offset by the block start
    class _Self:
        get_best_ids_all_metrics = lambda self, objs, opt_obj="rotate": Diffusion.get_best_ids_all_metrics(None, objs, opt_obj)
    objs = [metric2objective(m, "rotate_clockwise") for m in metrics[:6]]
    blocks = Diffusion.get_best_ids(_Self(), objs, 3, 2, opt_obj="rotate_clockwise")
    out["get_best_ids_rotate_clockwise_3x2"] = [{k: int(v) for k, v in b.items()} for b in blocks]
    with open(os.path.join(HERE, "golden_metrics.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("golden_metrics.json:", len(out["objectives"]), "objectives x", n_cand, "candidates")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    {"all": main, "bigG": main_big_g, "metrics": main_metrics}[what]()
