"""Deterministic synthetic weights, objects and noise for parity tests and the benchmark.

There is no network and the reference ships no checkpoints, so every
parity pin is produced from weights generated here.  Everything is drawn from
``numpy.random.RandomState`` (Mersenne Twister, bit-stable across platforms), never from
``torch.manual_seed``, so the GPU box, this container and the golden-fixture generator
(``tests/golden/make_golden.py``) all see bit-identical tensors.

The state-dict *names and shapes* mirror what the reference modules create:
  * ``ProfileForward2DModel``  -- /root/reference/dynamics/profile_forward_2d.py:78-135
  * ``ProfileForward3DModel``  -- /root/reference/dynamics/profile_forward_3d.py:13-65
  * ``PointNet2``              -- /root/reference/dynamics/models/pointnet2.py:11-19
  * ``ConditionalUnet1D``      -- /root/reference/generator/diffusion_utils.py:123-236
``make_golden.py`` loads them into the real reference modules with ``strict=True``, which
is the check that the tables below are right.

BatchNorm running statistics and affine parameters are randomised (a fresh BatchNorm is
the identity and would hide BN-folding bugs).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np
import torch

W = 256  # trunk width of both dynamics networks


def _linear(rs: np.random.RandomState, out_f: int, in_f: int, gain: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
    """torch.nn.Linear default init (kaiming_uniform a=sqrt(5)) => U(-1/sqrt(in), 1/sqrt(in))."""
    bound = gain / math.sqrt(in_f)
    w = rs.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
    b = rs.uniform(-bound, bound, size=(out_f,)).astype(np.float32)
    return w, b


def _bn(rs: np.random.RandomState, ch: int) -> Dict[str, np.ndarray]:
    return {
        "weight": rs.uniform(0.5, 1.5, size=(ch,)).astype(np.float32),
        "bias": (0.2 * rs.randn(ch)).astype(np.float32),
        "running_mean": (0.2 * rs.randn(ch)).astype(np.float32),
        "running_var": rs.uniform(0.3, 1.5, size=(ch,)).astype(np.float32),
    }


def _put_linear(sd, prefix, rs, out_f, in_f, gain=1.0):
    w, b = _linear(rs, out_f, in_f, gain)
    sd[prefix + ".weight"] = torch.from_numpy(w)
    sd[prefix + ".bias"] = torch.from_numpy(b)


def _put_bn(sd, prefix, rs, ch):
    for k, v in _bn(rs, ch).items():
        sd[f"{prefix}.{k}"] = torch.from_numpy(v)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.int64)


def _put_trunk(sd, prefix, rs, in_dim, widths, gain):
    """8 x (Linear, BatchNorm1d, ReLU) laid out as nn.Sequential indices 0,1,2, 3,4,5, ..."""
    last = in_dim
    for i, wd in enumerate(widths):
        _put_linear(sd, f"{prefix}linears.{3 * i}", rs, wd, last, gain)
        _put_bn(sd, f"{prefix}linears.{3 * i + 1}", rs, wd)
        last = wd


def dynamics2d_state_dict(seed: int = 0, params_ch: int = 14, object_ch: int = 200,
                          data_parallel_prefix: bool = True, gain: float = 1.7) -> "OrderedDict[str, torch.Tensor]":
    """State dict of the 2D dynamics network in the reference checkpoint format.

    ``data_parallel_prefix`` reproduces the ``module.`` key prefix the reference's
    ``torch.save(DataParallel.state_dict())`` produces (dynamics/trainer.py:105-106).
    ``gain`` > 1 keeps activations from shrinking through 8 BN'd layers so the guidance
    gradient is not vanishingly small with random weights.
    """
    rs = np.random.RandomState(10_000 + seed)
    p = "module." if data_parallel_prefix else ""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _put_linear(sd, p + "time_encoder.0", rs, W, W // 2)
    _put_linear(sd, p + "time_encoder.2", rs, W, W)
    _put_linear(sd, p + "object_encoder.0", rs, W, object_ch)
    _put_linear(sd, p + "object_encoder.2", rs, W, W)
    _put_linear(sd, p + "gripper_encoder.0", rs, W, params_ch)
    _put_linear(sd, p + "gripper_encoder.2", rs, W, W)
    _put_trunk(sd, p, rs, W + 27 + W + W, [W] * 8, gain)
    _put_linear(sd, p + "output", rs, 3, W)
    return sd


_PN2_SPECS = [  # (name, in_channel, mlp)
    ("sa1", 3, [64, 128]),
    ("sa2", 128 + 3, [128, 256]),
    ("sa3", 256 + 3, [256]),
]


def dynamics3d_state_dict(seed: int = 0, params_ch: int = 42, data_parallel_prefix: bool = True,
                          gain: float = 1.7) -> "OrderedDict[str, torch.Tensor]":
    """State dict of the 3D dynamics network (PointNet++ object encoder, 795->512 first layer).

    Includes the ``time_encoder.*`` tensors the reference defines but never uses
    (profile_forward_3d.py:26-30 vs :83) because real checkpoints carry them.
    """
    rs = np.random.RandomState(20_000 + seed)
    p = "module." if data_parallel_prefix else ""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _put_linear(sd, p + "time_encoder.0", rs, W, W // 2)
    _put_linear(sd, p + "time_encoder.2", rs, W, W)
    for name, cin, mlp in _PN2_SPECS:
        last = cin
        for i, co in enumerate(mlp):
            w, b = _linear(rs, co, last, gain=1.5)
            sd[f"{p}object_encoder.{name}.mlp_convs.{i}.weight"] = torch.from_numpy(w.reshape(co, last, 1, 1).copy())
            sd[f"{p}object_encoder.{name}.mlp_convs.{i}.bias"] = torch.from_numpy(b)
            last = co
        last = cin
        for i, co in enumerate(mlp):
            _put_bn(sd, f"{p}object_encoder.{name}.mlp_bns.{i}", rs, co)
    _put_linear(sd, p + "gripper_encoder.0", rs, W, params_ch)
    _put_linear(sd, p + "gripper_encoder.2", rs, W, W)
    _put_trunk(sd, p, rs, W + 27 + W + W, [2 * W] + [W] * 7, gain)
    _put_linear(sd, p + "output", rs, 3, W)
    return sd


def _conv1d(rs, co, ci, k, gain=1.0):
    bound = gain / math.sqrt(ci * k)
    w = rs.uniform(-bound, bound, size=(co, ci, k)).astype(np.float32)
    b = rs.uniform(-bound, bound, size=(co,)).astype(np.float32)
    return torch.from_numpy(w), torch.from_numpy(b)


def _put_gn(sd, prefix, rs, ch):
    sd[prefix + ".weight"] = torch.from_numpy(rs.uniform(0.5, 1.5, size=(ch,)).astype(np.float32))
    sd[prefix + ".bias"] = torch.from_numpy((0.2 * rs.randn(ch)).astype(np.float32))


def _put_resblock(sd, prefix, rs, ci, co, cond_dim, k=5):
    for j, (a, b) in enumerate([(ci, co), (co, co)]):
        w, bias = _conv1d(rs, b, a, k, gain=1.4)
        sd[f"{prefix}.blocks.{j}.block.0.weight"] = w
        sd[f"{prefix}.blocks.{j}.block.0.bias"] = bias
        _put_gn(sd, f"{prefix}.blocks.{j}.block.1", rs, b)
    _put_linear(sd, f"{prefix}.cond_encoder.1", rs, 2 * co, cond_dim)
    if ci != co:
        w, bias = _conv1d(rs, co, ci, 1)
        sd[f"{prefix}.residual_conv.weight"] = w
        sd[f"{prefix}.residual_conv.bias"] = bias


def unet1d_state_dict(seed: int = 0, input_dim: int = 1, down_dims: Tuple[int, ...] = (128, 256),
                      dsed: int = 32, kernel_size: int = 5) -> "OrderedDict[str, torch.Tensor]":
    """State dict of ``ConditionalUnet1D(input_dim=1, global_cond_dim=0, down_dims=[128,256],
    diffusion_step_embed_dim=32)`` as built at generator/train.py:80.

    Key order follows module registration order in diffusion_utils.py:156-236
    (mid_modules, then diffusion_step_encoder, up_modules, down_modules, final_conv).
    """
    assert len(down_dims) == 2, "the reference only ever builds two levels (train.py:80)"
    rs = np.random.RandomState(30_000 + seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    d0, d1 = down_dims
    cond = dsed
    _put_resblock(sd, "mid_modules.0", rs, d1, d1, cond, kernel_size)
    _put_resblock(sd, "mid_modules.1", rs, d1, d1, cond, kernel_size)
    _put_linear(sd, "diffusion_step_encoder.1", rs, dsed * 4, dsed)
    _put_linear(sd, "diffusion_step_encoder.3", rs, dsed, dsed * 4)
    # up level 0: Res(2*d1 -> d0), Res(d0 -> d0), ConvTranspose1d(d0, d0, 4, 2, 1)
    _put_resblock(sd, "up_modules.0.0", rs, 2 * d1, d0, cond, kernel_size)
    _put_resblock(sd, "up_modules.0.1", rs, d0, d0, cond, kernel_size)
    bound = 1.0 / math.sqrt(d0 * 4)
    sd["up_modules.0.2.conv.weight"] = torch.from_numpy(rs.uniform(-bound, bound, size=(d0, d0, 4)).astype(np.float32))
    sd["up_modules.0.2.conv.bias"] = torch.from_numpy(rs.uniform(-bound, bound, size=(d0,)).astype(np.float32))
    # down level 0: Res(in -> d0), Res(d0 -> d0), Conv1d(d0, d0, 3, 2, 1); level 1: Res(d0->d1), Res(d1->d1), Identity
    _put_resblock(sd, "down_modules.0.0", rs, input_dim, d0, cond, kernel_size)
    _put_resblock(sd, "down_modules.0.1", rs, d0, d0, cond, kernel_size)
    w, b = _conv1d(rs, d0, d0, 3)
    sd["down_modules.0.2.conv.weight"] = w
    sd["down_modules.0.2.conv.bias"] = b
    _put_resblock(sd, "down_modules.1.0", rs, d0, d1, cond, kernel_size)
    _put_resblock(sd, "down_modules.1.1", rs, d1, d1, cond, kernel_size)
    # final: Conv1dBlock(d0, d0, k) + Conv1d(d0, input_dim, 1)
    w, b = _conv1d(rs, d0, d0, kernel_size, gain=1.4)
    sd["final_conv.0.block.0.weight"] = w
    sd["final_conv.0.block.0.bias"] = b
    _put_gn(sd, "final_conv.0.block.1", rs, d0)
    w, b = _conv1d(rs, input_dim, d0, 1)
    sd["final_conv.1.weight"] = w
    sd["final_conv.1.bias"] = b
    return sd


def lightning_diffusion_checkpoint(unet_sd, classifier_sd=None) -> Dict[str, object]:
    """Wrap a UNet state dict the way the reference's Lightning checkpoints store it
    (generator/diffusion.py:730-753): ``state_dict['ema_nets.noise_pred_net.*']``, a nested
    ``state_dict['ema_model']`` dict, and ``classifier_model.module.*`` when guidance was on."""
    sd: "OrderedDict[str, object]" = OrderedDict()
    for k, v in unet_sd.items():
        sd["ema_nets.noise_pred_net." + k] = v
    if classifier_sd is not None:
        for k, v in classifier_sd.items():
            kk = k if k.startswith("module.") else "module." + k
            sd["classifier_model." + kk] = v
    sd["ema_model"] = {"noise_pred_net." + k: v.clone() for k, v in unet_sd.items()}
    return {"state_dict": sd, "epoch": 0, "global_step": 0}


# --------------------------------------------------------------------------------------
# Inputs
# --------------------------------------------------------------------------------------

def initial_noise(batch: int, num_points: int, seed: int = 0) -> torch.Tensor:
    """The only randomness of sampling (generator/diffusion.py:182-183)."""
    return torch.from_numpy(np.random.RandomState(seed).randn(batch, num_points, 1)).float()


def objects_2d(n_obj: int, n_vertices: int = 100, seed_base: int = 1000) -> torch.Tensor:
    """``(n_obj, n_vertices, 2)`` closed star-shaped contours in [-1,1]^2, resampled by arc length and
    snapped to the q/64-1 lattice of the reference's icon contours (SURVEY.md §8d;
    assets/icon_process.py:25,52 then generator/train.py:119-124)."""
    out = np.zeros((n_obj, n_vertices, 2), dtype=np.float32)
    for k in range(n_obj):
        rs = np.random.RandomState(seed_base + k)
        a = rs.uniform(0.0, 0.12, size=4)
        ph = rs.uniform(0.0, 2 * np.pi, size=4)
        th = np.linspace(0.0, 2 * np.pi, 2048, endpoint=False)
        r = 0.5 + sum(a[j] * np.cos((j + 1) * th + ph[j]) for j in range(4))
        pts = np.stack([r * np.cos(th), r * np.sin(th)], axis=-1)
        seg = np.linalg.norm(np.diff(np.concatenate([pts, pts[:1]], 0), axis=0), axis=1)
        s = np.concatenate([[0.0], np.cumsum(seg)])
        tgt = np.linspace(0.0, s[-1], n_vertices, endpoint=False)
        closed = np.concatenate([pts, pts[:1]], 0)
        x = np.interp(tgt, s, closed[:, 0])
        y = np.interp(tgt, s, closed[:, 1])
        c = np.stack([x, y], -1)
        lo, hi = c.min(0), c.max(0)
        c = (c - lo) / (hi - lo) * 2.0 - 1.0
        q = np.clip(np.round((c + 1.0) * 64.0), 0, 127)
        out[k] = (q / 64.0 - 1.0).astype(np.float32)
    return torch.from_numpy(out)


def objects_3d(n_obj: int, n_points: int = 512, seed_base: int = 2000) -> torch.Tensor:
    """``(n_obj, n_points, 3)`` "scanned-style" clouds in [-1,1]^3: area-uniform surface samples of a random
    superquadric / box / cylinder inside x,y in [-0.1,0.1], z in [0,0.12], normalised as
    generator/train.py:107-109.  A tiny seeded jitter keeps all points distinct (no FPS ties,
    SURVEY.md §8 a-9)."""
    out = np.zeros((n_obj, n_points, 3), dtype=np.float32)
    for k in range(n_obj):
        rs = np.random.RandomState(seed_base + k)
        kind = rs.randint(3)
        half = rs.uniform(0.03, 0.09, size=2)
        height = rs.uniform(0.04, 0.12)
        u = rs.uniform(-1, 1, size=(n_points, 3))
        if kind == 0:      # superquadric: project random directions
            e = rs.uniform(0.4, 1.6)
            d = rs.randn(n_points, 3)
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            p = np.sign(d) * np.abs(d) ** e
            pts = np.stack([p[:, 0] * half[0], p[:, 1] * half[1], (p[:, 2] * 0.5 + 0.5) * height], -1)
        elif kind == 1:    # box: pick a face by area
            dims = np.array([2 * half[0], 2 * half[1], height])
            areas = np.array([dims[1] * dims[2], dims[0] * dims[2], dims[0] * dims[1]])
            face = rs.choice(3, size=n_points, p=areas / areas.sum())
            sgn = rs.choice([-1.0, 1.0], size=n_points)
            p = u.copy()
            p[np.arange(n_points), face] = sgn
            pts = np.stack([p[:, 0] * half[0], p[:, 1] * half[1], (p[:, 2] * 0.5 + 0.5) * height], -1)
        else:              # cylinder: side + caps
            rad = half[0]
            a_side, a_cap = 2 * np.pi * rad * height, np.pi * rad * rad
            on_side = rs.uniform(size=n_points) < a_side / (a_side + 2 * a_cap)
            ang = rs.uniform(0, 2 * np.pi, size=n_points)
            rr = np.where(on_side, rad, rad * np.sqrt(rs.uniform(size=n_points)))
            zz = np.where(on_side, rs.uniform(0, height, size=n_points), rs.choice([0.0, height], size=n_points))
            pts = np.stack([rr * np.cos(ang), rr * np.sin(ang), zz], -1)
        pts = pts + 1e-4 * rs.randn(n_points, 3)
        pts[:, 0] = (pts[:, 0] + 0.1) / 0.2 * 2.0 - 1.0
        pts[:, 1] = (pts[:, 1] + 0.1) / 0.2 * 2.0 - 1.0
        pts[:, 2] = (pts[:, 2] - 0.0) / 0.12 * 2.0 - 1.0
        out[k] = pts.astype(np.float32)
    return torch.from_numpy(out)


def fps_starts(n_obj: int, n_points: int = 512, seed_base: int = 3000) -> torch.Tensor:
    """Host-supplied farthest-point-sampling start indices, one per (object, SA level).

    The reference draws them with ``torch.randint`` per batch row (pointnet2_utils.py:83), which makes
    stock 3D output irreproducible; parity is defined with these pinned starts.
    Level 0 indexes the ``n_points`` cloud, level 1 indexes sa1's 512 output points."""
    out = np.zeros((n_obj, 2), dtype=np.int64)
    for k in range(n_obj):
        rs = np.random.RandomState(seed_base + k)
        out[k, 0] = rs.randint(n_points)
        out[k, 1] = rs.randint(512)
    return torch.from_numpy(out)
