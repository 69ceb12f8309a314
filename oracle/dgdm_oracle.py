"""CPU ORACLE for the dynamics-guided diffusion sampling path of real-stanford/dgdm.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
package ``dgdm_b200`` never does; its path is hand-written CUDA behind a C ABI and fails
loudly when the extension is missing.

What it is: a plain, un-hoisted, un-fused restatement of the reference algorithm in
functional torch-CPU fp32 (the reference itself is torch; torch-CPU is its "numpy").  It
deliberately does the work the way the reference does it -- the layer-1 GEMM over every
guidance row, PointNet++ per row unless told otherwise, autograd for the input gradient --
so it is an independent check on every algebraic restructuring the CUDA path makes
(layer-1 split, BatchNorm folding, once-per-object encoders, ReLU-sign-bit backward).

Pinning status: **pinned against the reference run in the authoring container.**  The
reference has no tests or golden vectors of its own, so
``tests/golden/make_golden.py`` imports the reference's real modules from /root/reference
(third-party imports stubbed), feeds both the same synthetic weights/inputs, and commits the
reference's outputs under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this
file against those fixtures.  The one exception is the DDIM arithmetic, which lives in
diffusers==0.11.1 (requirements.txt:1; not vendored, not installable offline): it is
restated from that release's published algorithm and is **parity unpinned** against
diffusers itself (it is pinned only against the closed form, see ``ddim_*`` below).

Every function cites the reference lines it follows; paths are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

BN_EPS = 1e-5
GN_EPS = 1e-5

# generator/diffusion.py:30-33
SCALE_2D = 0.001
SCALE_2D_CONV = 10.0
SCALE_3D = 0.5
SCALE_3D_CONV = 0.8

# generator/diffusion.py:116-117 (same constants in dynamics/dataloader.py:11-15)
THRESHOLD_2D = (0.03, 0.002, 0.003)
STD_2D = (0.0565, 0.0026, 0.0047)
THRESHOLD_3D = (0.02, 0.001, 0.001)
STD_3D = (0.0312, 0.0016, 0.0026)

LINEAR_OBJECTIVES = {
    # generator/diffusion.py:433-468 ; coefficients of (d_theta, d_x, d_y)
    "rotate_clockwise": (-1.0, 0.0, 0.0),
    "rotate_counterclockwise": (1.0, 0.0, 0.0),
    "shift_up": (0.0, -1.0, 0.0),
    "shift_down": (0.0, 1.0, 0.0),
    "shift_left": (0.0, 0.0, -1.0),
    "shift_right": (0.0, 0.0, 1.0),
    "clockwise_up": (-1.0, -1.0, 0.0),
    "clockwise_down": (-1.0, 1.0, 0.0),
    "clockwise_left": (-1.0, 0.0, -1.0),
    "clockwise_right": (-1.0, 0.0, 1.0),
    "counterclockwise_up": (1.0, -1.0, 0.0),
    "counterclockwise_down": (1.0, 1.0, 0.0),
    "counterclockwise_left": (1.0, 0.0, -1.0),
    "counterclockwise_right": (1.0, 0.0, 1.0),
}
ALL_OBJECTIVES = ["rotate"] + list(LINEAR_OBJECTIVES) + ["convergence"]


def strip_prefix(sd: StateDict, prefix: str = "module.") -> StateDict:
    """Dynamics checkpoints are saved from nn.DataParallel (dynamics/trainer.py:105-106)."""
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in sd.items()}


# ======================================================================================
# DDIM scheduler  (diffusers==0.11.1 DDIMScheduler; call sites generator/train.py:83,
# generator/diffusion.py:103,571,575-576)
# ======================================================================================

def ddim_alphas_cumprod(num_train_timesteps: int) -> Tensor:
    """``squaredcos_cap_v2``: beta_i = min(1 - abar((i+1)/T)/abar(i/T), 0.999),
    abar(s) = cos^2((s+0.008)/1.008 * pi/2); alphas_cumprod = cumprod(1-beta) in fp32."""
    def abar(s: float) -> float:
        return math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
    betas = [min(1.0 - abar((i + 1) / num_train_timesteps) / abar(i / num_train_timesteps), 0.999)
             for i in range(num_train_timesteps)]
    betas_t = torch.tensor(betas, dtype=torch.float32)
    return torch.cumprod(1.0 - betas_t, dim=0)


def ddim_timesteps(num_train_timesteps: int, num_inference_steps: int) -> np.ndarray:
    """``set_timesteps``: (arange(n) * (T // n)).round()[::-1], steps_offset = 0."""
    ratio = num_train_timesteps // num_inference_steps
    return (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)


def ddim_step(model_output: Tensor, t: int, sample: Tensor, alphas_cumprod: Tensor,
              num_train_timesteps: int, num_inference_steps: int, clip_sample: bool = True) -> Tensor:
    """``DDIMScheduler.step`` with eta = 0, prediction_type = epsilon, clip_sample = True,
    use_clipped_model_output = False, set_alpha_to_one = True."""
    prev_t = t - num_train_timesteps // num_inference_steps
    a_t = alphas_cumprod[t]
    a_prev = alphas_cumprod[prev_t] if prev_t >= 0 else torch.tensor(1.0)
    b_t = 1 - a_t
    x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
    if clip_sample:
        x0 = torch.clamp(x0, -1, 1)
    direction = (1 - a_prev) ** 0.5 * model_output
    return a_prev ** 0.5 * x0 + direction


# ======================================================================================
# Dynamics networks
# ======================================================================================

def fourier_embed(x: Tensor, n_freq: int = 4) -> Tensor:
    """``get_embedder(d, 4)``: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^3 x), cos(2^3 x)]
    (dynamics/profile_forward_2d.py:9-38, 41-56; log-sampled bands 2**linspace(0,3,4))."""
    bands = 2.0 ** torch.linspace(0.0, float(n_freq - 1), steps=n_freq)
    parts = [x]
    for f in bands:
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, dim=-1)


def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """[cos(t f_i), sin(t f_i)], f_i = exp(-ln(max_period) i / half)  (profile_forward_2d.py:58-76)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _mlp2(sd: StateDict, name: str, x: Tensor, act) -> Tensor:
    h = act(F.linear(x, sd[f"{name}.0.weight"], sd[f"{name}.0.bias"]))
    return F.linear(h, sd[f"{name}.2.weight"], sd[f"{name}.2.bias"])


def _trunk(sd: StateDict, x: Tensor) -> Tensor:
    """8 x (Linear -> BatchNorm1d(eval) -> ReLU) then Linear(256,3)
    (profile_forward_2d.py:109-135,154-155; profile_forward_3d.py:39-65,84-85)."""
    for i in range(8):
        x = F.linear(x, sd[f"linears.{3*i}.weight"], sd[f"linears.{3*i}.bias"])
        p = f"linears.{3*i+1}"
        x = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                         training=False, eps=BN_EPS)
        x = F.relu(x)
    return F.linear(x, sd["output.weight"], sd["output.bias"])


def dynamics2d_forward(sd: StateDict, x_ctrl: Tensor, x_ori: Tensor, x_pos: Tensor, t: Tensor,
                       object_vertices: Tensor) -> Tensor:
    """``ProfileForward2DModel.forward`` (profile_forward_2d.py:137-156), eval mode.
    x_ctrl (N,P), x_ori (N,1), x_pos (N,2), t (N,) float in [0,1), object_vertices (N,2V) -> (N,3)."""
    g = _mlp2(sd, "gripper_encoder", x_ctrl, F.relu)
    pose = torch.cat([fourier_embed(x_ori), fourier_embed(x_pos)], dim=1)
    o = _mlp2(sd, "object_encoder", object_vertices, F.relu)
    te = _mlp2(sd, "time_encoder", timestep_embedding(t, 128), F.silu)
    return _trunk(sd, torch.cat([o, g, pose, te], dim=1))


# ---- PointNet++ (dynamics/models/pointnet2.py, pointnet2_utils.py) -----------------------

def _sqdist(a: Tensor, b: Tensor) -> Tensor:
    """-2 a.b^T + |a|^2 + |b|^2  (pointnet2_utils.py:27-48); the expanded form matters for which
    points fall inside a ball."""
    d = -2 * torch.matmul(a, b.transpose(1, 2))
    d += (a ** 2).sum(-1)[:, :, None]
    d += (b ** 2).sum(-1)[:, None, :]
    return d


def _gather(points: Tensor, idx: Tensor) -> Tensor:
    """points (n,N,C), idx (n,...) -> (n,...,C)  (pointnet2_utils.py:51-68)."""
    n = points.shape[0]
    bidx = torch.arange(n).view([n] + [1] * (idx.dim() - 1)).expand_as(idx)
    return points[bidx, idx, :]


def farthest_point_sample(xyz: Tensor, npoint: int, start: Tensor) -> Tensor:
    """pointnet2_utils.py:71-92 with the ``torch.randint`` start (:83) replaced by the host-supplied
    ``start`` (n,) -- the pinned-noise convention of SURVEY.md §8c."""
    n, N, _ = xyz.shape
    out = torch.zeros(n, npoint, dtype=torch.long)
    dist = torch.full((n, N), 1e10)
    far = start.clone().long()
    rows = torch.arange(n)
    for i in range(npoint):
        out[:, i] = far
        c = xyz[rows, far, :].view(n, 1, 3)
        d = ((xyz - c) ** 2).sum(-1)
        dist = torch.minimum(dist, d)          # == the masked assignment of :89-90
        far = dist.max(-1)[1]
    return out


def ball_query(radius: float, nsample: int, xyz: Tensor, new_xyz: Tensor) -> Tensor:
    """First ``nsample`` indices (ascending) within ``radius`` of each query, padded with the first hit
    (pointnet2_utils.py:95-115)."""
    n, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = torch.arange(N).view(1, 1, N).repeat(n, S, 1)
    idx[_sqdist(new_xyz, xyz) > radius ** 2] = N
    idx = idx.sort(dim=-1)[0][:, :, :nsample]
    first = idx[:, :, :1].expand(-1, -1, nsample)
    return torch.where(idx == N, first, idx)


def _sa_mlp(sd: StateDict, name: str, x: Tensor, n_layers: int) -> Tensor:
    """1x1 Conv2d -> BatchNorm2d(eval) -> ReLU stack then max over the sample axis
    (pointnet2_utils.py:203-208).  x: (n, C, nsample, npoint)."""
    for i in range(n_layers):
        x = F.conv2d(x, sd[f"{name}.mlp_convs.{i}.weight"], sd[f"{name}.mlp_convs.{i}.bias"])
        p = f"{name}.mlp_bns.{i}"
        x = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                         training=False, eps=BN_EPS)
        x = F.relu(x)
    return x.max(dim=2)[0]


def pointnet2_encode(sd: StateDict, xyz_cn: Tensor, fps_start: Tensor, prefix: str = "object_encoder") -> Tensor:
    """``PointNet2.forward`` (pointnet2.py:21-31): xyz_cn (n,3,N) -> (n,256).
    sa1: FPS 512, r=0.2, k=32, mlp 3->64->128; sa2: FPS 128, r=0.4, k=64, mlp 131->128->256;
    sa3: group-all, mlp 259->256.  ``fps_start`` (n,2) pins the two FPS starts."""
    xyz = xyz_cn.permute(0, 2, 1).contiguous()          # (n,N,3)
    n = xyz.shape[0]
    feats: Optional[Tensor] = None
    for li, (name, npoint, radius, nsample, nl) in enumerate(
            [("sa1", 512, 0.2, 32, 2), ("sa2", 128, 0.4, 64, 2)]):
        fidx = farthest_point_sample(xyz, npoint, fps_start[:, li])
        new_xyz = _gather(xyz, fidx)                                      # (n,S,3)
        gidx = ball_query(radius, nsample, xyz, new_xyz)                  # (n,S,k)
        g_xyz = _gather(xyz, gidx) - new_xyz[:, :, None, :]               # centroid-relative
        grouped = g_xyz if feats is None else torch.cat([g_xyz, _gather(feats, gidx)], dim=-1)
        out = _sa_mlp(sd, f"{prefix}.{name}", grouped.permute(0, 3, 2, 1), nl)   # (n,C',S)
        xyz, feats = new_xyz, out.permute(0, 2, 1)
    grouped = torch.cat([xyz, feats], dim=-1)[:, None]                    # (n,1,128,259)
    out = _sa_mlp(sd, f"{prefix}.sa3", grouped.permute(0, 3, 2, 1), 1)    # (n,256,1)
    return out.reshape(n, -1)


def dynamics3d_forward(sd: StateDict, x_ctrl: Tensor, x_ori: Tensor, x_pos: Tensor, t: Tensor,
                       object_vertices: Optional[Tensor] = None, fps_start: Optional[Tensor] = None,
                       object_code: Optional[Tensor] = None) -> Tensor:
    """``ProfileForward3DModel.forward`` (profile_forward_3d.py:67-86), eval mode.
    x_ctrl (N,3,P) of which only row 1 is read (:78); object_vertices (N,3,512).
    Raw 256-d time embedding -- the model's ``time_encoder`` is unused (:83).
    ``object_code`` short-circuits the encoder (it has no dependence on x or t)."""
    g = _mlp2(sd, "gripper_encoder", x_ctrl[:, 1, :], F.relu)
    pose = torch.cat([fourier_embed(x_ori), fourier_embed(x_pos)], dim=1)
    if object_code is None:
        object_code = pointnet2_encode(sd, object_vertices, fps_start)
    te = timestep_embedding(t, 256)
    return _trunk(sd, torch.cat([object_code, g, pose, te], dim=1))


# ======================================================================================
# Denoiser: ConditionalUnet1D (generator/diffusion_utils.py:123-285)
# ======================================================================================

def _sinusoidal_pos_emb(t: Tensor, dim: int) -> Tensor:
    """diffusion_utils.py:25-37: [sin(t w_i), cos(t w_i)], w_i = exp(-ln(1e4) i/(half-1))."""
    half = dim // 2
    w = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    a = t[:, None] * w[None, :]
    return torch.cat((a.sin(), a.cos()), dim=-1)


def _conv_block(sd: StateDict, p: str, x: Tensor, n_groups: int = 8) -> Tensor:
    """Conv1d(k, pad k//2) -> GroupNorm(8) -> Mish  (diffusion_utils.py:57-72)."""
    w = sd[p + ".block.0.weight"]
    x = F.conv1d(x, w, sd[p + ".block.0.bias"], padding=w.shape[-1] // 2)
    x = F.group_norm(x, n_groups, sd[p + ".block.1.weight"], sd[p + ".block.1.bias"], eps=GN_EPS)
    return F.mish(x)


def _res_block(sd: StateDict, p: str, x: Tensor, cond: Tensor) -> Tensor:
    """ConditionalResidualBlock1D.forward (diffusion_utils.py:100-120): FiLM scale*out+bias after block 0."""
    out = _conv_block(sd, p + ".blocks.0", x)
    emb = F.linear(F.mish(cond), sd[p + ".cond_encoder.1.weight"], sd[p + ".cond_encoder.1.bias"])
    c = out.shape[1]
    emb = emb.reshape(emb.shape[0], 2, c, 1)
    out = emb[:, 0] * out + emb[:, 1]
    out = _conv_block(sd, p + ".blocks.1", out)
    if (p + ".residual_conv.weight") in sd:
        res = F.conv1d(x, sd[p + ".residual_conv.weight"], sd[p + ".residual_conv.bias"])
    else:
        res = x
    return out + res


def unet1d_forward(sd: StateDict, sample: Tensor, timestep: Tensor) -> Tensor:
    """``ConditionalUnet1D.forward`` (diffusion_utils.py:238-285) for the one configuration the
    reference builds (train.py:80): two levels [128,256], no global_cond.
    sample (B,P,1), timestep (B,) -> (B,P,1)."""
    x = sample.moveaxis(-1, -2)
    t = timestep.expand(x.shape[0])
    e = _sinusoidal_pos_emb(t, sd["diffusion_step_encoder.1.weight"].shape[1])
    e = F.linear(e, sd["diffusion_step_encoder.1.weight"], sd["diffusion_step_encoder.1.bias"])
    cond = F.linear(F.mish(e), sd["diffusion_step_encoder.3.weight"], sd["diffusion_step_encoder.3.bias"])
    skips: List[Tensor] = []
    n_down = 2
    for lvl in range(n_down):
        x = _res_block(sd, f"down_modules.{lvl}.0", x, cond)
        x = _res_block(sd, f"down_modules.{lvl}.1", x, cond)
        skips.append(x)
        if lvl < n_down - 1:
            x = F.conv1d(x, sd[f"down_modules.{lvl}.2.conv.weight"], sd[f"down_modules.{lvl}.2.conv.bias"],
                         stride=2, padding=1)
    for m in range(2):
        x = _res_block(sd, f"mid_modules.{m}", x, cond)
    # one up level; note skips[0] is pushed but never popped (diffusion_utils.py:264-275)
    x = torch.cat((x, skips.pop()), dim=1)
    x = _res_block(sd, "up_modules.0.0", x, cond)
    x = _res_block(sd, "up_modules.0.1", x, cond)
    x = F.conv_transpose1d(x, sd["up_modules.0.2.conv.weight"], sd["up_modules.0.2.conv.bias"], stride=2, padding=1)
    x = _conv_block(sd, "final_conv.0", x)
    x = F.conv1d(x, sd["final_conv.1.weight"], sd["final_conv.1.bias"])
    return x.moveaxis(-1, -2)


def unet_from_lightning(ckpt: Dict[str, object]) -> StateDict:
    """Pull the sampling weights out of a Lightning checkpoint the way generator/diffusion.py:730-748
    does: ``ema_nets.noise_pred_net.*`` (raw weights, never ``ema_model``), ``_orig_mod.`` stripped."""
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    out = {}
    pre = "ema_nets.noise_pred_net."
    for k, v in sd.items():
        if isinstance(k, str) and k.startswith(pre) and isinstance(v, torch.Tensor):
            out[k[len(pre):].replace("_orig_mod.", "")] = v
    return out


# ======================================================================================
# Objectives, guidance gradient, sampler
# ======================================================================================

def slicer(a: Tensor, lower: int, upper: int) -> Tensor:
    """Wrap-around slice (dynamics/metrics.py:32-38)."""
    if lower < 0:
        return torch.cat((a[lower:], a[:upper]))
    if upper > len(a):
        return torch.cat((a[lower:], a[:upper - len(a)]))
    return a[lower:upper]


def convergence_mode(profile: Tensor) -> Tuple[Tensor, Tensor]:
    """Lengths of 1..10..0 runs (with wrap-around) and the index of each run's last 1
    (dynamics/metrics.py:4-21)."""
    prof = torch.where(profile > 0, 1.0, 0.0)
    n = len(prof)
    if torch.all(prof == 0):
        return torch.tensor([n]), torch.tensor([0])
    if torch.all(prof == 1):
        return torch.tensor([n]), torch.tensor([n - 1])
    prof = torch.cat((prof, prof), dim=0)
    d = torch.diff(prof)
    conv = torch.where(d < 0)[0]
    conv = conv[conv < n]
    start = torch.where(d > 0)[0]
    lens = torch.diff(torch.cat((torch.tensor([0]), start[start > conv[0]], torch.tensor([2 * n]))))
    return lens[:len(conv)], conv


def convergence_mode_three_class(profile: Tensor) -> Tuple[Tensor, Tensor]:
    """dynamics/metrics.py:23-30: drop the 'no rotation' class (1) and analyse the rest."""
    ids = torch.where(profile != 1)[0]
    if len(ids) == 0:
        return torch.tensor([0]), torch.tensor([0])
    lens, conv = convergence_mode(profile[profile != 1])
    return lens, ids[conv]


def deltas_to_objective(deltas: Tensor, opt_obj: str, centers: Optional[Tensor] = None,
                        grid_size: int = 0, num_pos: int = 0) -> Tensor:
    """generator/diffusion.py:430-471.  ``deltas`` (...,3) = (d_theta, d_x, d_y)/std."""
    if opt_obj == "rotate":
        return deltas[..., 0] ** 2
    if opt_obj in LINEAR_OBJECTIVES:
        c0, c1, c2 = LINEAR_OBJECTIVES[opt_obj]
        out = 0
        for c, j in ((c0, 0), (c1, 1), (c2, 2)):
            if c != 0.0:
                out = out + (deltas[..., j] if c > 0 else -deltas[..., j])
        return out
    if opt_obj == "convergence":
        # :445-452, replicated as-is including its row-slicing quirk
        G = grid_size * num_pos ** 2
        half = (grid_size // 2) * num_pos ** 2
        parts = []
        for i, center in enumerate(centers):
            c = int(center)
            dt = deltas[i * G:(i + 1) * G, 0]
            parts.append(torch.cat([slicer(dt, c * num_pos ** 2 - half, c * num_pos ** 2),
                                    slicer(-dt, c * num_pos ** 2, c * num_pos ** 2 + half)], dim=0))
        return torch.cat(parts, dim=0)
    raise ValueError("opt obj not supported")


def pose_grid(batch: int, grid_size: int, num_pos: int, ori_range: Sequence[float]) -> Tuple[Tensor, Tensor]:
    """Row r = g*B + b, g = (o*num_pos + ix)*num_pos + iy (generator/diffusion.py:478-482).
    Returns ori (B*G,1) and pos (B*G,2)."""
    ori, px, py = torch.meshgrid(torch.linspace(ori_range[0], ori_range[1], grid_size),
                                 torch.linspace(-1, 1, num_pos), torch.linspace(-1, 1, num_pos), indexing="ij")
    rep = lambda v: v.reshape(-1).repeat_interleave(batch)
    return rep(ori).reshape(-1, 1), torch.stack([rep(px), rep(py)], dim=-1)


class OracleSampler:
    """Restatement of the sampler core of ``Diffusion`` (generator/diffusion.py:430-576, 621-647).

    mode: 'point' (2D, P=14) or 'point_3d' (3D, P=42).  ``dyn_sd`` may carry the ``module.`` prefix.
    For 3D, ``fps_start`` (n_obj,2) pins PointNet++; ``hoist_object_code=True`` evaluates the object
    encoder once per object instead of once per guidance row -- bit-identical in exact arithmetic and
    **[probe]**-verified to <3e-11, and the only way a CPU finishes in seconds.
    """

    def __init__(self, mode: str, unet_sd: StateDict, dyn_sd: StateDict, object_vertices: Tensor,
                 grid_size: int, num_pos: int, num_train_timesteps: int = 15, num_inference_steps: int = 5,
                 sub_batch_size: int = 512, fps_start: Optional[Tensor] = None, hoist_object_code: bool = True):
        assert mode in ("point", "point_3d")
        self.mode = mode
        self.unet_sd = unet_sd
        self.dyn_sd = strip_prefix(dyn_sd)
        self.object_vertices = object_vertices
        self.grid_size, self.num_pos = grid_size, num_pos
        self.T, self.n_inf = num_train_timesteps, num_inference_steps
        self.alphas_cumprod = ddim_alphas_cumprod(num_train_timesteps)
        self.timesteps = ddim_timesteps(num_train_timesteps, num_inference_steps)
        self.sub_batch_size = sub_batch_size
        self.fps_start = fps_start
        self.hoist = hoist_object_code
        thr, std = (THRESHOLD_3D, STD_3D) if mode == "point_3d" else (THRESHOLD_2D, STD_2D)
        self.threshold_std = torch.tensor(thr) / torch.tensor(std)
        self._codes: Dict[int, Tensor] = {}

    # -- dynamics net over explicit rows -------------------------------------------------
    def _object_code(self, obj_idx: int) -> Tensor:
        if obj_idx not in self._codes:
            with torch.no_grad():
                cloud = self.object_vertices[obj_idx].t()[None]              # (1,3,512)
                self._codes[obj_idx] = pointnet2_encode(self.dyn_sd, cloud, self.fps_start[obj_idx:obj_idx + 1])
        return self._codes[obj_idx]

    def _logits(self, pts: Tensor, ori: Tensor, pos: Tensor, t_float: Tensor, obj_idx: int) -> Tensor:
        n = pts.shape[0]
        if self.mode == "point":
            obj = self.object_vertices[obj_idx].reshape(1, -1).expand(n, -1)
            return dynamics2d_forward(self.dyn_sd, pts, ori, pos, t_float, obj)
        if self.hoist:
            return dynamics3d_forward(self.dyn_sd, pts, ori, pos, t_float,
                                      object_code=self._object_code(obj_idx).expand(n, -1))
        cloud = self.object_vertices[obj_idx].t()[None].expand(n, -1, -1)
        return dynamics3d_forward(self.dyn_sd, pts, ori, pos, t_float, object_vertices=cloud,
                                  fps_start=self.fps_start[obj_idx:obj_idx + 1].expand(n, -1))

    def _tile_pts(self, x: Tensor, G: int) -> Tensor:
        B, P = x.shape[0], x.shape[1]
        if self.mode == "point":
            return x.repeat(G, 1, 1).reshape(B * G, P)                         # diffusion.py:484
        lin = torch.linspace(-1.0, 1.0, P // 2).repeat(B * 2, 1).reshape(B, 1, -1)
        pts = torch.cat([lin, x.moveaxis(-1, -2), lin], dim=1)                 # diffusion.py:489
        return pts.repeat(G, 1, 1)

    # -- cond_fn -------------------------------------------------------------------------
    def cond_fn(self, x: Tensor, t: int, opt_obj: str, obj_idx: int, ori_range=(-1.0, 1.0),
                convergence_centers: Optional[Tensor] = None, return_logits: bool = False):
        """generator/diffusion.py:473-504: d/dx of the objective summed over all G pose rows."""
        B = x.shape[0]
        G = self.grid_size * self.num_pos ** 2
        with torch.enable_grad():
            x = x.detach().clone().requires_grad_(True)
            ori, pos = pose_grid(B, self.grid_size, self.num_pos, ori_range)
            pts = self._tile_pts(x, G)
            tf = torch.full((B * G,), float(t)) / self.T
            kw = dict(centers=convergence_centers, grid_size=self.grid_size, num_pos=self.num_pos)
            if self.mode == "point_3d":
                grad = 0.0
                all_logits = []
                for i in range(0, B * G, self.sub_batch_size):                 # diffusion.py:493-499
                    s = slice(i, i + self.sub_batch_size)
                    logits = self._logits(pts[s], ori[s], pos[s], tf[s], obj_idx)
                    grad = grad + torch.autograd.grad(deltas_to_objective(logits, opt_obj, **kw).sum(), x)[0]
                    all_logits.append(logits.detach())
                logits = torch.cat(all_logits)
            else:
                logits = self._logits(pts, ori, pos, tf, obj_idx)
                grad = torch.autograd.grad(deltas_to_objective(logits, opt_obj, **kw).sum(), x)[0]
        return (grad, logits.detach()) if return_logits else grad

    # -- convergence centres (generator/diffusion.py:506-539) ------------------------------
    def profile_logits(self, designs: Tensor, obj_idx: int, ori_range=(-1.0, 1.0)) -> Tensor:
        """Forward-only profile pass at t=0, pos=(0,0), ``grid_size`` orientations, rows r = g*B + b
        (diffusion.py:509-516).  Returns (grid_size, B, 3)."""
        B = designs.shape[0]
        with torch.no_grad():
            ori = torch.linspace(ori_range[0], ori_range[1], self.grid_size).repeat_interleave(B).reshape(-1, 1)
            pos = torch.zeros(B * self.grid_size, 2)
            pts = self._tile_pts(designs, self.grid_size)
            tf = torch.zeros(B * self.grid_size)
            logits = self._logits(pts, ori, pos, tf, obj_idx)
        return logits.reshape(self.grid_size, B, 3)

    def convergence_centers(self, unguided: Tensor, obj_idx: int, ori_range=(-1.0, 1.0)) -> Tensor:
        lg = self.profile_logits(unguided, obj_idx, ori_range)[..., 0]            # (grid,B)
        thr = self.threshold_std[0]
        cls = torch.where(lg > thr, 2.0, torch.where(lg < -thr, 0.0, 1.0))      # diffusion.py:532
        out = []
        for b in range(unguided.shape[0]):
            lens, centers = convergence_mode_three_class(cls[:, b])
            out.append(centers[torch.argmax(lens)])
        return torch.stack(out)

    # -- sampler loops -------------------------------------------------------------------
    def classifier_scale(self, opt_obj: str, multi: bool = False) -> float:
        if self.mode == "point":
            return SCALE_2D_CONV if (opt_obj == "convergence" and not multi) else SCALE_2D
        return SCALE_3D_CONV if (opt_obj == "convergence" and not multi) else SCALE_3D

    def _step(self, sample: Tensor, t: int, grad: Tensor, scale: float) -> Tuple[Tensor, Tensor]:
        B = sample.shape[0]
        with torch.no_grad():
            eps = unet1d_forward(self.unet_sd, sample, torch.full((B,), int(t), dtype=torch.int64))
            eps_hat = eps - (1 - self.alphas_cumprod[t]).sqrt() * grad * scale      # diffusion.py:575
            nxt = ddim_step(eps_hat, int(t), sample, self.alphas_cumprod, self.T, self.n_inf)
        return nxt, eps

    def guided_sample(self, noise: Tensor, opt_obj: str = "rotate", ori_range=(-1.0, 1.0),
                      unguided_sample: Optional[Tensor] = None, trace: Optional[list] = None) -> Tensor:
        """generator/diffusion.py:561-576: every object restarts from the same noise.
        Returns designs (n_obj,B,P,1).  ``trace`` collects per-step (eps, grad, sample)."""
        scale = self.classifier_scale(opt_obj)
        outs = []
        for oi in range(self.object_vertices.shape[0]):
            centers = self.convergence_centers(unguided_sample, oi, ori_range) if opt_obj == "convergence" else None
            sample = noise.clone().detach()
            for t in self.timesteps:
                grad = self.cond_fn(sample, int(t), opt_obj, oi, ori_range, centers)
                nxt, eps = self._step(sample, int(t), grad, scale)
                if trace is not None:
                    trace.append(dict(obj=oi, t=int(t), eps=eps, grad=grad, sample=nxt))
                sample = nxt
            outs.append(sample)
        return torch.stack(outs)

    def guided_sample_multi_object(self, noise: Tensor, opt_obj: str = "rotate", ori_range=(-1.0, 1.0),
                                   trace: Optional[list] = None) -> Tensor:
        """generator/diffusion.py:637-647: one trajectory, gradient averaged over objects."""
        scale = self.classifier_scale(opt_obj, multi=True)
        n_obj = self.object_vertices.shape[0]
        sample = noise.clone().detach()
        for t in self.timesteps:
            grad = 0.0
            for oi in range(n_obj):
                grad = grad + self.cond_fn(sample, int(t), opt_obj, oi, ori_range)
            grad = grad / n_obj
            nxt, eps = self._step(sample, int(t), grad, scale)
            if trace is not None:
                trace.append(dict(t=int(t), eps=eps, grad=grad, sample=nxt))
            sample = nxt
        return sample

    def unguided_sample(self, noise: Tensor) -> Tensor:
        """generator/diffusion.py:249-256."""
        sample = noise.clone().detach()
        for t in self.timesteps:
            sample, _ = self._step(sample, int(t), torch.zeros_like(sample), 0.0)
        return sample

    def denoise_from_data(self, data: Tensor, seed: int = 0) -> Dict[str, float]:
        """generator/diffusion.py:179-201, 230-244: noise the data at t = num_inference_steps, run the unguided
        schedule, report the three validation metrics."""
        B = data.shape[0]
        noise = torch.from_numpy(np.random.RandomState(seed).randn(*data.shape)).float()
        a = self.alphas_cumprod[self.n_inf]
        sample = a ** 0.5 * data + (1 - a) ** 0.5 * noise               # DDIMScheduler.add_noise
        npl = 0.0
        with torch.no_grad():
            for t in self.timesteps:
                eps = unet1d_forward(self.unet_sd, sample, torch.full((B,), int(t), dtype=torch.int64))
                npl += F.mse_loss(eps, noise).item()
                sample = ddim_step(eps, int(t), sample, self.alphas_cumprod, self.T, self.n_inf)
        return {"val/noise pred loss": npl / self.n_inf, "val/denoise loss": F.mse_loss(sample, data).item(),
                "val/accuracy": torch.mean((torch.abs(sample - data) < 0.01).float()).item()}

    # -- predicted task score and best-of-N (SURVEY.md §8c last row) -----------------------
    def score(self, designs: Tensor, obj_idx: int, opt_obj: str, ori_range=(-1.0, 1.0)) -> Tensor:
        """score[cand] = mean over the ``grid_size`` profile orientations of the objective."""
        lg = self.profile_logits(designs, obj_idx, ori_range)                  # (grid,B,3)
        return deltas_to_objective(lg, opt_obj).mean(dim=0)


def best_of_n(scores: Tensor) -> Tensor:
    """argmax over candidates, first occurrence wins (np.argmax; generator/diffusion.py:356-386)."""
    return torch.as_tensor(np.argmax(scores.numpy(), axis=-1))


def top_k(scores: Tensor, k: int) -> Tensor:
    """Indices of the k largest scores, ties broken by lowest index (stable descending sort)."""
    order = np.argsort(-scores.numpy(), axis=-1, kind="stable")
    return torch.as_tensor(order[..., :k].copy())
