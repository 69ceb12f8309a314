"""Checkpoint ingestion and weight pre-packing (host logic; everything here is CPU-testable).

Accepts exactly the two checkpoint formats of the reference:
  * dynamics: flat ``state_dict`` saved from ``nn.DataParallel`` -- every key prefixed ``module.``,
    BatchNorm ``running_mean/var/num_batches_tracked`` present, 3D carries an unused ``time_encoder.*``
    (dynamics/trainer.py:105-106, loaded at generator/train.py:90);
  * diffusion: Lightning ``.ckpt`` whose ``state_dict`` holds ``ema_nets.noise_pred_net.*`` (the weights
    sampling uses), a nested ``ema_model`` dict, optionally ``classifier_model.module.*``, with
    ``_orig_mod.`` inserted by torch.compile (generator/diffusion.py:730-753).
Unknown keys are ignored, as ``load_state_dict(strict=False)`` does there.

Pre-packing (all exact algebra, done in float64 and rounded once to fp32):
  * BatchNorm (eval) folded into the preceding Linear / 1x1 conv;
  * trunk layer 1 split by input block [object | gripper | pose | time]
    (concatenation order of dynamics/profile_forward_2d.py:154);
  * transposed copies for the input-gradient GEMMs;
  * Conv1d weights re-laid tap-major ``[Cout][tap][Cin]`` for the implicit-GEMM convs; the
    ConvTranspose1d split into its even/odd output phases.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib

BN_EPS = 1e-5
POSE_DIM = 27   # 9 (orientation) + 18 (position) Fourier features, profile_forward_2d.py:86-91


def strip_module_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def load_dynamics_state_dict(path_or_sd) -> Dict[str, torch.Tensor]:
    """Reference format: generator/train.py:90 ``classifier_model.load_state_dict(torch.load(path))``."""
    sd = torch.load(path_or_sd, map_location="cpu") if isinstance(path_or_sd, (str, bytes)) else path_or_sd
    if "state_dict" in sd and not any(k.endswith("output.weight") for k in sd):
        sd = sd["state_dict"]
    # a Lightning guidance-time ckpt also embeds the classifier as classifier_model.module.*
    if any(k.startswith("classifier_model.") for k in sd if isinstance(k, str)):
        sd = {k[len("classifier_model."):]: v for k, v in sd.items()
              if isinstance(k, str) and k.startswith("classifier_model.")}
    return strip_module_prefix({k: v for k, v in sd.items() if isinstance(v, torch.Tensor)})


def load_unet_state_dict(path_or_ckpt) -> Dict[str, torch.Tensor]:
    """Reference format: generator/diffusion.py:730-748 (``ema_nets.noise_pred_net.*``, ``_orig_mod.`` stripped)."""
    ck = torch.load(path_or_ckpt, map_location="cpu") if isinstance(path_or_ckpt, (str, bytes)) else path_or_ckpt
    sd = ck["state_dict"] if "state_dict" in ck else ck
    pre = "ema_nets.noise_pred_net."
    if any(isinstance(k, str) and k.startswith(pre) for k in sd):
        return {k[len(pre):].replace("_orig_mod.", ""): v for k, v in sd.items()
                if isinstance(k, str) and k.startswith(pre) and isinstance(v, torch.Tensor)}
    return {k.replace("_orig_mod.", ""): v for k, v in sd.items() if isinstance(v, torch.Tensor)}


def _fold_bn(w: torch.Tensor, b: torch.Tensor, sd, bn: str):
    """(s*W, s*(b-mean)+beta), s = gamma/sqrt(var+eps); float64 in, fp32 out."""
    s = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + BN_EPS)
    wf = w.double().reshape(w.shape[0], -1) * s[:, None]
    bf = (b.double() - sd[bn + ".running_mean"].double()) * s + sd[bn + ".bias"].double()
    return wf, bf


def fold_dynamics(sd: Dict[str, torch.Tensor]) -> Dict[str, object]:
    """-> dict of CPU fp32 tensors + ints mirroring ``dgdm_dyn_weights`` (include/dgdm_b200.h)."""
    sd = strip_module_prefix(sd)
    is_3d = any(k.startswith("object_encoder.sa1") for k in sd)
    f32 = lambda t: t.to(torch.float32).contiguous()
    out: Dict[str, object] = {}
    out["P"] = int(sd["gripper_encoder.0.weight"].shape[1])
    out["is_3d"] = int(is_3d)
    out["obj_dim"] = 256 if is_3d else int(sd["object_encoder.0.weight"].shape[1])
    for nm, key in (("ge_w0", "gripper_encoder.0.weight"), ("ge_b0", "gripper_encoder.0.bias"),
                    ("ge_w1", "gripper_encoder.2.weight"), ("ge_b1", "gripper_encoder.2.bias")):
        out[nm] = f32(sd[key])
    if not is_3d:
        for nm, key in (("oe_w0", "object_encoder.0.weight"), ("oe_b0", "object_encoder.0.bias"),
                        ("oe_w1", "object_encoder.2.weight"), ("oe_b1", "object_encoder.2.bias"),
                        ("te_w0", "time_encoder.0.weight"), ("te_b0", "time_encoder.0.bias"),
                        ("te_w1", "time_encoder.2.weight"), ("te_b1", "time_encoder.2.bias")):
            out[nm] = f32(sd[key])
    # layer 1: columns are [object 256 | gripper 256 | pose 27 | time 256]
    w1, b1 = _fold_bn(sd["linears.0.weight"], sd["linears.0.bias"], sd, "linears.1")
    H1 = w1.shape[0]
    assert w1.shape[1] == 256 + 256 + POSE_DIM + 256, f"unexpected layer-1 fan-in {w1.shape[1]}"
    out["H1"] = int(H1)
    out["w1_obj"] = f32(w1[:, 0:256])
    out["w1_ctrl"] = f32(w1[:, 256:512])
    out["w1_ctrl_t"] = f32(w1[:, 256:512].t())
    out["w1_pose"] = f32(w1[:, 512:512 + POSE_DIM])
    out["w1_time"] = f32(w1[:, 512 + POSE_DIM:])
    out["b1"] = f32(b1)
    out["ge_w1_t"] = f32(sd["gripper_encoder.2.weight"].t())
    out["ge_w0_t"] = f32(sd["gripper_encoder.0.weight"].t())
    for i in range(7):   # trunk layers 2..8 = nn.Sequential indices 3,6,...,21 (+1 for their BatchNorm)
        w, b = _fold_bn(sd[f"linears.{3 * (i + 1)}.weight"], sd[f"linears.{3 * (i + 1)}.bias"], sd,
                        f"linears.{3 * (i + 1) + 1}")
        out[f"wl{i}"] = f32(w)
        out[f"wl_t{i}"] = f32(w.t())
        out[f"bl{i}"] = f32(b)
    out["w_out"] = f32(sd["output.weight"])
    out["b_out"] = f32(sd["output.bias"])
    return out


def fold_pointnet2(sd: Dict[str, torch.Tensor], prefix: str = "object_encoder.") -> Dict[str, torch.Tensor]:
    """1x1 Conv2d + BatchNorm2d(eval) of the three set-abstraction levels -> 5 (W [out,in], b [out]) pairs
    (dynamics/models/pointnet2_utils.py:176-183, 203-205)."""
    sd = strip_module_prefix(sd)
    out = {}
    i = 0
    for sa, nl in (("sa1", 2), ("sa2", 2), ("sa3", 1)):
        for l in range(nl):
            w, b = _fold_bn(sd[f"{prefix}{sa}.mlp_convs.{l}.weight"], sd[f"{prefix}{sa}.mlp_convs.{l}.bias"], sd,
                            f"{prefix}{sa}.mlp_bns.{l}")
            out[f"w{i}"] = w.to(torch.float32).contiguous()
            out[f"b{i}"] = b.to(torch.float32).contiguous()
            i += 1
    return out


_RES_BLOCKS = ["down_modules.0.0", "down_modules.0.1", "down_modules.1.0", "down_modules.1.1",
               "mid_modules.0", "mid_modules.1", "up_modules.0.0", "up_modules.0.1"]


def _tap_major(w: torch.Tensor) -> torch.Tensor:
    """Conv1d weight [Cout, Cin, k] -> [Cout, k, Cin] (K index = tap*Cin + ci)."""
    return w.permute(0, 2, 1).to(torch.float32).contiguous()


def fold_unet(sd: Dict[str, torch.Tensor]) -> Dict[str, object]:
    """-> dict mirroring ``dgdm_unet_weights``; checks the reference's only configuration (train.py:80)."""
    f32 = lambda t: t.to(torch.float32).contiguous()
    out: Dict[str, object] = {}
    out["se_w0"], out["se_b0"] = f32(sd["diffusion_step_encoder.1.weight"]), f32(sd["diffusion_step_encoder.1.bias"])
    out["se_w1"], out["se_b1"] = f32(sd["diffusion_step_encoder.3.weight"]), f32(sd["diffusion_step_encoder.3.bias"])
    if tuple(out["se_w0"].shape) != (128, 32):
        raise ValueError("unsupported UNet: diffusion_step_embed_dim must be 32 (generator/train.py:80)")
    blocks = []
    for p in _RES_BLOCKS:
        w0 = sd[p + ".blocks.0.block.0.weight"]
        b = {"cin": int(w0.shape[1]), "cout": int(w0.shape[0]),
             "conv0_w": _tap_major(w0), "conv0_b": f32(sd[p + ".blocks.0.block.0.bias"]),
             "gn0_w": f32(sd[p + ".blocks.0.block.1.weight"]), "gn0_b": f32(sd[p + ".blocks.0.block.1.bias"]),
             "conv1_w": _tap_major(sd[p + ".blocks.1.block.0.weight"]), "conv1_b": f32(sd[p + ".blocks.1.block.0.bias"]),
             "gn1_w": f32(sd[p + ".blocks.1.block.1.weight"]), "gn1_b": f32(sd[p + ".blocks.1.block.1.bias"]),
             "film_w": f32(sd[p + ".cond_encoder.1.weight"]), "film_b": f32(sd[p + ".cond_encoder.1.bias"]),
             "res_w": None, "res_b": None}
        if w0.shape[2] != 5:
            raise ValueError("unsupported UNet: kernel_size must be 5")
        if (p + ".residual_conv.weight") in sd:
            b["res_w"] = f32(sd[p + ".residual_conv.weight"][:, :, 0])
            b["res_b"] = f32(sd[p + ".residual_conv.bias"])
        blocks.append(b)
    out["blocks"] = blocks
    out["down_w"], out["down_b"] = _tap_major(sd["down_modules.0.2.conv.weight"]), f32(sd["down_modules.0.2.conv.bias"])
    wt = sd["up_modules.0.2.conv.weight"].to(torch.float32)    # ConvTranspose1d: [Cin, Cout, 4]
    # out[t] = sum_{j,k : t = 2j - 1 + k} in[j] . wt[:, :, k]; even t uses taps k=3 (j=m-1), k=1 (j=m);
    # odd t uses k=2 (j=m), k=0 (j=m+1)
    even = torch.stack([wt[:, :, 3].t(), wt[:, :, 1].t()], dim=1)    # [Cout, 2, Cin]
    odd = torch.stack([wt[:, :, 2].t(), wt[:, :, 0].t()], dim=1)
    out["up_w"], out["up_b"] = torch.stack([even, odd], dim=0).contiguous(), f32(sd["up_modules.0.2.conv.bias"])
    out["fin_w"], out["fin_b"] = _tap_major(sd["final_conv.0.block.0.weight"]), f32(sd["final_conv.0.block.0.bias"])
    out["fin_gn_w"], out["fin_gn_b"] = f32(sd["final_conv.0.block.1.weight"]), f32(sd["final_conv.0.block.1.bias"])
    out["out_w"], out["out_b"] = f32(sd["final_conv.1.weight"].reshape(1, -1)), f32(sd["final_conv.1.bias"])
    return out


# ---------------------------------------------------------------------------------------------------
# device-resident packs (hold the tensors alive and the ctypes struct that points at them)
# ---------------------------------------------------------------------------------------------------

class DynamicsPack:
    def __init__(self, sd, device, build_tc: bool = True):
        self.host = fold_dynamics(sd)
        self.device = torch.device(device)
        self.P, self.H1, self.is_3d, self.obj_dim = (self.host[k] for k in ("P", "H1", "is_3d", "obj_dim"))
        self.t: Dict[str, torch.Tensor] = {k: v.to(self.device) for k, v in self.host.items()
                                           if isinstance(v, torch.Tensor)}
        s = _lib.DynWeights()
        s.P, s.H1, s.obj_dim, s.is_3d = self.P, self.H1, self.obj_dim, self.is_3d
        for name, _ in _lib.DynWeights._fields_[4:24]:
            setattr(s, name, self.t[name].data_ptr() if name in self.t else None)
        for i in range(7):
            s.wl[i], s.wl_t[i], s.bl[i] = (self.t[f"{n}{i}"].data_ptr() for n in ("wl", "wl_t", "bl"))
        s.w_out, s.b_out = self.t["w_out"].data_ptr(), self.t["b_out"].data_ptr()
        s.tc_image = None
        self.struct = s
        self.tc_image: Optional[torch.Tensor] = None
        if build_tc:
            self.build_tc_image()

    def build_tc_image(self):
        l = _lib.lib()
        n = l.dgdm_dyn_tc_image_bytes(self.H1)
        self.tc_image = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(l.dgdm_dyn_pack_tc(C.byref(self.struct), self.tc_image.data_ptr(), _lib.stream_ptr(self.device)),
                       "dgdm_dyn_pack_tc")
        self.struct.tc_image = self.tc_image.data_ptr()


class PointNet2Pack:
    def __init__(self, sd, device):
        self.host = fold_pointnet2(sd)
        self.t = {k: v.to(device) for k, v in self.host.items()}
        s = _lib.PointNet2Weights()
        for i in range(5):
            s.w[i], s.b[i] = self.t[f"w{i}"].data_ptr(), self.t[f"b{i}"].data_ptr()
        self.struct = s


class UnetPack:
    def __init__(self, sd, device, build_tc: bool = True):
        self.host = fold_unet(sd)
        self.device = torch.device(device)
        dev = lambda v: None if v is None else v.to(device)
        self.t = {k: dev(v) for k, v in self.host.items() if k != "blocks"}
        self.blocks = [{k: (dev(v) if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
                       for b in self.host["blocks"]]
        s = _lib.UnetWeights()
        for k, v in self.t.items():
            setattr(s, k, v.data_ptr())
        for i, b in enumerate(self.blocks):
            s.blocks[i].cin, s.blocks[i].cout = b["cin"], b["cout"]
            for k, v in b.items():
                if k not in ("cin", "cout"):
                    setattr(s.blocks[i], k, None if v is None else v.data_ptr())
        s.tc_image = None
        self.struct = s
        self.tc_image: Optional[torch.Tensor] = None
        if build_tc:
            self.build_tc_image()

    def build_tc_image(self):
        l = _lib.lib()
        n = l.dgdm_unet_tc_image_bytes(C.byref(self.struct))
        self.tc_image = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(l.dgdm_unet_pack_tc(C.byref(self.struct), self.tc_image.data_ptr(), _lib.stream_ptr(self.device)),
                       "dgdm_unet_pack_tc")
        self.struct.tc_image = self.tc_image.data_ptr()
