"""``torch.library`` custom ops over the C ABI ("extern "C" functions ... wrapped as torch.library custom
ops so they compose with the host code").

    import dgdm_b200.ops                      # registers the namespace
    eps  = torch.ops.dgdm_b200.unet1d_forward(x, unet_handle, t, precision)
    grad = torch.ops.dgdm_b200.dyn_guidance(x, objects, dyn_handle, t_frac, grid_size, num_pos, ori_lo, ori_hi,
                                            c0, c1, c2, sq0, grad_mul, precision)
    x1   = torch.ops.dgdm_b200.ddim_guided_update(x, eps, grad, *scheduler.coefficients(t), scale, clip)
    idx, best = torch.ops.dgdm_b200.best_of_n(scores, k)

The ops are thin: contiguous fp32 CUDA tensors in, new tensors out, launched on the current stream of the tensors'
device; weights are passed as integer handles of packs registered with :func:`register_pack` (a pack owns the device
blob built once from a checkpoint: ``pack.DynamicsPack`` / ``pack.UnetPack``).  CUDA only -- there is no CPU kernel to
register, and a call with CPU tensors raises.  Each op has a fake (meta) implementation so that shape propagation
(``torch.compile`` tracing of the HOST code around the ops, FakeTensor) works without touching the GPU.

Reference call sites these replace: generator/diffusion.py:573 (``noise_pred_net``), :574 (``cond_fn``), :575-576
(guided epsilon + ``DDIMScheduler.step``), :346-428 (arg-max selection).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import _lib

_PACKS: Dict[int, object] = {}
_WS: Dict[Tuple[str, int], torch.Tensor] = {}          # (op, device index) -> workspace


def register_pack(pack) -> int:
    """Keep ``pack`` (DynamicsPack / UnetPack) alive and return the integer handle the ops take."""
    h = id(pack)
    _PACKS[h] = pack
    return h


def release_pack(handle: int) -> None:
    _PACKS.pop(handle, None)


def _pack(handle: int):
    try:
        return _PACKS[handle]
    except KeyError:
        raise ValueError(f"unknown pack handle {handle}: register the pack with dgdm_b200.ops.register_pack") from None


def _workspace(key: str, dev: torch.device, nbytes: int) -> torch.Tensor:
    k = (key, dev.index or 0)
    buf = _WS.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        _WS[k] = buf
    return buf


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"dgdm_b200 op: `{name}` must be a CUDA tensor (this path has no CPU implementation)")
    return t.to(torch.float32).contiguous()


# ---------------------------------------------------------------------------------------------- K3 denoiser
@torch.library.custom_op("dgdm_b200::unet1d_forward", mutates_args=())
def unet1d_forward(x: torch.Tensor, unet: int, t: int, precision: int) -> torch.Tensor:
    """ConditionalUnet1D.forward (diffusion_utils.py:238-285) for one timestep: x (n, P) -> eps (n, P)."""
    pk = _pack(unet)
    x = _cuda_f32(x, "x")
    n, P = x.shape
    out = torch.empty_like(x)
    lib = _lib.lib()
    ws = _workspace("unet", x.device, lib.dgdm_unet1d_workspace_bytes(n, P))
    with torch.cuda.device(x.device):
        _lib.check(lib.dgdm_unet1d_forward(C.byref(pk.struct), x.data_ptr(), n, P, int(t), out.data_ptr(), ws.data_ptr(),
                                           ws.numel(), int(precision), _lib.stream_ptr(x.device)), "dgdm_unet1d_forward")
    return out


@unet1d_forward.register_fake
def _(x, unet, t, precision):
    return torch.empty_like(x, dtype=torch.float32)


# ---------------------------------------------------------------------------------------------- K1 + K2 guidance
@torch.library.custom_op("dgdm_b200::dyn_guidance", mutates_args=())
def dyn_guidance(x: torch.Tensor, objects: torch.Tensor, dyn: int, t_frac: float, grid_size: int, num_pos: int,
                 ori_lo: float, ori_hi: float, c0: float, c1: float, c2: float, sq0: float, grad_mul: float,
                 precision: int) -> torch.Tensor:
    """cond_fn (diffusion.py:473-504) for designs x (n_obj * B, P), object-major, against objects (n_obj, obj_dim):
    gradient of sum over the pose grid of c . logits + sq0 * logits[0]^2 (deltas_to_objective, :430-471)."""
    pk = _pack(dyn)
    x, objects = _cuda_f32(x, "x"), _cuda_f32(objects, "objects")
    if x.shape[1] != pk.P or objects.shape[1] != pk.obj_dim:
        raise ValueError(f"designs must be (n, {pk.P}) and objects (n_obj, {pk.obj_dim})")
    nd, n_obj = x.shape[0], objects.shape[0]
    grid = _lib.PoseGrid()
    grid.ori_lo, grid.ori_hi, grid.grid_size, grid.num_pos = float(ori_lo), float(ori_hi), int(grid_size), int(num_pos)
    grid.pos_zero = 0
    obj = _lib.Objective()
    obj.c[0], obj.c[1], obj.c[2], obj.sq0, obj.row_coef = float(c0), float(c1), float(c2), float(sq0), None
    grad = torch.empty_like(x)
    lib = _lib.lib()
    ws = _workspace("dyn", x.device, lib.dgdm_dyn_guidance_workspace_bytes(C.byref(pk.struct), nd, n_obj, 1, C.byref(grid),
                                                                           int(precision)))
    with torch.cuda.device(x.device):
        _lib.check(lib.dgdm_dyn_guidance(C.byref(pk.struct), x.data_ptr(), nd, objects.data_ptr(), n_obj, 1, None,
                                         float(t_frac), C.byref(grid), C.byref(obj), float(grad_mul), grad.data_ptr(), None,
                                         ws.data_ptr(), ws.numel(), int(precision), _lib.stream_ptr(x.device)),
                   "dgdm_dyn_guidance")
    return grad


@dyn_guidance.register_fake
def _(x, objects, dyn, t_frac, grid_size, num_pos, ori_lo, ori_hi, c0, c1, c2, sq0, grad_mul, precision):
    return torch.empty_like(x, dtype=torch.float32)


# ---------------------------------------------------------------------------------------------- K4 update
@torch.library.custom_op("dgdm_b200::ddim_guided_update", mutates_args=())
def ddim_guided_update(sample: torch.Tensor, eps: torch.Tensor, grad: torch.Tensor, sqrt_1m_at: float, sqrt_at: float,
                       sqrt_aprev: float, sqrt_1m_aprev: float, scale: float, clip: bool) -> torch.Tensor:
    """eps_hat = eps - sqrt(1 - a_t) * grad * scale; DDIMScheduler.step (diffusion.py:575-576) with the four schedule
    coefficients (sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev)) of ``scheduler.DDIMScheduler.coefficients(t)``."""
    sample, eps, grad = _cuda_f32(sample, "sample"), _cuda_f32(eps, "eps"), _cuda_f32(grad, "grad")
    out = torch.empty_like(sample)
    with torch.cuda.device(sample.device):
        _lib.check(_lib.lib().dgdm_ddim_guided_update(out.data_ptr(), sample.data_ptr(), eps.data_ptr(), grad.data_ptr(),
                                                      sample.numel(), float(sqrt_1m_at), float(sqrt_at), float(sqrt_aprev),
                                                      float(sqrt_1m_aprev), float(scale), int(bool(clip)),
                                                      _lib.stream_ptr(sample.device)), "dgdm_ddim_guided_update")
    return out


@ddim_guided_update.register_fake
def _(sample, eps, grad, sqrt_1m_at, sqrt_at, sqrt_aprev, sqrt_1m_aprev, scale, clip):
    return torch.empty_like(sample, dtype=torch.float32)


# ---------------------------------------------------------------------------------------------- K6 selection
@torch.library.custom_op("dgdm_b200::best_of_n", mutates_args=())
def best_of_n(scores: torch.Tensor, k: int) -> List[torch.Tensor]:
    """Top-k candidates per object, ties to the lowest index (np.argmax, diffusion.py:356-386): scores (n_obj, B) ->
    [indices (n_obj, k) int64, values (n_obj, k)]."""
    s = _cuda_f32(scores, "scores")
    n_obj, n_cand = s.shape
    idx = torch.empty((n_obj, k), dtype=torch.int64, device=s.device)
    best = torch.empty((n_obj, k), dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib().dgdm_best_of_n(s.data_ptr(), n_obj, n_cand, int(k), idx.data_ptr(), best.data_ptr(),
                                             _lib.stream_ptr(s.device)), "dgdm_best_of_n")
    return [idx, best]


@best_of_n.register_fake
def _(scores, k):
    n_obj = scores.shape[0]
    return [scores.new_empty((n_obj, k), dtype=torch.int64), scores.new_empty((n_obj, k), dtype=torch.float32)]
