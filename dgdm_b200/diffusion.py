"""Host-side mirror of the sampler core of the reference's ``Diffusion`` LightningModule
(/root/reference/generator/diffusion.py:430-576, 621-647) -- same method names, argument meaning and error
behaviour -- driving the sm_100a kernels through the C ABI (include/dgdm_b200.h).

What stays on the host: argument checking, the 5-iteration step loop, the 15-entry noise schedule and the
tiny run-length analysis of ``get_convergence_centers``.  Everything numeric runs in the CUDA library:
  denoiser forward           -> dgdm_unet1d_forward          (K3)
  cond_fn (fwd + d/dx, sum)  -> dgdm_dyn_guidance            (K1 + K2)
  guidance + DDIM update     -> dgdm_ddim_guided_update      (K4)
  PointNet++ object codes    -> dgdm_pointnet2_encode        (K5, once per object)
  profile pass / scores      -> dgdm_dyn_score               (K1 forward only)
  best-of-N                  -> dgdm_best_of_n               (K6)
There is no PyTorch fallback: without the built library every call raises.

Differences from the reference that are deliberate and documented:
  * objects are processed as one batch (``n_obj * B`` trajectories per launch) instead of a Python loop;
    the per-object results are identical because trajectories are independent (diffusion.py:561-570);
  * "best design" selection uses the network-predicted task score, because the MuJoCo
    evaluator that the reference scores with is CPU-only and out of scope;
  * 3D farthest-point-sampling starts are host-supplied (``fps_starts``) instead of ``torch.randint``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .pack import DynamicsPack, PointNet2Pack, UnetPack, load_dynamics_state_dict, load_unet_state_dict
from .scheduler import DDIMScheduler

# generator/diffusion.py:30-33
SCALE_2D = 0.001
SCALE_2D_CONV = 10.0
SCALE_3D = 0.5
SCALE_3D_CONV = 0.8

# deltas_to_objective (generator/diffusion.py:430-471): (c_theta, c_x, c_y, sq0)
OBJECTIVES: Dict[str, tuple] = {
    "rotate": (0.0, 0.0, 0.0, 1.0),
    "rotate_clockwise": (-1.0, 0.0, 0.0, 0.0),
    "rotate_counterclockwise": (1.0, 0.0, 0.0, 0.0),
    "shift_up": (0.0, -1.0, 0.0, 0.0),
    "shift_down": (0.0, 1.0, 0.0, 0.0),
    "shift_left": (0.0, 0.0, -1.0, 0.0),
    "shift_right": (0.0, 0.0, 1.0, 0.0),
    "convergence": (1.0, 0.0, 0.0, 0.0),      # with a per-row sign table
    "clockwise_up": (-1.0, -1.0, 0.0, 0.0),
    "clockwise_down": (-1.0, 1.0, 0.0, 0.0),
    "clockwise_left": (-1.0, 0.0, -1.0, 0.0),
    "clockwise_right": (-1.0, 0.0, 1.0, 0.0),
    "counterclockwise_up": (1.0, -1.0, 0.0, 0.0),
    "counterclockwise_down": (1.0, 1.0, 0.0, 0.0),
    "counterclockwise_left": (1.0, 0.0, -1.0, 0.0),
    "counterclockwise_right": (1.0, 0.0, 1.0, 0.0),
}


def _objective_struct(opt_obj: str, row_coef: Optional[torch.Tensor] = None) -> _lib.Objective:
    if opt_obj not in OBJECTIVES:
        raise ValueError("opt obj not supported")            # same message as diffusion.py:470
    c0, c1, c2, sq = OBJECTIVES[opt_obj]
    o = _lib.Objective()
    o.c[0], o.c[1], o.c[2], o.sq0 = c0, c1, c2, sq
    o.row_coef = None if row_coef is None else row_coef.data_ptr()
    return o


def _slicer_idx(n: int, lower: int, upper: int) -> np.ndarray:
    """Index version of dynamics/metrics.py:32-38 ``slicer`` (wrap-around slice of arange(n))."""
    a = np.arange(n)
    if lower < 0:
        return np.concatenate((a[lower:], a[:upper]))
    if upper > n:
        return np.concatenate((a[lower:], a[:upper - n]))
    return a[lower:upper]


def _convergence_mode(profile: np.ndarray):
    """dynamics/metrics.py:4-21 on a {0,2} profile: lengths of 1..10..0 runs and index of each run's last 1."""
    prof = (profile > 0).astype(np.float64)
    n = len(prof)
    if np.all(prof == 0):
        return np.array([n]), np.array([0])
    if np.all(prof == 1):
        return np.array([n]), np.array([n - 1])
    two = np.concatenate((prof, prof))
    d = np.diff(two)
    conv = np.where(d < 0)[0]
    conv = conv[conv < n]
    start = np.where(d > 0)[0]
    lens = np.diff(np.concatenate(([0], start[start > conv[0]], [2 * n])))
    return lens[:len(conv)], conv


def _convergence_center(classes: np.ndarray) -> int:
    """convergence_mode_three_class (dynamics/metrics.py:23-30) + centers[argmax(lengths)] (diffusion.py:536-537)."""
    ids = np.where(classes != 1)[0]
    if len(ids) == 0:
        return 0
    lens, conv = _convergence_mode(classes[classes != 1])
    return int(ids[conv][int(np.argmax(lens))])


class Diffusion:
    """Drop-in for the sampling half of ``generator.diffusion.Diffusion``.

    ``noise_pred_net``: a ConditionalUnet1D state dict, a Lightning checkpoint (dict or path) or a ``UnetPack``.
    ``classifier_model``: a dynamics-network state dict in the reference checkpoint format (``module.`` prefix
    accepted), a path, or a ``DynamicsPack``.
    ``precision``: 'fp32' (tcgen05 bf16x3, <=1e-3), 'bf16' (tcgen05 single pass, <=2e-2) or 'fp32_simt'
    (CUDA-core FFMA, exact order).
    """

    def __init__(self, noise_pred_net, noise_scheduler: DDIMScheduler, num_inference_steps: int,
                 num_epochs: int = 10000, mode: str = "point", input_dim: int = 1, num_points: int = 14,
                 class_cond: bool = True, classifier_model=None, grid_size: int = 360, num_pos: int = 5,
                 object_vertices: Optional[torch.Tensor] = None, object_ids: Optional[List] = None,
                 num_cpus: int = 32, sub_batch_size: int = 1024, pts_x_dim: int = 7, pts_z_dim: int = 3,
                 render_video: bool = False, seed: int = 0, *, fps_starts: Optional[torch.Tensor] = None,
                 precision: str = "fp32", device: Union[str, torch.device] = "cuda:0", **_ignored):
        if mode not in ("point", "point_3d"):
            raise ValueError("model type not supported")        # diffusion.py:502
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        if input_dim != 1:
            raise ValueError("only input_dim=1 is supported (generator/train.py:74)")
        self.lib = _lib.lib()                                    # raises if the CUDA library is not built
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DgdmError("dgdm_b200 runs on CUDA devices only (no CPU fallback)")
        self.mode, self.num_points, self.seed = mode, num_points, seed
        self.pts_x_dim, self.pts_z_dim = pts_x_dim, pts_z_dim
        self.precision = _lib.PRECISIONS[precision]
        self.noise_scheduler = noise_scheduler
        self.num_inference_steps = num_inference_steps
        self.noise_scheduler.set_timesteps(num_inference_steps)   # diffusion.py:103
        self.class_cond = class_cond
        self.sub_batch_size = sub_batch_size                      # accepted, unused: rows never hit HBM
        with torch.cuda.device(self.device):
            self.unet = noise_pred_net if isinstance(noise_pred_net, UnetPack) else \
                UnetPack(load_unet_state_dict(noise_pred_net), self.device,
                         build_tc=self.precision != _lib.PREC_FP32_SIMT)
            self.dyn: Optional[DynamicsPack] = None
            self.pn2: Optional[PointNet2Pack] = None
            if class_cond:
                if classifier_model is None:
                    raise ValueError("classifier_model is required with class_cond=True")
                if isinstance(classifier_model, DynamicsPack):
                    self.dyn = classifier_model
                    sd = None
                else:
                    sd = load_dynamics_state_dict(classifier_model)
                    self.dyn = DynamicsPack(sd, self.device, build_tc=self.precision != _lib.PREC_FP32_SIMT)
                if bool(self.dyn.is_3d) != (mode == "point_3d"):
                    raise ValueError("classifier_model does not match mode")
                if self.dyn.P != num_points:
                    raise ValueError(f"classifier expects {self.dyn.P} control points, num_points={num_points}")
                if self.dyn.is_3d:
                    if sd is None:
                        raise ValueError("3D needs the raw state dict (PointNet++ weights) alongside DynamicsPack")
                    self.pn2 = PointNet2Pack(sd, self.device)
        self.grid_size, self.num_pos = grid_size, num_pos
        self.object_ids = object_ids
        thr, std = ((0.02, 0.001, 0.001), (0.0312, 0.0016, 0.0026)) if mode == "point_3d" else \
                   ((0.03, 0.002, 0.003), (0.0565, 0.0026, 0.0047))                       # diffusion.py:116-117
        self.threshold = torch.tensor(thr)
        self.std = torch.tensor(std)
        self.threshold_std = self.threshold / self.std
        self.max_designs_per_launch = 131072                     # guided_sample chunks objects beyond this (see there)
        self._ws: Dict[str, torch.Tensor] = {}
        self._graphs: Dict[tuple, tuple] = {}                    # captured passes (guided_sample(cuda_graph=True))
        self._obj_version = 0                                    # bumped by set_objects: captured graphs read its buffers
        self.object_vertices = None
        self.fps_starts = None
        self._obj_dev: Optional[torch.Tensor] = None
        if class_cond:
            self.set_objects(object_vertices, fps_starts)

    # ------------------------------------------------------------------------------------------
    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf

    def encode_objects(self, object_vertices: torch.Tensor, fps_starts: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device-side object representation the guidance kernels take: 2D flattened contours
        (x0,y0,x1,y1,...; diffusion.py:485) or 3D PointNet++ codes (K5)."""
        ov = object_vertices.to(device=self.device, dtype=torch.float32).contiguous()
        if self.mode == "point":
            flat = ov.reshape(ov.shape[0], -1).contiguous()
            if flat.shape[1] != self.dyn.obj_dim:
                raise ValueError(f"object has {flat.shape[1]} coordinates, the object encoder expects {self.dyn.obj_dim}")
            return flat
        n, npts = ov.shape[0], ov.shape[1]
        if fps_starts is None:
            fps_starts = torch.zeros((n, 2), dtype=torch.int64)
        st = fps_starts.to(device=self.device, dtype=torch.int64).contiguous()
        if st.shape != (n, 2):
            raise ValueError("fps_starts must be (n_obj, 2)")
        codes = torch.empty((n, 256), dtype=torch.float32, device=self.device)
        nb = self.lib.dgdm_pointnet2_workspace_bytes(n, npts)
        if nb == 0:
            raise ValueError(f"unsupported cloud size {npts} (object_max_num_vertices must be 512)")
        ws = self._workspace("pn2", nb)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dgdm_pointnet2_encode(C.byref(self.pn2.struct), ov.data_ptr(), n, npts, st.data_ptr(),
                                                      codes.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(self.device)),
                       "dgdm_pointnet2_encode")
        return codes

    def set_objects(self, object_vertices: Optional[torch.Tensor], fps_starts: Optional[torch.Tensor] = None) -> None:
        if object_vertices is None:
            raise ValueError("object vertices not provided")    # diffusion.py:295
        self.object_vertices = object_vertices
        self.fps_starts = fps_starts
        self._obj_dev = self.encode_objects(object_vertices, fps_starts)
        self._obj_version += 1                                   # graphs captured against the old object codes are stale
        self._graphs.clear()

    # ------------------------------------------------------------------------------------------
    def noise_pred_net(self, sample: torch.Tensor, timesteps) -> torch.Tensor:
        """``ConditionalUnet1D.forward`` (diffusion_utils.py:238-285): (B,P,1), (B,) int64 -> (B,P,1).
        All timesteps of a call must be equal, as they are at diffusion.py:572."""
        if torch.is_tensor(timesteps):
            t = int(timesteps.reshape(-1)[0])
            if timesteps.numel() > 1 and not bool((timesteps == t).all()):
                raise ValueError("per-sample timesteps are not supported: the sampler uses one t per step")
        else:
            t = int(timesteps)
        x = sample.to(device=self.device, dtype=torch.float32).reshape(sample.shape[0], -1).contiguous()
        n, P = x.shape
        eps = torch.empty_like(x)
        ws = self._workspace("unet", self.lib.dgdm_unet1d_workspace_bytes(n, P))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dgdm_unet1d_forward(C.byref(self.unet.struct), x.data_ptr(), n, P, t, eps.data_ptr(),
                                                    ws.data_ptr(), ws.numel(), self.precision, _lib.stream_ptr(self.device)),
                       "dgdm_unet1d_forward")
        return eps.reshape(sample.shape)

    def _check_widths(self, x: torch.Tensor, objects_dev: torch.Tensor) -> None:
        """The C ABI takes P and obj_dim from the weight struct: a tensor of another width would be read with the
        wrong stride (or out of bounds), so refuse it here."""
        if x.dim() != 2 or x.shape[1] != self.dyn.P:
            raise ValueError(f"designs must be (n, {self.dyn.P}), got {tuple(x.shape)}")
        if objects_dev.dim() != 2 or objects_dev.shape[1] != self.dyn.obj_dim:
            raise ValueError(f"objects must be (n, {self.dyn.obj_dim}) "
                             f"({'PointNet++ codes' if self.dyn.is_3d else 'flattened contours'}), "
                             f"got {tuple(objects_dev.shape)}")

    def _grid(self, ori_range: Sequence[float], profile: bool = False) -> _lib.PoseGrid:
        g = _lib.PoseGrid()
        g.ori_lo, g.ori_hi = float(ori_range[0]), float(ori_range[1])
        g.grid_size, g.num_pos, g.pos_zero = self.grid_size, self.num_pos, int(profile)
        return g

    def _t_frac(self, t: int) -> float:
        # timesteps.float() / num_train_timesteps in fp32 (diffusion.py:487)
        return float(torch.tensor(float(t), dtype=torch.float32) / self.noise_scheduler.config.num_train_timesteps)

    def guidance(self, x: torch.Tensor, t: int, objects_dev: torch.Tensor, objs_per_design: int, opt_obj: str,
                 ori_range=(-1.0, 1.0), grad_mul: float = 1.0, row_coef: Optional[torch.Tensor] = None,
                 pair_object: Optional[torch.Tensor] = None, want_logits: bool = False):
        """Batched cond_fn: x (n_designs,P) -> grad (n_designs,P) [, logits (n_pairs*G,3)]."""
        self._check_widths(x, objects_dev)
        nd, P = x.shape
        n_obj = objects_dev.shape[0]
        grid = self._grid(ori_range)
        obj = _objective_struct(opt_obj, row_coef)
        G = self.grid_size * self.num_pos ** 2
        grad = torch.empty_like(x)
        logits = torch.empty((nd * objs_per_design * G, 3), dtype=torch.float32, device=self.device) if want_logits else None
        nb = self.lib.dgdm_dyn_guidance_workspace_bytes(C.byref(self.dyn.struct), nd, n_obj, objs_per_design,
                                                        C.byref(grid), self.precision)
        ws = self._workspace("dyn", nb)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dgdm_dyn_guidance(
                C.byref(self.dyn.struct), x.data_ptr(), nd, objects_dev.data_ptr(), n_obj, objs_per_design,
                _lib.ptr(pair_object), self._t_frac(t), C.byref(grid), C.byref(obj), float(grad_mul), grad.data_ptr(),
                _lib.ptr(logits), ws.data_ptr(), ws.numel(), self.precision, _lib.stream_ptr(self.device)), "dgdm_dyn_guidance")
        return (grad, logits) if want_logits else grad

    def score(self, x: torch.Tensor, objects_dev: torch.Tensor, objs_per_design: int, opt_obj: str,
              ori_range=(-1.0, 1.0), want_logits: bool = False, t: int = 0):
        """Predicted task score per (design, object) pair: mean objective over the ``grid_size`` profile
        orientations at t = 0, pos = (0,0) -- the input convention of get_convergence_centers
        (diffusion.py:509-516)."""
        self._check_widths(x, objects_dev)
        nd, P = x.shape
        n_obj = objects_dev.shape[0]
        grid = self._grid(ori_range, profile=True)
        obj = _objective_struct(opt_obj)
        scores = torch.empty((nd * objs_per_design,), dtype=torch.float32, device=self.device)
        logits = torch.empty((nd * objs_per_design * self.grid_size, 3), dtype=torch.float32,
                             device=self.device) if want_logits else None
        nb = self.lib.dgdm_dyn_guidance_workspace_bytes(C.byref(self.dyn.struct), nd, n_obj, objs_per_design,
                                                        C.byref(grid), self.precision)
        ws = self._workspace("dyn", nb)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dgdm_dyn_score(
                C.byref(self.dyn.struct), x.data_ptr(), nd, objects_dev.data_ptr(), n_obj, objs_per_design, None,
                self._t_frac(t), C.byref(grid), C.byref(obj), scores.data_ptr(), _lib.ptr(logits), ws.data_ptr(),
                ws.numel(), self.precision, _lib.stream_ptr(self.device)), "dgdm_dyn_score")
        return (scores, logits) if want_logits else scores

    def classifier_model(self, pts: torch.Tensor, ori: torch.Tensor, pos: torch.Tensor, timesteps: torch.Tensor,
                         object_vertices: torch.Tensor = None, *, object_codes: Optional[torch.Tensor] = None,
                         fps_starts: Optional[torch.Tensor] = None, opt_obj: Optional[str] = None,
                         return_grad: bool = False):
        """The dynamics network called on explicit rows, with the reference's own signature
        (``ProfileForward2DModel.forward`` profile_forward_2d.py:137-156 / ``ProfileForward3DModel.forward``
        profile_forward_3d.py:67-86; call sites diffusion.py:487,496,500,516):
        pts (N,P) [3D: (N,3,P), only row 1 is read], ori (N,1), pos (N,2), timesteps (N,) float in [0,1),
        object_vertices 2D (N,2V) / 3D (N,3,512)  ->  logits (N,3).
        3D per-row clouds run PointNet++ per row (as the reference does); pass ``object_codes`` (N,256) to reuse
        codes.  ``return_grad`` also returns d/dpts of sum_r objective(logits_r) for ``opt_obj``."""
        f = lambda a: a.to(device=self.device, dtype=torch.float32).contiguous()
        if self.mode == "point_3d":
            x = f(pts[:, 1, :]) if pts.dim() == 3 else f(pts)
            if object_codes is None:
                if object_vertices is None:
                    raise ValueError("object vertices not provided")
                clouds = object_vertices.to(self.device).permute(0, 2, 1)        # (N,512,3)
                object_codes = self.encode_objects(clouds, fps_starts)
            objs = f(object_codes)
        else:
            x = f(pts)
            if object_vertices is None:
                raise ValueError("object vertices not provided")
            objs = f(object_vertices.reshape(object_vertices.shape[0], -1))
        self._check_widths(x, objs)
        n = x.shape[0]
        if objs.shape[0] != n:
            raise ValueError(f"classifier_model: {objs.shape[0]} object rows for {n} design rows")
        o, p, t = f(ori.reshape(n)), f(pos.reshape(n, 2)), f(timesteps.reshape(n))
        logits = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        grad = torch.empty_like(x) if return_grad else None
        obj = _objective_struct(opt_obj) if opt_obj is not None else None
        if return_grad and obj is None:
            raise ValueError("return_grad needs opt_obj")
        ws = self._workspace("dyn", self.lib.dgdm_dyn_rows_workspace_bytes(C.byref(self.dyn.struct), n, self.precision))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dgdm_dyn_forward_rows(
                C.byref(self.dyn.struct), x.data_ptr(), o.data_ptr(), p.data_ptr(), t.data_ptr(), objs.data_ptr(), n,
                C.byref(obj) if obj is not None else None, logits.data_ptr(), _lib.ptr(grad), ws.data_ptr(), ws.numel(),
                self.precision, _lib.stream_ptr(self.device)), "dgdm_dyn_forward_rows")
        return (logits, grad) if return_grad else logits

    def best_of_n(self, scores: torch.Tensor, k: int = 1):
        """scores (n_obj, n_cand) -> (idx int64 (n_obj,k), best (n_obj,k)); ties to the lowest index (np.argmax)."""
        s = scores.to(device=self.device, dtype=torch.float32).contiguous()
        n_obj, n_cand = s.shape
        idx = torch.empty((n_obj, k), dtype=torch.int64, device=self.device)
        best = torch.empty((n_obj, k), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dgdm_best_of_n(s.data_ptr(), n_obj, n_cand, k, idx.data_ptr(), best.data_ptr(),
                                               _lib.stream_ptr(self.device)), "dgdm_best_of_n")
        return idx, best

    # ------------------------------------------------------------------------------------------
    def _object_for(self, object_vertices: Optional[torch.Tensor]) -> torch.Tensor:
        if object_vertices is None:
            raise ValueError("object vertices not provided")
        if self.object_vertices is not None:
            for i in range(self.object_vertices.shape[0]):
                ov = self.object_vertices[i]
                if ov.data_ptr() == object_vertices.data_ptr() and ov.shape == object_vertices.shape:
                    return self._obj_dev[i:i + 1]
        return self.encode_objects(object_vertices[None])

    def _convergence_row_coef(self, centers: torch.Tensor, batch: int) -> torch.Tensor:
        """Sign table of the 'convergence' objective, reproducing diffusion.py:445-452 as written -- including
        that it slices the grid-major logits (row = g*B + b, :479-484) as if they were candidate-major
        (SURVEY.md §8 f-2).  Returned in this library's pair-major row order (b*G + g)."""
        G = self.grid_size * self.num_pos ** 2
        np2 = self.num_pos ** 2
        half = (self.grid_size // 2) * np2
        coef_ref = np.zeros(batch * G, dtype=np.float32)
        for i, c in enumerate(np.asarray(centers.cpu()).astype(np.int64).tolist()):
            base = i * G
            coef_ref[base + _slicer_idx(G, c * np2 - half, c * np2)] += 1.0
            coef_ref[base + _slicer_idx(G, c * np2, c * np2 + half)] -= 1.0
        mine = coef_ref.reshape(G, batch).T.copy()        # [b, g] <- ref row g*B + b
        return torch.from_numpy(mine.reshape(-1)).to(self.device)

    def cond_fn(self, x, t, opt_obj: str = "rotate", object_vertices=None, ori_range=[-1.0, 1.0],
                convergence_centers=None) -> torch.Tensor:
        """``Diffusion.cond_fn`` (diffusion.py:473-504): (B,P,1), t -> d/dx sum_g objective, (B,P,1)."""
        tt = int(t.reshape(-1)[0]) if torch.is_tensor(t) else int(t)
        xs = x.detach().to(device=self.device, dtype=torch.float32)
        x2 = xs.reshape(xs.shape[0], -1).contiguous()
        obj_dev = self._object_for(object_vertices)
        row_coef = None
        if opt_obj == "convergence":
            if convergence_centers is None:
                raise ValueError("convergence objective needs convergence_centers")
            if self.mode == "point_3d":
                raise ValueError("convergence guidance is ill-defined in 3D in the reference (sub-batch slicing, "
                                 "diffusion.py:497 vs :448); not supported")
            row_coef = self._convergence_row_coef(convergence_centers, x2.shape[0])
        g = self.guidance(x2, tt, obj_dev, 1, opt_obj, ori_range, 1.0, row_coef)
        return g.reshape(x.shape)

    def profile_logits(self, designs: torch.Tensor, object_vertices: torch.Tensor, ori_range=(-1.0, 1.0)) -> torch.Tensor:
        """Forward-only profile pass -> (B, grid_size, 3) logits."""
        x2 = designs.to(device=self.device, dtype=torch.float32).reshape(designs.shape[0], -1).contiguous()
        _, lg = self.score(x2, self._object_for(object_vertices), 1, "rotate", ori_range, want_logits=True)
        return lg.reshape(x2.shape[0], self.grid_size, 3)

    def get_convergence_centers(self, unguided_sample, object_vertices, batch_size, ori_range=[-1.0, 1.0]) -> torch.Tensor:
        """diffusion.py:506-539: classify d_theta of the profile pass into {0,1,2}, take the centre of the
        longest converging run per candidate."""
        lg = self.profile_logits(unguided_sample, object_vertices, ori_range)[..., 0].cpu().numpy()   # (B, grid)
        thr = float(self.threshold_std[0])
        cls = np.where(lg > thr, 2.0, np.where(lg < -thr, 0.0, 1.0))
        return torch.tensor([_convergence_center(cls[b]) for b in range(batch_size)], dtype=torch.int64)

    # ------------------------------------------------------------------------------------------
    def classifier_scale(self, opt_obj: str, multi: bool = False) -> float:
        conv = opt_obj == "convergence" and not multi          # :549-560 vs :631-636
        if self.mode == "point":
            return SCALE_2D_CONV if conv else SCALE_2D
        return SCALE_3D_CONV if conv else SCALE_3D

    def _graphed(self, key, noise: torch.Tensor, batch_size: int, body) -> Dict[str, torch.Tensor]:
        """Run ``body(noise)`` through a CUDA graph captured once per ``key``.  The first call runs the body eagerly
        (which sizes every cached workspace and performs the library's one-time kernel-attribute setup), then captures
        it on a side stream with a static copy of the noise; later calls copy the new noise in and replay.  Outputs
        are cloned so they survive the next replay."""
        nz = self._check_noise(noise, batch_size)
        ent = self._graphs.get(key)
        if ent is None:
            static_in = nz.clone()
            body(static_in)                                      # eager warm-up
            torch.cuda.synchronize(self.device)
            with torch.cuda.device(self.device):             # capture on the device that owns the buffers
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    static_out = body(static_in)
            ent = (g, static_in, static_out, dict(self._ws))     # keep the captured workspaces alive if the cache regrows
            self._graphs[key] = ent
        g, static_in, static_out = ent[:3]
        static_in.copy_(nz)
        with torch.cuda.device(self.device):
            g.replay()
        return {k: v.clone() for k, v in static_out.items()}

    def _check_noise(self, noise: torch.Tensor, batch_size: int) -> torch.Tensor:
        if noise.shape[0] != batch_size or noise.shape[1] != self.num_points:
            raise ValueError(f"noise must be ({batch_size},{self.num_points},1), got {tuple(noise.shape)}")
        return noise.to(device=self.device, dtype=torch.float32).reshape(batch_size, -1).contiguous()

    def unguided_sample(self, noise: torch.Tensor) -> torch.Tensor:
        """diffusion.py:249-256."""
        x = self._check_noise(noise, noise.shape[0]).clone()
        for t in self.noise_scheduler.timesteps.tolist():
            eps = self.noise_pred_net(x, t)
            x = self.noise_scheduler.guided_step(eps, t, x, None, 0.0, out=x)
        return x.reshape(noise.shape)

    def denoise_from_data(self, tensor_data: torch.Tensor, seed: Optional[int] = None) -> Dict[str, float]:
        """The denoise-from-data validation metrics of ``validation_step`` (diffusion.py:179-201, 230-244): noise the
        data at t = num_inference_steps with RandomState(seed) noise, run the unguided schedule, report
        'val/noise pred loss', 'val/denoise loss', 'val/accuracy'."""
        B = tensor_data.shape[0]
        data = self._check_noise(tensor_data, B)
        rs = np.random.RandomState(self.seed if seed is None else seed)
        noise = torch.from_numpy(rs.randn(B, self.num_points, 1)).float().to(self.device).reshape(B, -1).contiguous()
        a = self.noise_scheduler.alphas_cumprod[self.num_inference_steps]
        # add_noise: sqrt(a) x0 + sqrt(1-a) eps  ==  K4 with an identity x0 stage (sqrt_1m_at = 0, sqrt_at = 1, no clip)
        sample = torch.empty_like(data)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dgdm_ddim_guided_update(sample.data_ptr(), data.data_ptr(), noise.data_ptr(), None,
                                                        data.numel(), 0.0, 1.0, float(a ** 0.5), float((1 - a) ** 0.5),
                                                        0.0, 0, _lib.stream_ptr(self.device)), "dgdm_ddim_guided_update")
        npl = 0.0
        for t in self.noise_scheduler.timesteps.tolist():
            eps = self.noise_pred_net(sample, t)
            npl += float(torch.mean((eps - noise) ** 2))
            sample = self.noise_scheduler.guided_step(eps, t, sample, None, 0.0)
        return {"val/noise pred loss": npl / self.num_inference_steps,
                "val/denoise loss": float(torch.mean((sample - data) ** 2)),
                "val/accuracy": float(torch.mean((torch.abs(sample - data) < 0.01).float()))}

    def predicted_objectives(self, designs: torch.Tensor, object_vertices: torch.Tensor, opt_obj: str,
                             ori_range=(-1.0, 1.0)):
        """Prediction-side stand-in for the reference's MuJoCo tables (diffusion.py:583-598): per-candidate objective
        dicts from the profile pass and the best candidate per metric key (dgdm_b200.metrics)."""
        from . import metrics as M
        lg = self.profile_logits(designs, object_vertices, ori_range).cpu().numpy()
        objs = M.predicted_objectives(lg, opt_obj, self.mode)
        return objs, M.get_best_ids_all_metrics(objs, opt_obj)

    def guided_sample(self, batch_idx, batch_size, noise, save_dir=None, opt_obj="rotate", ori_range=[-1.0, 1.0],
                      unguided_sample=None, top_k: int = 1, trace: Optional[list] = None, cuda_graph: bool = False):
        """Per-object guided sampling (diffusion.py:541-576): every object restarts from the same ``noise``.
        Returns {'designs' (n_obj,B,P,1), 'scores' (n_obj,B), 'best_ids' (n_obj,top_k), 'best_scores'}.

        ``cuda_graph=True`` captures the whole pass (every denoiser forward, guidance step, DDIM update, the scoring
        pass and the selection: a few hundred launches) into one CUDA graph on first use and replays it afterwards;
        results are bit-identical to the eager path.  It pays when the pass is launch-bound (small batches)."""
        if opt_obj not in OBJECTIVES:
            raise ValueError("opt obj not supported")
        if cuda_graph and trace is None and opt_obj != "convergence":
            key = ("per_object", batch_size, self._obj_dev.shape[0], opt_obj, tuple(ori_range), top_k, self._obj_version)
            return self._graphed(key, noise, batch_size, lambda nz: self.guided_sample(
                batch_idx, batch_size, nz, save_dir, opt_obj, ori_range, None, top_k))
        n_obj = self._obj_dev.shape[0]
        B, P = batch_size, self.num_points
        x0 = self._check_noise(noise, B)
        # Bound the per-launch working set (U / dU / encoder intermediates: ~10 KB per design; with DGDM_TRUNK_SLOTS=1
        # also one H1-wide partial sum per (pair, tile), ~20 KB per design at G = 1125): a 1024-object x 512-candidate
        # sweep on one GPU runs as four launches per step.  Objects are independent (diffusion.py:561-570), so chunks of
        # objects give identical results (test_object_chunking_is_invisible).
        per_launch = max(1, self.max_designs_per_launch // B)
        if n_obj > per_launch and opt_obj != "convergence" and trace is None:
            parts = [self._guided_sample_objects(self._obj_dev[o:o + per_launch], x0, opt_obj, ori_range, None, top_k, None)
                     for o in range(0, n_obj, per_launch)]
            return {k: torch.cat([p[k] for p in parts], dim=0) for k in parts[0]}
        row_coef = None
        if opt_obj == "convergence":
            if unguided_sample is None:
                raise ValueError("convergence objective needs unguided_sample")
            if self.mode == "point_3d":
                raise ValueError("convergence guidance is not supported in 3D (see cond_fn)")
            row_coef = torch.cat([self._convergence_row_coef(
                self.get_convergence_centers(unguided_sample, self.object_vertices[o], B, ori_range), B)
                for o in range(n_obj)])
        return self._guided_sample_objects(self._obj_dev, x0, opt_obj, ori_range, row_coef, top_k, trace)

    def _guided_sample_objects(self, obj_dev: torch.Tensor, x0: torch.Tensor, opt_obj: str, ori_range, row_coef,
                               top_k: int, trace: Optional[list]) -> Dict[str, torch.Tensor]:
        """The step loop of guided_sample (diffusion.py:570-576) for the objects in ``obj_dev``, all at once."""
        n_obj = obj_dev.shape[0]
        B, P = x0.shape
        scale = self.classifier_scale(opt_obj)
        x = x0.repeat(n_obj, 1).contiguous()                     # design d = o*B + b (object-major)
        for t in self.noise_scheduler.timesteps.tolist():
            eps = self.noise_pred_net(x, t)
            grad = self.guidance(x, t, obj_dev, 1, opt_obj, ori_range, 1.0, row_coef)
            nxt = self.noise_scheduler.guided_step(eps, t, x, grad, scale)
            if trace is not None:
                trace.append(dict(t=t, eps=eps.reshape(n_obj, B, P, 1), grad=grad.reshape(n_obj, B, P, 1),
                                  sample=nxt.reshape(n_obj, B, P, 1)))
            x = nxt
        score_obj = "rotate_counterclockwise" if opt_obj == "convergence" else opt_obj
        scores = self.score(x, obj_dev, 1, score_obj, ori_range).reshape(n_obj, B)
        idx, best = self.best_of_n(scores, top_k)
        return {"designs": x.reshape(n_obj, B, P, 1), "scores": scores, "best_ids": idx, "best_scores": best}

    def guided_sample_multi_object(self, batch_idx, batch_size, noise, save_dir=None, opt_obj="rotate",
                                   ori_range=[-1.0, 1.0], top_k: int = 1, trace: Optional[list] = None,
                                   cuda_graph: bool = False):
        """One trajectory guided by the object-averaged gradient (diffusion.py:621-647).
        Returns {'designs' (B,P,1), 'scores' (B,) mean over objects, 'best_ids' (top_k,), 'best_scores'}.
        ``cuda_graph``: as in :meth:`guided_sample`."""
        if opt_obj not in OBJECTIVES or opt_obj == "convergence":
            raise ValueError("opt obj not supported")
        if cuda_graph and trace is None:
            key = ("multi_object", batch_size, self._obj_dev.shape[0], opt_obj, tuple(ori_range), top_k, self._obj_version)
            return self._graphed(key, noise, batch_size, lambda nz: self.guided_sample_multi_object(
                batch_idx, batch_size, nz, save_dir, opt_obj, ori_range, top_k))
        n_obj = self._obj_dev.shape[0]
        B, P = batch_size, self.num_points
        scale = self.classifier_scale(opt_obj, multi=True)
        x = self._check_noise(noise, B).clone()
        for t in self.noise_scheduler.timesteps.tolist():
            eps = self.noise_pred_net(x, t)
            grad = self.guidance(x, t, self._obj_dev, n_obj, opt_obj, ori_range, 1.0 / n_obj)
            nxt = self.noise_scheduler.guided_step(eps, t, x, grad, scale)
            if trace is not None:
                trace.append(dict(t=t, eps=eps.reshape(B, P, 1), grad=grad.reshape(B, P, 1), sample=nxt.reshape(B, P, 1)))
            x = nxt
        scores = self.score(x, self._obj_dev, n_obj, opt_obj, ori_range).reshape(B, n_obj).mean(dim=1)
        idx, best = self.best_of_n(scores[None], top_k)
        return {"designs": x.reshape(B, P, 1), "scores": scores, "best_ids": idx[0], "best_scores": best[0]}
