"""Prediction-side scoring: the reference ranks designs with MuJoCo roll-outs
(dynamics/sim_test_mj.py -> dynamics/metrics.py:metric2objective -> generator/diffusion.py:391-428); the simulator
is CPU-only and out of scope, so this module builds the same tables from the dynamics network's own profile pass
(logits of ``Diffusion.profile_logits``), with the reference's thresholds, units, key names and arg-min/max rules,
so the GPU best-of-N is a like-for-like stand-in.  Keys that only a roll-out can produce (``final_*``) appear when
the caller supplies the simulator's ``final_*`` fields (then the tables are the reference's key for key: pinned by
tests/test_metrics_golden.py against the reference's own functions); ``max_convergence_range_*`` is not emitted.

Also ``export_designs``: de-normalise control points to the metres the simulators consume.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

# generator/diffusion.py:116-117, dynamics/dataloader.py:11-15
THRESHOLD = {"point": (0.03, 0.002, 0.003), "point_3d": (0.02, 0.001, 0.001)}
STD = {"point": (0.0565, 0.0026, 0.0047), "point_3d": (0.0312, 0.0016, 0.0026)}


def predicted_metrics(logits: np.ndarray, mode: str = "point") -> Dict[str, np.ndarray]:
    """logits (..., n_rot, 3) = (d_theta, d_x, d_y)/std over the profile orientations of ONE candidate (or batched).
    Returns the fields of the simulator's metric dict that a prediction can fill (dynamics/sim_test_mj.py:207-216):
    ``profile``/``profile_x``/``profile_y`` in {0,1,2} (0: negative beyond threshold, 1: inside, 2: positive;
    generator/diffusion.py:532), ``delta_theta`` in degrees, ``delta_pos`` in metres."""
    lg = np.asarray(logits, dtype=np.float64)
    thr, std = np.asarray(THRESHOLD[mode]), np.asarray(STD[mode])
    phys = lg * std
    cls = np.where(phys > thr, 2, np.where(phys < -thr, 0, 1)).astype(np.int64)
    return {"profile": cls[..., 0], "profile_x": cls[..., 1], "profile_y": cls[..., 2],
            "delta_theta": phys[..., 0] * 180.0 / np.pi, "delta_pos": phys[..., 1:3]}


# objective -> (rotation class or None, axis ('x'|'y') or None, translation class or None)
_SPEC = {
    "rotate_clockwise": (0, None, None), "rotate_counterclockwise": (2, None, None),
    "shift_up": (None, "x", 0), "shift_down": (None, "x", 2), "shift_left": (None, "y", 0), "shift_right": (None, "y", 2),
    "clockwise_up": (0, "x", 0), "clockwise_down": (0, "x", 2), "clockwise_left": (0, "y", 0), "clockwise_right": (0, "y", 2),
    "counterclockwise_up": (2, "x", 0), "counterclockwise_down": (2, "x", 2), "counterclockwise_left": (2, "y", 0),
    "counterclockwise_right": (2, "y", 2),
}
_ROT = {0: "clockwise", 2: "counterclockwise"}
_TRN = {("x", 0): "up", ("x", 2): "down", ("y", 0): "left", ("y", 2): "right"}


def metric2objective(metric: Dict[str, np.ndarray], objective: str) -> Dict[str, float]:
    """dynamics/metrics.py:67-233 for one candidate (metric fields are (n_rot,) arrays): same keys, in the same order,
    same dtypes (class counts are ints, rates and means floats).  The roll-out keys (``final_delta_theta[_abs]``,
    ``final_pos_x|y``) are emitted when the metric dict carries the simulator's ``final_*`` fields and left out for a
    prediction-only dict (``predicted_metrics``)."""
    p, px, py = metric["profile"], metric["profile_x"], metric["profile_y"]
    has_final = "final_delta_theta" in metric
    if objective in ("rotate", "rotate_in_place"):
        out = {"success_rate": float(np.mean((p == 0) | (p == 2), dtype=np.float32)),
               "num_zero_classes": int(np.sum(p == 1)), "delta_theta_abs": float(np.mean(np.abs(metric["delta_theta"])))}
        if has_final:
            out["final_delta_theta_abs"] = float(np.mean(np.abs(metric["final_delta_theta"])))
        return out
    if objective not in _SPEC:
        raise ValueError("opt obj not supported")
    rot, axis, trn = _SPEC[objective]
    out: Dict[str, float] = {}
    ok = np.ones_like(p, dtype=bool)
    if rot is not None:
        n_rot = int(np.sum(p == rot))
        ok &= p == rot
    if axis is not None:
        prof = px if axis == "x" else py
        n_trn = int(np.sum(prof == trn))
        ok &= prof == trn
    out["success_rate"] = float(np.mean(ok, dtype=np.float32))
    if rot is not None and axis is not None:
        out[f"num_{_ROT[rot]}_{_TRN[(axis, trn)]}_classes"] = n_rot + n_trn
    if rot is not None:
        out[f"num_{_ROT[rot]}_classes"] = n_rot
        out["delta_theta"] = float(np.mean(metric["delta_theta"]))
        if has_final:
            out["final_delta_theta"] = float(np.mean(metric["final_delta_theta"]))
    if axis is not None:
        col = 0 if axis == "x" else 1
        out[f"num_{_TRN[(axis, trn)]}_classes"] = n_trn
        out["delta_pos_" + axis] = float(np.mean(metric["delta_pos"][..., col]))
        if "final_pos" in metric:
            out["final_pos_" + axis] = float(np.mean(np.asarray(metric["final_pos"])[..., col]))
    return out


def _direction(key: str, objective: str) -> str:
    """arg-min or arg-max per key, as generator/diffusion.py:391-428."""
    if key == "num_zero_classes":
        return "min"
    if key.startswith("num_") or key in ("success_rate", "delta_theta_abs", "final_delta_theta_abs"):
        return "max"
    rot, axis, trn = _SPEC[objective]
    if key in ("delta_theta", "final_delta_theta"):
        return "min" if rot == 0 else "max"              # clockwise = negative d_theta
    if key in ("delta_pos_x", "delta_pos_y", "final_pos_x", "final_pos_y"):
        return "min" if trn == 0 else "max"              # up / left = negative
    raise KeyError(key)


def _check_objective(opt_obj: str) -> None:
    if opt_obj not in _SPEC and opt_obj not in ("rotate", "rotate_in_place"):
        raise ValueError("opt obj not supported")        # generator/diffusion.py:388,425


def get_best_ids_all_metrics(objectives: Sequence[Dict[str, float]], opt_obj: str = "rotate") -> Dict[str, int]:
    """generator/diffusion.py:391-428: index of the best candidate per metric key, first occurrence on ties
    (np.argmax / np.argmin); keys in the reference's order, ``success_rate`` last."""
    _check_objective(opt_obj)
    best = {}
    for key in [k for k in objectives[0] if k != "success_rate"] + ["success_rate"]:
        vals = np.asarray([o[key] for o in objectives])
        best[key] = int(np.argmin(vals) if _direction(key, opt_obj) == "min" else np.argmax(vals))
    return best


def get_average_best_ids(objectives: Sequence[Dict[str, float]], opt_obj: str = "rotate") -> int:
    """generator/diffusion.py:354-389: the single ranking key of each objective -- fewest in-threshold classes for
    'rotate', the most classes of the wanted kind otherwise (the combined count for the 8 combination objectives)."""
    _check_objective(opt_obj)
    if opt_obj in ("rotate", "rotate_in_place"):
        return int(np.argmin([o["num_zero_classes"] for o in objectives]))
    rot, axis, trn = _SPEC[opt_obj]
    name = "_".join(([_ROT[rot]] if rot is not None else []) + ([_TRN[(axis, trn)]] if axis is not None else []))
    return int(np.argmax([o[f"num_{name}_classes"] for o in objectives]))


def get_best_ids(objectives: Sequence[Dict[str, float]], num_grippers: int, num_objects: int,
                 opt_obj: str = "rotate") -> List[Dict[str, int]]:
    """generator/diffusion.py:346-352: per-object blocks of ``num_grippers`` candidates, ids offset by the block start."""
    out = []
    for idx in range(num_objects):
        best = get_best_ids_all_metrics(objectives[idx * num_grippers:(idx + 1) * num_grippers], opt_obj)
        out.append({k: v + idx * num_grippers for k, v in best.items()})
    return out


def predicted_objectives(logits: np.ndarray, opt_obj: str, mode: str = "point") -> List[Dict[str, float]]:
    """logits (B, n_rot, 3) -> one objective dict per candidate."""
    m = predicted_metrics(logits, mode)
    return [metric2objective({k: v[b] for k, v in m.items()}, opt_obj) for b in range(np.asarray(logits).shape[0])]


def export_designs(designs: np.ndarray, mode: str = "point") -> np.ndarray:
    """Normalised control points (B,P,1) in [-1,1] -> the arrays the simulators consume, in metres.
    2D (dynamics/sim_test_mj.py:257-262): (B,P,2) with x = linspace(-0.12,0.12,P/2) twice, y = p*0.03 - 0.015.
    3D (dynamics/sim_test_mj_3d.py:235-237): (B,P) with y = p*0.05 - 0.05."""
    d = np.asarray(designs, dtype=np.float64)
    d = d.reshape(d.shape[0], -1)
    if mode == "point_3d":
        return (d * 0.05 - 0.05).astype(np.float32)
    half = d.shape[1] // 2
    x = np.linspace(-0.12, 0.12, half)
    x = np.concatenate([x, x])[None, :, None].repeat(d.shape[0], 0)
    return np.concatenate([x, (d * 0.03 - 0.015)[..., None]], axis=-1).astype(np.float32)
