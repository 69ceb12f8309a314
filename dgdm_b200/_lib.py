"""ctypes binding of the C ABI in include/dgdm_b200.h.

There is no CPU fallback: if ``libdgdm_b200.so`` is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdgdm_b200.so")

PREC_FP32_SIMT, PREC_BF16X3, PREC_BF16, PREC_FP16, PREC_FP16X3 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32_simt": PREC_FP32_SIMT, "fp32": PREC_BF16X3, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16, "fp16": PREC_FP16, "fp16x3": PREC_FP16X3}

fp = C.c_void_p  # device pointers travel as integers


class DynWeights(C.Structure):
    _fields_ = ([("P", C.c_int32), ("H1", C.c_int32), ("obj_dim", C.c_int32), ("is_3d", C.c_int32)] +
                [(n, fp) for n in ("ge_w0", "ge_b0", "ge_w1", "ge_b1", "oe_w0", "oe_b0", "oe_w1", "oe_b1",
                                   "te_w0", "te_b0", "te_w1", "te_b1", "w1_obj", "w1_ctrl", "w1_ctrl_t", "w1_pose",
                                   "w1_time", "b1", "ge_w1_t", "ge_w0_t")] +
                [("wl", fp * 7), ("wl_t", fp * 7), ("bl", fp * 7), ("w_out", fp), ("b_out", fp), ("tc_image", fp)])


class PoseGrid(C.Structure):
    _fields_ = [("ori_lo", C.c_float), ("ori_hi", C.c_float), ("grid_size", C.c_int32), ("num_pos", C.c_int32),
                ("pos_zero", C.c_int32)]


class Objective(C.Structure):
    _fields_ = [("c", C.c_float * 3), ("sq0", C.c_float), ("row_coef", fp)]


class UnetResBlock(C.Structure):
    _fields_ = ([("cin", C.c_int32), ("cout", C.c_int32)] +
                [(n, fp) for n in ("conv0_w", "conv0_b", "gn0_w", "gn0_b", "conv1_w", "conv1_b", "gn1_w", "gn1_b",
                                   "film_w", "film_b", "res_w", "res_b")])


class UnetWeights(C.Structure):
    _fields_ = ([(n, fp) for n in ("se_w0", "se_b0", "se_w1", "se_b1")] + [("blocks", UnetResBlock * 8)] +
                [(n, fp) for n in ("down_w", "down_b", "up_w", "up_b", "fin_w", "fin_b", "fin_gn_w", "fin_gn_b",
                                   "out_w", "out_b", "tc_image")])


class PointNet2Weights(C.Structure):
    _fields_ = [("w", fp * 5), ("b", fp * 5)]


EXPORTS = {
    # name: (restype, argtypes)
    "dgdm_last_error": (C.c_char_p, []),
    "dgdm_abi_version": (C.c_int, []),
    "dgdm_launch_count": (C.c_uint64, []),
    "dgdm_trunk_timing": (C.c_int, [C.c_int32]),
    "dgdm_trunk_timing_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "dgdm_trunk_timing_read_split": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "dgdm_ddim_guided_update": (C.c_int, [fp, fp, fp, fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                          C.c_float, C.c_int, fp]),
    "dgdm_dyn_tc_image_bytes": (C.c_size_t, [C.c_int32]),
    "dgdm_dyn_pack_tc": (C.c_int, [C.POINTER(DynWeights), fp, fp]),
    "dgdm_dyn_guidance_workspace_bytes": (C.c_size_t, [C.POINTER(DynWeights), C.c_int32, C.c_int32, C.c_int32,
                                                       C.POINTER(PoseGrid), C.c_int32]),
    "dgdm_dyn_guidance": (C.c_int, [C.POINTER(DynWeights), fp, C.c_int32, fp, C.c_int32, C.c_int32, fp, C.c_float,
                                    C.POINTER(PoseGrid), C.POINTER(Objective), C.c_float, fp, fp, fp, C.c_size_t,
                                    C.c_int32, fp]),
    "dgdm_dyn_score": (C.c_int, [C.POINTER(DynWeights), fp, C.c_int32, fp, C.c_int32, C.c_int32, fp, C.c_float,
                                 C.POINTER(PoseGrid), C.POINTER(Objective), fp, fp, fp, C.c_size_t, C.c_int32, fp]),
    "dgdm_dyn_rows_workspace_bytes": (C.c_size_t, [C.POINTER(DynWeights), C.c_int64, C.c_int32]),
    "dgdm_dyn_forward_rows": (C.c_int, [C.POINTER(DynWeights), fp, fp, fp, fp, fp, C.c_int64, C.POINTER(Objective), fp, fp,
                                        fp, C.c_size_t, C.c_int32, fp]),
    "dgdm_unet1d_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "dgdm_unet1d_forward": (C.c_int, [C.POINTER(UnetWeights), fp, C.c_int32, C.c_int32, C.c_int32, fp, fp,
                                      C.c_size_t, C.c_int32, fp]),
    "dgdm_unet_tc_image_bytes": (C.c_size_t, [C.POINTER(UnetWeights)]),
    "dgdm_unet_pack_tc": (C.c_int, [C.POINTER(UnetWeights), fp, fp]),
    "dgdm_linear_tc_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "dgdm_linear_tc": (C.c_int, [fp, C.c_int64, fp, fp, fp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, fp, C.c_size_t, fp]),
    "dgdm_pointnet2_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "dgdm_pointnet2_encode": (C.c_int, [C.POINTER(PointNet2Weights), fp, C.c_int32, C.c_int32, fp, fp, fp,
                                        C.c_size_t, fp]),
    "dgdm_best_of_n": (C.c_int, [fp, C.c_int32, C.c_int32, C.c_int32, fp, fp, fp]),
    "dgdm_linear_f32": (C.c_int, [fp, C.c_int64, fp, fp, fp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                  fp]),
}

_LIB: Optional[C.CDLL] = None


class DgdmError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the CUDA library; raise loudly if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise DgdmError(f"{LIB_PATH} is missing: build it with `python -m dgdm_b200.build` "
                            "(there is no CPU or PyTorch fallback for this path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            f = getattr(l, name)          # AttributeError if the symbol is not exported
            f.restype, f.argtypes = res, args
        if l.dgdm_abi_version() != 1:
            raise DgdmError("libdgdm_b200.so ABI version mismatch")
        _LIB = l
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dgdm_last_error().decode("utf-8", "replace")
        raise DgdmError(f"{what or 'dgdm call'} failed (code {rc}): {msg}")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "dgdm_b200 kernels need contiguous CUDA tensors"
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    """Current stream of ``device`` (default: the current device).  Callers launch inside
    ``torch.cuda.device(device)`` so that the library's cudaGetDevice()/launch context matches the pointers."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream
