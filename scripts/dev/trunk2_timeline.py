"""In-kernel clock64 timeline of one tile PAIR of tc_trunk2_kernel (CTA 0, second pair).

    DGDM_NVCC_EXTRA=-DDGDM_TRUNK_TRACE python -m dgdm_b200.build -f
    python scripts/dev/trunk2_timeline.py {bf16|fp16} [2d|3d]
Trace slots (dynamics_tc.cu, TR2()): issuer item i at i*16 + {0: start, 1: accumulator free, 3+3kb: stage kb full,
4+3kb: k-block issued + committed, 14: end}; epilogue warp 0 at 2048 + i*16 + {0: start,
1: accumulator complete (d_ready), 2..5: k-block kb handed over, 6: end, 8..11: ring slot for k-block kb free}.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [R, R + "/tests"]
from dgdm_b200 import _lib, synthetic as syn  # noqa: E402
from dgdm_b200.diffusion import Diffusion  # noqa: E402
from dgdm_b200.scheduler import DDIMScheduler  # noqa: E402

prec = sys.argv[1]
is3d = len(sys.argv) > 2 and sys.argv[2] == "3d"
n_obj = 64
if is3d:
    dm = Diffusion(syn.unet1d_state_dict(0), DDIMScheduler(15), 5, mode="point_3d", num_points=42,
                   classifier_model=syn.dynamics3d_state_dict(0), grid_size=45, num_pos=5,
                   object_vertices=syn.objects_3d(n_obj), object_ids=list(range(n_obj)), fps_starts=syn.fps_starts(n_obj),
                   precision=prec, device="cuda:0")
    x = syn.initial_noise(128, 42)[..., 0].cuda().repeat(n_obj, 1).contiguous()
else:
    dm = Diffusion(syn.unet1d_state_dict(0), DDIMScheduler(15), 5, mode="point", num_points=14,
                   classifier_model=syn.dynamics2d_state_dict(0), grid_size=36, num_pos=5,
                   object_vertices=syn.objects_2d(n_obj), object_ids=list(range(n_obj)), precision=prec, device="cuda:0")
    x = syn.initial_noise(256, 14)[..., 0].cuda().repeat(n_obj, 1).contiguous()
for _ in range(3):
    dm.guidance(x, 6, dm._obj_dev, 1, "rotate_clockwise")
torch.cuda.synchronize()
lib = _lib.lib()
buf = (C.c_longlong * 8192)()
lib.dgdm_trunk_trace_read.argtypes = [C.POINTER(C.c_longlong), C.c_int32]
assert lib.dgdm_trunk_trace_read(buf, 8192) == 0
tr = np.array(buf[:], dtype=np.int64)
print("DGDM_TRUNK2 =", os.environ.get("DGDM_TRUNK2"))
n_m = (34 if is3d else 30)
n_e = (36 if is3d else 32)
t0 = min(int(v) for v in tr[:4096] if v > 0)
print(f"{prec} {'3D' if is3d else '2D'} two-tile kernel; cycles relative to the first stamp of the pair")
print("MMA issuer items: start | wait d_free | per k-block: wait full, issue + commit | dur")
prev = None
for i in range(n_m):
    r = tr[i * 16: i * 16 + 16]
    if r[14] == 0:
        continue
    kb = [(int(r[3 + 3 * k] - (r[1] if k == 0 else r[4 + 3 * (k - 1)])), int(r[4 + 3 * k] - r[3 + 3 * k])) for k in range(4)]
    gap = 0 if prev is None else int(r[0] - prev)
    print(f"  M{i:2d}: start {int(r[0]-t0):7d} (gap {gap:5d}) dfree {int(r[1]-r[0]):5d} | " + " ".join(f"[{w:4d},{s:4d}]" for w, s in kb)
          + f" | dur {int(r[14]-r[0]):5d}")
    prev = r[14]
print(f"issuer: {int(prev - tr[0])} cycles for the pair = {int(prev - tr[0]) // 2} per tile")
print("epilogue warp 0 items: start | wait d_ready | k-block hand-overs rel. to d_ready (ring-slot waits) | dur")
for i in range(n_e):
    r = tr[2048 + i * 16: 2048 + i * 16 + 16]
    if r[0] == 0:
        continue
    ref = r[1] if r[1] > 0 else r[0]
    hand = [int(r[2 + k] - ref) if r[2 + k] > 0 else -1 for k in range(4)]
    slotw = [int(r[8 + k] - ref) if r[8 + k] > 0 else -1 for k in range(4)]
    print(f"  E{i:2d}: start {int(r[0]-t0):7d} wait_d {int(ref-r[0]):5d} | handed {hand} slot-free-at {slotw} | dur {int(r[6]-r[0]) if r[6] > 0 else -1:5d}")
