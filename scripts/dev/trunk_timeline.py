"""In-kernel clock64 timeline of one tile of tc_trunk_kernel (CTA 0, third tile).

Build with the trace hooks, run on a GPU, print per segment what the MMA issuer, two epilogue warps and the producer did:
    DGDM_NVCC_EXTRA=-DDGDM_TRUNK_TRACE python -m dgdm_b200.build -f
    python scripts/dev/trunk_timeline.py {fp32|bf16|fp16} [2d|3d]
Trace slots (dynamics_tc.cu, TR()): issuer sg*64 + j*4 + {0: before the a_ready wait of weight tile j, 1: after it,
2: ring slot full, 3: the tile's MMAs issued and committed}; epilogue warp 0 at 2048 + sg*16 + {0: accumulator
complete, 1..4: k-block 0..3 handed over, 6..10: k-block 0 chain (tcgen05.ld done, converted, tcgen05.st issued,
st complete, arrived)}, warp 15 at 4096 + ...; producer 6144 + sg*8 + slot: ring slot free.
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
x3 = prec == "fp32"
ntile = 8 if x3 else 4                     # weight tiles (ring slots) per 256-wide segment
n_seg = (17 if is3d else 15) - (1 if x3 else 0)
t0 = tr[0]                                 # issuer, segment 0, first tile, before its a_ready wait
print(f"{prec} {'3D' if is3d else '2D'}: {n_seg} segments; one weight tile = 4 (hi.hi) [+4 lo.hi | 4 hi.lo] N=256 MMAs of 128 cycles")
print("MMA issuer, cycles relative to the tile's first stamp; per weight tile: wait for the operand k-block, wait for "
      "the weights, issue + commit")
prev_end = 0
for sg in range(n_seg):
    b = sg * 64
    row = tr[b: b + ntile * 4].reshape(ntile, 4)
    if row[0, 3] == 0:
        continue
    wa = (row[:, 1] - row[:, 0]).tolist(); wf = (row[:, 2] - row[:, 1]).tolist(); iss = (row[:, 3] - row[:, 2]).tolist()
    start, end = int(row[0, 0] - t0), int(row[-1, 3] - t0)
    print(f"seg {sg:2d}: start {start:7d} end {end:7d} dur {end - start:6d} gap_before {start - prev_end:5d} | wait_a {wa} | "
          f"wait_full {wf} | issue {iss}")
    prev_end = end
print(f"tile: {prev_end} cycles from first stamp to last commit")
for name, base in (("epilogue warp 0", 2048), ("epilogue warp 15", 4096)):
    print(name, ": accumulator complete (rel. to tile start), then k-block hand-overs relative to it"
          + ("; k-block 0 chain: ld done, converted, st issued, st complete, arrived" if base == 2048 else ""))
    for sg in range(n_seg):
        r = tr[base + sg * 16: base + sg * 16 + 11]
        if r[1] == 0:
            print(f"  seg {sg:2d}: wake {int(r[0] - t0):7d}")
            continue
        fine = f" | kb0 chain +{[int(v - r[0]) for v in r[6:11]]}" if base == 2048 and r[6] else ""
        # when did the issuer see k-block 0 of the NEXT segment?  (its stamp 1 of tile 0)
        nxt = tr[(sg + 1) * 64 + 1] if sg + 1 < n_seg else 0
        seen = f" | issuer past a_ready[0] +{int(nxt - r[0])}" if nxt else ""
        print(f"  seg {sg:2d}: wake {int(r[0] - t0):7d} | kb0..3 done +{[int(v - r[0]) for v in r[1:5]]}{fine}{seen}")
