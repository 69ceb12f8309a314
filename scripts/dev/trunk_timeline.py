import sys, ctypes as C, numpy as np, torch
import os; R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path[:0] = [R, R + "/tests", R + "/oracle"]
from dgdm_b200 import synthetic as syn, _lib
from test_gpu_parity import make2d
prec = sys.argv[1]
objs = syn.objects_2d(64)
dm = make2d(prec, objs, 36, 5)
x = syn.initial_noise(256, 14)[..., 0].cuda().repeat(64, 1).contiguous()
for _ in range(3):
    dm.guidance(x, 6, dm._obj_dev, 1, "rotate_clockwise")
torch.cuda.synchronize()
l = _lib.lib()
buf = (C.c_longlong * 8192)()
l.dgdm_trunk_trace_read.argtypes = [C.POINTER(C.c_longlong), C.c_int32]
assert l.dgdm_trunk_trace_read(buf, 8192) == 0
tr = np.array(buf[:], dtype=np.int64)
t0 = tr[0]
x3 = prec == "fp32"
ntile = 8 if x3 else 4
print("MMA issuer (cycles rel. to segment 0 start): per segment: start, [per W tile: wait_a, wait_full, issue] ... ")
prev_end = None
for sg in range(15):
    row = tr[sg * 64: sg * 64 + ntile * 4].reshape(ntile, 4) - t0
    seg_start = row[0, 0]; seg_end = row[-1, 3]
    wa = (row[:, 1] - row[:, 0]); wf = (row[:, 2] - row[:, 1]); iss = (row[:, 3] - row[:, 2])
    print(f"seg {sg:2d}: start {seg_start:7d} dur {seg_end - seg_start:6d} | wait_a {wa.tolist()} | wait_full {wf.tolist()} | issue {iss.tolist()}")
for name, base in (("epi warp0", 2048), ("epi warp15", 4096)):
    print(name, "(d_ready wake, then signal kb0..3), rel. to segment-0 start")
    for sg in range(15):
        r = tr[base + sg * 16: base + sg * 16 + 5] - t0
        print(f"  seg {sg:2d}: wake {r[0]:7d} | kb done +{(r[1:] - r[0]).tolist()}")
print("producer: time each ring slot became free (MMAs that read it completed), per segment")
for sg in range(15):
    r = tr[6144 + sg * 8: 6144 + sg * 8 + ntile] - t0
    print(f"  seg {sg:2d}: {r.tolist()}")
