import sys, torch
import os; R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path[:0] = [R, R + "/tests", R + "/oracle"]
from dgdm_b200 import synthetic as syn
from test_gpu_parity import make2d
prec = sys.argv[1]
objs = syn.objects_2d(64)
dm = make2d(prec, objs, 36, 5)
x = syn.initial_noise(256, 14)[..., 0].cuda().repeat(64, 1).contiguous()
for _ in range(5):
    dm.guidance(x, 6, dm._obj_dev, 1, "rotate_clockwise")
torch.cuda.synchronize()
