import sys, torch
import os; R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path[:0] = [R, R + "/tests"]
from dgdm_b200 import synthetic as syn
from dgdm_b200.diffusion import Diffusion
from dgdm_b200.scheduler import DDIMScheduler
prec, P, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dm = Diffusion(syn.unet1d_state_dict(0), DDIMScheduler(15), 5, mode="point", num_points=14, class_cond=False, precision=prec)
dm.num_points = P
x = syn.initial_noise(n, P).cuda()
for _ in range(2):
    dm.noise_pred_net(x, 9)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    dm.noise_pred_net(x, 9)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
fl = n * (54.83e6 if P == 14 else 164.06e6)
print(f"unet {prec} P={P} n={n}: {ms:.3f} ms, {fl/ms/1e9:.1f} TFLOP/s algorithmic")
