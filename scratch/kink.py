import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import dgdm_oracle as orc
from dgdm_b200 import synthetic as syn
g2 = dict(np.load("tests/golden/golden_2d.npz"))
objs = torch.from_numpy(g2["objects"])
samp = orc.OracleSampler("point", syn.unet1d_state_dict(0), syn.dynamics2d_state_dict(0), objs, 6, 2)
x = torch.from_numpy(g2["loop_rotate_sample_o1_s2"])   # input of step 3
want = g2["loop_rotate_grad_o1_s3"]
g = samp.cond_fn(x, 3, "rotate", 1)
print("oracle vs golden", np.abs(g.numpy()-want).reshape(4,-1).max(1))
for eps in (1e-7, -1e-7, 3e-7, 1e-6):
    gp = samp.cond_fn(x*(1+eps), 3, "rotate", 1)
    print("perturb", eps, np.abs(gp.numpy()-want).reshape(4,-1).max(1))
# fp64 oracle
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in syn.dynamics2d_state_dict(0).items()}
s64 = orc.OracleSampler("point", syn.unet1d_state_dict(0), sd64, objs.double(), 6, 2)
torch.set_default_dtype(torch.float64)
g64 = s64.cond_fn(x.double(), 3, "rotate", 1)
print("fp64 vs golden", np.abs(g64.numpy()-want).reshape(4,-1).max(1))
